"""Frame-dataset input boundary: enumerate and gather 21^3 x C voxel frames from an
aposteriori ``.hdf5`` file (schema documented at /root/reference/design_utils/utils.py:238-251).

Same names and return conventions as ``create_flat_dataset_map`` (utils.py:318-407) and
``load_batch`` (utils.py:487-530); the file is parsed by this package's own HDF5 reader (h5py
is not available) and is opened ONCE per dataset instead of once per batch.
"""
from __future__ import annotations

import typing as t
import warnings
from pathlib import Path

import numpy as np

from .hdf5 import File
from .postprocess import standard_amino_acids

# The reference maps non-standard labels through aposteriori.config.UNCOMMON_RESIDUE_DICT
# (aposteriori==2.4.0, not vendored, not installable here).  This is NOT that table: only the
# widely documented modified-residue -> parent pairs are listed; anything else fails the same
# assertion the reference raises (utils.py:385-389).  Extend via ``UNCOMMON_RESIDUE_DICT.update``.
UNCOMMON_RESIDUE_DICT = {
    "MSE": "MET", "SEP": "SER", "TPO": "THR", "PTR": "TYR", "HYP": "PRO", "CSO": "CYS", "CME": "CYS",
    "CSD": "CYS", "OCS": "CYS", "KCX": "LYS", "MLY": "LYS", "M3L": "LYS", "LLP": "LYS", "PCA": "GLU",
    "FME": "MET", "DAL": "ALA",
}

_open_files: t.Dict[str, File] = {}
_MIN_FRAMES_PER_THREAD = 4
_pool_state: dict = {}


def _loader_threads() -> int:
    """Threads used by ``load_batch`` (env TIMED_B200_LOADER_THREADS; default: the host cores, at most 32)."""
    import os
    try:
        return max(1, int(os.environ.get("TIMED_B200_LOADER_THREADS", min(32, os.cpu_count() or 1))))
    except ValueError:
        return 1


def _pool():
    from concurrent.futures import ThreadPoolExecutor
    n = _loader_threads()
    if _pool_state.get("n") != n:
        if "pool" in _pool_state:
            _pool_state["pool"].shutdown(wait=False)
        _pool_state["pool"] = ThreadPoolExecutor(max_workers=n, thread_name_prefix="frame-loader")
        _pool_state["n"] = n
    return _pool_state["pool"]


def _open(path) -> File:
    key = str(Path(path).resolve())
    f = _open_files.get(key)
    if f is None or Path(key).stat().st_mtime_ns != getattr(f, "_mtime", None):
        f = File(key)
        f._mtime = Path(key).stat().st_mtime_ns
        _open_files[key] = f
    return f


def _as_str(x) -> str:
    return x.decode("utf-8") if isinstance(x, (bytes, np.bytes_)) else str(x)


def create_flat_dataset_map(frame_dataset: Path, filter_list: t.List[str] = [],
                            remove_blacklist_silently: bool = False):
    """utils.py:318-407.  Order: pdb codes and chains in the file's (name) order, residue ids
    sorted as INTEGERS (utils.py:367-371).  Returns ([(pdb, chain, res_id, label)], {pdbs})."""
    standard = set(standard_amino_acids.values())
    f = _open(frame_dataset)
    flat: t.List[t.Tuple[str, str, str, str]] = []
    pdbs = set()
    for pdb_code in f:
        if pdb_code[:4] in filter_list:
            if remove_blacklist_silently:
                warnings.warn(f"PDB code {pdb_code} was found in benchmark dataset. It was automatically removed.")
                continue
            raise ValueError(
                f"PDB code {pdb_code} was found in benchmark dataset. Turn on remove_blacklist_silently=True "
                f"if you want to ignore these structures for training.")
        for chain_id in f[pdb_code].keys():
            chain = f[pdb_code][chain_id]
            for residue_id in sorted(chain.keys(), key=int):
                label = _as_str(chain[residue_id].attrs["label"])
                if label not in standard:
                    if label in UNCOMMON_RESIDUE_DICT:
                        warnings.warn(f"{label} is not a standard residue.")
                        label = UNCOMMON_RESIDUE_DICT[label]
                        warnings.warn(f"Residue converted to {label}.")
                    assert label in standard, f"Expected natural amino acid, but got {label}."
                flat.append((pdb_code, chain_id, str(int(residue_id)), label))
                pdbs.add(pdb_code)
    return flat, pdbs


def dataset_metadata(frame_dataset: Path) -> dict:
    """Root attributes of the dataset (utils.py:230-281 reads them into DatasetMetadata)."""
    f = _open(frame_dataset)
    return {k: f.attrs[k] for k in f.attrs.keys()}


def _native_load(f: File, rows, dims, X: np.ndarray, y: np.ndarray) -> bool:
    """Fast path of ``load_batch``: every frame of the batch is a chunked dataset filtered by deflate (+ shuffle) only ->
    one call into libtimed_b200's host-side inflater (zlib on a thread pool, scatter + cast to the batch array without the
    GIL).  Returns False (nothing written) when the file is stored any other way; the Python reader then does the work."""
    import ctypes as C
    from . import _lib
    try:
        lib = _lib.load()
    except Exception:
        return False
    src_off, src_size, dst_frame, origin = [], [], [], []
    common = None
    objs = []
    # one native call for the whole batch when every frame is one deflate chunk (the usual layout); else frame by frame
    import os
    fast = None
    if X.dtype != np.bool_ and len(rows) and not os.environ.get("TIMED_B200_NO_FAST_INDEX"):
        fast = _fast_frame_index(f, rows, dims)
    if fast is not None:
        offs, sizes, y_fast, dtype = fast
        n = len(rows)
        a_base = np.frombuffer(f.buf, dtype=np.uint8)            # (named: the arrays must outlive the call)
        a_frame = np.arange(n, dtype=np.int64)
        a_org = np.zeros((n, len(dims)), np.int32)
        a_dims = np.asarray(dims, dtype=np.int32)
        rc = lib.timed_b200_inflate_chunks(
            C.c_void_p(a_base.ctypes.data), n, C.c_void_p(offs.ctypes.data), C.c_void_p(sizes.ctypes.data),
            C.c_void_p(a_frame.ctypes.data), C.c_void_p(a_org.ctypes.data), len(dims), C.c_void_p(a_dims.ctypes.data),
            C.c_void_p(a_dims.ctypes.data), 1, 0, {4: _lib.DTYPE_F32, 8: _lib.DTYPE_F64, 1: _lib.DTYPE_U8}[dtype.itemsize],
            _lib.DTYPE_F32, C.c_void_p(X.ctypes.data), _loader_threads())
        _lib.check(rc)
        y[:] = y_fast
        return True
    for i, row in enumerate(rows):
        pdb_code, chain_id, residue_id = (str(v) for v in row[:3])
        ds = f[pdb_code][chain_id][residue_id]
        objs.append(ds)
        info = ds.chunk_table() if hasattr(ds, "chunk_table") else None
        if info is None or tuple(ds.shape) != tuple(dims):
            return False
        cdims, table, deflate, shuffle, dtype = info
        key = (cdims, deflate, shuffle, dtype)
        if common is None:
            common = key
        elif key != common:
            return False
        for org, off, size in table:
            origin.append(org)
            src_off.append(off)
            src_size.append(size)
            dst_frame.append(i)
    cdims, deflate, shuffle, dtype = common
    to_bool = X.dtype == np.bool_
    if to_bool and dtype.itemsize != 1:
        return False                                   # float frames in a boolean dataset: let the reader decide
    code = {4: _lib.DTYPE_F32, 8: _lib.DTYPE_F64, 1: _lib.DTYPE_U8}[dtype.itemsize]
    a_off = np.asarray(src_off, dtype=np.int64)
    a_size = np.asarray(src_size, dtype=np.int64)
    a_frame = np.asarray(dst_frame, dtype=np.int64)
    a_org = np.ascontiguousarray(np.asarray(origin, dtype=np.int32).reshape(len(origin), len(dims)))
    a_cd = np.asarray(cdims, dtype=np.int32)
    a_fd = np.asarray(dims, dtype=np.int32)
    base = np.frombuffer(f.buf, dtype=np.uint8)
    out = X.view(np.uint8) if to_bool else X
    rc = lib.timed_b200_inflate_chunks(
        C.c_void_p(base.ctypes.data), len(a_off), C.c_void_p(a_off.ctypes.data), C.c_void_p(a_size.ctypes.data),
        C.c_void_p(a_frame.ctypes.data), C.c_void_p(a_org.ctypes.data), len(dims), C.c_void_p(a_cd.ctypes.data),
        C.c_void_p(a_fd.ctypes.data), int(deflate), int(shuffle), code, _lib.DTYPE_U8 if to_bool else _lib.DTYPE_F32,
        C.c_void_p(out.ctypes.data), _loader_threads())
    _lib.check(rc)
    if to_bool:
        np.not_equal(out, 0, out=X)                    # stored enum bytes -> booleans
    for i, ds in enumerate(objs):
        y[i] = ds.attrs["encoded_residue"]
    return True


def load_batch(dataset_path: Path, data_point_batch: t.Sequence[t.Tuple]) -> t.Tuple[np.ndarray, np.ndarray]:
    """utils.py:487-530.  X: (B, *frame_dims) float32 when ``voxels_as_gaussian`` else bool --
    the reference allocates float64 and Keras casts it to float32 on entry (predict.py:142), so
    the values the network sees are identical; y: (B, 20) float64 labels."""
    f = _open(dataset_path)
    dims = tuple(int(d) for d in f.attrs["frame_dims"])
    gaussian = bool(f.attrs["voxels_as_gaussian"])
    n = len(data_point_batch)
    X = np.zeros((n, *dims), dtype=np.float32 if gaussian else np.bool_)
    y = np.zeros((n, 20), dtype=float)
    if n and _native_load(f, data_point_batch, dims, X, y):
        return X, y

    def one(i):
        pdb_code, chain_id, residue_id = (str(v) for v in data_point_batch[i][:3])
        ds = f[pdb_code][chain_id][residue_id]
        X[i] = ds[()]                         # gzip inflate releases the GIL: frames decode in parallel
        y[i] = ds.attrs["encoded_residue"]

    if n >= 2 * _MIN_FRAMES_PER_THREAD and _loader_threads() > 1:
        list(_pool().map(one, range(n)))
    else:
        for i in range(n):
            one(i)
    return X, y


def _attr_header(d: bytes) -> t.Tuple[bytes, int]:
    """(name, offset of the data) of a raw HDF5 attribute message (versions 1-3), as hdf5/reader.py::_Attrs walks it."""
    import struct
    version = d[0]
    nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
    if version == 1:
        p, pad = 8, (lambda x: (x + 7) // 8 * 8)
    elif version in (2, 3):
        p, pad = (8 if version == 2 else 9), (lambda x: x)
    else:
        raise ValueError("attribute message version")
    name = d[p:p + nsz].split(b"\x00")[0]
    return name, p + pad(nsz) + pad(tsz) + pad(ssz)


def _fast_frame_index(f: File, rows, dims):
    """(chunk offsets, stored sizes, labels (n, 20), stored dtype) of the frames of ``rows`` through ONE native call
    (``timed_b200_hdf5_frame_index``: object headers / chunk B-trees walked on host threads), the first frame parsed by the
    Python reader as the template every other frame must match byte for byte.  None when the file is not stored as one
    unshuffled deflate chunk per gaussian frame in version-1 object headers, or any frame deviates: the caller then walks the
    batch frame by frame."""
    import ctypes as C
    from . import _lib
    n = len(rows)
    groups: t.Dict[t.Tuple[str, str], dict] = {}
    addrs = np.empty(n, np.int64)
    try:
        for i, row in enumerate(rows):
            key = (str(row[0]), str(row[1]))
            links = groups.get(key)
            if links is None:
                links = groups[key] = f[key[0]][key[1]]._load()
            addrs[i] = links[str(row[2])]
        ds0 = f[str(rows[0][0])][str(rows[0][1])][str(rows[0][2])]
        info = ds0.chunk_table() if hasattr(ds0, "chunk_table") else None
        if info is None or tuple(ds0.shape) != tuple(dims):
            return None
        cdims, table, deflate, shuffle, dtype = info
        if len(table) != 1 or tuple(cdims) != tuple(dims) or not deflate or shuffle or any(int(o) for o in table[0][0]):
            return None
        raw = {m.type: bytes(m.data) for m in ds0._msgs if m.type in (0x01, 0x03, 0x0B)}
        hdr = None
        for m in ds0._msgs:
            if m.type == 0x0C:
                d = bytes(m.data)
                name, p = _attr_header(d)
                if name == b"encoded_residue":
                    hdr, data_len = d[:p], len(d) - p
                    want = np.asarray(ds0.attrs["encoded_residue"], dtype=np.float64)
                    if data_len != 160 or want.shape != (20,) or not np.array_equal(np.frombuffer(d[p:], "<f8"), want):
                        return None
        if hdr is None or 0x01 not in raw or 0x03 not in raw or 0x0B not in raw:
            return None
    except (KeyError, AttributeError, ValueError, IndexError):
        return None
    base = np.frombuffer(f.buf, dtype=np.uint8)
    offs, sizes = np.empty(n, np.int64), np.empty(n, np.int64)
    labels = np.empty((n, 20), np.float64)
    status = np.empty(n, np.int32)
    rc = _lib.load().timed_b200_hdf5_frame_index(
        C.c_void_p(base.ctypes.data), len(base), int(f.base_addr), n, C.c_void_p(addrs.ctypes.data),
        raw[0x01], len(raw[0x01]), raw[0x03], len(raw[0x03]), raw[0x0B], len(raw[0x0B]), hdr, len(hdr), 160, len(dims),
        C.c_void_p(offs.ctypes.data), C.c_void_p(sizes.ctypes.data), C.c_void_p(labels.ctypes.data),
        C.c_void_p(status.ctypes.data), _loader_threads())
    _lib.check(rc)
    if status.any():
        return None
    return offs, sizes, labels, dtype


def load_batch_device(dataset_path: Path, data_point_batch: t.Sequence[t.Tuple], device: int = 0):
    """``load_batch`` with the frames left ON THE DEVICE: for datasets whose gaussian frames are stored as one
    deflate-filtered chunk each (what aposteriori / h5py write with ``compression='gzip'``), the STORED bytes of the batch go
    to the GPU (18 KB instead of 222 KB per frame on real structures) and every chunk is inflated there by its own warp
    (``timed_b200_inflate_device``, csrc/inflate.cuh) straight into the batch tensor the network reads.  Returns
    ``(frames, y)`` -- frames a CUDA tensor (B, *frame_dims) of the stored dtype, the values ``load_batch`` returns, byte for
    byte -- or ``None`` when the file is stored any other way, a stream does not inflate cleanly, or
    ``TIMED_B200_NO_DEVICE_INFLATE`` is set: the caller then takes ``load_batch``."""
    import ctypes as C
    import os
    if os.environ.get("TIMED_B200_NO_DEVICE_INFLATE"):
        return None
    import torch
    from . import _lib
    if not torch.cuda.is_available():
        return None
    f = _open(dataset_path)
    dims = tuple(int(d) for d in f.attrs["frame_dims"])
    n = len(data_point_batch)
    if not n or not bool(f.attrs["voxels_as_gaussian"]):
        return None                                     # boolean datasets are 1 byte per voxel already: host path
    fast = None if os.environ.get("TIMED_B200_NO_FAST_INDEX") else _fast_frame_index(f, data_point_batch, dims)
    offs, sizes, objs, common = np.empty(n, np.int64), np.empty(n, np.int64), [], None
    if fast is not None:
        offs, sizes, y_fast, common = fast
    for i, row in enumerate(data_point_batch if fast is None else ()):
        pdb_code, chain_id, residue_id = (str(v) for v in row[:3])
        ds = f[pdb_code][chain_id][residue_id]
        info = ds.chunk_table() if hasattr(ds, "chunk_table") else None
        if info is None or tuple(ds.shape) != dims:
            return None
        cdims, table, deflate, shuffle, dtype = info
        if len(table) != 1 or tuple(cdims) != dims or not deflate or shuffle or any(int(o) for o in table[0][0]):
            return None
        if common is None:
            common = dtype
        elif dtype != common:
            return None
        offs[i], sizes[i] = table[0][1], table[0][2]
        objs.append(ds)
    tdt = {("f", 4): torch.float32, ("f", 8): torch.float64, ("u", 1): torch.uint8}.get((common.kind, common.itemsize))
    if tdt is None or common.byteorder == ">":
        return None
    base = np.frombuffer(f.buf, dtype=np.uint8)
    lo, hi = int(offs.min()), int((offs + sizes).max())
    if hi - lo <= 4 * int(sizes.sum()) + (1 << 20):     # the batch's chunks are (nearly) contiguous in the file: one copy
        with warnings.catch_warnings():                 # (a read-only view of the mapped file; it is only read)
            warnings.filterwarnings("ignore", message="The given NumPy array is not writable")
            comp = torch.from_numpy(base[lo:hi]).to(f"cuda:{device}")
        rel = offs - lo
    else:
        comp = torch.from_numpy(np.concatenate([base[o:o + s] for o, s in zip(offs, sizes)])).to(f"cuda:{device}")
        rel = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    d_off = torch.from_numpy(rel).to(comp.device)
    d_size = torch.from_numpy(sizes).to(comp.device)
    frames = torch.empty((n, *dims), dtype=tdt, device=comp.device)
    status = torch.empty(n, dtype=torch.int32, device=comp.device)
    out_bytes = int(np.prod(dims)) * common.itemsize
    with torch.cuda.device(comp.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.load().timed_b200_inflate_device(
            C.c_void_p(comp.data_ptr()), n, C.c_void_p(d_off.data_ptr()), C.c_void_p(d_size.data_ptr()), out_bytes,
            C.c_void_p(frames.data_ptr()), C.c_void_p(status.data_ptr()), C.c_void_p(stream)))
        if int(status.ne(0).sum().item()):
            warnings.warn(f"{Path(dataset_path).name}: {int(status.ne(0).sum().item())} frame chunk(s) did not inflate on the "
                          "device; reading the batch on the host", RuntimeWarning, stacklevel=2)
            return None
    if fast is not None:
        return frames, y_fast
    y = np.zeros((n, 20), dtype=float)
    for i, ds in enumerate(objs):
        y[i] = ds.attrs["encoded_residue"]
    return frames, y
