"""Structure -> residue frames on the GPU: the step BEFORE the inference path.

The reference builds its frame datasets with aposteriori's ``make-frame-dataset`` (/root/reference/README.md:84-97: 21 A
edge, 21 voxels per side, gaussian voxels, C-beta injected, codec CNOCBCA / CNOCBCAQ / CNOCBCAP; the same call with
keyword arguments at /root/reference/ui.py:63-87,114-127).  aposteriori is an un-vendored dependency that is absent
offline, so this module restates its published behaviour; **parity with aposteriori's voxel values is unpinned** and the
functions say so.  What IS pinned: the local frame (C-alpha at the origin, N on +y, C in the xy plane at x > 0) is the one
under which the reference's hard-coded C-beta, (-0.741287356, -0.53937931, -1.224287356) -- "the average position of all
beta-Carbon in the protein 1QYS after the aforementioned rotations", README.md:242 -- coincides with the mean C-beta of
real residues (0.03 A on 1ubq; every other axis assignment is >= 0.29 A off), and the dataset schema
(/root/reference/design_utils/utils.py:238-251) that ``predict.py`` reads back.

All voxel arithmetic runs in ``timed_b200_voxelise`` (csrc/voxelise.cuh): one CTA per residue, every atom of the structure
tested against the frame, 3x3x3 unit-mass gaussian stamps accumulated in fixed point (order independent).  There is no CPU
path; ``oracle/voxelise_oracle.py`` is the checker.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import typing as t
import warnings
from pathlib import Path

import numpy as np

from . import _lib
from .postprocess import standard_amino_acids

IDEAL_CB = (-0.741287356, -0.53937931, -1.224287356)
# van der Waals radii (A); a stamp's gaussian has FWHM = the radius' diameter / 2, i.e. sigma = r / 2.3548 -- OUR choice,
# aposteriori's exact kernel is not available offline
VDW = {"C": 1.70, "N": 1.55, "O": 1.52, "S": 1.80}
CODECS = {
    "CNOCBCA": ["C", "N", "O", "CB", "CA"],
    "CNOCACB": ["C", "N", "O", "CA", "CB"],
    "CNOCBCAQ": ["C", "N", "O", "CB", "CA", "Q"],
    "CNOCBCAP": ["C", "N", "O", "CB", "CA", "P"],
    "CNOCACBQ": ["C", "N", "O", "CA", "CB", "Q"],
    "CNOCACBP": ["C", "N", "O", "CA", "CB", "P"],
}
_ATOM_LABEL = {"N": "N", "CA": "CA", "C": "C", "O": "O", "OXT": "O", "CB": "CB"}      # keep_sidechain_cb_atom_filter
_CHARGE = {"D": -1, "E": -1, "K": 1, "R": 1, "H": 1}                                   # residue_charge in {-1, 0, +1} (utils.py:86,97)
_POLAR = set("RDEHK")                                                                  # polarity_Zimmerman >= 20 (utils.py:95)
_THREE_TO_ONE = {v: k for k, v in standard_amino_acids.items()}


class Residue(t.NamedTuple):
    chain: str
    res_id: str
    label: str                    # three-letter code
    atoms: dict                   # name -> (x, y, z)


def parse_pdb(path, all_states: bool = False) -> t.List[t.List[Residue]]:
    """ATOM records of a (possibly gzipped) PDB file -> one residue list per state (first MODEL only unless all_states).
    Alternate locations other than the first are dropped; insertion-code duplicates of a residue number are skipped."""
    path = Path(path)
    opener = gzip.open if path.suffix == ".gz" else open
    states: t.List[dict] = [{}]
    with opener(path, "rt") as fh:
        for line in fh:
            rec = line[:6]
            if rec == "ENDMDL":
                if not all_states:
                    break
                states.append({})
            if rec != "ATOM  ":
                continue
            alt = line[16]
            name = line[12:16].strip()
            key = (line[21], line[22:27])                     # chain, resSeq + iCode
            res = states[-1].setdefault(key, {"label": line[17:20].strip(), "atoms": {}, "alt": alt})
            if alt not in (" ", res["alt"]) and res["alt"] != " ":
                continue
            if name not in res["atoms"]:
                res["atoms"][name] = (float(line[30:38]), float(line[38:46]), float(line[46:54]))
    out = []
    for st in states:
        if not st:
            continue
        seen, residues = set(), []
        for (chain, num), r in st.items():
            rid = num[:4].strip()
            if (chain, rid) in seen:
                warnings.warn(f"{path.name}: residue {chain}{num.strip()} repeats number {rid} (insertion code); skipped")
                continue
            seen.add((chain, rid))
            residues.append(Residue(chain if chain.strip() else "A", rid, r["label"], r["atoms"]))
        out.append(residues)
    return out


def residue_frame(n, ca, c) -> np.ndarray:
    """(12,) float32: origin = C-alpha, then the rows (x, y, z) of the rotation into the residue's local frame:
    y along CA -> N, x the component of CA -> C orthogonal to y (C in the xy plane at x > 0), z = x cross y."""
    n, ca, c = (np.asarray(v, dtype=np.float64) for v in (n, ca, c))
    ey = n - ca
    ey /= np.linalg.norm(ey)
    v = c - ca
    ex = v - ey * (v @ ey)
    ex /= np.linalg.norm(ex)
    ez = np.cross(ex, ey)
    return np.concatenate([ca, ex, ey, ez]).astype(np.float32)


class AtomTables(t.NamedTuple):
    atoms: np.ndarray             # (n_atoms, 4) float32 x, y, z, sigma (voxels)
    channel: np.ndarray           # (n_atoms,) int32
    residue: np.ndarray           # (n_atoms,) int32 index into `residues`
    is_cb: np.ndarray             # (n_atoms,) int32
    frames: np.ndarray            # (n_res, 12) float32
    prop: t.Optional[np.ndarray]  # (n_res,) float32 or None
    valid: np.ndarray             # indices of the residues that have N, CA and C (they get a frame)
    channels: t.List[str]
    atom_range: t.Optional[np.ndarray] = None     # (n_res, 2) int32: atoms a residue's frame looks at (batched structures)


class ResidueInfo(t.NamedTuple):
    """Per-residue arrays of one state of one structure, in file order (what the vectorised parser returns)."""
    chain: np.ndarray             # (n_res,) str
    res_id: np.ndarray            # (n_res,) str
    label: np.ndarray             # (n_res,) str, three-letter code


_ENC_NAMES = [b"N", b"CA", b"C", b"O", b"OXT", b"CB"]


def fast_tables(path, codec: str, voxel_edge: float, all_states: bool = False):
    """Vectorised PDB -> [(AtomTables, ResidueInfo)] per state: the same tables ``build_tables(parse_pdb(...))`` produces
    (tests/test_voxelise_cpu.py checks them equal), ~10x faster -- the structure route of predict.py is bound by this."""
    if codec not in CODECS:
        raise ValueError(f"unknown codec {codec!r} (known: {sorted(CODECS)})")
    channels = CODECS[codec]
    prop_kind = channels[-1] if channels[-1] in ("Q", "P") else None
    path = Path(path)
    with (gzip.open(path, "rb") if path.suffix == ".gz" else open(path, "rb")) as fh:
        raw = fh.read()
    lines = raw.split(b"\n")
    # state boundaries: ENDMDL closes a state
    out = []
    start = 0
    bounds = [i for i, ln in enumerate(lines) if ln.startswith(b"ENDMDL")] or []
    segments = []
    for b in bounds:
        segments.append((start, b))
        start = b + 1
    segments.append((start, len(lines)))
    if not all_states:
        segments = [next((sg for sg in segments if any(ln.startswith(b"ATOM  ") for ln in lines[sg[0]:sg[1]])), segments[0])]
    for a, b in segments:
        atom_lines = [ln[:54].ljust(54) for ln in lines[a:b] if ln.startswith(b"ATOM  ")]
        if not atom_lines:
            continue
        arr = np.frombuffer(b"".join(atom_lines), dtype=np.uint8).reshape(len(atom_lines), 54)

        def col(lo, hi):
            return np.ascontiguousarray(arr[:, lo:hi]).view(f"S{hi - lo}")[:, 0]

        names = np.char.strip(col(12, 16))
        alt = col(16, 17)
        resname = np.char.strip(col(17, 20))
        reskey = col(21, 27)                                       # chain + resSeq + iCode
        xyz = np.stack([col(30, 38).astype(np.float64), col(38, 46).astype(np.float64), col(46, 54).astype(np.float64)], axis=1)
        # residues in order of first appearance
        uniq, first, inv = np.unique(reskey, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[order] = np.arange(len(uniq))
        res_of_atom = rank[inv]
        n_res = len(uniq)
        first_atom = first[order]
        # alternate locations: keep atoms whose altLoc is blank or equals the first altLoc seen in their residue
        res_alt = alt[first_atom][res_of_atom]
        keep = (alt == b" ") | (alt == res_alt) | (res_alt == b" ")
        # first occurrence of an atom name within a residue wins
        combo = res_of_atom.astype(np.int64) * 4096 + np.unique(names, return_inverse=True)[1]
        _, first_combo = np.unique(np.where(keep, combo, -1 - np.arange(len(combo))), return_index=True)
        is_first = np.zeros(len(combo), dtype=bool)
        is_first[first_combo] = True
        keep &= is_first
        chain = np.array([k[:1].decode() if k[:1].strip() else "A" for k in uniq[order]])
        res_id = np.array([k[1:5].decode().strip() for k in uniq[order]])
        label = np.char.decode(resname[first_atom], "ascii")
        # insertion-code duplicates of a residue number: keep the first, drop the rest (as parse_pdb)
        seen, dup_res = set(), np.zeros(n_res, dtype=bool)
        for i in range(n_res):
            key = (chain[i], res_id[i])
            dup_res[i] = key in seen
            seen.add(key)
        if dup_res.any():
            warnings.warn(f"{path.name}: {int(dup_res.sum())} residue(s) repeat a residue number (insertion code); skipped")
            keep &= ~dup_res[res_of_atom]
            live = np.nonzero(~dup_res)[0]
            remap = np.full(n_res, -1, dtype=np.int64)
            remap[live] = np.arange(len(live))
            res_of_atom = remap[res_of_atom]
            chain, res_id, label, n_res = chain[live], res_id[live], label[live], len(live)
        # backbone coordinates per residue
        bb = {}
        for nm in (b"N", b"CA", b"C"):
            m = keep & (names == nm)
            c = np.full((n_res, 3), np.nan)
            c[res_of_atom[m]] = xyz[m]
            bb[nm] = c
        has = ~(np.isnan(bb[b"N"][:, 0]) | np.isnan(bb[b"CA"][:, 0]) | np.isnan(bb[b"C"][:, 0]))
        frames = np.zeros((n_res, 12), dtype=np.float32)
        if has.any():
            n_, ca, c_ = bb[b"N"][has], bb[b"CA"][has], bb[b"C"][has]
            ey = n_ - ca
            ey /= np.linalg.norm(ey, axis=1, keepdims=True)
            v = c_ - ca
            ex = v - ey * np.sum(v * ey, axis=1, keepdims=True)
            ex /= np.linalg.norm(ex, axis=1, keepdims=True)
            ez = np.cross(ex, ey)
            frames[has] = np.concatenate([ca, ex, ey, ez], axis=1).astype(np.float32)
        prop = None
        if prop_kind:
            one = [_THREE_TO_ONE.get(l, "X") for l in label]
            prop = np.array([_CHARGE.get(o, 0) if prop_kind == "Q" else (1.0 if o in _POLAR else 0.0) for o in one], dtype=np.float32)
        # encoded atoms, in file order (what build_tables produces)
        enc = keep & np.isin(names, _ENC_NAMES)
        nm = names[enc]
        lab = np.where(nm == b"OXT", b"O", nm)
        ch = np.array([channels.index(x.decode()) for x in lab], dtype=np.int32) if len(lab) else np.zeros(0, np.int32)
        first_letter = np.array([x[:1].decode() for x in lab]) if len(lab) else np.zeros(0, "U1")
        sigma = np.array([VDW[f] for f in first_letter], dtype=np.float64) / 2.3548 / voxel_edge
        atoms = np.concatenate([xyz[enc], sigma[:, None]], axis=1).astype(np.float32)
        tab = AtomTables(atoms, ch, res_of_atom[enc].astype(np.int32), (lab == b"CB").astype(np.int32), frames, prop,
                         np.nonzero(has)[0].astype(np.int64), channels)
        out.append((tab, ResidueInfo(chain, res_id, label)))
    if not out:
        raise ValueError(f"{path}: no ATOM records")
    return out


_NATIVE_WARNED = False


def native_tables(paths, codec: str, voxel_edge: float, all_states: bool = False, n_threads: t.Optional[int] = None):
    """``[fast_tables(p, codec, voxel_edge, all_states) for p in paths]`` with the file reading, gunzip and record parsing done
    by libtimed_b200's host-side PDB reader on a thread pool (``timed_b200_pdb_parse``, csrc/pdb_parse.cuh) and the float
    arithmetic (frames, sigmas) by the same numpy expressions as ``fast_tables`` over ALL files at once: identical tables
    (tests/test_voxelise_cpu.py), ~20x faster per structure -- the structure route of predict.py was bound by the parser.
    A file the native reader reports as unusual (unreadable, no ATOM record, malformed field) goes through ``fast_tables``,
    which raises / warns as before; so does everything when the library is not built (one RuntimeWarning)."""
    global _NATIVE_WARNED
    import ctypes as C
    if codec not in CODECS:
        raise ValueError(f"unknown codec {codec!r} (known: {sorted(CODECS)})")
    paths = [Path(p) for p in paths]
    try:
        from . import _lib
        lib = _lib.load()
    except Exception as e:  # noqa: BLE001 -- host-side helper: same tables from the Python parser, much slower
        if not _NATIVE_WARNED:
            warnings.warn(f"timed_b200_pdb_parse unavailable ({type(e).__name__}: {e}); using the Python PDB parser "
                          "(identical tables, ~20x slower)", RuntimeWarning, stacklevel=2)
            _NATIVE_WARNED = True
        return [fast_tables(p, codec, voxel_edge, all_states) for p in paths]
    if not paths:
        return []
    channels = CODECS[codec]
    prop_kind = channels[-1] if channels[-1] in ("Q", "P") else None
    arr = (C.c_char_p * len(paths))(*[os.fsencode(str(p)) for p in paths])
    handle = C.c_void_p()
    threads = n_threads or min(len(paths), max(1, os.cpu_count() or 1), 16)
    _lib.check(lib.timed_b200_pdb_parse(arr, len(paths), int(bool(all_states)), int(threads), C.byref(handle)))
    try:
        ns, nr, na = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(lib.timed_b200_pdb_sizes(handle, C.byref(ns), C.byref(nr), C.byref(na)))
        ns, nr, na = ns.value, nr.value, na.value
        status = np.zeros(len(paths), np.int32)
        st_file = np.zeros(ns, np.int32)
        st_r = np.zeros(ns + 1, np.int64)
        st_a = np.zeros(ns + 1, np.int64)
        st_dup = np.zeros(ns, np.int32)
        r_chain = np.zeros(nr, "S1")
        r_id = np.zeros(nr, "S4")
        r_label = np.zeros(nr, "S3")
        r_has = np.zeros(nr, np.uint8)
        r_bb = np.zeros((nr, 9), np.float64)
        a_xyz = np.zeros((na, 3), np.float64)
        a_name = np.zeros(na, np.int32)
        a_res = np.zeros(na, np.int32)
        ptr = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        _lib.check(lib.timed_b200_pdb_export(handle, ptr(status), ptr(st_file), ptr(st_r), ptr(st_a), ptr(st_dup), ptr(r_chain),
                                             ptr(r_id), ptr(r_label), ptr(r_has), ptr(r_bb), ptr(a_xyz), ptr(a_name), ptr(a_res)))
    finally:
        lib.timed_b200_pdb_free(handle)
    # ---- the float arithmetic of fast_tables, once over all residues / atoms (row-wise: independent of the batch)
    has = r_has.astype(bool)
    frames = np.zeros((nr, 12), dtype=np.float32)
    if has.any():
        n_, ca, c_ = r_bb[has, 0:3], r_bb[has, 3:6], r_bb[has, 6:9]
        ey = n_ - ca
        ey /= np.linalg.norm(ey, axis=1, keepdims=True)
        v = c_ - ca
        ex = v - ey * np.sum(v * ey, axis=1, keepdims=True)
        ex /= np.linalg.norm(ex, axis=1, keepdims=True)
        ez = np.cross(ex, ey)
        frames[has] = np.concatenate([ca, ex, ey, ez], axis=1).astype(np.float32)
    chain = np.char.decode(r_chain, "ascii")
    res_id = np.char.strip(np.char.decode(r_id, "ascii"))
    label = np.char.strip(np.char.decode(r_label, "ascii"))
    prop_all = None
    if prop_kind:
        one = [_THREE_TO_ONE.get(l, "X") for l in label]
        prop_all = np.array([_CHARGE.get(o, 0) if prop_kind == "Q" else (1.0 if o in _POLAR else 0.0) for o in one], dtype=np.float32)
    lab_of = ["N", "CA", "C", "O", "O", "CB"]                     # OXT is encoded as O
    ch_lut = np.array([channels.index(x) for x in lab_of], dtype=np.int32)
    sig_lut = np.array([VDW[x[0]] for x in lab_of], dtype=np.float64) / 2.3548 / voxel_edge
    atoms = np.concatenate([a_xyz, sig_lut[a_name][:, None]], axis=1).astype(np.float32)
    a_ch = ch_lut[a_name]
    a_cb = (a_name == 5).astype(np.int32)
    out: t.List[t.Optional[list]] = [[] if status[i] == 0 else None for i in range(len(paths))]
    for s in range(ns):
        r0, r1, a0, a1 = int(st_r[s]), int(st_r[s + 1]), int(st_a[s]), int(st_a[s + 1])
        path = paths[int(st_file[s])]
        if st_dup[s]:
            warnings.warn(f"{path.name}: {int(st_dup[s])} residue(s) repeat a residue number (insertion code); skipped")
        tab = AtomTables(atoms[a0:a1], a_ch[a0:a1], a_res[a0:a1], a_cb[a0:a1], frames[r0:r1],
                         prop_all[r0:r1] if prop_all is not None else None, np.nonzero(has[r0:r1])[0].astype(np.int64), channels)
        out[int(st_file[s])].append((tab, ResidueInfo(chain[r0:r1], res_id[r0:r1], label[r0:r1])))
    for i, p in enumerate(paths):
        if out[i] is None or not out[i]:
            out[i] = fast_tables(p, codec, voxel_edge, all_states)          # raises / warns as the Python reader does
    return out


def build_tables(residues: t.Sequence[Residue], codec: str, voxel_edge: float) -> AtomTables:
    if codec not in CODECS:
        raise ValueError(f"unknown codec {codec!r} (known: {sorted(CODECS)})")
    channels = CODECS[codec]
    prop_kind = channels[-1] if channels[-1] in ("Q", "P") else None
    rows, ch, ri, cb = [], [], [], []
    frames = np.zeros((len(residues), 12), dtype=np.float32)
    prop = np.zeros(len(residues), dtype=np.float32) if prop_kind else None
    valid = []
    for i, r in enumerate(residues):
        one = _THREE_TO_ONE.get(r.label, "X")
        if prop_kind == "Q":
            prop[i] = _CHARGE.get(one, 0)
        elif prop_kind == "P":
            prop[i] = 1.0 if one in _POLAR else 0.0
        if all(k in r.atoms for k in ("N", "CA", "C")):
            frames[i] = residue_frame(r.atoms["N"], r.atoms["CA"], r.atoms["C"])
            valid.append(i)
        for name, xyz in r.atoms.items():
            label = _ATOM_LABEL.get(name)
            if label is None:
                continue
            sigma = VDW[label[0]] / 2.3548 / voxel_edge
            rows.append((*xyz, sigma))
            ch.append(channels.index(label))
            ri.append(i)
            cb.append(1 if label == "CB" else 0)
    return AtomTables(np.asarray(rows, dtype=np.float32).reshape(-1, 4), np.asarray(ch, dtype=np.int32),
                      np.asarray(ri, dtype=np.int32), np.asarray(cb, dtype=np.int32), frames, prop,
                      np.asarray(valid, dtype=np.int64), channels)


def voxelise_tables(tab: AtomTables, residues_idx: np.ndarray, voxels_per_side: int = 21, frame_edge_length: float = 21.0,
                    voxels_as_gaussian: bool = True, encode_cb: bool = True, dtype=np.float32, device: int = 0,
                    return_device: bool = False, chunk: int = 2048):
    """Frames of the residues ``residues_idx`` (indices into the structure's residue list) -> (n, V, V, V, C) array of
    ``dtype`` (float32 / float16; bool for boolean voxels).  Runs on the GPU in chunks of ``chunk`` residues."""
    import torch
    lib = _lib.load()
    _lib.require_device()
    V, Cn = int(voxels_per_side), len(tab.channels)
    edge = float(frame_edge_length) / V
    dtype = np.dtype(np.bool_ if not voxels_as_gaussian else dtype)
    code = {np.dtype(np.float32): _lib.DTYPE_F32, np.dtype(np.float16): _lib.DTYPE_F16, np.dtype(np.bool_): _lib.DTYPE_U8}[dtype]
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float16): torch.float16, np.dtype(np.bool_): torch.uint8}[dtype]
    dev = torch.device("cuda", device)
    residues_idx = np.asarray(residues_idx, dtype=np.int64)
    n = len(residues_idx)
    out = torch.empty((n, V, V, V, Cn), dtype=tdt, device=dev)
    if n == 0 or len(tab.atoms) == 0:
        out.zero_()
        return out if return_device else out.cpu().numpy().astype(dtype)
    with torch.cuda.device(dev):
        d_atoms = torch.from_numpy(tab.atoms).to(dev)
        d_ch = torch.from_numpy(tab.channel).to(dev)
        d_ri = torch.from_numpy(tab.residue).to(dev)
        d_cb = torch.from_numpy(tab.is_cb).to(dev)
        prop_ch = Cn - 1 if tab.prop is not None else -1
        cb_ch = tab.channels.index("CB")
        cbx = (C.c_float * 4)(*IDEAL_CB, VDW["C"] / 2.3548 / edge)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        scratch = torch.empty((min(chunk, n), V, V, V, Cn), dtype=torch.int32, device=dev)
        d_fr = torch.from_numpy(tab.frames).to(dev)
        d_pr = torch.from_numpy(tab.prop).to(dev) if tab.prop is not None else None
        d_idx = torch.from_numpy(residues_idx.astype(np.int32)).to(dev)
        d_rng = torch.from_numpy(np.ascontiguousarray(tab.atom_range, dtype=np.int32)).to(dev) if tab.atom_range is not None else None
        for r0 in range(0, n, chunk):
            m = min(chunk, n - r0)
            _lib.check(lib.timed_b200_voxelise(
                C.c_void_p(d_atoms.data_ptr()), C.c_void_p(d_ch.data_ptr()), C.c_void_p(d_ri.data_ptr()),
                C.c_void_p(d_cb.data_ptr()), len(tab.atoms), C.c_void_p(d_fr.data_ptr()),
                C.c_void_p(d_pr.data_ptr()) if d_pr is not None else None, C.c_void_p(d_idx.data_ptr()),
                C.c_void_p(d_rng.data_ptr()) if d_rng is not None else None, r0, m, V, edge, Cn,
                int(voxels_as_gaussian), int(encode_cb), cbx, cb_ch, prop_ch, C.c_void_p(scratch.data_ptr()),
                C.c_void_p(out[r0:].data_ptr()), code, stream))
        torch.cuda.synchronize(dev)
    if return_device:
        return out
    host = out.cpu().numpy()
    return host.astype(np.bool_) if dtype == np.bool_ else host


_STD_LABELS = np.array(sorted(_THREE_TO_ONE))


def _dataset_order(tab: AtomTables, info: ResidueInfo) -> t.List[int]:
    """Residues that get a frame, in dataset order: chains as they appear, residue ids sorted as integers
    (utils.py:367-371); residues without N/CA/C or with a non-standard name are left out."""
    valid = np.asarray(tab.valid, dtype=np.int64)
    if len(valid):
        valid = valid[np.isin(np.asarray(info.label)[valid], _STD_LABELS)]
    if not len(valid):
        return []
    chains = np.asarray(info.chain)[valid]
    uniq, first, inv = np.unique(chains, return_index=True, return_inverse=True)
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))           # chains in order of first appearance
    res_no = np.asarray(info.res_id)[valid].astype(np.int64)                 # ValueError for a non-integer id, as int() gave
    return valid[np.lexsort((res_no, rank[inv]))].tolist()                   # stable: equal numbers keep the file order


class State(t.NamedTuple):
    """One state of one structure, parsed and ready to voxelise."""
    code: str                     # pdb code (``code_{state}`` for NMR ensembles voxelised state by state)
    tab: AtomTables
    info: ResidueInfo
    order: t.List[int]            # residues that get a frame, in dataset order


def load_states(paths, codec: str = "CNOCBCA", voxels_per_side: int = 21, frame_edge_length: float = 21.0,
                voxelise_all_states: bool = False) -> t.List[State]:
    """Parse structure files (host only, ~2 ms per 100-residue structure) into per-state tables."""
    edge = float(frame_edge_length) / voxels_per_side
    out = []
    paths = [Path(p) for p in paths]
    parsed = native_tables(paths, codec, edge, all_states=voxelise_all_states)
    for path, states in zip(paths, parsed):
        pdb_code = path.name.split(".pdb")[0]
        for si, (tab, info) in enumerate(states):
            code = f"{pdb_code}_{si}" if voxelise_all_states and len(states) > 1 else pdb_code
            idx = _dataset_order(tab, info)
            skipped = len(info.label) - len(idx)
            if skipped:
                warnings.warn(f"{path.name}: {skipped} residue(s) without N/CA/C or with a non-standard name were skipped")
            out.append(State(code, tab, info, idx))
    return out


def flat_map_of(states: t.Sequence[State]) -> t.List[t.Tuple[str, str, str, str]]:
    return [(st.code, str(st.info.chain[i]), str(st.info.res_id[i]), str(st.info.label[i])) for st in states for i in st.order]


def voxelise_states(states: t.Sequence[State], voxels_per_side: int = 21, frame_edge_length: float = 21.0,
                    voxels_as_gaussian: bool = True, encode_cb: bool = True, dtype=np.float32, device: int = 0,
                    return_device: bool = False):
    """Frames of all residues of ``states`` (in flat-map order).  The atom and residue tables of the states are
    concatenated -- every residue carries the range of atoms of its own state -- so the whole set is voxelised by one launch
    per 2048 frames."""
    atoms, chan, resi, iscb, frames, props, ranges, order = [], [], [], [], [], [], [], []
    a0 = r0 = 0
    for st in states:
        tab = st.tab
        atoms.append(tab.atoms)
        chan.append(tab.channel)
        resi.append(tab.residue + r0)
        iscb.append(tab.is_cb)
        frames.append(tab.frames)
        if tab.prop is not None:
            props.append(tab.prop)
        ranges.append(np.tile(np.array([[a0, a0 + len(tab.atoms)]], dtype=np.int32), (len(tab.frames), 1)))
        order.extend(i + r0 for i in st.order)
        a0 += len(tab.atoms)
        r0 += len(tab.frames)
    if not atoms:
        raise ValueError("no structure to voxelise")
    big = AtomTables(np.concatenate(atoms), np.concatenate(chan), np.concatenate(resi).astype(np.int32), np.concatenate(iscb),
                     np.concatenate(frames), np.concatenate(props) if props else None, np.zeros(0, np.int64),
                     states[0].tab.channels, np.concatenate(ranges))
    return voxelise_tables(big, np.asarray(order, dtype=np.int64), voxels_per_side, frame_edge_length, voxels_as_gaussian,
                           encode_cb, dtype, device, return_device)


def voxelise_structures(paths, codec: str = "CNOCBCA", voxels_per_side: int = 21, frame_edge_length: float = 21.0,
                        voxels_as_gaussian: bool = True, encode_cb: bool = True, voxelise_all_states: bool = False,
                        dtype=np.float32, device: int = 0, return_device: bool = False):
    """Several structure files -> (frames (n, V, V, V, C), flat map [(pdb_code, chain, res_id, label)])."""
    states = load_states(paths, codec, voxels_per_side, frame_edge_length, voxelise_all_states)
    fr = voxelise_states(states, voxels_per_side, frame_edge_length, voxels_as_gaussian, encode_cb, dtype, device, return_device)
    return fr, flat_map_of(states)


def voxelise_structure(path, codec: str = "CNOCBCA", voxels_per_side: int = 21, frame_edge_length: float = 21.0,
                       voxels_as_gaussian: bool = True, encode_cb: bool = True, voxelise_all_states: bool = False,
                       dtype=np.float32, device: int = 0, return_device: bool = False):
    """One structure file -> (frames (n, V, V, V, C), flat map [(pdb_code, chain, res_id, label)]) in dataset order
    (chains as they appear, residue ids sorted as integers -- utils.py:367-371).  NMR states get ``pdb_code_{state}``."""
    return voxelise_structures([path], codec, voxels_per_side, frame_edge_length, voxels_as_gaussian, encode_cb,
                               voxelise_all_states, dtype, device, return_device)


def make_frame_dataset(structure_files, output_folder, name: str, frame_edge_length: float = 21.0, voxels_per_side: int = 21,
                       codec: str = "CNOCBCA", processes: int = 1, is_pdb_gzipped: bool = False,
                       require_confirmation: bool = False, voxels_as_gaussian: bool = True, voxelise_all_states: bool = False,
                       verbosity: int = 1, encode_cb: bool = True, compression: t.Optional[str] = "gzip", device: int = 0) -> Path:
    """Keyword-compatible stand-in for ``aposteriori.data_prep.create_frame_data_set.make_frame_dataset`` as the
    reference calls it (ui.py:73-86): writes ``{output_folder}/{name}.hdf5`` in the schema ``predict.py`` reads
    (utils.py:238-251) and returns its path.  ``processes`` / ``is_pdb_gzipped`` / ``require_confirmation`` are accepted
    for compatibility (gzip is detected from the suffix; all frames are computed on the GPU)."""
    from .hdf5 import write_frame_dataset
    if hasattr(codec, "name"):                       # an aposteriori Codec object
        codec = str(codec.name)
    warnings.warn("voxel values restate aposteriori's published behaviour and are UNVERIFIED against aposteriori 2.4.0 "
                  "(absent offline); the frame alignment is pinned by the reference's hard-coded C-beta", RuntimeWarning,
                  stacklevel=2)
    out = Path(output_folder) / f"{name}.hdf5"
    tree: dict = {}
    n_ch = len(CODECS[codec])
    for path in structure_files:
        frames, flat = voxelise_structure(path, codec, voxels_per_side, frame_edge_length, voxels_as_gaussian, encode_cb,
                                          voxelise_all_states, np.float32, device)
        for fr, (pdb, chain, rid, label) in zip(frames, flat):
            tree.setdefault(pdb, {}).setdefault(chain, {})[rid] = (fr, label)
        if verbosity > 1:
            print(f"{Path(path).name}: {len(flat)} frames")
    write_frame_dataset(out, tree, (voxels_per_side,) * 3 + (n_ch,), voxels_as_gaussian=voxels_as_gaussian,
                        atom_encoder=tuple(CODECS[codec]), frame_edge_length=frame_edge_length, compression=compression)
    return out
