"""CPU: the pure-Python HDF5 writer/reader pair round-trips Keras-style model files and
aposteriori-style frame datasets (real-file parity is unpinned: no h5py/libhdf5 offline)."""
import json
import struct

import numpy as np
import pytest

from timed_design_b200 import standins
from timed_design_b200.hdf5 import File, Hdf5FormatError, read_keras_h5, write_frame_dataset, write_keras_h5
from timed_design_b200.hdf5.writer import Writer


def test_keras_h5_roundtrip(tmp_path):
    cfg, w = standins.tiny_standin(calib_frames=0)
    p = tmp_path / "TIMED_tiny.h5"
    write_keras_h5(p, cfg, w)
    assert p.read_bytes()[:8] == b"\x89HDF\r\n\x1a\n"
    cfg2, w2 = read_keras_h5(p)
    assert cfg2 == json.loads(json.dumps(cfg))
    assert list(w2) == [k for k in (l["name"] for l in cfg["config"]["layers"]) if k in w]
    for layer, ws in w.items():
        for k, v in ws.items():
            np.testing.assert_array_equal(w2[layer][f"{layer}/{k}"], v)
    f = File(p)
    assert f.attrs["backend"] == "tensorflow" and f.attrs["keras_version"] == "2.13.1"
    assert f["model_weights/conv3d/conv3d/kernel:0"].shape == (3, 3, 3, 6, 16)
    assert f["model_weights"]["conv3d"]["conv3d/bias:0"].dtype == np.float32
    with pytest.raises(KeyError):
        f["model_weights/nope"]


def test_many_children_and_deep_btree(tmp_path):
    w = Writer()
    for i in range(300):                      # > 8 per SNOD and > 32 leaves -> two B-tree levels
        w.root.dataset(f"g/{i}", np.arange(i % 7 + 1, dtype=np.int64))
    w.save(tmp_path / "many.h5")
    f = File(tmp_path / "many.h5")
    keys = f["g"].keys()
    assert len(keys) == 300 and keys == sorted(str(i) for i in range(300))
    for i in (0, 7, 8, 31, 32, 255, 299):
        np.testing.assert_array_equal(f[f"g/{i}"][()], np.arange(i % 7 + 1))


def test_frame_dataset_roundtrip_gzip_and_bool(tmp_path):
    rng = np.random.default_rng(0)
    dims = (5, 5, 5, 6)
    frames = {"1ubq": {"A": {str(i): (rng.random(dims).astype(np.float32), ["MET", "GLN", "ILE"][i % 3])
                             for i in (1, 2, 10, 11)}},
              "2abc": {"B": {"7": (rng.random(dims).astype(np.float32), "TRP")}}}
    p = tmp_path / "data.hdf5"
    write_frame_dataset(p, frames, dims)
    with File(p) as f:
        assert list(f) == ["1ubq", "2abc"] and f["1ubq"].keys() == ["A"]
        assert sorted(f["1ubq"]["A"].keys(), key=int) == ["1", "2", "10", "11"]
        assert tuple(f.attrs["frame_dims"]) == dims and f.attrs["voxels_as_gaussian"] == True  # noqa: E712
        assert f.attrs["make_frame_dataset_ver"] == "2.0.0"
        assert [bytes(x).decode() for x in f.attrs["atom_encoder"]] == ["C", "N", "O", "CB", "CA", "Q"]
        for pdb, chains in frames.items():
            for chain, residues in chains.items():
                for res, (arr, label) in residues.items():
                    ds = f[pdb][chain][res]
                    np.testing.assert_array_equal(ds[()], arr)
                    assert ds.attrs["label"] == label
                    assert ds.attrs["encoded_residue"].shape == (20,) and ds.attrs["encoded_residue"].sum() == 1
    bframes = {"1ubq": {"A": {"1": (rng.random(dims) > 0.5, "MET")}}}
    write_frame_dataset(tmp_path / "b.hdf5", bframes, dims, voxels_as_gaussian=False, compression=None)
    f = File(tmp_path / "b.hdf5")
    assert f.attrs["voxels_as_gaussian"] == False  # noqa: E712
    got = f["1ubq/A/1"][()]
    assert got.dtype == np.bool_
    np.testing.assert_array_equal(got, bframes["1ubq"]["A"]["1"][0])


def test_reader_fails_loudly(tmp_path):
    (tmp_path / "x.h5").write_bytes(b"not hdf5" * 100)
    with pytest.raises(Hdf5FormatError):
        File(tmp_path / "x.h5")
    cfg, w = standins.tiny_standin(calib_frames=0)
    write_keras_h5(tmp_path / "m.h5", cfg, w)
    raw = bytearray((tmp_path / "m.h5").read_bytes())
    raw[8] = 9                                 # unknown superblock version
    (tmp_path / "bad.h5").write_bytes(bytes(raw))
    with pytest.raises(Hdf5FormatError):
        File(tmp_path / "bad.h5")


def test_npz_container_roundtrip(tmp_path):
    from timed_design_b200.model import read_model_file, save_npz
    cfg, w = standins.tiny_standin(calib_frames=0)
    save_npz(tmp_path / "m.npz", cfg, w)
    cfg2, w2 = read_model_file(tmp_path / "m.npz")
    assert cfg2 == json.loads(json.dumps(cfg))
    np.testing.assert_array_equal(w2["conv3d"]["kernel:0"], w["conv3d"]["kernel:0"])


def _frames_file(tmp_path, name, frames, dims, gaussian, compression, chunks, labels_onehot):
    """aposteriori-style dataset written through the low-level Writer so that chunking can be chosen."""
    w = Writer()
    w.root.attrs.update({"make_frame_dataset_ver": "2.0.0", "frame_dims": np.asarray(dims, dtype=np.int64),
                         "voxels_as_gaussian": bool(gaussian), "frame_edge_length": 21.0})
    for i, fr in enumerate(frames):
        ds = w.root.dataset(f"1abc/A/{i + 1}", fr, compression=compression, chunks=chunks)
        ds.attrs["label"] = "ALA"
        ds.attrs["encoded_residue"] = labels_onehot[i]
    p = tmp_path / name
    w.save(p)
    return p


@pytest.mark.parametrize("dtype,gaussian,compression,chunks", [
    (np.float64, True, "gzip", None), (np.float32, True, "gzip", (4, 5, 9, 2)), (np.float64, True, None, (5, 5, 5, 3)),
    (np.bool_, False, "gzip", None), (np.bool_, False, "gzip", (9, 4, 3, 6))])
def test_native_frame_inflater_matches_python_reader(tmp_path, dtype, gaussian, compression, chunks, monkeypatch):
    """timed_b200_inflate_chunks (zlib + scatter + cast on host threads, no device needed) fills load_batch's arrays
    exactly as the pure-Python reader does: whole-frame and ragged multi-chunk layouts, stored float64 / float32 /
    boolean frames, compressed or not, any thread count."""
    from timed_design_b200 import frames as F
    rng = np.random.default_rng(4)
    dims = (9, 9, 9, 6)
    n = 23
    raw = rng.random((n, *dims))
    data = (raw > 0.7) if dtype is np.bool_ else raw.astype(dtype)
    onehot = np.eye(20)[rng.integers(0, 20, n)]
    p = _frames_file(tmp_path, "f.hdf5", data, dims, gaussian, compression, chunks, onehot)
    rows = [("1abc", "A", str(i + 1), "ALA") for i in range(n)][::-1]             # any order
    calls = []
    real = F._native_load

    def spy(*a, **k):
        calls.append(real(*a, **k))
        return calls[-1]

    monkeypatch.setattr(F, "_native_load", spy)
    for threads in ("1", "5"):
        monkeypatch.setenv("TIMED_B200_LOADER_THREADS", threads)
        X, y = F.load_batch(p, rows)
        assert calls[-1] is (compression is not None)     # contiguous (uncompressed) datasets stay on the Python reader
        assert X.dtype == (np.float32 if gaussian else np.bool_)
        np.testing.assert_array_equal(X, data[::-1].astype(X.dtype))
        np.testing.assert_array_equal(y, onehot[::-1])
    monkeypatch.setattr(F, "_native_load", lambda *a, **k: False)                   # pure-Python reader
    X2, y2 = F.load_batch(p, rows)
    np.testing.assert_array_equal(X2, X)
    np.testing.assert_array_equal(y2, y)


def test_native_frame_inflater_reports_corrupt_chunks(tmp_path):
    from timed_design_b200 import _lib, frames as F
    dims = (6, 6, 6, 4)
    data = np.random.default_rng(0).random((3, *dims))
    p = _frames_file(tmp_path, "c.hdf5", data, dims, True, "gzip", None, np.eye(20)[:3])
    f = File(p)
    _, table, *_ = f["1abc/A/2"].chunk_table()
    blob = bytearray(p.read_bytes())
    off = table[0][1]
    blob[off + 8:off + 16] = bytes([255] * 8)                                       # damage the deflate stream
    bad = tmp_path / "bad.hdf5"
    bad.write_bytes(bytes(blob))
    with pytest.raises(_lib.TimedB200Error):
        F.load_batch(bad, [("1abc", "A", "2", "ALA")])


def test_native_frame_index_equals_the_python_walk(tmp_path):
    """timed_b200_hdf5_frame_index (one native call per batch: object headers, chunk B-trees, the encoded_residue attribute)
    returns the chunk offsets / stored sizes / labels the Python reader finds frame by frame; files stored any other way,
    or with one deviating frame, are refused (the caller then walks the batch with the Python reader)."""
    from timed_design_b200 import frames as fr
    from timed_design_b200.hdf5 import write_frame_dataset
    from timed_design_b200.postprocess import standard_amino_acids
    labs = list(standard_amino_acids.values())
    rng = np.random.default_rng(0)
    dims = (9, 9, 9, 6)
    tree, k = {}, 0
    for c in range(3):
        for r in range(40):
            x = np.where(rng.random(dims) < 0.05, rng.random(dims), 0).astype(np.float32)
            tree.setdefault("1abc" if c < 2 else "2xyz", {}).setdefault("ABC"[c], {})[str(r - 3)] = (x, labs[k % 20])
            k += 1
    p = tmp_path / "x.hdf5"
    write_frame_dataset(p, tree, dims, compression="gzip")
    flat, _ = fr.create_flat_dataset_map(p)
    rows = [flat[i] for i in rng.permutation(len(flat))]                       # any order, any subset
    f = fr._open(p)
    fast = fr._fast_frame_index(f, rows, dims)
    assert fast is not None
    offs, sizes, y, dtype = fast
    assert dtype == np.float32 and y.shape == (len(rows), 20) and (y.sum(1) == 1).all()
    for i, row in enumerate(rows):
        ds = f[str(row[0])][str(row[1])][str(row[2])]
        (org, off, size), = ds.chunk_table()[1]
        assert (off, size) == (offs[i], sizes[i]) and np.array_equal(ds.attrs["encoded_residue"], y[i])
    # stored without compression: not this route
    q = tmp_path / "plain.hdf5"
    write_frame_dataset(q, {"1abc": {"A": {"1": tree["1abc"]["A"]["1"]}}}, dims, compression=None)
    fq, _ = fr.create_flat_dataset_map(q)
    assert fr._fast_frame_index(fr._open(q), fq, dims) is None
    # one frame's object header damaged: the whole batch is refused
    raw = bytearray(p.read_bytes())
    addr = f["1abc"]["B"]._load()["5"] + f.base_addr
    assert raw[addr] == 1
    raw[addr] = 7
    bad = tmp_path / "bad.hdf5"
    bad.write_bytes(bytes(raw))
    assert fr._fast_frame_index(fr._open(bad), [r for r in flat if not (r[1] == "B" and r[2] == "5")][:20], dims) is not None
    assert fr._fast_frame_index(fr._open(bad), flat, dims) is None
    # a row that does not exist
    assert fr._fast_frame_index(f, [("1abc", "A", "999", "ALA")], dims) is None


def test_native_frame_index_survives_corrupt_files(tmp_path):
    """Random byte damage in and around the object headers, truncated lengths: the native index refuses frames (status != 0)
    or returns chunks that lie inside the file -- it never reads outside the mapping."""
    import ctypes as C

    from timed_design_b200 import _lib
    from timed_design_b200 import frames as fr
    from timed_design_b200.hdf5 import write_frame_dataset
    rng = np.random.default_rng(0)
    dims = (7, 7, 7, 6)
    tree = {"1abc": {ch: {str(r + 1): (np.where(rng.random(dims) < 0.05, rng.random(dims), 0).astype(np.float32), "ALA")
                          for r in range(30)} for ch in "AB"}}
    p = tmp_path / "x.hdf5"
    write_frame_dataset(p, tree, dims, compression="gzip")
    flat, _ = fr.create_flat_dataset_map(p)
    f = fr._open(p)
    assert fr._fast_frame_index(f, flat, dims) is not None
    ds0 = f[flat[0][0]][flat[0][1]][flat[0][2]]
    raw_m = {m.type: bytes(m.data) for m in ds0._msgs if m.type in (1, 3, 0x0B)}
    hdr = next(bytes(m.data)[:fr._attr_header(bytes(m.data))[1]] for m in ds0._msgs
               if m.type == 0x0C and fr._attr_header(bytes(m.data))[0] == b"encoded_residue")
    addrs = np.array([f[r[0]][r[1]]._load()[r[2]] for r in flat], np.int64)
    raw0, n, lib = p.read_bytes(), len(flat), _lib.load()
    refused = 0
    for _ in range(300):
        raw = bytearray(raw0)
        for _ in range(rng.integers(1, 6)):
            a = int(addrs[rng.integers(n)]) + int(rng.integers(0, 400)) if rng.random() < 0.7 else int(rng.integers(0, len(raw)))
            if a < len(raw):
                raw[a] = int(rng.integers(0, 256))
        buf = np.frombuffer(bytes(raw), np.uint8)
        flen = len(buf) if rng.random() < 0.8 else int(rng.integers(1, len(buf)))
        offs, sizes, lab, st = np.empty(n, np.int64), np.empty(n, np.int64), np.empty((n, 20)), np.empty(n, np.int32)
        rc = lib.timed_b200_hdf5_frame_index(
            C.c_void_p(buf.ctypes.data), flen, 0, n, C.c_void_p(addrs.ctypes.data), raw_m[1], len(raw_m[1]), raw_m[3],
            len(raw_m[3]), raw_m[0x0B], len(raw_m[0x0B]), hdr, len(hdr), 160, len(dims), C.c_void_p(offs.ctypes.data),
            C.c_void_p(sizes.ctypes.data), C.c_void_p(lab.ctypes.data), C.c_void_p(st.ctypes.data), 2)
        assert rc == 0
        ok = st == 0
        assert ((offs[ok] >= 0) & (sizes[ok] > 0) & (offs[ok] + sizes[ok] <= flen)).all()
        refused += int((~ok).sum())
    assert refused > 0
