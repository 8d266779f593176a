"""GPU: the device-side zlib inflater (csrc/inflate.cuh, one warp per stream) against zlib itself, bit for bit."""
import ctypes as C
import zlib

import numpy as np
import pytest

from timed_design_b200 import _lib


def _inflate_device(streams, out_bytes):
    import torch
    lib = _lib.load()
    sizes = np.array([len(s) for s in streams], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    comp = torch.from_numpy(np.frombuffer(b"".join(streams), dtype=np.uint8).copy()).cuda()
    d_off, d_size = torch.from_numpy(offs).cuda(), torch.from_numpy(sizes).cuda()
    out = torch.full((len(streams), out_bytes), 0xAB, dtype=torch.uint8, device="cuda")
    status = torch.full((len(streams),), -1, dtype=torch.int32, device="cuda")
    _lib.check(lib.timed_b200_inflate_device(C.c_void_p(comp.data_ptr()), len(streams), C.c_void_p(d_off.data_ptr()),
                                             C.c_void_p(d_size.data_ptr()), out_bytes, C.c_void_p(out.data_ptr()),
                                             C.c_void_p(status.data_ptr()), None))
    torch.cuda.synchronize()
    return out.cpu().numpy(), status.cpu().numpy()


def _payloads(n_bytes, rng):
    """Byte strings of one length that drive every block type and match shape."""
    sparse = np.zeros(max(1, n_bytes // 4), np.float32)
    idx = rng.choice(len(sparse), max(1, len(sparse) // 40), replace=False)
    sparse[idx] = rng.random(len(idx)).astype(np.float32)
    text = (b"the quick brown fox jumps over the lazy dog. " * (n_bytes // 45 + 1))[:n_bytes]
    ramp = (np.arange(n_bytes) % 251).astype(np.uint8).tobytes()
    return {"zeros": bytes(n_bytes), "random": rng.integers(0, 256, n_bytes, dtype=np.uint8).tobytes(),
            "sparse_f32": sparse.tobytes().ljust(n_bytes, b"\0")[:n_bytes], "text": text, "ramp": ramp,
            "few_symbols": rng.choice(np.frombuffer(b"ab", np.uint8), n_bytes).tobytes()}


@pytest.mark.gpu
@pytest.mark.parametrize("n_bytes", [1, 7, 300, 4096, 70000, 222264])
def test_device_inflate_equals_zlib(n_bytes):
    """Stored (level 0, incl. multi-block above 64 KB), fixed-Huffman (tiny inputs) and dynamic-Huffman streams; literals only,
    long overlapping matches (runs of zeros), far matches; every stream of a launch inflates to the same length."""
    rng = np.random.default_rng(n_bytes)
    data = _payloads(n_bytes, rng)
    streams, want = [], []
    for name, raw in data.items():
        for level in (0, 1, 6, 9):
            streams.append(zlib.compress(raw, level))
            want.append(raw)
        co = zlib.compressobj(6, zlib.DEFLATED, 15, 9, zlib.Z_FIXED)            # fixed Huffman codes whatever the size
        streams.append(co.compress(raw) + co.flush())
        want.append(raw)
    out, status = _inflate_device(streams, n_bytes)
    assert (status == 0).all(), status
    for i, raw in enumerate(want):
        assert out[i].tobytes() == raw, f"stream {i}"


@pytest.mark.gpu
def test_device_inflate_flags_bad_streams():
    rng = np.random.default_rng(5)
    raw = _payloads(5000, rng)["sparse_f32"]
    good = zlib.compress(raw, 6)
    bad_header = b"\x79" + good[1:]
    truncated = good[: len(good) // 2]
    wrong_len = zlib.compress(raw[:-10], 6)                     # inflates to fewer bytes than asked for
    too_long = zlib.compress(raw + b"x" * 10, 6)
    corrupt = bytearray(good)
    corrupt[len(corrupt) // 2] ^= 0x5A
    out, status = _inflate_device([good, bad_header, truncated, wrong_len, too_long, good], 5000)
    assert status[0] == 0 and status[5] == 0 and out[0].tobytes() == raw and out[5].tobytes() == raw
    assert status[1] == 1 and status[2] != 0 and status[3] == 5 and status[4] == 3
    _, st = _inflate_device([bytes(corrupt)], 5000)             # either detected or (rarely) a different valid stream: never a hang
    assert st[0] in (0, 2, 3, 4, 5)


@pytest.mark.gpu
def test_device_inflate_route_of_predict_equals_the_host_route(tmp_path, monkeypatch):
    """frames.load_batch_device (stored gzip chunks -> GPU -> one warp per chunk) returns load_batch's frames and labels
    byte for byte, and predict.py writes identical files whichever route reads the dataset
    (TIMED_B200_NO_DEVICE_INFLATE = host route)."""
    import json
    from pathlib import Path

    from timed_design_b200 import frames, predict, standins
    from timed_design_b200.hdf5 import write_frame_dataset, write_keras_h5
    ubq = json.loads((Path(__file__).parent / "golden" / "1ubq_chainA.json").read_text())
    X = standins.synthetic_frames(76, seed=3)
    X[X < 0.6] = 0.0                                              # sparse, as voxelised structures are
    residues = {str(rid): (X[i], ubq["labels"][i]) for i, rid in enumerate(ubq["residue_ids"])}
    write_frame_dataset(tmp_path / "data.hdf5", {"1ubq": {"A": residues}, "2xyz": {"B": dict(list(residues.items())[:9])}},
                        (21, 21, 21, 6), compression="gzip")
    flat, _ = frames.create_flat_dataset_map(tmp_path / "data.hdf5")
    Xh, yh = frames.load_batch(tmp_path / "data.hdf5", flat)
    dev = frames.load_batch_device(tmp_path / "data.hdf5", flat)
    assert dev is not None, "gzip single-chunk float frames must take the device route"
    assert np.array_equal(dev[0].cpu().numpy(), Xh) and np.array_equal(dev[1], yh)
    some = [flat[i] for i in (80, 3, 41)]                         # scattered rows: chunks far apart in the file
    d2 = frames.load_batch_device(tmp_path / "data.hdf5", some)
    assert np.array_equal(d2[0].cpu().numpy(), frames.load_batch(tmp_path / "data.hdf5", some)[0])
    cfg, w = standins.timed_standin(20, filters=(8, 16, 16, 24, 32), calib_frames=4)
    write_keras_h5(tmp_path / "TIMED.h5", cfg, w)
    monkeypatch.chdir(tmp_path)
    outs = []
    for name, env in (("dev", None), ("host", "1")):
        if env:
            monkeypatch.setenv("TIMED_B200_NO_DEVICE_INFLATE", env)
        out = tmp_path / name
        predict.cli(["--path_to_dataset", str(tmp_path / "data.hdf5"), "--path_to_model", str(tmp_path / "TIMED.h5"),
                     "--path_to_output", str(out), "--path_to_datasetmap", str(out / "datasetmap.txt"), "--batch_size", "32",
                     "--yes"])
        outs.append({p.name: p.read_bytes() for p in sorted(out.iterdir())})
    assert outs[0].keys() == outs[1].keys() and "TIMED.csv" in outs[0]
    for k in outs[0]:
        assert outs[0][k] == outs[1][k], k
