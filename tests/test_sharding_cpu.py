"""CPU (gloo, world_size 2): the N>1 host logic -- contiguous frame sharding + one all-gather of the
probability blocks reproduces the single-process matrix, including uneven and empty shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from timed_design_b200.dist import gather_rows, shard_range, shard_samples


def _fake_probs(idx: np.ndarray, c: int) -> np.ndarray:
    z = np.sin(np.outer(idx + 1, np.arange(1, c + 1)) * 0.37)
    e = np.exp(z)
    return (e / e.sum(1, keepdims=True)).astype(np.float32)


def _worker(rank, world, port, n_total, c, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop = shard_range(n_total, rank, world)
        local = torch.from_numpy(_fake_probs(np.arange(start, stop), c)) if stop > start else torch.zeros((0, c))
        full = gather_rows(local, n_total)
        q.put((rank, full.numpy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_total,c", [(76, 20), (7, 338), (1, 20)])
def test_gather_rows_world2_matches_single_process(n_total, c):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, c, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _fake_probs(np.arange(n_total), c)
    for r in range(2):
        np.testing.assert_array_equal(results[r], ref)


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 76, 4096, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= s <= e for s, e in spans)
            assert max(e - s for s, e in spans) == (-(-n // world) if n else 0)
    assert shard_samples(10, 1, 4) == (3, 3) and shard_samples(10, 3, 4) == (9, 1)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_gather_rows_without_process_group_is_identity():
    x = torch.arange(12, dtype=torch.float32).reshape(6, 2)
    assert torch.equal(gather_rows(x, 6), x)


# ----------------------------------------------------------------------------- predict driver, world_size 2 (gloo)
class _FakeModel:
    """Deterministic stand-in for the device model (the CPU suite has no GPU): probabilities from channel sums."""
    n_classes = 20

    def predict(self, X, batch_size=None):
        X = np.asarray(X, dtype=np.float64)
        z = X.reshape(len(X), -1, X.shape[-1]).sum(1) @ np.sin(np.arange(X.shape[-1] * 20).reshape(X.shape[-1], 20))
        e = np.exp(z - z.max(1, keepdims=True))
        return (e / e.sum(1, keepdims=True)).astype(np.float32)

    def close(self):
        pass


def _write_dataset(path, n=37):
    from timed_design_b200 import standins
    from timed_design_b200.hdf5 import write_frame_dataset
    frames = standins.synthetic_frames(n, side=7, seed=3)
    labels = ["ALA", "GLY", "LEU", "LYS", "SER"]
    chains = {"A": {str(i + 1): (frames[i], labels[i % 5]) for i in range(n - 10)},
              "B": {str(i + 1): (frames[n - 10 + i], labels[i % 5]) for i in range(10)}}
    write_frame_dataset(path, {"1abc": chains}, (7, 7, 7, 6))


def _predict_worker(rank, world, port, data, out_dir, start_batch=0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world), TIMED_B200_DIST_BACKEND="gloo")
    from timed_design_b200 import predict
    predict.load_model = lambda path, device=0, **kw: _FakeModel()
    predict.load_dataset_and_predict([__import__("pathlib").Path("TIMED.h5")], data, batch_size=8, start_batch=start_batch,
                                     dataset_map_path=os.path.join(out_dir, "datasetmap.txt"), path_to_output=out_dir)
    dist.destroy_process_group()


def test_predict_driver_world2_resumes_from_start_batch(tmp_path, monkeypatch):
    """start_batch under sharding (predict.py:32,54-57,126): a run that stopped after three batches is resumed by two
    ranks; the appended files equal those of an uninterrupted single-process run byte for byte."""
    from pathlib import Path
    from timed_design_b200 import predict
    data = tmp_path / "d.hdf5"
    _write_dataset(data)
    monkeypatch.setattr(predict, "load_model", lambda path, device=0, **kw: _FakeModel())
    monkeypatch.chdir(tmp_path)
    full = tmp_path / "full"
    full.mkdir()
    predict.load_dataset_and_predict([Path("TIMED.h5")], data, batch_size=8, dataset_map_path=full / "datasetmap.txt",
                                     path_to_output=full)
    # the interrupted run: keep what the first three batches appended (24 rows of each per-frame file + the map)
    part = tmp_path / "part"
    part.mkdir()
    for name in ("TIMED.csv", "encoded_labels.csv"):
        (part / name).write_text("".join((full / name).read_text().splitlines(keepends=True)[:24]))
    (part / "datasetmap.txt").write_bytes((full / "datasetmap.txt").read_bytes())
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_predict_worker, args=(r, 2, port, str(data), str(part), 3)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    for name in sorted(f.name for f in full.iterdir()):
        assert (full / name).read_bytes() == (part / name).read_bytes(), name


def test_predict_driver_world2_writes_the_single_process_files(tmp_path, monkeypatch):
    """torchrun-style launch of load_dataset_and_predict with two ranks (gloo): frames sharded by flat index, one
    all-gather, rank 0 writes -- every output file is byte-identical to the single-process run."""
    from pathlib import Path
    from timed_design_b200 import predict
    data = tmp_path / "d.hdf5"
    _write_dataset(data)
    single = tmp_path / "single"
    single.mkdir()
    monkeypatch.setattr(predict, "load_model", lambda path, device=0, **kw: _FakeModel())
    monkeypatch.chdir(tmp_path)
    predict.load_dataset_and_predict([Path("TIMED.h5")], data, batch_size=8,
                                     dataset_map_path=single / "datasetmap.txt", path_to_output=single)
    multi = tmp_path / "multi"
    multi.mkdir()
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_predict_worker, args=(r, 2, port, str(data), str(multi))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    names = sorted(f.name for f in single.iterdir())
    assert names == sorted(f.name for f in multi.iterdir()) and "TIMED.csv" in names and "TIMED.fasta" in names
    for name in names:
        assert (single / name).read_bytes() == (multi / name).read_bytes(), name


# ----------------------------------------------------------------------------- sampler fan-out, world_size 2 (gloo)
def _fake_sample_chains(prob_list, sample_n, cats=None, *, seed=None, stream_id0=0, first_sample=0, temperature=None,
                        return_metrics=False):
    """Counter-based like the device sampler: the letter of (chain, sample, residue) depends on the GLOBAL sample index."""
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    seqs, mets = [], []
    for c, p in enumerate(prob_list):
        n = len(p)
        s_idx = first_sample + np.arange(sample_n)[:, None]
        seqs.append(letters[(7 * s_idx + 3 * np.arange(n)[None, :] + c) % 20].astype(np.uint8))
        mets.append(np.stack([s_idx[:, 0] + c, s_idx[:, 0] * 0.5, s_idx[:, 0] + 100.0 * n, s_idx[:, 0] * 0 + c], axis=1).astype(np.float64))
    return (seqs, mets) if return_metrics else seqs


def _sample_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world), TIMED_B200_DIST_BACKEND="gloo")
    from timed_design_b200 import sampling_utils as su
    su.sample_chains = _fake_sample_chains
    probs = {"1abcA": np.ones((11, 20)) / 20, "1abcB": np.ones((4, 20)) / 20, "2xyzA": np.ones((1, 20)) / 20}
    out = su.sample_with_multiprocessing(8, list(probs), 7, probs, None)
    q.put((rank, out))
    dist.destroy_process_group()


def test_sampler_fanout_world2_equals_single_process(monkeypatch):
    """sample_with_multiprocessing under a two-rank launch: each rank draws a block of the sample index, the gathered
    result on every rank equals the single-process one (letters and metrics, chain by chain)."""
    from timed_design_b200 import sampling_utils as su
    monkeypatch.setattr(su, "sample_chains", _fake_sample_chains)
    probs = {"1abcA": np.ones((11, 20)) / 20, "1abcB": np.ones((4, 20)) / 20, "2xyzA": np.ones((1, 20)) / 20}
    ref = su.sample_with_multiprocessing(8, list(probs), 7, probs, None)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sample_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(ref["1abcA"]) == 7 and len(ref["1abcA"][0][0]) == 11
    for r in range(2):
        assert results[r] == ref
