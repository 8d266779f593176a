"""GPU parity of the Monte-Carlo sampler (timed_b200_sample & co through the host API) against
golden vectors produced by the reference's own functions and against the numpy oracle.
Indices/letters: bit-exact under injected uniforms.  Temperature: <= 4 ulp (CUDA pow is 2 ulp)."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import sampler_oracle as so

G = Path(__file__).parent / "golden"
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def su():
    from timed_design_b200 import sampling_utils
    return sampling_utils


@pytest.fixture(scope="module")
def golden():
    return np.load(G / "sampler.npz")


def _rot_letters():
    from timed_design_b200.postprocess import _THREE_TO_ONE
    cats = json.loads((G / "rotamer_codec.json").read_text())["flat_categories"]
    return [_THREE_TO_ONE[c.split("_")[0]] for c in cats]


@pytest.mark.parametrize("c", [20, 338])
def test_sample_block_matches_reference_golden(su, golden, c):
    probs, r = golden[f"probs_{c}"], golden[f"r_{c}"]
    cats = None if c == 20 else _rot_letters()
    seqs, idx = su.sample_block(probs, r.shape[0], cats, uniforms=r, return_idx=True)
    np.testing.assert_array_equal(idx, golden[f"idx_{c}"])
    assert [row.tobytes().decode() for row in seqs] == list(golden[f"seq_{c}"])


def test_cumsum_bit_identical_to_numpy(su, golden):
    import ctypes as C
    import torch
    from timed_design_b200 import _lib
    for c in (20, 338):
        p = torch.from_numpy(golden[f"probs_{c}"]).cuda()
        out = torch.empty_like(p)
        _lib.check(_lib.load().timed_b200_cumsum_rows(C.c_void_p(p.data_ptr()), p.shape[0], p.shape[1],
                                                     C.c_void_p(out.data_ptr()), None))
        np.testing.assert_array_equal(out.cpu().numpy(), golden[f"cumsum_{c}"])


def test_random_choice_prob_index_api(su, golden):
    probs, r = golden["probs_20"], golden["r_20"]
    idx = su.random_choice_prob_index(probs, return_seq=False, uniforms=r[3])
    np.testing.assert_array_equal(idx, golden["idx_20"][3])
    seq = su.random_choice_prob_index(probs, return_seq=True, uniforms=r[3])
    assert "".join(seq) == str(golden["seq_20"][3])
    seq = su.random_choice_prob_index(golden["probs_338"], return_seq=True, rotamer_categories=_rot_letters(),
                                      uniforms=golden["r_338"][5])
    assert "".join(seq) == str(golden["seq_338"][5])
    # axis=0: categories along rows
    idx0 = su.random_choice_prob_index(probs.T.copy(), axis=0, return_seq=False, uniforms=r[2])
    np.testing.assert_array_equal(idx0, golden["idx_20"][2])


def test_temperature_matches_reference_golden(su):
    t = np.load(G / "temperature.npz")
    for c in (20, 338):
        for temp in (0.1, 0.5, 1, 2.0, 5.0, 100):
            got = su.apply_temp_to_probs(t[f"probs_{c}"], temp)
            ref = t[f"out_{c}_t{temp}"]
            np.testing.assert_allclose(got, ref, rtol=1e-14, atol=1e-300)
    # properties the reference's own test pins (tests/test_sampling_utils.py:47-62)
    row = t["test_row"]
    assert np.allclose(su.apply_temp_to_probs(row, 1), row)
    cold = su.apply_temp_to_probs(row, 0.01)
    assert np.argmax(cold) == np.argmax(row) and np.isclose(cold[:, np.argmax(cold)], 1.0)
    assert np.allclose(su.apply_temp_to_probs(row, 100), 0.05, rtol=0.01, atol=0.01)


def test_philox_uniforms_bit_exact(su):
    import ctypes as C
    import torch
    from timed_design_b200 import _lib
    n_res, n_s, first, seed, sid = 13, 9, 1234567, 42, 5
    out = torch.empty((n_s, n_res), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().timed_b200_sample_uniforms(n_res, n_s, first, seed, sid, C.c_void_p(out.data_ptr()), None))
    np.testing.assert_array_equal(out.cpu().numpy(), so.philox_uniforms(n_res, n_s, first, seed, sid))


def test_philox_draws_equal_injected_uniforms_and_shard_invariant(su, golden):
    probs = golden["probs_20"]
    n_res = probs.shape[0]
    u = so.philox_uniforms(n_res, 12, 0, 7, 3)
    a, ia = su.sample_block(probs, 12, seed=7, stream_id=3, return_idx=True)
    b, ib = su.sample_block(probs, 12, uniforms=u, return_idx=True)
    np.testing.assert_array_equal(ia, ib)
    np.testing.assert_array_equal(a, b)
    # two shards (as two GPUs would draw them) concatenate to the single-device result
    s0, _ = su.sample_block(probs, 5, seed=7, stream_id=3, first_sample=0)
    s1, _ = su.sample_block(probs, 7, seed=7, stream_id=3, first_sample=5)
    np.testing.assert_array_equal(np.concatenate([s0, s1]), a)


def test_distribution_of_reference_test(su, golden):
    """tests/test_sampling_utils.py:31-44: 1e6 draws of the golden row, frequencies within 0.01."""
    row = golden["test_row"]
    seqs, idx = su.sample_block(np.repeat(row, 1000, axis=0), 1000, seed=123, return_idx=True)
    freq = np.bincount(idx.ravel(), minlength=20) / idx.size
    assert np.isclose(freq.sum(), row.sum(), rtol=0.01)
    assert np.allclose(row[0], freq, rtol=0.01, atol=0.01)


def test_sample_with_multiprocessing_and_save_as(su, golden, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    probs = {"1abcA": golden["probs_20"].tolist(), "2xyzA": golden["probs_20"][:11].tolist()}
    su.set_seed(42)
    out = su.sample_with_multiprocessing(8, list(probs), 6, probs, None)
    assert list(out) == ["1abcA", "2xyzA"]
    assert all(len(v) == 6 and len(v[0]) == 5 for v in out.values())
    assert len(out["1abcA"][0][0]) == 37 and len(out["2xyzA"][0][0]) == 11
    again = su.sample_with_multiprocessing(1, list(probs), 6, probs, None)
    assert out == again                                  # seeded and worker-count independent
    paths = su.save_as(out, "TIMED_temp_1_n_6_1abcA", "all")
    assert paths == ["TIMED_temp_1_n_6_1abcA.json", "TIMED_temp_1_n_6_1abcA.fasta", "TIMED_temp_1_n_6_1abcA_metrics.csv"]
    d = json.loads((tmp_path / paths[0]).read_text())
    assert d["1abcA"][0][0] == out["1abcA"][0][0]
    fasta = (tmp_path / paths[1]).read_text().splitlines()
    assert fasta[0] == ">1abcA_0" and fasta[1] == out["1abcA"][0][0]
    assert (tmp_path / paths[2]).read_text().splitlines()[0] == \
        "pdb,sequence,charge,isoelectric_point,molecular_weight,molar_extinction"


def test_save_as_byte_identical_to_reference(su, tmp_path, monkeypatch):
    files = json.loads((G / "files.json").read_text())
    monkeypatch.chdir(tmp_path)
    sampled = {"1abcA": [("ACDKW", 1.0, 7.0, 550.0, 5500.0), ("AAAAA", 0.0, 7.0, 550.0, 0.0)],
               "2xyzA": [("KRDEW", 0.0, 7.0, 550.0, 5500.0)]}
    su.save_as(sampled, "TIMED_temp_0.5_n_2_1abcA", "all")
    for ext in (".json", ".fasta", "_metrics.csv"):
        name = "TIMED_temp_0.5_n_2_1abcA" + ext
        assert (tmp_path / name).read_text() == files[name], name


def test_argmax_fp16_kernel():
    import ctypes as C
    import torch
    from timed_design_b200 import _lib
    rng = np.random.default_rng(5)
    p = rng.dirichlet(np.ones(338) * 0.3, size=4097).astype(np.float32)
    p[7, 100] = p[7, 200] = 0.75        # fp16 tie -> first index
    d = torch.from_numpy(p).cuda()
    out = torch.empty(len(p), dtype=torch.int32, device="cuda")
    _lib.check(_lib.load().timed_b200_argmax_fp16(C.c_void_p(d.data_ptr()), len(p), 338, C.c_void_p(out.data_ptr()), None))
    np.testing.assert_array_equal(out.cpu().numpy(), np.argmax(p.astype(np.float16), axis=1))


def test_sample_chains_equals_per_chain_launches(su):
    """One launch over all chains (timed_b200_sample_chains) draws byte-for-byte what per-chain launches keyed
    (seed, stream_id0 + chain) draw -- ragged lengths, a one-residue chain, blocks that are not multiples of four."""
    rng = np.random.default_rng(17)
    for n_cls in (20, 338):
        cats = None if n_cls == 20 else list(("ACDEFGHIKLMNPQRSTVWY" * 17)[:338])
        chains = [rng.dirichlet(np.ones(n_cls), size=n).astype(np.float16).astype(np.float64) for n in (57, 1, 130, 3, 76)]
        got, metrics = su.sample_chains(chains, 11, cats, seed=5, stream_id0=40, return_metrics=True)
        for i, p in enumerate(chains):
            ref, _ = su.sample_block(p, 11, cats, seed=5, stream_id=40 + i)
            np.testing.assert_array_equal(got[i], ref)
            assert metrics[i].shape == (11, 4)
        cold = su.sample_chains(chains, 4, cats, seed=5, temperature=0.01)
        for i, p in enumerate(chains):
            ref, _ = su.sample_block(p, 4, cats, seed=5, stream_id=i, temperature=0.01)
            np.testing.assert_array_equal(cold[i], ref)
