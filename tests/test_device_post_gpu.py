"""GPU parity of the device-side post-processing: NMR consensus and the metrics of
sampled sequences, both against restatements that the CPU suite pins to the reference's own code."""
import json
from pathlib import Path

import numpy as np
import pytest

from timed_design_b200 import device_post, postprocess, seq_metrics

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def test_consensus_matches_reference_golden():
    """The golden NMR case was produced by the reference's extract_sequence_from_pred_matrix (make_golden.py)."""
    z = np.load(GOLDEN / "postprocess.npz", allow_pickle=False)
    meta = json.loads((GOLDEN / "postprocess.json").read_text())
    pm = z["pm_nmr"].astype(np.float16)
    out = postprocess.extract_sequence_from_pred_matrix(z["nmr_map"], pm, None, is_consensus=False)
    cons, cons_prob = device_post.nmr_consensus(out[1], None)
    assert cons == meta["nmr"]["consensus"]
    for k, v in cons_prob.items():
        assert v.dtype == np.float16
        np.testing.assert_array_equal(v.astype(np.float64), z[f"nmr_consensus_prob_{k}"])


@pytest.mark.parametrize("n_cls", [20, 338])
def test_consensus_random_ensembles_bit_exact(n_cls):
    """Several structures with 1..5 states, re-appearing names, near-ties: fp16 consensus arrays bit-identical to the
    numpy restatement (float16 arithmetic, running pairwise mean) and identical consensus sequences."""
    rng = np.random.default_rng(n_cls)
    cats = None if n_cls == 20 else list(("ACDEFGHIKLMNPQRSTVWY" * 17)[:338])
    pdb_to_prob = {}
    for name, n_states, n_res in (("1abc", 3, 57), ("2xyz", 1, 11), ("3nmr", 5, 130), ("4one", 2, 1), ("1abc", 2, 57)):
        for s in range(n_states):
            logits = rng.standard_normal((n_res, n_cls)) * 2
            p = np.exp(logits - logits.max(1, keepdims=True))
            p = (p / p.sum(1, keepdims=True)).astype(np.float16)
            if s and n_res > 3:
                p[:3] = pdb_to_prob[f"{name}_{s - 1}A"][:3]            # identical rows: exact ties survive the mean
            key = f"{name}_{s}A"
            while key in pdb_to_prob:
                key += "x"
            pdb_to_prob[key] = p
    # numpy restatement of utils.py:694-713
    ref_prob, last = {}, ""
    for key, p in pdb_to_prob.items():
        cur = key.split("_")[0]
        if cur != last:
            ref_prob[cur] = p
            last = cur
        else:
            ref_prob[cur] = (ref_prob[cur] + p) / 2
    letters = postprocess._letters_for(cats)
    cons, cons_prob = device_post.nmr_consensus(pdb_to_prob, cats)
    assert list(cons) == list(ref_prob)
    for k, v in ref_prob.items():
        assert v.dtype == np.float16
        np.testing.assert_array_equal(cons_prob[k].view(np.uint16), v.view(np.uint16))
        assert cons[k] == "".join(letters[np.argmax(v, axis=1)])


def test_seq_metrics_device_matches_host_restatement():
    rng = np.random.default_rng(3)
    letters = np.frombuffer(seq_metrics.LETTERS.encode(), np.uint8)
    for n_res in (1, 7, 76, 333, 1000):
        seqs = letters[rng.integers(0, 20, size=(257, n_res))]
        seqs[0] = letters[0]                                  # poly-A
        seqs[1] = letters[rng.integers(0, 20)]
        got = device_post.seq_metrics_of(seqs)
        charge, pi, mw, ext = seq_metrics.metrics_from_composition(seq_metrics.composition(seqs))
        np.testing.assert_allclose(got[:, 0], charge, rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(got[:, 2], mw, rtol=1e-12)
        np.testing.assert_array_equal(got[:, 3], ext)
        # the isoelectric point is an argmin over a 0.1-pH grid: equal unless two grid points tie to rounding
        assert (np.abs(got[:, 1] - pi) < 1e-9).mean() >= 0.995
        assert np.abs(got[:, 1] - pi).max() <= 0.1 + 1e-9


def test_seq_metrics_rejects_unknown_letters():
    seqs = np.frombuffer(b"ACDXF", np.uint8)[None, :].copy()
    with pytest.raises(ValueError):
        device_post.seq_metrics_of(seqs)


def test_sample_from_sequences_carries_device_metrics():
    from timed_design_b200 import sampling_utils as su
    rng = np.random.default_rng(9)
    p = rng.dirichlet(np.ones(20), size=40)
    out = su.sample_from_sequences("1ubqA", 25, {"1ubqA": p}, None)
    rows = out["1ubqA"]
    assert len(rows) == 25 and all(len(r) == 5 and len(r[0]) == 40 for r in rows)
    for seq, charge, pi, mw, ext in rows[:5]:
        c2, p2, m2, e2 = seq_metrics.calculate_seq_metrics(seq)
        assert abs(charge - c2) < 1e-9 and abs(mw - m2) < 1e-6 and ext == e2 and abs(pi - p2) <= 0.1 + 1e-9
