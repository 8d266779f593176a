"""CPU: the oracle restatements and the product's host-side post-processing, both pinned to
golden vectors produced by executing the reference's own functions
(tests/golden/make_golden.py) and to the reference's own test properties
(/root/reference/tests/test_sampling_utils.py, tests/test_utils.py)."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import postprocess_oracle as po
from oracle import sampler_oracle as so
from timed_design_b200 import postprocess as pp

G = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def sampler():
    return np.load(G / "sampler.npz")


@pytest.fixture(scope="module")
def post():
    return np.load(G / "postprocess.npz"), json.loads((G / "postprocess.json").read_text())


# ------------------------------------------------------------------ sampler oracle
@pytest.mark.parametrize("c", [20, 338])
def test_oracle_choice_index_matches_reference(sampler, c):
    probs, r, idx = sampler[f"probs_{c}"], sampler[f"r_{c}"], sampler[f"idx_{c}"]
    for s in range(r.shape[0]):
        np.testing.assert_array_equal(so.choice_index(probs, r[s]), idx[s])
    # the "no cumsum entry exceeds r" quirk is in the fixture (sampling_utils.py:82)
    no_true = probs.cumsum(1)[:, -1] <= r[0]
    assert no_true.sum() >= 2 and (idx[0][no_true] == 0).all()


def test_oracle_sequences_match_reference(sampler):
    cats = json.loads((G / "rotamer_codec.json").read_text())["flat_categories"]
    letters = [pp._THREE_TO_ONE[c.split("_")[0]] for c in cats]
    assert so.sample_sequences(sampler["probs_20"], sampler["r_20"]) == list(sampler["seq_20"])
    assert so.sample_sequences(sampler["probs_338"], sampler["r_338"], letters) == list(sampler["seq_338"])


def test_oracle_temperature_matches_reference():
    t = np.load(G / "temperature.npz")
    for c in (20, 338):
        for temp in (0.01, 0.1, 0.5, 1, 2.0, 5.0, 100):
            np.testing.assert_array_equal(so.apply_temp_to_probs(t[f"probs_{c}"], temp), t[f"out_{c}_t{temp}"])


def test_temperature_properties_of_reference_test():
    """tests/test_sampling_utils.py:47-62 restated on the oracle."""
    row = np.load(G / "sampler.npz")["test_row"]
    assert np.allclose(so.apply_temp_to_probs(row, 1), row)
    cold = so.apply_temp_to_probs(row, 0.01)
    assert np.argmax(cold) == np.argmax(row) and np.isclose(cold[:, np.argmax(cold)], 1.0)
    assert np.allclose(so.apply_temp_to_probs(row, 100), 0.05, rtol=0.01, atol=0.01)


def test_oracle_distribution_of_reference_test():
    """tests/test_sampling_utils.py:31-44 (1e6 draws, tol 0.01), vectorised over the draws."""
    row = np.load(G / "sampler.npz")["test_row"]
    r = np.random.default_rng(0).random(1_000_000)
    idx = (row.cumsum(axis=1) > r[:, None]).argmax(axis=1)
    freq = np.bincount(idx, minlength=20) / len(idx)
    assert np.allclose(row[0], freq, rtol=0.01, atol=0.01)


def test_philox_known_answer():
    """Random123 known-answer vectors for philox4x32-10."""
    assert so.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert so.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert so.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = so.philox_uniforms(5, 3, 7, 42, 1)
    assert ((u >= 0) & (u < 1)).all() and len(np.unique(u)) == 15


# ------------------------------------------------------------------ codec / post-processing
def test_rotamer_codec_matches_reference():
    g = json.loads((G / "rotamer_codec.json").read_text())
    for impl in (po, pp):
        codec, labels, guide = impl.get_rotamer_codec(True) if impl is pp else impl.get_rotamer_codec()
        assert labels == g["flat_categories"] and guide == g["reduction_guide"]
        assert [int(np.argmax(codec[i])) for i in range(338)] == g["class_to_residue"]
        assert all(codec[i].sum() == 1 and codec[i].shape == (20,) for i in range(338))
    assert g["reduction_guide"] == [0, 1, 4, 13, 40, 49, 50, 59, 68, 149, 158, 185, 194, 203, 230, 311,
                                    314, 317, 320, 329]          # utils.py:425
    assert labels[:5] == ["ALA_0", "CYS_1", "CYS_2", "CYS_3", "ASP_11"]   # utils.py:422


def test_compress_rotamers(post):
    z, _ = post
    for impl in (po, pp):
        np.testing.assert_array_equal(impl.compress_rotamer_predictions_to_20(z["compress_in"]), z["compress_out"])
    # /root/reference/tests/test_utils.py:6-11
    assert pp.compress_rotamer_predictions_to_20(np.ones((1, 338))).shape[-1] == 20


@pytest.mark.parametrize("impl", [po, pp])
def test_extract_sequences_match_reference(post, impl):
    z, meta = post
    cats = json.loads((G / "rotamer_codec.json").read_text())["flat_categories"]
    pm20 = z["pm20"].astype(np.float16)
    s, p, real, _, _ = impl.extract_sequence_from_pred_matrix(z["old_map"], pm20, None)
    assert s == meta["old"]["seq"] and real == meta["old"]["real"]
    assert {k: list(np.array(v).shape) for k, v in p.items()} == meta["old"]["prob_shapes"]
    assert list(s) == list(meta["old"]["seq"])            # insertion order too
    np.testing.assert_array_equal(np.array(p["1abcB"]), pm20[7:12])
    s, _, real, _, _ = impl.extract_sequence_from_pred_matrix(z["new_map"], pm20, None)
    assert s == meta["new"]["seq"] and real == meta["new"]["real"]
    s, _, _, _, _ = impl.extract_sequence_from_pred_matrix(z["old_map"], z["pm338"].astype(np.float16), cats)
    assert s == meta["rot"]["seq"]
    s, _, _, cons, cons_p = impl.extract_sequence_from_pred_matrix(
        z["nmr_map"], z["pm_nmr"].astype(np.float16), None, is_consensus=True)
    assert s == meta["nmr"]["seq"] and cons == meta["nmr"]["consensus"]
    for k, v in cons_p.items():
        np.testing.assert_array_equal(np.asarray(v, dtype=np.float64), z[f"nmr_consensus_prob_{k}"])


def test_fp16_tie_breaks_to_first_index(post):
    z, meta = post
    pm20 = z["pm20"].astype(np.float16)
    assert pm20[3, 4] == pm20[3, 9] == pm20[3].max()
    assert meta["old"]["seq"]["1abcA"][3] == "F"          # class 4, the first of the tied pair


def test_file_writers_byte_identical(tmp_path, monkeypatch, post):
    z, _ = post
    files = json.loads((G / "files.json").read_text())
    pred32 = np.array(files["__pred32__"], dtype=np.float32)
    old_map = z["old_map"]
    labels = list(pp.standard_amino_acids.values())
    y_true = np.eye(20)[[labels.index(l) for l in old_map[:, 3]]]
    monkeypatch.chdir(tmp_path)
    pp.save_outputs_to_file(list(y_true[:12]), {0: list(pred32[:12])}, old_map, 0, "TIMED", tmp_path)
    pp.save_outputs_to_file(list(y_true[12:]), {0: list(pred32[12:])}, old_map, 0, "TIMED", tmp_path)
    pp.convert_dataset_map_for_srb(old_map, "TIMED", tmp_path)
    seqs = pp.extract_sequence_from_pred_matrix(
        old_map, np.genfromtxt(tmp_path / "TIMED.csv", delimiter=",", dtype=np.float16), None)
    pp.save_dict_to_fasta(seqs[0], "TIMED", tmp_path)
    pp.save_dict_to_fasta(seqs[2], "dataset", tmp_path)
    for name in ("TIMED.csv", "encoded_labels.csv", "datasetmap.txt", "TIMED.txt", "TIMED.fasta", "dataset.fasta"):
        assert (tmp_path / name).read_text() == files[name], name
    # oracle text builders agree as well
    assert po.predictions_csv_text(pred32) == files["TIMED.csv"]
    assert po.srb_map_text(old_map) == files["TIMED.txt"]
    assert po.labels_csv_text(y_true) == files["encoded_labels.csv"]
    # the datasetmap round-trips through the loaders
    m = pp.load_datasetmap(tmp_path / "TIMED.txt")
    assert m.tolist() == [["1abcA", "7"], ["1abcB", "5"], ["2xyzA", "9"]]
    assert pp.load_datasetmap(tmp_path / "datasetmap.txt", is_old=True).shape == (21, 4)
