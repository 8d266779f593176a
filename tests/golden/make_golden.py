"""Generate golden vectors by running the REFERENCE's own functions (read from /root/reference,
never copied) on seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The reference modules import third-party packages that are absent here (ampal, h5py,
aposteriori, logomaker, ...).  The functions on the hot path only need two small ampal tables,
which are stubbed below from knowledge of ampal 1.5.1 and *checked* against the values the
reference pins in-repo: the 20-letter order and the rotamer offsets
[0,1,4,13,40,49,50,59,68,149,158,185,194,203,230,311,314,317,320,329]
(/root/reference/design_utils/utils.py:425).  Everything else executed is the reference's code.

Outputs (committed): sampler.npz, temperature.npz, rotamer_codec.json, postprocess.npz/json,
files.json.  /root/reference does not exist on the GPU box, so tests only read these files.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

STANDARD_AMINO_ACIDS = {
    "A": "ALA", "C": "CYS", "D": "ASP", "E": "GLU", "F": "PHE", "G": "GLY", "H": "HIS", "I": "ILE",
    "K": "LYS", "L": "LEU", "M": "MET", "N": "ASN", "P": "PRO", "Q": "GLN", "R": "ARG", "S": "SER",
    "T": "THR", "V": "VAL", "W": "TRP", "Y": "TYR"}
N_CHI = {"ARG": 4, "ASN": 2, "ASP": 2, "CYS": 1, "GLN": 3, "GLU": 3, "HIS": 2, "ILE": 2, "LEU": 2,
         "LYS": 4, "MET": 3, "PHE": 2, "PRO": 2, "SER": 1, "THR": 1, "TRP": 2, "TYR": 2, "VAL": 1}
GUIDE = [0, 1, 4, 13, 40, 49, 50, 59, 68, 149, 158, 185, 194, 203, 230, 311, 314, 317, 320, 329]


def _stub_modules():
    ampal = types.ModuleType("ampal")
    aa = types.ModuleType("ampal.amino_acids")
    aa.standard_amino_acids = dict(STANDARD_AMINO_ACIDS)
    aa.side_chain_dihedrals = {k: [None] * v for k, v in N_CHI.items()}
    aa.polarity_Zimmerman = {}
    aa.residue_charge = {}
    ampal.amino_acids = aa
    ampal.Assembly = ampal.Polypeptide = ampal.Residue = object   # only used in annotations
    sys.modules["ampal"] = ampal
    sys.modules["ampal.amino_acids"] = aa
    sys.modules["h5py"] = types.ModuleType("h5py")
    apo = types.ModuleType("aposteriori")
    cfg = types.ModuleType("aposteriori.config")
    cfg.MAKE_FRAME_DATASET_VER = "2.0.0"
    cfg.UNCOMMON_RESIDUE_DICT = {}
    dp = types.ModuleType("aposteriori.data_prep")
    cfd = types.ModuleType("aposteriori.data_prep.create_frame_data_set")
    cfd.DatasetMetadata = object
    sys.modules.update({"aposteriori": apo, "aposteriori.config": cfg, "aposteriori.data_prep": dp,
                        "aposteriori.data_prep.create_frame_data_set": cfd})
    du = types.ModuleType("design_utils")
    du.__path__ = []
    au = types.ModuleType("design_utils.analyse_utils")
    # calculate_seq_metrics wraps four ampal functions (absent); the sampler's file writers only
    # need *a* 4-tuple, so use a deterministic composition-only stand-in.
    au.calculate_seq_metrics = lambda seq: (float(seq.count("K") + seq.count("R") - seq.count("D") - seq.count("E")),
                                            7.0, float(len(seq)) * 110.0, float(seq.count("W")) * 5500.0)
    sys.modules["design_utils"] = du
    sys.modules["design_utils.analyse_utils"] = au


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, REF / rel)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def fp16_softmax_rows(rng, n, c):
    z = rng.standard_normal((n, c)) * 2.0
    e = np.exp(z - z.max(1, keepdims=True))
    p = e / e.sum(1, keepdims=True)
    return p.astype(np.float16).astype(np.float64)   # what genfromtxt(.csv) yields in sample.py:32


def main():
    _stub_modules()
    ru = _load("ref_utils", "design_utils/utils.py")
    sys.modules["design_utils.utils"] = ru
    rs = _load("ref_sampling_utils", "design_utils/sampling_utils.py")

    # ------------------------------------------------------------------ rotamer codec (a10)
    codec, flat_categories, guide = ru.get_rotamer_codec(return_reduction_guide=True)
    assert guide == GUIDE, "ampal stub disagrees with the offsets pinned at utils.py:425"
    assert len(flat_categories) == 338
    with open(OUT / "rotamer_codec.json", "w") as f:
        json.dump({"flat_categories": flat_categories, "reduction_guide": guide,
                   "class_to_residue": [int(np.argmax(codec[i])) for i in range(338)]}, f)
    res_to_r = {v: k for k, v in STANDARD_AMINO_ACIDS.items()}
    rot_letters = [res_to_r[c.split("_")[0]] for c in flat_categories]

    # ------------------------------------------------------------------ sampler (a12)
    rng = np.random.default_rng(20240517)
    blob = {}
    test_row = np.array([[0.01] * 5 + [0.20] + [0.01] * 5 + [0.50, 0.10] + [0.01] * 6 + [0.04]])
    blob["test_row"] = test_row     # golden vector of tests/test_sampling_utils.py:5-28
    for c, cats in ((20, None), (338, rot_letters)):
        probs = fp16_softmax_rows(rng, 37, c)
        S = 16
        r_all = np.empty((S, 37))
        idx_all = np.empty((S, 37), dtype=np.int64)
        seq_all = []
        orig_rand = np.random.rand
        for s in range(S):
            np.random.seed(1000 * c + s)
            r = orig_rand(37)
            if s == 0:
                # rows whose fp16-rounded probabilities sum to < 1: r just below 1 exceeds every
                # cumsum entry -> (cumsum > r) is all False -> argmax returns 0 (the quirk)
                short = np.where(probs.cumsum(axis=1)[:, -1] < 1.0)[0]
                assert len(short) >= 2
                r[short] = 1.0 - 2 ** -53
                r[(short[0] + 1) % 37] = 0.0
            r_all[s] = r
            np.random.rand = lambda n, _r=r: _r.copy()
            try:
                idx_all[s] = rs.random_choice_prob_index(probs, return_seq=False)
                seq = rs.random_choice_prob_index(probs, return_seq=True, rotamer_categories=cats)
            finally:
                np.random.rand = orig_rand
            seq_all.append("".join(seq))
        blob[f"probs_{c}"] = probs
        blob[f"r_{c}"] = r_all
        blob[f"idx_{c}"] = idx_all
        blob[f"seq_{c}"] = np.array(seq_all)
        blob[f"cumsum_{c}"] = probs.cumsum(axis=1)
    np.savez_compressed(OUT / "sampler.npz", **blob)

    # ------------------------------------------------------------------ temperature (a11)
    tb = {}
    for c in (20, 338):
        probs = blob[f"probs_{c}"]
        tb[f"probs_{c}"] = probs
        for t in (0.01, 0.1, 0.5, 1, 2.0, 5.0, 100):
            tb[f"out_{c}_t{t}"] = rs.apply_temp_to_probs(probs, t=t)
    tb["test_row"] = test_row
    for t in (1, 0.01, 100):
        tb[f"out_test_row_t{t}"] = rs.apply_temp_to_probs(test_row, t=t)
    np.savez_compressed(OUT / "temperature.npz", **tb)

    # ------------------------------------------------------------------ post-processing (a5,a8,a9)
    pp = {}
    meta = {}
    labels = list(STANDARD_AMINO_ACIDS.values())
    # old-style 4-column map: two structures, one with two chains
    old_map = []
    for pdb, chain, n in (("1abc", "A", 7), ("1abc", "B", 5), ("2xyz", "A", 9)):
        for i in range(n):
            old_map.append((pdb, chain, str(i + 1), labels[(i * 3 + len(old_map)) % 20]))
    old_map = np.array(old_map)
    pm20 = fp16_softmax_rows(rng, len(old_map), 20).astype(np.float16)
    pm20[3, 4] = pm20[3, 9] = pm20[3].max() + np.float16(0.01)    # exact fp16 tie -> first index
    out = ru.extract_sequence_from_pred_matrix(old_map, pm20, rotamers_categories=None, old_datasetmap=True)
    pp["old_map"] = old_map
    pp["pm20"] = pm20.astype(np.float32)
    meta["old"] = {"seq": out[0], "real": out[2],
                   "prob_shapes": {k: list(np.array(v).shape) for k, v in out[1].items()}}
    # new-style 2-column map
    new_map = np.array([("1abcA", "7"), ("1abcB", "5"), ("2xyzA", "9")])
    out = ru.extract_sequence_from_pred_matrix(new_map, pm20, rotamers_categories=None)
    meta["new"] = {"seq": out[0], "real": out[2]}
    pp["new_map"] = new_map
    # rotamer categories
    pm338 = fp16_softmax_rows(rng, len(old_map), 338).astype(np.float16)
    out = ru.extract_sequence_from_pred_matrix(old_map, pm338, rotamers_categories=flat_categories)
    pp["pm338"] = pm338.astype(np.float32)
    meta["rot"] = {"seq": out[0]}
    pp["compress_in"] = rng.random((5, 338))
    pp["compress_out"] = ru.compress_rotamer_predictions_to_20(pp["compress_in"])
    # NMR consensus: states 1nmr_0 .. 1nmr_2, same chain length
    nmr_map = []
    for state in range(3):
        for i in range(6):
            nmr_map.append((f"1nmr_{state}", "A", str(i + 1), labels[i]))
    nmr_map = np.array(nmr_map)
    pm_nmr = fp16_softmax_rows(rng, len(nmr_map), 20).astype(np.float16)
    out = ru.extract_sequence_from_pred_matrix(nmr_map, pm_nmr, rotamers_categories=None, is_consensus=True)
    pp["nmr_map"] = nmr_map
    pp["pm_nmr"] = pm_nmr.astype(np.float32)
    meta["nmr"] = {"seq": out[0], "consensus": out[3]}
    for k, v in out[4].items():
        pp[f"nmr_consensus_prob_{k}"] = np.asarray(v, dtype=np.float64)
    np.savez_compressed(OUT / "postprocess.npz", **pp)
    with open(OUT / "postprocess.json", "w") as f:
        json.dump(meta, f, indent=1)

    # ------------------------------------------------------------------ file writers (App. A)
    files = {}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            y_true = np.eye(20)[[labels.index(l) for l in old_map[:, 3]]]
            pred32 = pm20.astype(np.float32) * np.float32(1.0009765625)   # not fp16-exact on purpose
            pp_pred = pred32
            # two batches, as predict.py:125-158 would append them
            ru.save_outputs_to_file(list(y_true[:12]), {0: list(pred32[:12])}, old_map, 0, "TIMED", tmp)
            ru.save_outputs_to_file(list(y_true[12:]), {0: list(pred32[12:])}, old_map, 0, "TIMED", tmp)
            ru.convert_dataset_map_for_srb(old_map, "TIMED", tmp)
            seqs = ru.extract_sequence_from_pred_matrix(
                old_map, np.genfromtxt(tmp / "TIMED.csv", delimiter=",", dtype=np.float16), None)
            ru.save_dict_to_fasta(seqs[0], "TIMED", tmp)
            ru.save_dict_to_fasta(seqs[2], "dataset", tmp)
            # sampler writers
            sampled = {"1abcA": [("ACDKW", 1.0, 7.0, 550.0, 5500.0), ("AAAAA", 0.0, 7.0, 550.0, 0.0)],
                       "2xyzA": [("KRDEW", 0.0, 7.0, 550.0, 5500.0)]}
            rs.save_as(sampled, "TIMED_temp_0.5_n_2_1abcA", "all")
            for p in sorted(tmp.iterdir()):
                files[p.name] = p.read_text()
        finally:
            os.chdir(cwd)
    files["__pred32__"] = pp_pred.tolist()
    with open(OUT / "files.json", "w") as f:
        json.dump(files, f)
    print("golden vectors written to", OUT)
    for p in sorted(OUT.iterdir()):
        print(f"  {p.name:24s} {p.stat().st_size:8d} B")


if __name__ == "__main__":
    main()
