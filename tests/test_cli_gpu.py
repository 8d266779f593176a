"""GPU end-to-end of the drop-in CLIs (BASELINE.json configs[0]: predict.py on the 1ubq test
structure, then sample.py on its outputs).  aposteriori is not installable, so the 76 frames are
synthetic blobs; residue ids, labels and the true sequence come from the reference's own fixture
tests/testing_files/1ubq.pdb1.gz (committed as tests/golden/1ubq_chainA.json)."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import keras_oracle as ko
from timed_design_b200 import standins
from timed_design_b200.hdf5 import write_frame_dataset, write_keras_h5

pytestmark = pytest.mark.gpu
G = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    ubq = json.loads((G / "1ubq_chainA.json").read_text())
    frames = standins.synthetic_frames(76, seed=0)
    residues = {str(rid): (frames[i], ubq["labels"][i]) for i, rid in enumerate(ubq["residue_ids"])}
    write_frame_dataset(d / "data.hdf5", {"1ubq": {"A": residues}}, (21, 21, 21, 6))
    cfg, w = standins.timed_standin(20, filters=(8, 16, 16, 24, 32), calib_frames=4)
    write_keras_h5(d / "TIMED.h5", cfg, w)
    cfg_r, w_r = standins.timed_standin(338, filters=(8, 16, 16, 24, 32), seed=9, calib_frames=4)
    write_keras_h5(d / "TIMED_rotamer.h5", cfg_r, w_r)
    return d, ubq, frames, (cfg, w), (cfg_r, w_r)


def test_predict_cli_writes_reference_file_set(workdir, monkeypatch):
    d, ubq, frames, (cfg, w), _ = workdir
    from timed_design_b200 import predict
    out = d / "out"
    monkeypatch.chdir(d)
    predict.cli(["--path_to_dataset", str(d / "data.hdf5"), "--path_to_model", str(d / "TIMED.h5"),
                 "--path_to_output", str(out), "--path_to_datasetmap", str(out / "datasetmap.txt"), "--yes"])
    for name in ("TIMED.csv", "encoded_labels.csv", "datasetmap.txt", "TIMED.txt", "TIMED.fasta", "dataset.fasta"):
        assert (out / name).exists(), name
    assert (out / "dataset.fasta").read_text() == f">1ubqA\n{ubq['sequence']}\n"
    assert (out / "TIMED.txt").read_text() == "ignore_uncommon False\ninclude_pdbs\n##########\n1ubqA 76\n"
    dm = (out / "datasetmap.txt").read_text().splitlines()
    assert len(dm) == 76 and dm[0] == "1ubq,A,1,MET" and dm[-1] == "1ubq,A,76,GLY"
    labels = np.genfromtxt(out / "encoded_labels.csv", delimiter=",")
    assert labels.shape == (76, 20) and (labels.sum(1) == 1).all()
    # probabilities: the CSV holds the float16 cast printed with '%.18e'
    text = (out / "TIMED.csv").read_text().splitlines()
    assert len(text) == 76 and len(text[0].split(",")[0]) == len("1.234741210937500000e-01")
    got = np.genfromtxt(out / "TIMED.csv", delimiter=",")
    ref = ko.forward_torch(cfg, w, frames)
    assert np.array_equal(got.astype(np.float16).astype(np.float64), got)          # exactly fp16-representable
    assert np.abs(got - ref).max() <= 1e-4 + 2 ** -11
    # sequence: fp16 argmax, identical to the oracle outside near-ties
    seq = (out / "TIMED.fasta").read_text().splitlines()
    assert seq[0] == ">1ubqA" and len(seq[1]) == 76
    letters = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
    safe = ~ko.near_tie_rows(ref)
    assert (np.array(list(seq[1]))[safe] == letters[ko.fp16_argmax(ref)][safe]).all()
    assert seq[1] == "".join(letters[np.argmax(got.astype(np.float16), axis=1)])


def test_sample_cli_on_predict_outputs(workdir, monkeypatch):
    d, ubq, *_ = workdir
    from timed_design_b200 import sample
    out = d / "out"
    monkeypatch.chdir(d)
    paths = sample.cli(["--path_to_pred_matrix", str(out / "TIMED.csv"), "--path_to_datasetmap", str(out / "TIMED.txt"),
                        "--sample_n", "50", "--temperature", "0.5", "--seed", "7"])
    assert paths == ["TIMED_temp_0.5_n_50_1ubqA.json", "TIMED_temp_0.5_n_50_1ubqA.fasta",
                     "TIMED_temp_0.5_n_50_1ubqA_metrics.csv"]
    samples = json.loads((d / paths[0]).read_text())["1ubqA"]
    assert len(samples) == 50 and all(len(s[0]) == 76 and len(s) == 5 for s in samples)
    first = (d / paths[1]).read_text()
    sample.cli(["--path_to_pred_matrix", str(out / "TIMED.csv"), "--path_to_datasetmap", str(out / "TIMED.txt"),
                "--sample_n", "50", "--temperature", "0.5", "--seed", "7", "--workers", "3"])
    assert (d / paths[1]).read_text() == first                        # seeded: reproducible
    sample.cli(["--path_to_pred_matrix", str(out / "TIMED.csv"), "--path_to_datasetmap", str(out / "TIMED.txt"),
                "--sample_n", "50", "--temperature", "0.5", "--seed", "8"])
    assert (d / paths[1]).read_text() != first
    # a cold temperature collapses onto the argmax sequence (tests/test_sampling_utils.py:52-58)
    sample.cli(["--path_to_pred_matrix", str(out / "TIMED.csv"), "--path_to_datasetmap", str(out / "TIMED.txt"),
                "--sample_n", "5", "--temperature", "0.01", "--save_as", "fasta"])
    cold = (d / "TIMED_temp_0.01_n_5_1ubqA.fasta").read_text().splitlines()
    argmax_seq = (out / "TIMED.fasta").read_text().splitlines()[1]
    assert sum(a != b for a, b in zip(cold[1], argmax_seq)) <= 3


def test_predict_cli_rotamer_mode(workdir, monkeypatch):
    d, ubq, frames, _, (cfg_r, w_r) = workdir
    from timed_design_b200 import predict
    out = d / "out_rot"
    monkeypatch.chdir(d)
    predict.cli(["--path_to_dataset", str(d / "data.hdf5"), "--path_to_model", str(d / "TIMED_rotamer.h5"),
                 "--path_to_output", str(out), "--path_to_datasetmap", str(out / "datasetmap.txt"),
                 "--predict_rotamers", "True", "--batch_size", "20", "--yes"])
    raw = np.genfromtxt(out / "TIMED_rotamer_rot.csv", delimiter=",")
    assert raw.shape == (76, 338)
    ref = ko.forward_torch(cfg_r, w_r, frames)
    assert np.abs(raw - ref).max() <= 1e-4
    onehot = np.genfromtxt(out / "TIMED_rotamer.csv", delimiter=",")
    assert onehot.shape == (76, 20) and (onehot.sum(1) == 1).all()
    from timed_design_b200.postprocess import rotamer_class_to_residue
    np.testing.assert_array_equal(onehot.argmax(1), rotamer_class_to_residue()[raw.argmax(1)])
    assert len((out / "TIMED_rotamer.fasta").read_text().splitlines()[1]) == 76


def test_binary_outputs_and_npy_sampling(workdir, monkeypatch):
    """predict.py --binary_outputs writes {model}.npy = the float16 matrix of {model}.csv; sample.py on the .npy draws
    exactly what it draws from the .csv."""
    d, *_ = workdir
    from timed_design_b200 import predict, sample
    out = d / "out_bin"
    monkeypatch.chdir(d)
    predict.cli(["--path_to_dataset", str(d / "data.hdf5"), "--path_to_model", str(d / "TIMED.h5"),
                 "--path_to_output", str(out), "--path_to_datasetmap", str(out / "datasetmap.txt"), "--yes",
                 "--binary_outputs", "--batch_size", "32"])
    m = np.load(out / "TIMED.npy")
    assert m.dtype == np.float16 and m.shape == (76, 20)
    np.testing.assert_array_equal(m.astype(np.float64), np.genfromtxt(out / "TIMED.csv", delimiter=","))
    a = sample.cli(["--path_to_pred_matrix", str(out / "TIMED.csv"), "--path_to_datasetmap", str(out / "TIMED.txt"),
                    "--sample_n", "20", "--temperature", "0.7", "--save_as", "fasta"])
    from_csv = (d / a[0]).read_text()
    b = sample.cli(["--path_to_pred_matrix", str(out / "TIMED.npy"), "--path_to_datasetmap", str(out / "TIMED.txt"),
                    "--sample_n", "20", "--temperature", "0.7", "--save_as", "fasta"])
    assert a == b and (d / b[0]).read_text() == from_csv


def test_predict_cli_nmr_consensus(tmp_path, monkeypatch):
    """--is_structure_nmr: three states of one structure -> consensus files; the device consensus equals the numpy
    restatement of utils.py:694-713 on the float16 matrix."""
    from timed_design_b200 import predict, postprocess
    ubq = json.loads((G / "1ubq_chainA.json").read_text())
    n = 24
    states = {}
    for s in range(3):
        frames = standins.synthetic_frames(n, seed=10 + s)
        states[f"1nmr_{s}"] = {"A": {str(rid): (frames[i], ubq["labels"][i]) for i, rid in enumerate(ubq["residue_ids"][:n])}}
    write_frame_dataset(tmp_path / "nmr.hdf5", states, (21, 21, 21, 6))
    cfg, w = standins.timed_standin(20, filters=(8, 16, 16, 24, 32), calib_frames=4)
    write_keras_h5(tmp_path / "TIMED.h5", cfg, w)
    out = tmp_path / "out"
    monkeypatch.chdir(tmp_path)
    predict.cli(["--path_to_dataset", str(tmp_path / "nmr.hdf5"), "--path_to_model", str(tmp_path / "TIMED.h5"),
                 "--path_to_output", str(out), "--path_to_datasetmap", str(out / "datasetmap.txt"), "--yes",
                 "--is_structure_nmr", "True", "--batch_size", "40"])
    pm = np.genfromtxt(out / "TIMED.csv", delimiter=",", dtype=np.float16)
    dmap = np.genfromtxt(out / "datasetmap.txt", delimiter=",", dtype=str)
    ref = postprocess.extract_sequence_from_pred_matrix(dmap, pm, None, is_consensus=True)
    # the reference's writers default to `Path.cwd()` evaluated at import time (utils.py:595-613): that is where the
    # consensus .fasta lands; the consensus .csv is a bare relative filename (utils.py:587) -> the CWD at call time
    import inspect
    fasta_dir = Path(inspect.signature(postprocess.save_dict_to_fasta).parameters["path_to_output"].default)
    fasta_path = fasta_dir / "TIMED_consensus.fasta"
    fasta = fasta_path.read_text().splitlines()
    fasta_path.unlink()
    assert len(ref[3]) == 1
    (name, seq), = ref[3].items()
    assert fasta == [f">{name}", seq] and len(seq) == n
    cons_csv = np.genfromtxt(tmp_path / "TIMED_consensus.csv", delimiter=",")
    np.testing.assert_array_equal(cons_csv, np.asarray(ref[4][name], dtype=np.float64))
