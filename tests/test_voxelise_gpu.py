"""Voxeliser on the GPU (timed_b200_voxelise) against the numpy oracle, and BASELINE config 1 end to end from the REAL
structure file: 1ubq.pdb1.gz -> frames on the device -> frame dataset (.hdf5) -> predict CLI -> the reference's file set."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import keras_oracle as ko
from oracle import voxelise_oracle as vo
from timed_design_b200 import standins
from timed_design_b200 import voxelise as vx

pytestmark = pytest.mark.gpu
G = Path(__file__).parent / "golden"
PDB = G / "1ubq.pdb1.gz"


def _oracle(tab, idx, gaussian, encode_cb=True):
    cb = (*vx.IDEAL_CB, vx.VDW["C"] / 2.3548)
    prop_ch = len(tab.channels) - 1 if tab.prop is not None else -1
    return vo.voxelise(tab.atoms, tab.channel, tab.residue, tab.is_cb, tab.frames, tab.prop, list(idx), 21, 1.0,
                       len(tab.channels), gaussian, encode_cb, cb, tab.channels.index("CB"), prop_ch)


@pytest.mark.parametrize("codec,gaussian", [("CNOCBCA", True), ("CNOCBCA", False), ("CNOCBCAQ", True), ("CNOCACBP", True)])
def test_kernel_matches_oracle(codec, gaussian):
    (residues,) = vx.parse_pdb(PDB)
    tab = vx.build_tables(residues, codec, 1.0)
    idx = np.array([0, 7, 26, 40, 75, 33, 12], dtype=np.int64)             # any order, both termini
    got = vx.voxelise_tables(tab, idx, voxels_as_gaussian=gaussian, chunk=3)   # three launches
    want = _oracle(tab, idx, gaussian)
    assert got.shape == want.shape == (7, 21, 21, 21, len(tab.channels))
    if gaussian:
        # fixed-point accumulation: at most the last 2^-24 step differs (device exp vs numpy exp, <= 1 ulp apart)
        assert np.abs(got - want).max() <= 2.0 ** -23
        assert (got != want).mean() < 1e-3
    else:
        np.testing.assert_array_equal(got.astype(np.uint8), want)


def test_float16_frames_and_determinism():
    (residues,) = vx.parse_pdb(PDB)
    tab = vx.build_tables(residues, "CNOCBCA", 1.0)
    a = vx.voxelise_tables(tab, tab.valid)
    b = vx.voxelise_tables(tab, tab.valid)
    np.testing.assert_array_equal(a, b)                                   # integer atomics: no run-to-run noise
    h = vx.voxelise_tables(tab, tab.valid, dtype=np.float16)
    np.testing.assert_array_equal(h, a.astype(np.float16))


def test_config1_from_the_real_structure(tmp_path, monkeypatch):
    """BASELINE.json configs[0]: predict.py TIMED on tests/testing_files (1 PDB), starting from the PDB file itself."""
    from timed_design_b200 import predict
    from timed_design_b200.hdf5 import write_keras_h5
    from timed_design_b200.model import Model
    gold = json.loads((G / "1ubq_chainA.json").read_text())
    with pytest.warns(RuntimeWarning, match="UNVERIFIED against aposteriori"):
        data = vx.make_frame_dataset([PDB], tmp_path, "data", codec="CNOCBCAQ", voxels_as_gaussian=True)
    frames, flat = vx.voxelise_structure(PDB, "CNOCBCAQ")
    assert frames.shape == (76, 21, 21, 21, 6)
    assert [f[3] for f in flat] == gold["labels"] and [int(f[2]) for f in flat] == gold["residue_ids"]
    cfg, w = standins.timed_standin(20, filters=(8, 16, 16, 24, 32), calib_frames=4)
    write_keras_h5(tmp_path / "TIMED.h5", cfg, w)
    out = tmp_path / "out"
    monkeypatch.chdir(tmp_path)
    predict.cli(["--path_to_dataset", str(data), "--path_to_model", str(tmp_path / "TIMED.h5"), "--path_to_output", str(out),
                 "--path_to_datasetmap", str(out / "datasetmap.txt"), "--yes"])
    assert (out / "dataset.fasta").read_text() == f">1ubqA\n{gold['sequence']}\n"
    got = np.genfromtxt(out / "TIMED.csv", delimiter=",")
    assert got.shape == (76, 20)
    ref = ko.forward_torch(cfg, w, frames)                                 # the dataset round trip is lossless (float32)
    assert np.abs(got - ref).max() <= 1e-4 + 2 ** -11
    # frames handed over on the device (no HDF5, no host copy) give the same probabilities
    import torch
    dfr, _ = vx.voxelise_structure(PDB, "CNOCBCAQ", return_device=True)
    m = Model(cfg, w)
    probs = torch.empty((76, 20), dtype=torch.float32, device="cuda")
    ws = torch.empty(m.workspace_bytes(76), dtype=torch.uint8, device="cuda")
    m.forward_device(dfr, probs, ws, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(probs.cpu().numpy(), m.predict(frames))


def test_predict_cli_straight_from_structure_files(tmp_path, monkeypatch):
    """Extension: --path_to_dataset may be a structure file (or a directory of them).  Frames are voxelised on the GPU and
    fed to the network without touching the host; the outputs equal those of the .hdf5 route byte for byte."""
    import shutil
    from timed_design_b200 import predict
    from timed_design_b200.hdf5 import write_keras_h5
    cfg, w = standins.timed_standin(20, c_in=5, filters=(8, 16, 16, 24, 32), calib_frames=0)
    write_keras_h5(tmp_path / "TIMED.h5", cfg, w)
    pdbs = tmp_path / "pdbs"
    pdbs.mkdir()
    shutil.copy(PDB, pdbs / "1ubq.pdb1.gz")
    monkeypatch.chdir(tmp_path)
    a, b = tmp_path / "from_pdb", tmp_path / "from_hdf5"
    predict.cli(["--path_to_dataset", str(pdbs), "--path_to_model", str(tmp_path / "TIMED.h5"), "--path_to_output", str(a),
                 "--path_to_datasetmap", str(a / "datasetmap.txt"), "--yes", "--batch_size", "500"])
    with pytest.warns(RuntimeWarning):
        data = vx.make_frame_dataset([PDB], tmp_path, "data", codec="CNOCBCA")
    predict.cli(["--path_to_dataset", str(data), "--path_to_model", str(tmp_path / "TIMED.h5"), "--path_to_output", str(b),
                 "--path_to_datasetmap", str(b / "datasetmap.txt"), "--yes", "--batch_size", "500"])
    names = sorted(f.name for f in b.iterdir())
    assert names == sorted(f.name for f in a.iterdir()) and "TIMED.fasta" in names
    for name in names:
        assert (a / name).read_bytes() == (b / name).read_bytes(), name
