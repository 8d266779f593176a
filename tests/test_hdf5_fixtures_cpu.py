"""The HDF5 reader against byte fixtures assembled from the format specification by a script that shares no code with
the package (tests/golden/make_hdf5_fixtures.py): version-2 superblock / object headers / compact-link groups /
version-3 attributes / filter pipeline v2 with fletcher32 / two-level chunk B-tree / global-heap strings (spec_latest.h5)
and the version-0 superblock + symbol-table layout h5py writes by default (spec_earliest.h5)."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from timed_design_b200 import frames
from timed_design_b200.hdf5 import File, Hdf5FormatError, read_keras_h5

G = Path(__file__).resolve().parent / "golden"
EXP = json.loads((G / "spec_expected.json").read_text())


def test_generator_is_independent_and_reproducible(tmp_path):
    src = (G / "make_hdf5_fixtures.py").read_text()
    code = src.split('"""', 2)[2]                                                   # everything after the module docstring
    assert not [ln for ln in code.splitlines() if ln.lstrip().startswith(("import ", "from ")) and
                ("timed_design_b200" in ln or "h5py" in ln or "hdf5" in ln)]
    work = tmp_path / "golden"
    work.mkdir()
    (work / "make_hdf5_fixtures.py").write_text(src)
    subprocess.run([sys.executable, str(work / "make_hdf5_fixtures.py")], check=True, capture_output=True)
    for name in ("spec_latest.h5", "spec_earliest.h5"):
        assert (work / name).read_bytes() == (G / name).read_bytes(), name


def test_latest_layout_keras_model():
    """superblock v2, OHDR + OCHK continuation, link messages, attribute v3, vlen strings in GCOL (41 KB model_config),
    chunked + shuffle + deflate + fletcher32 with a level-1 chunk B-tree, contiguous and compact datasets."""
    e = EXP["latest"]
    cfg, w = read_keras_h5(G / "spec_latest.h5")
    assert cfg["config"]["name"] == "spec_fixture" and len(cfg["config"]["layers"]) == e["model_config_layers"]
    assert list(w) == ["conv3d", "batch_normalization"]
    np.testing.assert_array_equal(w["conv3d"]["conv3d/kernel:0"], np.array(e["kernel"], np.float32))
    np.testing.assert_array_equal(w["conv3d"]["conv3d/bias:0"], np.array(e["bias"], np.float32))
    np.testing.assert_array_equal(w["batch_normalization"]["batch_normalization/gamma:0"], np.array(e["gamma"], np.float32))
    f = File(G / "spec_latest.h5")
    assert len(f.attrs["model_config"]) == e["model_config_len"] > 32768
    assert f.attrs["keras_version"] == "2.13.1"
    assert list(f.attrs["atom_encoder"]) == ["C", "N", "Cα"]                     # UTF-8 vlen array
    assert f.attrs["voxels_as_gaussian"] is np.True_ or bool(f.attrs["voxels_as_gaussian"]) is True
    assert float(f.attrs["frame_edge_length"]) == 21.0
    np.testing.assert_array_equal(f.attrs["frame_dims"], [5, 5, 5, 6])
    assert f["model_weights"].attrs["backend"] == "tensorflow"


def test_latest_layout_frames_and_native_inflater():
    e = EXP["latest"]
    f = File(G / "spec_latest.h5")
    d7, d12 = f["1abc/A/7"], f["1abc/A/12"]
    assert d7.dtype == np.bool_ and d7.attrs["label"] == "GLY" and int(np.argmax(d7.attrs["encoded_residue"])) == 5
    np.testing.assert_array_equal(d7[()], np.array(e["frame7"], bool))
    np.testing.assert_array_equal(d12[()], np.array(e["frame12"]))
    # the native chunk inflater (timed_b200_inflate_chunks) reads the same chunks through chunk_table()
    X = np.zeros((1, 5, 5, 5, 6), np.float32)
    y = np.zeros((1, 20))
    assert frames._native_load(f, [("1abc", "A", "12")], (5, 5, 5, 6), X, y)
    np.testing.assert_array_equal(X[0], np.array(e["frame12"]).astype(np.float32))
    assert int(np.argmax(y[0])) == 18
    kernel = f["model_weights/conv3d/conv3d/kernel:0"]
    assert kernel.chunk_table() is None                                          # fletcher32: left to the Python reader


def test_earliest_layout_frame_dataset():
    """superblock v0, symbol-table groups whose B-tree spans three SNODs, v1 object headers with continuation,
    version-1 attributes and filter pipeline; residue ids come back in INTEGER order (utils.py:367-371)."""
    e = EXP["earliest"]
    path = G / "spec_earliest.h5"
    f = File(path)
    assert list(f["2xyz/A"].keys()) == sorted(e["res_ids"])                      # name (B-tree) order, as h5py iterates
    flat, pdbs = frames.create_flat_dataset_map(path)
    assert pdbs == {"2xyz"}
    assert [r[2] for r in flat] == sorted(e["res_ids"], key=int)
    assert {r[3] for r in flat} == {"ALA", "SER"}
    X, y = frames.load_batch(path, flat)
    assert X.dtype == np.float32 and X.shape == (10, 4, 4, 4, 2)
    for i, r in enumerate(flat):
        np.testing.assert_array_equal(X[i], np.array(e["frames"][r[2]]).astype(np.float32))
        assert int(np.argmax(y[i])) == (0 if r[3] == "ALA" else 15)


def test_truncated_file_raises_format_error(tmp_path):
    """A chunk whose stored bytes lie past the end of the file must raise, not fault inside the native inflater."""
    raw = (G / "spec_earliest.h5").read_bytes()
    p = tmp_path / "cut.h5"
    p.write_bytes(raw[:len(raw) - 3000])
    with pytest.raises(Hdf5FormatError):
        flat, _ = frames.create_flat_dataset_map(p)
        frames.load_batch(p, flat)
