"""CPU: the C-ABI library loads without a GPU/driver and exports every symbol the header declares;
host-side graph lowering and error behaviour."""
import re
from pathlib import Path

import numpy as np
import pytest

from timed_design_b200 import _lib, standins
from timed_design_b200.keras_graph import (OP_CONV3D, OP_GPOOL, OP_POOL3D, OP_SOFTMAX, UnsupportedLayerError,
                                           parse_model_config)

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "timed_b200.h").read_text()
    declared = set(re.findall(r"\b(timed_b200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.timed_b200_abi_version() == _lib.ABI_VERSION == 4


def test_no_cpu_fallback_without_device(lib):
    if lib.timed_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    from timed_design_b200.model import Model
    cfg, w = standins.tiny_standin(calib_frames=0)
    with pytest.raises(_lib.TimedB200Error):
        Model(cfg, w)
    from timed_design_b200 import sampling_utils
    with pytest.raises(_lib.TimedB200Error):
        sampling_utils.apply_temp_to_probs(np.full((2, 20), 0.05), 0.5)


def test_timed_standin_lowers_to_fused_ops():
    cfg, w = standins.timed_standin(calib_frames=0)
    g = parse_model_config(cfg, w)
    kinds = [op.kind for op in g.ops]
    assert kinds.count(OP_CONV3D) == 6 and kinds.count(OP_POOL3D) == 2
    assert kinds[-2:] == [OP_GPOOL, OP_SOFTMAX]
    conv = [op for op in g.ops if op.kind == OP_CONV3D]
    assert all(op.fused == ["Conv3D", "ELU", "BatchNormalization"] for op in conv)
    assert all(op.act1 == 2 and op.scale is not None and op.act2 == 0 for op in conv)
    # SURVEY.md App. E / BASELINE.md: algorithmic FLOPs per frame
    assert abs(g.flops_per_frame() - 2.3692e9) < 1e5
    cfg, w = standins.timed_standin(338, calib_frames=0)
    assert abs(parse_model_config(cfg, w).flops_per_frame() - 4.2683e9) < 1e5
    cfg, w = standins.densecpd_standin(calib_frames=0)
    assert abs(parse_model_config(cfg, w).flops_per_frame() - 17.0986e9) < 1e5


def test_bn_fold_matches_definition():
    cfg, w = standins.tiny_standin(calib_frames=0)
    g = parse_model_config(cfg, w)
    op = [o for o in g.ops if o.kind == OP_CONV3D][0]
    bn = w["batch_normalization"]
    x = np.linspace(-2, 2, 7)[:, None]
    ref = bn["gamma:0"] * (x - bn["moving_mean:0"]) / np.sqrt(bn["moving_variance:0"] + 1e-3) + bn["beta:0"]
    np.testing.assert_allclose(x * op.scale + op.shift, ref, rtol=1e-5, atol=1e-6)


def test_sequential_config_and_unsupported_layers():
    cfg, w = standins.tiny_standin(calib_frames=0)
    layers = [{"class_name": l["class_name"], "config": l["config"]} for l in cfg["config"]["layers"]]
    seq = {"class_name": "Sequential", "config": {"name": "seq", "layers": layers}}
    g = parse_model_config(seq, w)
    assert g.n_classes == 20 and len(g.ops) == len(parse_model_config(cfg, w).ops)
    bad = {"class_name": "Sequential", "config": {"name": "bad", "layers": layers[:2] + [
        {"class_name": "LSTM", "config": {"name": "lstm"}}]}}
    with pytest.raises(UnsupportedLayerError):
        parse_model_config(bad, w)
    with pytest.raises(KeyError):
        parse_model_config(cfg, {})            # missing weights fail loudly


def test_synthetic_frames_are_index_deterministic():
    a = standins.synthetic_frames(6, side=9)
    b = standins.synthetic_frames(3, side=9, first_index=3)
    np.testing.assert_array_equal(a[3:], b)


def test_device_post_host_helpers():
    """Host-side pieces of the device post-processing: grouping of NMR states as utils.py:696-705 walks the dict, and the
    packed metric table layout timed_b200_seq_metrics documents."""
    from timed_design_b200 import device_post, seq_metrics
    groups = device_post.consensus_groups(["1abc_0A", "1abc_1A", "2xyzA", "1abc_2A"], [5, 5, 3, 5])
    assert groups == [("1abc", 0, 2), ("2xyzA", 2, 1), ("1abc", 3, 1)]
    import pytest
    with pytest.raises(ValueError):
        device_post.consensus_groups(["1abc_0A", "1abc_1A"], [5, 6])
    t = seq_metrics.device_tables()
    n_grid = int(t[62])
    assert n_grid == 120 and len(t) == 63 + 2 * n_grid + 20 * n_grid
    assert abs(t[63] - 1.0) < 1e-12 and abs(t[63 + n_grid - 1] - 12.9) < 1e-9
    # table-driven charge at pH 7.4 == the restated formula
    counts = np.arange(20)[None, :]
    c, pi, mw, ext = seq_metrics.metrics_from_composition(counts)
    assert abs((counts[0] * t[40:60]).sum() + t[60] - c[0]) < 1e-10
    assert abs((counts[0] * t[:20]).sum() + t[61] - mw[0]) < 1e-9


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU port timed on the host cores) needs no GPU and prints one JSON line with the
    keys the driver reads."""
    import json
    import subprocess
    import sys
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_bench_reference_arm_of_the_sampler_config_and_rank_gating():
    """`--config sampler --impl reference` times the reference's numpy loop verbatim under a process pool; under a
    multi-rank launch only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys
    root = Path(__file__).resolve().parents[1]
    cmd = [sys.executable, str(root / "bench.py"), "--impl", "reference", "--config", "sampler", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "residues/s" and line["value"] > 1e5
    assert line["cpu_baseline"]["kind"] == "reference" and "sampled residues/sec" in line["metric"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_configs_name_the_baseline_workloads():
    import importlib.util
    root = Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("bench_mod", root / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert set(mod.CONFIGS) == {"timed20", "timed338", "densecpd", "sampler"}
    assert "20 classes" in mod.CONFIGS["timed20"]["metric"] and "338" in mod.CONFIGS["timed338"]["metric"]
    assert "DenseCPD" in mod.CONFIGS["densecpd"]["metric"] and [c["baseline_config"] for c in mod.CONFIGS.values()].count(1) == 1
    chains = mod.sampler_chains(20)
    assert len(chains) == 59 and all(60 <= len(c) <= 400 and c.shape[1] == 20 for c in chains)
    assert len(mod.SAMPLER_TEMPS) == 20 and mod.SAMPLER_TEMPS[0] == 0.1 and mod.SAMPLER_TEMPS[-1] == 2.0
