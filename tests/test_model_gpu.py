"""GPU parity of the whole inference graph against the CPU oracle (numpy fp64 and torch fp32).

Tolerance from BASELINE.json north_star: per-residue softmax probabilities within 1e-4 of the
fp32 CPU path; fp16-rounded argmax identical outside the reported near-tie set."""
import numpy as np
import pytest

from oracle import keras_oracle as ko
from timed_design_b200 import standins

PROB_TOL = 1e-4


def _check(model_fn, frames_fn, n, tol=PROB_TOL, use_numpy64=True):
    from timed_design_b200.model import Model
    cfg, w = model_fn()
    X = frames_fn(n)
    m = Model(cfg, w)
    p = m.predict(X)
    assert p.shape == (n, m.n_classes) and p.dtype == np.float32
    ref32 = ko.forward_torch(cfg, w, X)
    d32 = float(np.abs(p - ref32).max())
    assert np.isfinite(p).all()
    assert d32 <= tol, f"max|dp| vs torch-fp32 oracle = {d32:.3e}"
    if use_numpy64:
        ref64 = ko.forward_numpy(cfg, w, X)
        assert float(np.abs(p - ref64).max()) <= tol
    safe = ~ko.near_tie_rows(ref32)
    assert (ko.fp16_argmax(p)[safe] == ko.fp16_argmax(ref32)[safe]).all()
    np.testing.assert_allclose(p.sum(1), 1.0, atol=1e-5)
    return p, ref32


@pytest.mark.gpu
def test_tiny_timed_parity():
    _check(lambda: standins.tiny_standin(), lambda n: standins.synthetic_frames(n, side=9), 37)


@pytest.mark.gpu
def test_tiny_timed_parity_bool_and_f64_inputs():
    from timed_design_b200.model import Model
    cfg, w = standins.tiny_standin()
    m = Model(cfg, w)
    X = standins.synthetic_frames(9, side=9)
    Xb = X > 0.2
    for arr in (X.astype(np.float64), Xb, Xb.astype(np.uint8)):
        p = m.predict(arr)
        ref = ko.forward_torch(cfg, w, arr)
        assert np.abs(p - ref).max() <= PROB_TOL


@pytest.mark.gpu
def test_timed_standin_parity_20():
    """TIMED-20 at its benchmark size (21^3 x 6) on 256 frames -- enough rows for the CTA-pair tiles the bench runs."""
    p, ref = _check(lambda: standins.timed_standin(20), lambda n: standins.synthetic_frames(n, seed=77), 256,
                    use_numpy64=False, tol=5e-5)
    assert len(set(ko.fp16_argmax(ref))) > 1      # the stand-in is not a constant predictor


@pytest.mark.gpu
def test_timed_standin_parity_338():
    _check(lambda: standins.timed_standin(338, seed=8), lambda n: standins.synthetic_frames(n, seed=77), 256,
           use_numpy64=False, tol=5e-5)


@pytest.mark.gpu
def test_prodconn_standin_parity():
    _check(lambda: standins.prodconn_standin(side=13, calib_frames=3),
           lambda n: standins.synthetic_frames(n, side=13), 5)


@pytest.mark.gpu
def test_prodconn_full_size_parity():
    """ProDCoNN stand-in at 21^3 (k = 3/5/7 branches, Concatenate, valid conv + pools, Flatten, two Dense)."""
    _check(lambda: standins.prodconn_standin(), lambda n: standins.synthetic_frames(n, seed=31), 8, use_numpy64=False)


@pytest.mark.gpu
def test_densecpd_small_parity():
    _check(lambda: standins.densecpd_standin(side=12, n_layers=2, calib_frames=3),
           lambda n: standins.synthetic_frames(n, side=12), 5)


@pytest.mark.gpu
def test_densecpd_full_size_parity():
    """The FULL DenseCPD stand-in of BASELINE config 4: 21^3, three dense blocks of six layers, 17.1 GFLOP/frame."""
    _check(lambda: standins.densecpd_standin(), lambda n: standins.synthetic_frames(n, seed=31), 8, use_numpy64=False)


@pytest.mark.gpu
@pytest.mark.parametrize("gain,tol", [(16.0, 5e-5), (32.0, 1e-4)])
def test_logit_gain_stress(gain, tol):
    """Probability errors of any finite-precision evaluation scale with the logit gain of the last BatchNorm (a trained,
    sharper network = a larger gain).  At 2x and 4x the stand-in's gain the contract must still hold -- with the
    corrections of the bf16 split in the main accumulator (round 1's default) 2x measured 1.35e-4."""
    _check(lambda: standins.timed_standin(20, logit_gain=gain), lambda n: standins.synthetic_frames(n, seed=77), 256,
           use_numpy64=False, tol=tol)


@pytest.mark.gpu
def test_default_accumulation_is_tighter_than_fast_and_consistent():
    """Default: separate correction accumulator (epilogue drains TMEM into registers before the math).  The A/B switch
    Model(precise=False) accumulates the corrections into the main accumulator: same answers within tolerance,
    measurably further from the oracle on a batch large enough to engage the CTA-pair tiles."""
    from timed_design_b200.model import Model
    cfg, w = standins.timed_standin(20)
    uniq = standins.synthetic_frames(48, seed=5)
    X = np.tile(uniq, (8, 1, 1, 1, 1))                      # 384 frames: enough rows for pair tiles
    ref = ko.forward_torch(cfg, w, uniq)
    m = Model(cfg, w)
    assert any("pair" in m.op_kernel(i, 384) for i in range(len(m.graph.ops)))
    prec = m.predict(X, batch_size=4096)
    fast = Model(cfg, w, precise=False).predict(X, batch_size=4096)
    e_fast = np.abs(fast[:48] - ref).max()
    e_prec = np.abs(prec[:48] - ref).max()
    assert e_fast <= PROB_TOL and e_prec <= 5e-5
    assert e_prec < e_fast
    np.testing.assert_array_equal(prec[:48], prec[48:96])    # batch-position independent
    np.testing.assert_array_equal(prec[:48], m.predict(uniq))  # ... and tile-configuration independent (pair vs single CTA)
    safe = ~ko.near_tie_rows(ref)
    assert (ko.fp16_argmax(prec[:48])[safe] == ko.fp16_argmax(ref)[safe]).all()


def _pool_graph(side, c1, pool_pad, relu, seed=21):
    """input(6) -> Conv3D(c1, k3, same)[+ReLU | ELU -> BN] -> MaxPool(2, pool_pad) -> GAP -> Dense softmax: the pooled tensor
    is plain fp32 (read by the global pool), so the fused conv+pool epilogue's NDHWC output path is exercised."""
    b = standins._Builder(f"pool_{side}_{c1}_{pool_pad}", (side, side, side, 6), seed, standins.synthetic_frames(3, side, 6, seed=99))
    if relu:
        x = b.conv3d(b.input_name, c1, 3, "same", activation="relu")
    else:
        x = b.bn(b.elu(b.conv3d(b.input_name, c1, 3, "same")))
    x = b.pool(x, "max", 2, pool_pad)
    x = b.gap(x)
    x = b.dense(x, 20, activation="softmax")
    return b.finish(x)


@pytest.mark.gpu
@pytest.mark.parametrize("side,c1,pool_pad,relu", [(21, 32, "same", False), (21, 24, "valid", True), (10, 16, "valid", False),
                                                   (13, 32, "same", True), (4, 8, "same", False)])
def test_fused_conv_maxpool_epilogue(side, c1, pool_pad, relu, monkeypatch):
    """MaxPool(2,2,2) fused into thinz_conv_kernel's epilogue (z pairs from TMEM, in-plane 2x2 through a staged tile):
    odd and even volumes, TF 'same' partial windows and 'valid' truncation, against the oracle and the unfused path."""
    from timed_design_b200.model import Model
    X = standins.synthetic_frames(7, side=side, seed=4)
    cfg, w = _pool_graph(side, c1, pool_pad, relu)
    ref = ko.forward_torch(cfg, w, X)
    m = Model(cfg, w)                                       # default: the whole max-pool runs in the conv epilogue
    fused = m.predict(X)
    assert any("thinz_conv_kernel(+maxpool)" in m.op_kernel(i, 7) for i in range(len(m.graph.ops)))
    monkeypatch.setenv("TIMED_B200_NO_POOLFUSE", "1")
    m2 = Model(cfg, w)                                      # only the z direction is pooled in the epilogue
    zfused = m2.predict(X)
    assert m2.launches_per_forward == m.launches_per_forward + 1
    monkeypatch.setenv("TIMED_B200_NO_ZPOOL", "1")
    unfused = Model(cfg, w).predict(X)
    for got in (fused, zfused, unfused):
        assert np.abs(got - ref).max() <= PROB_TOL
    # (the fused instantiation accumulates the bf16-split corrections in the main TMEM columns, the others in their own:
    # the same products summed in a different order)
    assert np.abs(fused - unfused).max() <= 5e-6 and np.abs(zfused - unfused).max() <= 2e-6


@pytest.mark.gpu
def test_fused_pool_into_cpv_matches_unfused(monkeypatch):
    """TIMED stand-in with the first max-pool fused into the first conv's epilogue, pooled tensor written straight in the
    chunk-plane padded-volume layout the slab conv reads."""
    from timed_design_b200.model import Model
    cfg, w = standins.timed_standin(20)
    X = standins.synthetic_frames(5, seed=12)
    m = Model(cfg, w)
    fused = m.predict(X)
    assert "maxpool" in m.op_kernel(1, 5) and "slab" in m.op_kernel(3, 5)
    monkeypatch.setenv("TIMED_B200_NO_POOLFUSE", "1")
    base = Model(cfg, w).predict(X)
    assert np.abs(fused - base).max() <= 2e-6
    assert np.abs(fused - ko.forward_torch(cfg, w, X)).max() <= PROB_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("classes", [20, 338])
def test_fused_head_matches_unfused_bit_for_bit(classes, monkeypatch):
    """Network head in one launch: col2im gather + bias/ELU/BN + GlobalAveragePooling + Softmax after the head GEMM
    (20 classes, tap-to-N), or pooling + softmax (338 classes, direct conv).  Same summation orders as the separate
    col2im / gpool / softmax kernels => identical bits, two (20) / one (338) launches fewer per forward."""
    from timed_design_b200.model import Model
    cfg, w = standins.timed_standin(classes, seed=7 if classes == 20 else 8)
    X = standins.synthetic_frames(9, seed=41)
    m = Model(cfg, w)
    fused = m.predict(X)
    names = [m.op_kernel(i, 9) for i in range(len(m.graph.ops))]
    assert any("head_" in n for n in names), names
    monkeypatch.setenv("TIMED_B200_NO_HEADFUSE", "1")
    m2 = Model(cfg, w)
    unfused = m2.predict(X)
    assert m2.launches_per_forward == m.launches_per_forward + (2 if classes == 20 else 1)
    np.testing.assert_array_equal(fused, unfused)
    assert np.abs(fused - ko.forward_torch(cfg, w, X)).max() <= PROB_TOL


@pytest.mark.gpu
def test_voxel_stationary_tiles_do_not_change_the_probabilities(monkeypatch):
    """256 frames per pass is where the wide convs switch to voxel-stationary tiles (skip the taps in the zero padding):
    same bits as the im2col tiling, whatever the pass size."""
    from timed_design_b200.model import Model
    cfg, w = standins.timed_standin(20)
    X = standins.synthetic_frames(600, seed=12)
    m = Model(cfg, w)
    a = m.predict(X, batch_size=4096)                     # passes of 512 + 88 frames: the first with voxel-stationary tiles
    monkeypatch.setenv("TIMED_B200_NO_VOX", "1")
    b = m.predict(X, batch_size=4096)
    monkeypatch.delenv("TIMED_B200_NO_VOX")
    c = m.predict(X, batch_size=100)                      # six passes, none reaches 256 frames
    m.close()
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a, c)


@pytest.mark.gpu
@pytest.mark.parametrize("side,k,cin,classes", [(6, 3, 64, 20), (5, 2, 48, 338), (4, 3, 20, 7)])
def test_linear_head_conv_gap_collapse(side, k, cin, classes, monkeypatch):
    """A head conv WITHOUT activation followed by GlobalAveragePooling is evaluated as box sums of its input + one dense GEMM
    (the average commutes with the conv, ConvPlan::gap_collapse): same probabilities as convolving every voxel
    (TIMED_B200_NO_GAPFOLD) and as the oracle.  TIMED's own head has ELU + BatchNorm before the pooling and keeps the
    per-voxel path."""
    from timed_design_b200.model import Model
    b = standins._Builder(f"linear_head_{side}_{k}_{cin}_{classes}", (side, side, side, 6), 21,
                          standins.synthetic_frames(4, side, 6, seed=99))
    x = b.conv3d(b.input_name, cin, 3, "same")
    x = b.bn(b.elu(x))
    x = b.conv3d(x, classes, k, "same")                    # linear head
    x = b.softmax(b.gap(x))
    cfg, w = b.finish(x)
    X = standins.synthetic_frames(70, side=side, seed=5)
    m = Model(cfg, w)
    p = m.predict(X)
    kernels = [m.op_kernel(i, 70) for i in range(len(m.graph.ops))]
    monkeypatch.setenv("TIMED_B200_NO_GAPFOLD", "1")
    m2 = Model(cfg, w)
    p2 = m2.predict(X)
    m.close(); m2.close()
    ref = ko.forward_torch(cfg, w, X)
    assert np.abs(p - ref).max() <= PROB_TOL
    assert np.abs(p - p2).max() <= 2e-5
    assert any("gap_boxsum_kernel" in kk for kk in kernels), kernels


@pytest.mark.gpu
@pytest.mark.parametrize("side,n_layers,frames", [(12, 2, 9), (21, 6, 3)])
def test_densenet_preactivation_in_the_conv_operand_path_is_bit_identical(side, n_layers, frames, monkeypatch):
    """DenseNet bottlenecks (BatchNorm -> ReLU -> 1x1x1 conv): the affine + ReLU + bf16 split of the concatenated fp32 tensor
    run inside the conv's operand path (bnrelu_conv1x1_kernel: TMA-staged fp32 tile, transform warps, UMMA), the BatchNorm
    launch and its split-plane copy of the tensor disappear.  Same arithmetic per element and the same MMA sequence as
    affine_act_kernel + conv_umma_kernel => identical probabilities, one launch fewer per dense layer; and the growth convs'
    col2im in the GEMM epilogue on top (the shipped DenseCPD path) against the Z-matrix route."""
    from timed_design_b200.model import Model
    cfg, w = standins.densecpd_standin(side=side, n_layers=n_layers, calib_frames=3)
    X = standins.synthetic_frames(frames, side=side, seed=5)
    m = Model(cfg, w)
    fused = m.predict(X)
    names = [m.op_kernel(i, frames) for i in range(len(m.graph.ops))]
    n_fused = sum("bnrelu_conv1x1_kernel" in n for n in names)
    assert n_fused >= 3 * n_layers - 1, names
    monkeypatch.setenv("TIMED_B200_NO_XFORM", "1")
    m2 = Model(cfg, w)
    unfused = m2.predict(X)
    assert m2.launches_per_forward == m.launches_per_forward + n_fused
    monkeypatch.setenv("TIMED_B200_NO_C2I_FUSE", "1")
    plain = Model(cfg, w).predict(X)
    np.testing.assert_array_equal(fused, unfused)
    np.testing.assert_array_equal(fused, plain)
    assert np.abs(fused - ko.forward_torch(cfg, w, X)).max() <= PROB_TOL
