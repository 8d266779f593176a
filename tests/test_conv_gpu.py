"""GPU parity of single Conv3D launches through the C ABI (timed_b200_conv3d_fwd) against the numpy
fp64 oracle: every operand layout (64/32/16-channel K blocks, W-folded and padded-volume thin inputs),
'same' and 'valid' padding, kernel sizes 1-7, one and two M sub-tiles, one and two N tiles, the
tap-to-N head path and the fused bias/ELU/ReLU/BatchNorm epilogues.

Tolerance: 1e-4 of the output range (the bf16 hi/lo split carries ~16 mantissa bits per operand and the
tensor core's accumulator adds ~2^-26 per accumulation); typical measured 5e-6..5e-5."""
import numpy as np
import pytest

from oracle import keras_oracle as ko
from tests.helpers import run_conv_gpu

pytestmark = pytest.mark.gpu

# (name, n, side, c_in, c_out, k, padding)
CASES = [
    ("k1_c64_n32", 2, 4, 64, 32, 1, "same"),
    ("k1_c16_n32", 2, 4, 16, 32, 1, "same"),
    ("k1_c32_n32", 2, 4, 32, 32, 1, "same"),
    ("k1_c128_n64", 2, 4, 128, 64, 1, "same"),
    ("k1_c6_n20", 3, 5, 6, 20, 1, "same"),
    ("k3_c64_n128", 2, 6, 64, 128, 3, "same"),
    ("k3_c6_n32_21", 1, 21, 6, 32, 3, "same"),
    ("k3_c5_n32_valid", 2, 9, 5, 32, 3, "valid"),
    ("k3_c32_n64_11", 2, 11, 32, 64, 3, "same"),
    ("k3_c256_n512", 4, 6, 256, 512, 3, "same"),
    ("k3_c512_n20_tap2n", 4, 6, 512, 20, 3, "same"),
    ("k3_c128_n16_tap2n_valid", 3, 7, 128, 16, 3, "valid"),
    ("k5_c64_n8_tap2n", 2, 6, 64, 8, 5, "same"),
    ("k3_c256_n20_tap2n_many", 300, 6, 256, 20, 3, "same"),
    ("k3_c512_n338", 2, 6, 512, 338, 3, "same"),
    ("k3_valid_c16", 2, 8, 16, 48, 3, "valid"),
    ("k5_c6_n16", 1, 9, 6, 16, 5, "same"),
    ("k7_c6_n16", 1, 9, 6, 16, 7, "same"),
    ("k4_valid_dense", 5, 4, 64, 128, 4, "valid"),
    ("k2_even_kernel", 2, 6, 32, 48, 2, "same"),
    ("k3_c128_n256_mt2", 700, 6, 128, 256, 3, "same"),
    ("k3_c64_n128_mt2", 700, 6, 64, 128, 3, "same"),
    ("k3_c256_n512_pair", 360, 6, 256, 512, 3, "same"),
    ("k3_c128_n338_pair", 400, 6, 128, 338, 3, "same"),
    ("k1_c96_n540_pair", 300, 6, 96, 540, 1, "same"),
    ("k3_c6_n32_many_tiles", 40, 13, 6, 32, 3, "same"),
]


def _torch_ref(x, w, b, k, padding):
    import torch
    import torch.nn.functional as F
    xt = torch.from_numpy(x).permute(0, 4, 1, 2, 3).double()
    wt = torch.from_numpy(w).permute(4, 3, 0, 1, 2).double()
    if padding == "same":
        p0, p1 = (k - 1) // 2, k - 1 - (k - 1) // 2
        xt = F.pad(xt, (p0, p1, p0, p1, p0, p1))
    return F.conv3d(xt, wt, torch.from_numpy(b).double()).permute(0, 2, 3, 4, 1).numpy()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv3d_matches_oracle(case):
    name, n, side, ci, co, k, padding = case
    rng = np.random.default_rng(abs(hash(name)) % 2 ** 31)
    x = rng.standard_normal((n, side, side, side, ci)).astype(np.float32)
    w = (rng.standard_normal((k, k, k, ci, co)) * np.sqrt(2.0 / (k ** 3 * ci))).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    y = run_conv_gpu(x, w, bias=b, padding=padding)
    assert np.isfinite(y).all()
    nref = min(n, 3)
    ref = ko.np_conv3d(x[:nref].astype(np.float64), w.astype(np.float64), b.astype(np.float64), padding)
    assert y.shape[1:] == ref.shape[1:]
    scale = np.abs(ref).max()
    assert np.abs(y[:nref] - ref).max() <= 1e-4 * scale
    if n > nref:                       # large batches: middle and tail frames against an independent conv
        for sl in (slice(n // 2, n // 2 + 4), slice(n - 3, n)):
            r2 = _torch_ref(x[sl], w, b, k, padding)
            assert np.abs(y[sl] - r2).max() <= 1e-4 * np.abs(r2).max()


def test_onehot_kernel_selects_the_right_taps():
    """Kernel with a single 1 per output channel: the conv must return shifted copies of the input, which
    pins the tap <-> coordinate mapping and the zero padding exactly."""
    rng = np.random.default_rng(5)
    for ci, side in ((64, 6), (6, 7)):
        x = rng.standard_normal((2, side, side, side, ci)).astype(np.float32)
        co = 32
        w = np.zeros((3, 3, 3, ci, co), np.float32)
        for o in range(co):
            t = o % 27
            w[t // 9, (t // 3) % 3, t % 3, (o * 7) % ci, o] = 1.0
        y = run_conv_gpu(x, w, padding="same")
        ref = ko.np_conv3d(x.astype(np.float64), w.astype(np.float64), None, "same")
        assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("act1,act2,affine", [("elu", None, True), ("relu", None, False), (None, "relu", True),
                                             (None, None, True), ("tanh", None, True), ("elu", "relu", True)])
def test_fused_epilogues(act1, act2, affine):
    rng = np.random.default_rng(11)
    for ci in (6, 32):
        x = rng.standard_normal((2, 7, 7, 7, ci)).astype(np.float32)
        w = (rng.standard_normal((3, 3, 3, ci, 24)) * np.sqrt(2.0 / (27 * ci))).astype(np.float32)
        b = (rng.standard_normal(24) * 0.2).astype(np.float32)
        sc = rng.uniform(0.5, 1.5, 24).astype(np.float32) if affine else None
        sh = (rng.standard_normal(24) * 0.3).astype(np.float32) if affine else None
        y = run_conv_gpu(x, w, bias=b, scale=sc, shift=sh, padding="same", act1=act1, act2=act2)
        ref = ko.np_conv3d(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), "same")
        ref = ko.np_activation(ref, act1) if act1 else ref
        if affine:
            ref = ref * sc + sh
        ref = ko.np_activation(ref, act2) if act2 else ref
        assert np.abs(y - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1.0)


@pytest.mark.parametrize("ci,co", [(128, 256), (256, 512)])
def test_cluster_mode_conv(ci, co, monkeypatch):
    """The default path (separate TMEM accumulator for the correction MMAs of the bf16 split, CTA-pair kernel with the
    register-drained epilogue): parity, and a smaller error than the A/B path that accumulates the corrections into
    the main accumulator (TIMED_B200_FAST_ACCUM) on the same data."""
    rng = np.random.default_rng(ci)
    x = rng.standard_normal((700, 6, 6, 6, ci)).astype(np.float32)
    w = (rng.standard_normal((3, 3, 3, ci, co)) * np.sqrt(2.0 / (27 * ci))).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    ref_head = ko.np_conv3d(x[:2].astype(np.float64), w.astype(np.float64), b.astype(np.float64), "same")
    ref_tail = _torch_ref(x[-3:], w, b, 3, "same")
    y = run_conv_gpu(x, w, bias=b)
    monkeypatch.setenv("TIMED_B200_FAST_ACCUM", "1")
    y_fast = run_conv_gpu(x, w, bias=b)
    for got, ref in ((y[:2], ref_head), (y[-3:], ref_tail), (y_fast[:2], ref_head)):
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()
    rms = lambda a, r: np.sqrt(((a - r) ** 2).mean())
    assert rms(y[:2], ref_head) < 0.8 * rms(y_fast[:2], ref_head)


@pytest.mark.parametrize("case", [("small_c256_n512", 4, 6, 256, 512, 3, "same"), ("small_c64_n160_k5", 3, 7, 64, 160, 5, "same"),
                                  ("small_c32_n192_valid", 2, 9, 32, 192, 3, "valid"), ("one_frame_c64_n256", 1, 4, 64, 256, 3, "same")],
                         ids=lambda c: c[0])
def test_pair_mode_small_shapes(case, monkeypatch):
    """tcgen05.mma.cta_group::2 path (conv_pair.cuh) forced on shapes with fewer tiles than CTAs: ragged last
    pair-tile (dummy peer half), 64B- and 128B-swizzled K blocks, odd N tiles, and agreement with the single-CTA path."""
    name, n, side, ci, co, k, padding = case
    rng = np.random.default_rng(abs(hash(name)) % 2 ** 31)
    x = rng.standard_normal((n, side, side, side, ci)).astype(np.float32)
    w = (rng.standard_normal((k, k, k, ci, co)) * np.sqrt(2.0 / (k ** 3 * ci))).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    y_single = run_conv_gpu(x, w, bias=b, padding=padding)
    monkeypatch.setenv("TIMED_B200_FORCE_PAIR", "1")
    for kc in ("64", "32"):
        monkeypatch.setenv("TIMED_B200_PAIR_KC", kc)
        y = run_conv_gpu(x, w, bias=b, padding=padding)
        ref = ko.np_conv3d(x[:2].astype(np.float64), w.astype(np.float64), b.astype(np.float64), padding)
        assert np.isfinite(y).all()
        assert np.abs(y[:2] - ref).max() <= 1e-4 * np.abs(ref).max()
        assert np.abs(y - y_single).max() <= 1e-4 * np.abs(ref).max()


VOX_CASES = [
    # name, frames, side, c_in, c_out, k, padding          kernel the planner picks
    ("vox_pair_c256_n512", 300, 6, 256, 512, 3, "same"),   # conv_pair: 2 blocks of 256 frames, the second 17 % full
    ("vox_pair_k5_c128_n256", 130, 5, 128, 256, 5, "same"),  # 5^3 filter on 5^3 voxels: 8 .. 125 valid taps per voxel
    ("vox_umma_c64_n128", 260, 6, 64, 128, 3, "same"),     # conv_umma, N-folded, two sub-tiles of 128 frames
    ("vox_umma_even_k2", 140, 4, 64, 96, 2, "same"),       # even filter: padding on the far side only
]


@pytest.mark.parametrize("case", VOX_CASES, ids=lambda c: c[0])
def test_voxel_stationary_tiles_are_bit_identical(case, monkeypatch):
    """Voxel-stationary tiles (ConvKernelParams::vox: the rows of a tile are frames at ONE output voxel, and filter taps
    that fall into the zero padding are skipped) must give the same BITS as the im2col tiling -- a skipped tap would have
    added exact zeros and the others keep their order -- and agree with the fp64 oracle."""
    name, n, side, ci, co, k, padding = case
    rng = np.random.default_rng(abs(hash(name)) % 2 ** 31)
    x = rng.standard_normal((n, side, side, side, ci)).astype(np.float32)
    w = (rng.standard_normal((k, k, k, ci, co)) * np.sqrt(2.0 / (k ** 3 * ci))).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    monkeypatch.setenv("TIMED_B200_NO_VOX", "1")
    y_im2col = run_conv_gpu(x, w, bias=b, padding=padding, act1="elu")
    monkeypatch.delenv("TIMED_B200_NO_VOX")
    monkeypatch.setenv("TIMED_B200_FORCE_VOX", "1")
    y_vox = run_conv_gpu(x, w, bias=b, padding=padding, act1="elu")
    np.testing.assert_array_equal(y_vox, y_im2col)
    ref = ko.np_activation(ko.np_conv3d(x[-2:].astype(np.float64), w.astype(np.float64), b.astype(np.float64), padding), "elu")
    assert np.abs(y_vox[-2:] - ref).max() <= 1e-4 * np.abs(ref).max()


SLAB_CASES = [
    ("slab_c32_n64_11_many", 40, 11, 32, 64, 3, "same"),
    ("slab_c64_n128_13", 3, 13, 64, 128, 3, "same"),
    ("slab_c16_n24_k5", 2, 12, 16, 24, 5, "same"),
    ("slab_c48_n96_k2_even", 2, 10, 48, 96, 2, "same"),
    ("slab_c20_n40_valid", 3, 10, 20, 40, 3, "valid"),
    ("slab_c32_n64_k1", 2, 9, 32, 64, 1, "same"),
    ("slab_c32_n16_small_forced", 2, 5, 32, 16, 3, "same"),
]


@pytest.mark.parametrize("case", SLAB_CASES, ids=[c[0] for c in SLAB_CASES])
def test_slab_conv_matches_oracle_and_im2col_path(case, monkeypatch):
    """slab_conv_kernel (chunk-plane padded-volume input, taps as descriptor shifts on one staged slab): parity with
    the fp64 oracle and agreement with the im2col kernel on the same data; odd/even kernels, 'valid' windows, channel
    counts that are not a multiple of 16, tiles that straddle frames."""
    name, n, side, ci, co, k, padding = case
    rng = np.random.default_rng(abs(hash(name)) % 2 ** 31)
    x = rng.standard_normal((n, side, side, side, ci)).astype(np.float32)
    w = (rng.standard_normal((k, k, k, ci, co)) * np.sqrt(2.0 / (k ** 3 * ci))).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, co).astype(np.float32)
    sh = (rng.standard_normal(co) * 0.3).astype(np.float32)
    monkeypatch.setenv("TIMED_B200_NO_SLAB", "1")
    y_ref_path = run_conv_gpu(x, w, bias=b, scale=sc, shift=sh, padding=padding, act1="elu")
    monkeypatch.delenv("TIMED_B200_NO_SLAB")
    monkeypatch.setenv("TIMED_B200_FORCE_SLAB", "1")
    y = run_conv_gpu(x, w, bias=b, scale=sc, shift=sh, padding=padding, act1="elu")
    assert np.isfinite(y).all()
    nref = min(n, 2)
    ref = ko.np_conv3d(x[:nref].astype(np.float64), w.astype(np.float64), b.astype(np.float64), padding)
    ref = ko.np_activation(ref, "elu") * sc + sh
    scale = max(np.abs(ref).max(), 1.0)
    assert np.abs(y[:nref] - ref).max() <= 1e-4 * scale
    assert np.abs(y - y_ref_path).max() <= 1e-4 * scale


@pytest.mark.parametrize("case", [("zfold_c6_n32_21", 3, 21, 6, 32, 3, "same"), ("zfold_c8_n16_k5", 2, 11, 8, 16, 5, "same"),
                                  ("zfold_c3_n20_valid", 2, 9, 3, 20, 3, "valid"), ("zfold_c6_n32_ragged_z", 2, 6, 6, 32, 3, "same"),
                                  ("zfold_c4_n48_k2", 2, 8, 4, 48, 2, "same")], ids=lambda c: c[0])
def test_zfold_thin_conv(case, monkeypatch):
    """thinz_conv_kernel (kd filter slices folded into the MMA N dimension, zt output planes per tile): parity with
    the oracle and with thin_conv_kernel; ragged last z group, 'valid' windows, even kernels, kd = 5."""
    name, n, side, ci, co, k, padding = case
    rng = np.random.default_rng(abs(hash(name)) % 2 ** 31)
    x = rng.standard_normal((n, side, side, side, ci)).astype(np.float32)
    w = (rng.standard_normal((k, k, k, ci, co)) * np.sqrt(2.0 / (k ** 3 * ci))).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    monkeypatch.setenv("TIMED_B200_NO_ZFOLD", "1")
    y_thin = run_conv_gpu(x, w, bias=b, padding=padding, act1="elu")
    monkeypatch.delenv("TIMED_B200_NO_ZFOLD")
    y = run_conv_gpu(x, w, bias=b, padding=padding, act1="elu")
    ref = ko.np_activation(ko.np_conv3d(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), padding), "elu")
    scale = max(np.abs(ref).max(), 1.0)
    assert np.isfinite(y).all()
    assert np.abs(y - ref).max() <= 1e-4 * scale
    assert np.abs(y - y_thin).max() <= 1e-4 * scale


def test_tap_to_n_variants_agree(monkeypatch):
    """Head conv as 1x1 GEMM + col2im: the variant that keeps the kw taps in K (Z matrix kw times smaller) against the
    all-taps-in-N variant and the oracle."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((5, 6, 6, 6, 512)).astype(np.float32)
    w = (rng.standard_normal((3, 3, 3, 512, 20)) * np.sqrt(2.0 / (27 * 512))).astype(np.float32)
    b = (rng.standard_normal(20) * 0.1).astype(np.float32)
    y_kw = run_conv_gpu(x, w, bias=b, act1="elu")
    monkeypatch.setenv("TIMED_B200_TAP2N_FULL", "1")
    y_full = run_conv_gpu(x, w, bias=b, act1="elu")
    monkeypatch.setenv("TIMED_B200_NO_TAP2N", "1")
    y_direct = run_conv_gpu(x, w, bias=b, act1="elu")
    ref = ko.np_activation(ko.np_conv3d(x[:2].astype(np.float64), w.astype(np.float64), b.astype(np.float64), "same"), "elu")
    for y in (y_kw, y_full, y_direct):
        assert np.abs(y[:2] - ref).max() <= 1e-4 * np.abs(ref).max()
    assert np.abs(y_kw - y_full).max() <= 5e-5 * np.abs(ref).max()


@pytest.mark.parametrize("shape", [(3, 9, 128, 32, 3, "same"), (2, 7, 64, 24, 3, "valid"), (2, 6, 512, 20, 3, "same"),
                                   (2, 8, 128, 32, 5, "same")], ids=lambda s: f"s{s[1]}_c{s[2]}_n{s[3]}_k{s[4]}_{s[5]}")
def test_tap_to_n_kw_in_n_variant(shape, monkeypatch):
    """Third tap-to-N formulation: the kw taps in the GEMM's N dimension, the (kd,kh) taps in K (im2col over D and H only),
    col2im summing the kw shifted copies along W -- what DenseCPD's 128 -> 32 growth convs run on.  Against the oracle and
    the direct conv, 'same' (W margins dropped by the gather) and 'valid' (narrower output) alike."""
    n, side, ci, co, k, padding = shape
    rng = np.random.default_rng(side * ci)
    x = rng.standard_normal((n, side, side, side, ci)).astype(np.float32)
    w = (rng.standard_normal((k, k, k, ci, co)) * np.sqrt(2.0 / (k ** 3 * ci))).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, co).astype(np.float32)
    sh = (rng.standard_normal(co) * 0.2).astype(np.float32)
    monkeypatch.setenv("TIMED_B200_TAP2N_W", "1")
    monkeypatch.setenv("TIMED_B200_TAP2N_MARGIN", "100")         # take the variant whatever the cost model says
    y_w = run_conv_gpu(x, w, bias=b, scale=sc, shift=sh, padding=padding, act1="relu")
    monkeypatch.delenv("TIMED_B200_TAP2N_W")
    monkeypatch.setenv("TIMED_B200_NO_TAP2N", "1")
    y_direct = run_conv_gpu(x, w, bias=b, scale=sc, shift=sh, padding=padding, act1="relu")
    ref = ko.np_activation(ko.np_conv3d(x[:1].astype(np.float64), w.astype(np.float64), b.astype(np.float64), padding), "relu")
    ref = ref * sc + sh
    assert y_w.shape == y_direct.shape
    assert np.abs(y_w[:1] - ref).max() <= 1e-4 * np.abs(ref).max()
    assert np.abs(y_w - y_direct).max() <= 5e-5 * np.abs(ref).max()


@pytest.mark.parametrize("shape", [(3, 21, 128, 32, "linear"), (3, 9, 128, 32, "relu"), (5, 6, 64, 16, "elu"), (2, 7, 128, 32, "linear"),
                                   (2, 11, 64, 32, "relu"), (1, 21, 160, 32, "linear")],
                         ids=lambda s: f"n{s[0]}_s{s[1]}_c{s[2]}_n{s[3]}_{s[4]}")
def test_col2im_over_kw_in_the_epilogue_is_bit_identical(shape, monkeypatch):
    """kw-in-N tap-to-N convs with kw = 3, 'same', C_out = 16 / 32 (DenseCPD's growth convs) sum the three kw-shifted
    column groups inside the GEMM epilogue on tiles of whole volume rows (ConvKernelParams::c2i) instead of writing the Z
    matrix and launching col2im_kernel: same operations in the same order, so the two routes must agree to the last bit --
    at tile seams (side 21: 6 rows per tile, tiles straddle planes and frames), lane-quadrant seams, the last partial tile,
    and with the layer's bias / activation / BatchNorm affine applied by the GEMM epilogue.  And both against the oracle."""
    n, side, ci, co, act = shape
    rng = np.random.default_rng(side * ci + co)
    x = rng.standard_normal((n, side, side, side, ci)).astype(np.float32)
    w = (rng.standard_normal((3, 3, 3, ci, co)) * np.sqrt(2.0 / (27 * ci))).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, co).astype(np.float32)
    sh = (rng.standard_normal(co) * 0.2).astype(np.float32)
    monkeypatch.setenv("TIMED_B200_TAP2N_W", "1")
    monkeypatch.setenv("TIMED_B200_TAP2N_MARGIN", "100")         # take the variant whatever the cost model says
    y_fused = run_conv_gpu(x, w, bias=b, scale=sc, shift=sh, padding="same", act1=act)
    monkeypatch.setenv("TIMED_B200_NO_C2I_FUSE", "1")
    y_z = run_conv_gpu(x, w, bias=b, scale=sc, shift=sh, padding="same", act1=act)
    assert y_fused.shape == y_z.shape == (n, side, side, side, co)
    if act == "elu":        # the GEMM epilogue's branch-free ELU (SFU exp / series) vs col2im_kernel's expm1f: ~1e-7 apart
        assert np.abs(y_fused - y_z).max() <= 1e-6
    else:
        assert np.array_equal(y_fused, y_z)
    ref = ko.np_conv3d(x[:1].astype(np.float64), w.astype(np.float64), b.astype(np.float64), "same")
    if act != "linear":
        ref = ko.np_activation(ref, act)
    ref = ref * sc + sh
    assert np.abs(y_fused[:1] - ref).max() <= 1e-4 * np.abs(ref).max()
