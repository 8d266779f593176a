"""Known-answer vectors for the Keras 2.13 layer semantics the conv stack depends on, derived BY HAND from the Keras /
TensorFlow documentation (tf.keras.layers.Conv3D / MaxPooling3D / AveragePooling3D / BatchNormalization / ELU / Flatten,
tf.nn "SAME" padding: out = ceil(n / s), pad_total = max((out - 1) * s + k - n, 0), pad_before = pad_total // 2) --
NOT computed with oracle/keras_oracle.py.  Every expected number below is written out as a literal with its derivation,
so the oracle's two restatements (numpy fp64, torch fp32) and the CUDA path are each pinned to the documented
semantics where they are easiest to get wrong: cross-correlation (no kernel flip), the asymmetric 'same' padding of even
kernels, max-pooling that ignores padding, average pooling that divides by the VALID count, BatchNorm's default
epsilon 1e-3 inside the square root, DHWIO kernel layout and NDHWC flatten order.

TensorFlow itself cannot run offline, so this is a pin to the documented behaviour, not to TensorFlow's binary.
"""
import math

import numpy as np
import pytest

from oracle import keras_oracle as ko
from timed_design_b200 import standins


def _line(values):
    """(1, n, 1, 1, 1) tensor: n voxels along the depth axis, one channel."""
    return np.asarray(values, dtype=np.float32).reshape(1, -1, 1, 1, 1)


# ------------------------------------------------------------------------------------------------ the cases
# name -> (input NDHWC, kernel DHWIO, padding, expected NDHWC output)
CONV_CASES = {
    # Conv3D is cross-correlation: out[i] = sum_t x[i + t - pad_before] * k[t].  k = 3, 'same': pad_before = 1.
    # x = [1, 2, 3, 4], k = [1, 10, 100]:
    #   out[0] = 0*1 + 1*10 + 2*100 = 210; out[1] = 1 + 20 + 300 = 321; out[2] = 2 + 30 + 400 = 432; out[3] = 3 + 40 + 0 = 43
    # (a flipped kernel -- true convolution -- would give 12, 123, 234, 340)
    "k3_same_no_flip": (_line([1, 2, 3, 4]), np.array([1, 10, 100], np.float32).reshape(3, 1, 1, 1, 1), "same",
                        _line([210, 321, 432, 43])),
    # EVEN kernel, 'same': pad_total = k - 1 = 1, pad_before = 1 // 2 = 0, pad_after = 1 (TensorFlow pads at the END).
    # x = [1, 2, 3], k = [10, 1]: out[i] = 10 x[i] + x[i+1] -> [12, 23, 30]   (padding in front would give [1, 12, 23])
    "k2_same_pads_at_end": (_line([1, 2, 3]), np.array([10, 1], np.float32).reshape(2, 1, 1, 1, 1), "same", _line([12, 23, 30])),
    # k = 4, 'same': pad_total = 3, before = 1, after = 2.  x = [1, 2, 3, 4, 5], k = [1, 10, 100, 1000]:
    #   out[0] = 0 + 10 + 200 + 3000 = 3210; out[1] = 1 + 20 + 300 + 4000 = 4321; out[2] = 2 + 30 + 400 + 5000 = 5432;
    #   out[3] = 3 + 40 + 500 + 0 = 543;   out[4] = 4 + 50 + 0 + 0 = 54
    "k4_same_1_before_2_after": (_line([1, 2, 3, 4, 5]), np.array([1, 10, 100, 1000], np.float32).reshape(4, 1, 1, 1, 1), "same",
                                 _line([3210, 4321, 5432, 543, 54])),
    # 'valid': no padding, out = n - k + 1.  x = [1, 2, 3], k = [10, 1] -> [12, 23]
    "k2_valid": (_line([1, 2, 3]), np.array([10, 1], np.float32).reshape(2, 1, 1, 1, 1), "valid", _line([12, 23])),
    # DHWIO layout: kernel[0,0,0,ci,co].  One voxel, channels x = (2, 3); W[ci, co] = [[1, 10], [100, 1000]]:
    #   out[co=0] = 2*1 + 3*100 = 302; out[co=1] = 2*10 + 3*1000 = 3020     (an OIDHW reading would give 32 and 3200)
    "k1_channel_order": (np.array([2, 3], np.float32).reshape(1, 1, 1, 1, 2),
                         np.array([[1, 10], [100, 1000]], np.float32).reshape(1, 1, 1, 2, 2), "same",
                         np.array([302, 3020], np.float32).reshape(1, 1, 1, 1, 2)),
}
# The axes are ordered (depth, height, width): a kernel that differs along each axis on a 2x2x2 volume.
#   x[d,h,w] = 1 + 4d + 2h + w (values 1..8), kernel k[a,b,c] = 100a + 10b + c for a,b,c in {0,1}, 'valid' -> one output:
#   sum x*k = sum over (d,h,w) of (1 + 4d + 2h + w)(100d + 10h + w)
#     (0,0,0): 1*0 = 0; (0,0,1): 2*1 = 2; (0,1,0): 3*10 = 30; (0,1,1): 4*11 = 44;
#     (1,0,0): 5*100 = 500; (1,0,1): 6*101 = 606; (1,1,0): 7*110 = 770; (1,1,1): 8*111 = 888      total = 2840
_x222 = np.array([1 + 4 * d + 2 * h + w for d in range(2) for h in range(2) for w in range(2)], np.float32).reshape(1, 2, 2, 2, 1)
_k222 = np.array([100 * a + 10 * b + c for a in range(2) for b in range(2) for c in range(2)], np.float32).reshape(2, 2, 2, 1, 1)
CONV_CASES["k222_axis_order"] = (_x222, _k222, "valid", np.array([2840], np.float32).reshape(1, 1, 1, 1, 1))

# name -> (input, kind, padding, expected); pool 2, stride 2
POOL_CASES = {
    # 'same' on n = 5: out = ceil(5 / 2) = 3, pad_total = (3-1)*2 + 2 - 5 = 1, before = 0: windows [1,5] [2,4] [9]
    "max_same_odd": (_line([1, 5, 2, 4, 9]), "max", "same", _line([5, 4, 9])),
    # the padding must not take part in the maximum: all-negative input, last window holds only -9
    "max_same_ignores_padding": (_line([-1, -5, -2, -4, -9]), "max", "same", _line([-1, -2, -9])),
    # 'valid' on n = 5: out = floor((5 - 2) / 2) + 1 = 2, the last voxel is dropped
    "max_valid_odd": (_line([1, 5, 2, 4, 9]), "max", "valid", _line([5, 4])),
    # average 'same' divides by the number of VALID voxels: [2,4] -> 3, [6,8] -> 7, [10] -> 10 (not 5)
    "avg_same_valid_count": (_line([2, 4, 6, 8, 10]), "avg", "same", _line([3, 7, 10])),
    "avg_valid": (_line([2, 4, 6, 8, 10]), "avg", "valid", _line([3, 7])),
}

# BatchNormalization (inference): y = gamma * (x - mean) / sqrt(var + eps) + beta, eps = 1e-3 by default, INSIDE the root.
#   x = 2, gamma = 3, beta = 0.5, mean = 1, var = 0.25: 3 * 1 / sqrt(0.251) + 0.5 = 5.98802...+0.5
BN_EXPECTED = 3.0 / math.sqrt(0.251) + 0.5           # 0.501^2 = 0.251001 -> sqrt(0.251) = 0.500999, 3 / 0.500999 = 5.988036 -> 6.488036 (eps = 1e-5 would give 6.49988)
# ELU(alpha = 1): x > 0 -> x; x <= 0 -> e^x - 1.   ELU(-1) = 1/e - 1 = -0.6321205588285577
ELU_NEG1 = 1.0 / math.e - 1.0


def _conv_graph(x, k, padding):
    b = standins._Builder("kat_conv", x.shape[1:], 0)
    name = b.conv3d(b.input_name, k.shape[-1], 1, padding, use_bias=False)
    b.layers[-1]["config"]["kernel_size"] = list(k.shape[:3])
    b.weights[name] = {"kernel:0": k}
    return b.finish(name), name


@pytest.mark.parametrize("case", sorted(CONV_CASES))
def test_oracle_conv_matches_hand_derived(case):
    x, k, padding, want = CONV_CASES[case]
    (cfg, w), name = _conv_graph(x, k, padding)
    _, vals = ko.forward_numpy(cfg, w, x, np.float64, return_all=True)
    np.testing.assert_array_equal(vals[name], want.astype(np.float64))
    got = ko.np_conv3d(x.astype(np.float64), k.astype(np.float64), None, padding)
    np.testing.assert_array_equal(got, want.astype(np.float64))


@pytest.mark.parametrize("case", sorted(POOL_CASES))
def test_oracle_pool_matches_hand_derived(case):
    x, kind, padding, want = POOL_CASES[case]
    got = ko.np_pool3d(x.astype(np.float64), (2, 1, 1), (2, 1, 1), padding, kind)
    np.testing.assert_allclose(got, want.astype(np.float64), rtol=0, atol=1e-12)


def _bn_elu_graph():
    """input (1 voxel, 2 channels) -> ELU -> BatchNormalization (default epsilon) -> Flatten -> Dense identity."""
    b = standins._Builder("kat_bn", (1, 1, 1, 2), 0)
    x = b.elu(b.input_name)
    x = b.bn(x)
    b.weights[x] = {"gamma:0": np.array([3, 1], np.float32), "beta:0": np.array([0.5, 0], np.float32),
                    "moving_mean:0": np.array([1, 0], np.float32), "moving_variance:0": np.array([0.25, 0.999], np.float32)}
    bn = x
    x = b.flatten(x)
    x = b.dense(x, 2)
    b.weights[x] = {"kernel:0": np.eye(2, dtype=np.float32), "bias:0": np.zeros(2, np.float32)}
    return b.finish(x), bn


def test_oracle_batchnorm_epsilon_and_elu():
    (cfg, w), bn = _bn_elu_graph()
    x = np.array([2.0, -1.0], np.float32).reshape(1, 1, 1, 1, 2)
    for fwd in (lambda: ko.forward_numpy(cfg, w, x, np.float64), lambda: ko.forward_torch(cfg, w, x, dtype="float64")):
        y = np.asarray(fwd(), dtype=np.float64)[0]
        # channel 0: ELU(2) = 2 -> BN = 6.488...; channel 1: ELU(-1) = 1/e - 1, var + eps = 1.0 exactly -> unchanged
        assert abs(y[0] - BN_EXPECTED) < 1e-6 and abs(y[0] - 6.488036) < 1e-5
        assert abs(y[1] - ELU_NEG1) < 1e-6


def _flatten_graph():
    """Flatten is NDHWC row-major: (D=2, H=1, W=1, C=2) -> [d0c0, d0c1, d1c0, d1c1]; Dense kernel rows follow that order."""
    b = standins._Builder("kat_flat", (2, 1, 1, 2), 0)
    x = b.flatten(b.input_name)
    x = b.dense(x, 1)
    b.weights[x] = {"kernel:0": np.array([[1], [10], [100], [1000]], np.float32), "bias:0": np.array([0.5], np.float32)}
    return b.finish(x)


def test_oracle_flatten_order_and_dense():
    cfg, w = _flatten_graph()
    x = np.array([[1, 2], [3, 4]], np.float32).reshape(1, 2, 1, 1, 2)      # d0 = (1, 2), d1 = (3, 4)
    # 1*1 + 2*10 + 3*100 + 4*1000 + 0.5 = 4321.5      (a channels-first flatten would give 1 + 30 + 200 + 4000 = 4231.5)
    assert float(ko.forward_numpy(cfg, w, x, np.float64)[0, 0]) == 4321.5
    assert float(ko.forward_torch(cfg, w, x, dtype="float64")[0, 0]) == 4321.5


# ------------------------------------------------------------------------------------------------ the CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CONV_CASES))
def test_gpu_conv_matches_hand_derived(case):
    from tests.helpers import run_conv_gpu
    x, k, padding, want = CONV_CASES[case]
    got = run_conv_gpu(x, k, padding=padding)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=0)       # small integers: the bf16 hi/lo split is exact here


def _pool_graph(n, kind, padding):
    b = standins._Builder("kat_pool", (n, 1, 1, 1), 0)
    x = b.pool(b.input_name, kind, 2, padding)
    for key in ("pool_size", "strides"):
        b.layers[-1]["config"][key] = [2, 1, 1]
    d = b.shapes[x][0]
    b.shapes[x] = (d, 1, 1, 1)
    x = b.flatten(x)
    x = b.dense(x, d)
    b.weights[x] = {"kernel:0": np.eye(d, dtype=np.float32), "bias:0": np.zeros(d, np.float32)}
    return b.finish(x)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(POOL_CASES))
def test_gpu_pool_matches_hand_derived(case):
    from timed_design_b200.model import Model
    x, kind, padding, want = POOL_CASES[case]
    cfg, w = _pool_graph(x.shape[1], kind, padding)
    got = Model(cfg, w).predict(x)
    np.testing.assert_allclose(got[0], want.reshape(-1), rtol=2e-6, atol=0)


@pytest.mark.gpu
def test_gpu_batchnorm_elu_flatten():
    from timed_design_b200.model import Model
    (cfg, w), _ = _bn_elu_graph()
    y = Model(cfg, w).predict(np.array([2.0, -1.0], np.float32).reshape(1, 1, 1, 1, 2))[0]
    assert abs(y[0] - BN_EXPECTED) < 2e-5 and abs(y[1] - ELU_NEG1) < 2e-6
    cfg, w = _flatten_graph()
    y = Model(cfg, w).predict(np.array([[1, 2], [3, 4]], np.float32).reshape(1, 2, 1, 1, 2))
    assert abs(float(y[0, 0]) - 4321.5) < 5e-3
