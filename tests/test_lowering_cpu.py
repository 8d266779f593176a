"""CPU: the Keras-graph lowering (timed_design_b200/keras_graph.py: conv + activation + BatchNorm absorbed into one fused
op, Dropout elided, Flatten+Dense turned into a 'valid' conv, DenseNet pre-activation kept as stand-alone affine ops) preserves
the network's function.  The fused op list -- exactly what the C ABI receives as tb_op_desc[] -- is evaluated by a small numpy
interpreter of the ABI's op semantics (include/timed_b200.h) and compared with the oracle's layer-by-layer evaluation of the
original Keras config."""
import numpy as np
import pytest

from oracle import keras_oracle as ko
from timed_design_b200 import standins
from timed_design_b200.keras_graph import (OP_ADD, OP_AFFINE, OP_CONCAT, OP_CONV3D, OP_GPOOL, OP_INPUT, OP_POOL3D, OP_SOFTMAX,
                                           parse_model_config)

ACT = {0: None, 1: "relu", 2: "elu", 3: "sigmoid", 4: "tanh"}


def _act(x, code, alpha):
    return x if not code else ko.np_activation(x, ACT[code], alpha)


def eval_fused(graph, X):
    """Reference semantics of the fused ops: CONV3D = act2(scale * act1(conv(x) + bias) + shift), etc."""
    vals = []
    for op in graph.ops:
        ins = [vals[i] for i in op.inputs]
        if op.kind == OP_INPUT:
            v = np.asarray(X, dtype=np.float64)
        elif op.kind == OP_CONV3D:
            v = ko.np_conv3d(ins[0], op.kernel_w.astype(np.float64),
                             None if op.bias is None else op.bias.astype(np.float64), "same" if op.pad_same else "valid")
            v = _act(v, op.act1, op.alpha1)
            if op.scale is not None or op.shift is not None:
                v = v * (1.0 if op.scale is None else op.scale.astype(np.float64)) + \
                    (0.0 if op.shift is None else op.shift.astype(np.float64))
            v = _act(v, op.act2, op.alpha2)
        elif op.kind == OP_POOL3D:
            v = ko.np_pool3d(ins[0], op.kernel, op.stride, "same" if op.pad_same else "valid", "avg" if op.pool_kind else "max")
        elif op.kind == OP_AFFINE:
            v = _act(ins[0], op.act1, op.alpha1)
            v = v * (1.0 if op.scale is None else op.scale.astype(np.float64)) + \
                (0.0 if op.shift is None else op.shift.astype(np.float64))
            v = _act(v, op.act2, op.alpha2)
        elif op.kind == OP_GPOOL:
            v = (ins[0].mean(axis=(1, 2, 3)) if op.pool_kind else ins[0].max(axis=(1, 2, 3)))[:, None, None, None, :]
        elif op.kind == OP_SOFTMAX:
            z = ins[0] - ins[0].max(axis=-1, keepdims=True)
            e = np.exp(z)
            v = e / e.sum(axis=-1, keepdims=True)
        elif op.kind == OP_CONCAT:
            v = np.concatenate(ins, axis=-1)
        elif op.kind == OP_ADD:
            v = ins[0] + ins[1]
        else:
            raise AssertionError(f"unknown op kind {op.kind}")
        assert tuple(v.shape[1:]) == tuple(op.out_shape), (op.name, v.shape, op.out_shape)
        vals.append(v)
    return vals[-1].reshape(len(X), -1)


@pytest.mark.parametrize("name,build,side", [
    ("timed", lambda: standins.tiny_standin(), 9),
    ("timed338", lambda: standins.tiny_standin(338, filters=(8, 8)), 9),
    ("densecpd", lambda: standins.densecpd_standin(side=8, n_layers=2, growth=8, bottleneck=16, stem=8, calib_frames=2), 8),
    ("prodconn", lambda: standins.prodconn_standin(side=9, branch=4, calib_frames=2), 9),
])
def test_lowered_graph_computes_what_the_keras_config_says(name, build, side):
    cfg, w = build()
    g = parse_model_config(cfg, w)
    X = standins.synthetic_frames(3, side=side, seed=5)
    ref = ko.forward_numpy(cfg, w, X)
    got = eval_fused(g, X)
    assert got.shape == ref.shape == (3, g.n_classes)
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-6)       # folded BN is float32: a few 1e-7 on probabilities
    np.testing.assert_allclose(got.sum(1), 1.0, atol=1e-6)
    # the lowering fuses: fewer ops than Keras layers, and no stand-alone activation / dropout / flatten survives
    assert len(g.ops) < len(cfg["config"]["layers"])
    assert abs(g.flops_per_frame() - standins.conv_flops_per_frame(cfg)) < 1e-6
