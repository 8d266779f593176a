"""Keras 2.x ``model_config`` -> fused inference ops for libtimed_b200.

The reference never defines its networks: ``predict.py:121`` does
``tf.keras.models.load_model(path)`` on an opaque ``.h5`` and ``predict.py:142`` calls
``.predict``.  This module is the host half of the replacement for that pair: it interprets the
``model_config`` JSON stored in the ``.h5`` (Sequential or Functional; layer set of SURVEY.md
App. D) and lowers it to the small fused-op vocabulary of ``include/timed_b200.h``:

* ``Conv3D``/``Dense``/``Flatten+Dense`` -> one implicit-GEMM op whose epilogue absorbs the
  bias, a following activation (ELU/ReLU), a following inference-mode BatchNormalization
  (folded to per-channel scale/shift) and one more activation -- so TIMED's
  Conv3D -> ELU -> BatchNorm block is a single kernel;
* Dropout / SpatialDropout3D are identities at inference and disappear;
* pooling, global pooling, softmax, concat, add map one-to-one.

Anything outside this vocabulary raises ``UnsupportedLayerError`` -- there is no fallback.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

OP_INPUT, OP_CONV3D, OP_POOL3D, OP_AFFINE, OP_GPOOL, OP_SOFTMAX, OP_CONCAT, OP_ADD = range(8)
ACT_CODES = {None: 0, "linear": 0, "relu": 1, "elu": 2, "sigmoid": 3, "tanh": 4}
IDENTITY_LAYERS = {"Dropout", "SpatialDropout3D", "SpatialDropout2D", "SpatialDropout1D",
                   "GaussianNoise", "GaussianDropout", "AlphaDropout", "ActivityRegularization"}


class UnsupportedLayerError(NotImplementedError):
    pass


@dataclasses.dataclass
class Op:
    kind: int
    inputs: List[int]
    name: str = ""
    kernel: Tuple[int, int, int] = (1, 1, 1)
    stride: Tuple[int, int, int] = (1, 1, 1)
    pad_same: bool = False
    c_out: int = 0
    pool_kind: int = 0            # 0 max, 1 avg
    act1: int = 0
    act2: int = 0
    alpha1: float = 1.0
    alpha2: float = 1.0
    kernel_w: Optional[np.ndarray] = None   # (kd,kh,kw,Cin,Cout) float32
    bias: Optional[np.ndarray] = None
    scale: Optional[np.ndarray] = None
    shift: Optional[np.ndarray] = None
    out_shape: Tuple[int, int, int, int] = (1, 1, 1, 0)   # (D,H,W,C)
    flops: float = 0.0
    fused: List[str] = dataclasses.field(default_factory=list)


@dataclasses.dataclass
class _KLayer:
    name: str
    cls: str
    cfg: dict
    inputs: List[str]


def _same_out(n: int, s: int) -> int:
    return -(-n // s)


def _weight(weights: Dict[str, Dict[str, np.ndarray]], layer: str, key: str) -> np.ndarray:
    if layer not in weights:
        raise KeyError(f"no weights stored for layer '{layer}'")
    for k, v in weights[layer].items():
        if k.split("/")[-1] == key:
            return np.asarray(v, dtype=np.float32)
    raise KeyError(f"layer '{layer}' has no weight '{key}' (has {list(weights[layer])})")


def _klayers(model_config: dict) -> Tuple[List[_KLayer], str]:
    cls = model_config.get("class_name")
    cfg = model_config["config"]
    layers = cfg["layers"] if isinstance(cfg, dict) else cfg
    out: List[_KLayer] = []
    if cls == "Sequential":
        prev = None
        for layer in layers:
            lcfg = layer["config"]
            name = lcfg["name"]
            if layer["class_name"] == "InputLayer":
                out.append(_KLayer(name, "InputLayer", lcfg, []))
            else:
                if prev is None:
                    shape = lcfg.get("batch_input_shape")
                    if shape is None:
                        raise UnsupportedLayerError("Sequential model without an input shape")
                    out.append(_KLayer("__input__", "InputLayer", {"batch_input_shape": shape}, []))
                    prev = "__input__"
                out.append(_KLayer(name, layer["class_name"], lcfg, [prev]))
            prev = name
        return out, prev
    if cls not in ("Functional", "Model"):
        raise UnsupportedLayerError(f"model class {cls}")
    for layer in layers:
        ins: List[str] = []
        inb = layer.get("inbound_nodes", [])
        if len(inb) > 1:
            raise UnsupportedLayerError(f"layer {layer['name']} is applied more than once (shared layers)")
        if inb:
            for item in inb[0]:
                ins.append(item[0])
        out.append(_KLayer(layer["name"], layer["class_name"], layer["config"], ins))
    outs = cfg.get("output_layers", [])
    if len(outs) != 1 or len(cfg.get("input_layers", [])) != 1:
        raise UnsupportedLayerError("exactly one input and one output are supported")
    return out, outs[0][0]


class Graph:
    """Fused op list + shape/FLOP bookkeeping."""

    def __init__(self, ops: List[Op], input_shape: Tuple[int, int, int, int], name: str = "model"):
        self.ops = ops
        self.input_shape = input_shape
        self.name = name

    @property
    def n_classes(self) -> int:
        return self.ops[-1].out_shape[3]

    def flops_per_frame(self) -> float:
        """sum 2*Do*Ho*Wo*k^3*Cin*Cout + sum 2*in*out  (SURVEY.md 8(d))."""
        return float(sum(op.flops for op in self.ops))

    def describe(self) -> str:
        kinds = ["input", "conv3d", "pool3d", "affine", "gpool", "softmax", "concat", "add"]
        lines = []
        for i, op in enumerate(self.ops):
            lines.append(f"{i:3d} {kinds[op.kind]:8s} in={op.inputs} out={op.out_shape} "
                         f"k={op.kernel} fused={'+'.join(op.fused)}")
        return "\n".join(lines)


def _bn_fold(layer: _KLayer, weights, c: int) -> Tuple[np.ndarray, np.ndarray]:
    cfg = layer.cfg
    axis = cfg.get("axis", -1)
    if isinstance(axis, (list, tuple)):
        axis = axis[0]
    # channels_last tensors only: axis must be the last one (-1, 4 for 5-D, 1 for 2-D)
    if axis not in (-1, 4, 1):
        raise UnsupportedLayerError(f"BatchNormalization axis {axis} (only the channel axis)")
    eps = float(cfg.get("epsilon", 1e-3))
    gamma = _weight(weights, layer.name, "gamma:0").astype(np.float64) if cfg.get("scale", True) \
        else np.ones(c)
    beta = _weight(weights, layer.name, "beta:0").astype(np.float64) if cfg.get("center", True) \
        else np.zeros(c)
    mean = _weight(weights, layer.name, "moving_mean:0").astype(np.float64)
    var = _weight(weights, layer.name, "moving_variance:0").astype(np.float64)
    if not (len(gamma) == len(beta) == len(mean) == len(var) == c):
        raise ValueError(f"{layer.name}: BatchNormalization vectors do not match {c} channels")
    scale = gamma / np.sqrt(var + eps)
    shift = beta - mean * scale
    return scale.astype(np.float32), shift.astype(np.float32)


def _act_of(layer: _KLayer) -> Optional[Tuple[str, float]]:
    """(activation name, alpha) if the layer is a pure elementwise activation we can fuse."""
    if layer.cls == "ELU":
        return "elu", float(layer.cfg.get("alpha", 1.0))
    if layer.cls == "ReLU":
        c = layer.cfg
        if c.get("max_value") is not None or c.get("negative_slope", 0.0) or c.get("threshold", 0.0):
            raise UnsupportedLayerError("ReLU with max_value/negative_slope/threshold")
        return "relu", 1.0
    if layer.cls == "Activation":
        fn = layer.cfg["activation"]
        if fn in ACT_CODES and fn not in (None, "linear"):
            return fn, 1.0
        if fn == "linear":
            return "linear", 1.0
        return None
    return None


def parse_model_config(model_config: dict, weights: Optional[Dict[str, Dict[str, np.ndarray]]] = None,
                       ) -> Graph:
    """Lower a Keras model_config (+ weights) to fused ops.  With ``weights=None`` only shapes
    and FLOPs are produced (zero-filled parameters)."""
    klayers, out_name = _klayers(model_config)
    by_name = {l.name: l for l in klayers}
    consumers: Dict[str, List[str]] = {l.name: [] for l in klayers}
    for l in klayers:
        for i in l.inputs:
            consumers[i].append(l.name)

    def w(layer, key, shape=None):
        if weights is None:
            return np.zeros(shape, dtype=np.float32)
        arr = _weight(weights, layer, key)
        if shape is not None and tuple(arr.shape) != tuple(shape):
            raise ValueError(f"{layer}/{key}: expected shape {tuple(shape)}, file has {arr.shape}")
        return arr

    ops: List[Op] = []
    tensor_of: Dict[str, int] = {}      # keras layer name -> op id producing its output
    absorbed: set = set()

    def sole_consumer(name: str) -> Optional[_KLayer]:
        cs = consumers[name]
        if len(cs) != 1 or name == out_name:
            return None
        return by_name[cs[0]]

    def absorb_chain(op: Op, tail: str) -> str:
        """Fold identity / activation / BatchNorm layers that follow ``tail`` into ``op``'s
        epilogue  act2(scale * act1(x) + shift).  Returns the last absorbed layer name."""
        while True:
            nxt = sole_consumer(tail)
            if nxt is None:
                return tail
            if nxt.cls in IDENTITY_LAYERS:
                absorbed.add(nxt.name)
                tail = nxt.name
                continue
            a = _act_of(nxt)
            if a is not None:
                fn, alpha = a
                if fn != "linear":
                    if op.act1 == 0 and op.scale is None and op.act2 == 0:
                        op.act1, op.alpha1 = ACT_CODES[fn], alpha
                    elif op.act2 == 0:
                        op.act2, op.alpha2 = ACT_CODES[fn], alpha
                    else:
                        return tail
                    op.fused.append(nxt.cls if nxt.cls != "Activation" else fn)
                absorbed.add(nxt.name)
                tail = nxt.name
                continue
            if nxt.cls == "BatchNormalization" and op.scale is None and op.act2 == 0:
                c = op.out_shape[3]
                if weights is None:
                    op.scale, op.shift = np.ones(c, np.float32), np.zeros(c, np.float32)
                else:
                    op.scale, op.shift = _bn_fold(nxt, weights, c)
                op.fused.append("BatchNormalization")
                absorbed.add(nxt.name)
                tail = nxt.name
                continue
            return tail

    flattened: Dict[str, bool] = {}
    input_shape = None
    for layer in klayers:
        if layer.name in absorbed:
            continue
        ins = [tensor_of[i] for i in layer.inputs]
        in_shape = ops[ins[0]].out_shape if ins else None
        cls, cfg = layer.cls, layer.cfg
        if cls == "InputLayer":
            shp = cfg.get("batch_input_shape") or cfg.get("batch_shape")
            if shp is None or len(shp) != 5:
                raise UnsupportedLayerError(f"input must be (None,D,H,W,C), got {shp}")
            input_shape = tuple(int(x) for x in shp[1:])
            op = Op(OP_INPUT, [], layer.name, kernel=input_shape[:3], c_out=input_shape[3],
                    out_shape=input_shape)
            ops.append(op)
            tensor_of[layer.name] = len(ops) - 1
            continue
        if cls in IDENTITY_LAYERS:
            tensor_of[layer.name] = ins[0]
            flattened[layer.name] = flattened.get(layer.inputs[0], False)
            continue
        if cls == "Conv3D":
            if cfg.get("data_format", "channels_last") != "channels_last":
                raise UnsupportedLayerError("Conv3D channels_first")
            if tuple(cfg.get("dilation_rate", (1, 1, 1))) != (1, 1, 1) or cfg.get("groups", 1) != 1:
                raise UnsupportedLayerError("Conv3D dilation/groups")
            if tuple(cfg.get("strides", (1, 1, 1))) != (1, 1, 1):
                raise UnsupportedLayerError("Conv3D strides != 1")
            k = tuple(int(x) for x in cfg["kernel_size"])
            d, h, wd, c = in_shape
            co = int(cfg["filters"])
            same = cfg["padding"] == "same"
            if cfg["padding"] not in ("same", "valid"):
                raise UnsupportedLayerError(f"Conv3D padding {cfg['padding']}")
            o = (d, h, wd) if same else (d - k[0] + 1, h - k[1] + 1, wd - k[2] + 1)
            op = Op(OP_CONV3D, ins, layer.name, kernel=k, pad_same=same, c_out=co,
                    kernel_w=w(layer.name, "kernel:0", (*k, c, co)),
                    bias=w(layer.name, "bias:0", (co,)) if cfg.get("use_bias", True) else None,
                    out_shape=(*o, co), flops=2.0 * o[0] * o[1] * o[2] * k[0] * k[1] * k[2] * c * co,
                    fused=["Conv3D"])
            tail = _finish_contraction(op, layer, cfg.get("activation", "linear"), ops, tensor_of,
                                       absorb_chain)
            continue
        if cls == "Flatten":
            # NDHWC row-major flatten == the memory order: alias; the Dense below is lowered to a
            # 'valid' convolution whose kernel spans the whole (D,H,W) volume.
            tensor_of[layer.name] = ins[0]
            flattened[layer.name] = True
            continue
        if cls == "Dense":
            d, h, wd, c = in_shape
            units = int(cfg["units"])
            if (d, h, wd) != (1, 1, 1) and not flattened.get(layer.inputs[0], False):
                raise UnsupportedLayerError("Dense on a spatial tensor without Flatten")
            feat = d * h * wd * c
            kern = w(layer.name, "kernel:0", (feat, units)).reshape(d, h, wd, c, units)
            if max(d, h, wd) > 16:
                raise UnsupportedLayerError("Flatten+Dense over more than 16 voxels per side")
            op = Op(OP_CONV3D, ins, layer.name, kernel=(d, h, wd), pad_same=False, c_out=units,
                    kernel_w=np.ascontiguousarray(kern),
                    bias=w(layer.name, "bias:0", (units,)) if cfg.get("use_bias", True) else None,
                    out_shape=(1, 1, 1, units), flops=2.0 * feat * units, fused=["Dense"])
            _finish_contraction(op, layer, cfg.get("activation", "linear"), ops, tensor_of, absorb_chain)
            continue
        if cls == "BatchNormalization" or _act_of(layer) is not None:
            c = in_shape[3]
            op = Op(OP_AFFINE, ins, layer.name, out_shape=in_shape, c_out=c, fused=[])
            if cls == "BatchNormalization":
                if weights is None:
                    op.scale, op.shift = np.ones(c, np.float32), np.zeros(c, np.float32)
                else:
                    op.scale, op.shift = _bn_fold(layer, weights, c)
                op.fused.append("BatchNormalization")
                tail = absorb_chain(op, layer.name)
            else:
                fn, alpha = _act_of(layer)
                if fn == "linear":
                    tensor_of[layer.name] = ins[0]
                    continue
                op.act1, op.alpha1 = ACT_CODES[fn], alpha
                op.fused.append(fn)
                tail = absorb_chain(op, layer.name)
            ops.append(op)
            tensor_of[layer.name] = tensor_of[tail] = len(ops) - 1
            continue
        if cls == "Softmax" or (cls == "Activation" and cfg.get("activation") == "softmax"):
            ax = cfg.get("axis", -1)
            if ax not in (-1, 1, 4):
                raise UnsupportedLayerError(f"Softmax axis {ax}")
            if in_shape[:3] != (1, 1, 1):
                raise UnsupportedLayerError("Softmax over a spatial tensor")
            ops.append(Op(OP_SOFTMAX, ins, layer.name, out_shape=in_shape, c_out=in_shape[3],
                          fused=["Softmax"]))
            tensor_of[layer.name] = len(ops) - 1
            continue
        if cls in ("MaxPooling3D", "AveragePooling3D"):
            size = tuple(int(x) for x in cfg["pool_size"])
            st = tuple(int(x) for x in (cfg.get("strides") or size))
            same = cfg["padding"] == "same"
            d, h, wd, c = in_shape
            dims = (d, h, wd)
            o = tuple(_same_out(dims[i], st[i]) if same else (dims[i] - size[i]) // st[i] + 1
                      for i in range(3))
            ops.append(Op(OP_POOL3D, ins, layer.name, kernel=size, stride=st, pad_same=same,
                          pool_kind=0 if cls.startswith("Max") else 1, out_shape=(*o, c), c_out=c,
                          fused=[cls]))
            tensor_of[layer.name] = len(ops) - 1
            continue
        if cls in ("GlobalAveragePooling3D", "GlobalMaxPooling3D"):
            if cfg.get("keepdims", False):
                raise UnsupportedLayerError("global pooling with keepdims")
            ops.append(Op(OP_GPOOL, ins, layer.name, pool_kind=1 if "Average" in cls else 0,
                          out_shape=(1, 1, 1, in_shape[3]), c_out=in_shape[3], fused=[cls]))
            tensor_of[layer.name] = len(ops) - 1
            continue
        if cls == "Concatenate":
            if cfg.get("axis", -1) not in (-1, 4):
                raise UnsupportedLayerError("Concatenate on a non-channel axis")
            if len(ins) > 8:
                raise UnsupportedLayerError("Concatenate of more than 8 tensors")
            c = sum(ops[i].out_shape[3] for i in ins)
            ops.append(Op(OP_CONCAT, ins, layer.name, out_shape=(*in_shape[:3], c), c_out=c,
                          fused=["Concatenate"]))
            tensor_of[layer.name] = len(ops) - 1
            continue
        if cls == "Add":
            cur = ins[0]
            for other in ins[1:]:
                ops.append(Op(OP_ADD, [cur, other], layer.name, out_shape=in_shape, c_out=in_shape[3],
                              fused=["Add"]))
                cur = len(ops) - 1
            tensor_of[layer.name] = cur
            continue
        raise UnsupportedLayerError(f"layer class {cls} ({layer.name}) is not supported")

    if input_shape is None:
        raise UnsupportedLayerError("model has no InputLayer")
    out_id = tensor_of[out_name]
    if out_id != len(ops) - 1:
        raise UnsupportedLayerError("the model output must be the last layer of the graph")
    if ops[-1].out_shape[:3] != (1, 1, 1):
        raise UnsupportedLayerError(f"model output must be (None, classes), got {ops[-1].out_shape}")
    name = model_config["config"].get("name", "model") if isinstance(model_config["config"], dict) else "model"
    return Graph(ops, input_shape, name)


def _finish_contraction(op: Op, layer: _KLayer, activation: str, ops, tensor_of, absorb_chain):
    """Handle the Keras ``activation=`` argument of Conv3D/Dense, then fuse what follows."""
    softmax_after = False
    if activation == "softmax":
        softmax_after = True
    elif activation in ACT_CODES:
        op.act1 = ACT_CODES[activation]
        if op.act1:
            op.fused.append(activation)
    else:
        raise UnsupportedLayerError(f"activation '{activation}' on {layer.name}")
    if softmax_after:
        ops.append(op)
        ops.append(Op(OP_SOFTMAX, [len(ops) - 1], layer.name + "/softmax", out_shape=op.out_shape,
                      c_out=op.out_shape[3], fused=["softmax"]))
        if op.out_shape[:3] != (1, 1, 1):
            raise UnsupportedLayerError("softmax activation on a spatial tensor")
        tensor_of[layer.name] = len(ops) - 1
        return layer.name
    tail = absorb_chain(op, layer.name)
    ops.append(op)
    tensor_of[layer.name] = tensor_of[tail] = len(ops) - 1
    return tail
