"""Synthetic stand-in Keras graphs -- NOT the authors' weights.

The reference ships no network definition: ``predict.py:121`` loads an opaque
Keras ``.h5``.  Nothing in ``/root/reference`` carries the layer graph, so the
benchmarks and parity tests run on stand-ins generated here, in exactly the
container format Keras 2.x writes (``model_config`` JSON + per-layer weight
arrays), so that the same loader handles a real ``TIMED.h5``.

Every generator returns ``(model_config: dict, weights: dict)`` where
``weights[layer_name]`` is an ordered ``{weight_name: float32 ndarray}`` using
Keras' names (``kernel:0``, ``bias:0``, ``gamma:0``, ``beta:0``,
``moving_mean:0``, ``moving_variance:0``).
"""
from __future__ import annotations

import numpy as np


def _inbound(*names):
    return [[[n, 0, 0, {}] for n in names]]


class _Builder:
    """Tiny Keras-functional-config builder with deterministic weights."""

    def __init__(self, name: str, input_shape, seed: int, calib: np.ndarray | None = None):
        self.name = name
        self.rng = np.random.default_rng(seed)
        # Calibration batch pushed through the graph as it is built (fixture generation only,
        # plain torch-CPU): BatchNorm moving statistics are set to the batch statistics of
        # their input, as a trained network's would be, so activations stay standardised and
        # the outputs are frame-dependent instead of a constant class bias.
        self.cal = {}
        self._calib0 = None
        if calib is not None:
            import torch
            self._calib0 = torch.from_numpy(np.ascontiguousarray(calib, dtype=np.float32)) \
                .permute(0, 4, 1, 2, 3).contiguous()
        self.layers = []
        self.weights = {}
        self.counts = {}
        self.shapes = {}
        self.input_name = "input_1"
        self.layers.append({
            "class_name": "InputLayer",
            "config": {"batch_input_shape": [None, *input_shape], "dtype": "float32",
                       "sparse": False, "ragged": False, "name": self.input_name},
            "name": self.input_name, "inbound_nodes": [],
        })
        self.shapes[self.input_name] = tuple(input_shape)
        if self._calib0 is not None:
            self.cal[self.input_name] = self._calib0

    def _cal(self, name, fn, inputs):
        """Apply ``fn`` to the calibration tensors of ``inputs`` (no-op without calibration)."""
        if self._calib0 is None:
            return
        import torch
        with torch.no_grad():
            self.cal[name] = fn(*[self.cal[i] for i in inputs])

    @staticmethod
    def _tf_same_pad(t, ks, st, value=0.0):
        import torch.nn.functional as F
        pads = []
        for dim, k, s_ in zip(t.shape[2:], ks, st):
            out = -(-dim // s_)
            tot = max((out - 1) * s_ + k - dim, 0)
            pads.append((tot // 2, tot - tot // 2))
        return F.pad(t, (*pads[2], *pads[1], *pads[0]), value=value)

    @staticmethod
    def _act(t, fn):
        import torch
        import torch.nn.functional as F
        if fn in (None, "linear"):
            return t
        if fn == "relu":
            return F.relu(t)
        if fn == "elu":
            return F.elu(t)
        if fn == "softmax":
            return torch.softmax(t, dim=-1)
        raise NotImplementedError(fn)

    def _name(self, base):
        k = self.counts.get(base, 0)
        self.counts[base] = k + 1
        return base if k == 0 else f"{base}_{k}"

    def _add(self, cls, base, cfg, inputs, out_shape):
        name = self._name(base)
        cfg = dict(cfg, name=name, trainable=True, dtype="float32")
        self.layers.append({"class_name": cls, "config": cfg, "name": name,
                            "inbound_nodes": _inbound(*inputs)})
        self.shapes[name] = tuple(out_shape)
        return name

    @staticmethod
    def _same_out(n, s):
        return -(-n // s)

    def conv3d(self, x, filters, k, padding="same", use_bias=True, activation="linear"):
        d, h, w, c = self.shapes[x]
        if padding == "same":
            od, oh, ow = d, h, w
        else:
            od, oh, ow = d - k + 1, h - k + 1, w - k + 1
        name = self._add("Conv3D", "conv3d", {
            "filters": filters, "kernel_size": [k, k, k], "strides": [1, 1, 1],
            "padding": padding, "data_format": "channels_last", "dilation_rate": [1, 1, 1],
            "groups": 1, "activation": activation, "use_bias": use_bias}, [x], (od, oh, ow, filters))
        fan_in = k * k * k * c
        wts = {"kernel:0": (self.rng.standard_normal((k, k, k, c, filters)) *
                            np.sqrt(2.0 / fan_in)).astype(np.float32)}
        if use_bias:
            wts["bias:0"] = (self.rng.standard_normal(filters) * 0.05).astype(np.float32)
        self.weights[name] = wts

        def fn(t):
            import torch
            import torch.nn.functional as F
            kk = torch.from_numpy(wts["kernel:0"]).permute(4, 3, 0, 1, 2).contiguous()
            bb = torch.from_numpy(wts["bias:0"]) if use_bias else None
            if padding == "same":
                t = self._tf_same_pad(t, (k, k, k), (1, 1, 1))
            return self._act(F.conv3d(t, kk, bb), activation)
        self._cal(name, fn, [x])
        return name

    def dense(self, x, units, activation="linear"):
        (f,) = self.shapes[x]
        name = self._add("Dense", "dense", {"units": units, "activation": activation,
                                            "use_bias": True}, [x], (units,))
        self.weights[name] = {
            "kernel:0": (self.rng.standard_normal((f, units)) * np.sqrt(2.0 / f)).astype(np.float32),
            "bias:0": (self.rng.standard_normal(units) * 0.05).astype(np.float32)}
        wd = self.weights[name]

        def fn(t):
            import torch
            return self._act(t @ torch.from_numpy(wd["kernel:0"]) + torch.from_numpy(wd["bias:0"]),
                             activation)
        self._cal(name, fn, [x])
        return name

    def bn(self, x, eps=1e-3, gain=1.0):
        shp = self.shapes[x]
        c = shp[-1]
        name = self._add("BatchNormalization", "batch_normalization", {
            "axis": [len(shp)], "momentum": 0.99, "epsilon": eps, "center": True, "scale": True},
            [x], shp)
        r = self.rng
        self.weights[name] = {
            "gamma:0": (r.uniform(0.5, 1.5, c) * gain).astype(np.float32),
            "beta:0": (r.standard_normal(c) * 0.1).astype(np.float32),
            "moving_mean:0": (r.standard_normal(c) * 0.1).astype(np.float32),
            "moving_variance:0": r.uniform(0.5, 1.5, c).astype(np.float32)}
        wb = self.weights[name]
        if self._calib0 is not None:
            t = self.cal[x]
            red = [0] + list(range(2, t.dim()))
            wb["moving_mean:0"] = t.mean(dim=red).numpy().astype(np.float32)
            wb["moving_variance:0"] = np.maximum(
                t.var(dim=red, unbiased=False).numpy(), 1e-4).astype(np.float32)

        def fn(t):
            import torch
            shp = [1, c] + [1] * (t.dim() - 2)
            sc = torch.from_numpy(wb["gamma:0"] / np.sqrt(wb["moving_variance:0"] + np.float32(eps)))
            return (t - torch.from_numpy(wb["moving_mean:0"]).view(shp)) * sc.view(shp) \
                + torch.from_numpy(wb["beta:0"]).view(shp)
        self._cal(name, fn, [x])
        return name

    def elu(self, x, alpha=1.0):
        name = self._add("ELU", "elu", {"alpha": alpha}, [x], self.shapes[x])
        self._cal(name, lambda t: self._act(t, "elu"), [x])
        return name

    def relu(self, x):
        name = self._add("ReLU", "re_lu", {"max_value": None, "negative_slope": 0.0,
                                           "threshold": 0.0}, [x], self.shapes[x])
        self._cal(name, lambda t: self._act(t, "relu"), [x])
        return name

    def activation(self, x, fn):
        name = self._add("Activation", "activation", {"activation": fn}, [x], self.shapes[x])
        self._cal(name, lambda t: self._act(t, fn), [x])
        return name

    def softmax(self, x):
        name = self._add("Softmax", "softmax", {"axis": -1}, [x], self.shapes[x])
        self._cal(name, lambda t: self._act(t, "softmax"), [x])
        return name

    def dropout(self, x, spatial=False, rate=0.3):
        cls = "SpatialDropout3D" if spatial else "Dropout"
        base = "spatial_dropout3d" if spatial else "dropout"
        name = self._add(cls, base, {"rate": rate, "noise_shape": None, "seed": None},
                         [x], self.shapes[x])
        self._cal(name, lambda t: t, [x])
        return name

    def pool(self, x, kind="max", size=2, padding="same"):
        d, h, w, c = self.shapes[x]
        if padding == "same":
            o = [self._same_out(n, size) for n in (d, h, w)]
        else:
            o = [(n - size) // size + 1 for n in (d, h, w)]
        cls = "MaxPooling3D" if kind == "max" else "AveragePooling3D"
        base = "max_pooling3d" if kind == "max" else "average_pooling3d"
        name = self._add(cls, base, {"pool_size": [size] * 3, "strides": [size] * 3,
                                     "padding": padding, "data_format": "channels_last"},
                         [x], (*o, c))

        def fn(t):
            import torch
            import torch.nn.functional as F
            ks = st = (size,) * 3
            if kind == "max":
                if padding == "same":
                    t = self._tf_same_pad(t, ks, st, float("-inf"))
                return F.max_pool3d(t, size, size)
            if padding == "same":
                ones = torch.ones((1, 1, *t.shape[2:]))
                return F.avg_pool3d(self._tf_same_pad(t, ks, st), size, size) / \
                    F.avg_pool3d(self._tf_same_pad(ones, ks, st), size, size)
            return F.avg_pool3d(t, size, size)
        self._cal(name, fn, [x])
        return name

    def gap(self, x):
        name = self._add("GlobalAveragePooling3D", "global_average_pooling3d",
                         {"data_format": "channels_last", "keepdims": False}, [x],
                         (self.shapes[x][-1],))
        self._cal(name, lambda t: t.mean(dim=(2, 3, 4)), [x])
        return name

    def flatten(self, x):
        name = self._add("Flatten", "flatten", {"data_format": "channels_last"}, [x],
                         (int(np.prod(self.shapes[x])),))
        self._cal(name, lambda t: t.permute(0, 2, 3, 4, 1).reshape(t.shape[0], -1), [x])
        return name

    def concat(self, xs):
        shp = self.shapes[xs[0]]
        c = sum(self.shapes[x][-1] for x in xs)
        name = self._add("Concatenate", "concatenate", {"axis": -1}, xs, (*shp[:-1], c))
        if self._calib0 is not None:
            import torch
            self.cal[name] = torch.cat([self.cal[i] for i in xs], dim=1)
        return name

    def add(self, xs):
        name = self._add("Add", "add", {}, xs, self.shapes[xs[0]])
        if self._calib0 is not None:
            self.cal[name] = sum(self.cal[i] for i in xs)
        return name

    def finish(self, out):
        cfg = {"class_name": "Functional",
               "config": {"name": self.name, "layers": self.layers,
                          "input_layers": [[self.input_name, 0, 0]],
                          "output_layers": [[out, 0, 0]]},
               "keras_version": "2.13.1", "backend": "tensorflow"}
        return cfg, self.weights


def timed_standin(n_classes: int = 20, c_in: int = 6, seed: int = 7,
                  filters=(32, 64, 128, 256, 512), side: int = 21, logit_gain: float = 8.0,
                  calib_frames: int = 6):
    """TIMED stand-in (README.md:254 prose): six Conv3D(k3, same)+bias -> ELU -> BN blocks,
    MaxPool(2, same) after blocks 1 and 2, SpatialDropout (identity), last block's
    C_out = n_classes, GlobalAveragePooling -> Softmax.  2.3692 GFLOP/frame at 20 classes.
    ``logit_gain`` scales the last BatchNorm's gamma so that random weights still give peaked,
    frame-dependent probabilities (otherwise GAP of a unit-variance BN output is ~beta for every
    frame and the argmax parity test would be vacuous).  Probability errors of any finite-precision
    evaluation scale linearly with this gain (DESIGN.md "Numerics"): 8 gives max-probabilities of
    0.2-0.9 across frames; the B200 path measures max |dp| 1.35e-4 at gain 16 and half that at 8."""
    b = _Builder(f"TIMED_standin_{n_classes}", (side, side, side, c_in), seed,
                 synthetic_frames(calib_frames, side, c_in, seed=99) if calib_frames else None)
    x = b.input_name
    for i, f in enumerate(list(filters) + [n_classes]):
        x = b.conv3d(x, f, 3, "same")
        x = b.elu(x)
        x = b.bn(x, gain=logit_gain if i == len(filters) else 1.0)
        if i < 2:
            x = b.pool(x, "max", 2, "same")
        x = b.dropout(x, spatial=True)
    x = b.gap(x)
    x = b.softmax(x)
    return b.finish(x)


def densecpd_standin(n_classes: int = 20, c_in: int = 6, seed: int = 11, growth: int = 32,
                     n_layers: int = 6, side: int = 21, bottleneck: int = 128, stem: int = 64,
                     calib_frames: int = 4):
    """DenseNet-BC style stand-in: stem conv, 3 dense blocks of ``n_layers``
    [BN->ReLU->Conv1x1x1(bottleneck) -> BN->ReLU->Conv3x3x3(growth)] with channel concat,
    transitions BN->ReLU->Conv1x1x1(C/2)->AvgPool(2, valid); BN->ReLU->GAP->Dense softmax."""
    b = _Builder(f"DenseCPD_standin_{n_classes}", (side, side, side, c_in), seed,
                 synthetic_frames(calib_frames, side, c_in, seed=99) if calib_frames else None)
    x = b.conv3d(b.input_name, stem, 3, "same", use_bias=False)
    for blk in range(3):
        for _ in range(n_layers):
            y = b.relu(b.bn(x))
            y = b.conv3d(y, bottleneck, 1, "same", use_bias=False)
            y = b.relu(b.bn(y))
            y = b.conv3d(y, growth, 3, "same", use_bias=False)
            x = b.concat([x, y])
        if blk < 2:
            y = b.relu(b.bn(x))
            y = b.conv3d(y, b.shapes[x][-1] // 2, 1, "same", use_bias=False)
            x = b.pool(y, "avg", 2, "valid")
    x = b.relu(b.bn(x))
    x = b.gap(x)
    x = b.dense(x, n_classes, activation="softmax")
    return b.finish(x)


def prodconn_standin(n_classes: int = 20, c_in: int = 6, seed: int = 13, side: int = 21,
                     branch: int = 16, calib_frames: int = 4):
    """ProDCoNN-style stand-in for layer-type coverage: parallel Conv3D branches
    k in {3,5,7} (same, ReLU fused as Keras ``activation='relu'``) -> Concatenate ->
    Conv3D(valid) -> MaxPool(valid) -> Dropout -> Flatten -> Dense(relu) -> Dense(softmax)."""
    b = _Builder(f"ProDCoNN_standin_{n_classes}", (side, side, side, c_in), seed,
                 synthetic_frames(calib_frames, side, c_in, seed=99) if calib_frames else None)
    br = [b.conv3d(b.input_name, branch, k, "same", activation="relu") for k in (3, 5, 7)]
    x = b.concat(br)
    x = b.pool(x, "max", 2, "valid")            # 21 -> 10
    x = b.conv3d(x, 64, 3, "valid", activation="relu")   # 10 -> 8
    x = b.pool(x, "max", 2, "valid")            # 8 -> 4
    x = b.dropout(x)
    x = b.flatten(x)
    x = b.dense(x, 128, activation="relu")
    x = b.dense(x, n_classes, activation="softmax")
    return b.finish(x)


def tiny_standin(n_classes: int = 20, c_in: int = 6, seed: int = 3, side: int = 9,
                 filters=(16, 32), **kw):
    """Small TIMED-shaped graph for fast parity tests (oracle runs in milliseconds)."""
    return timed_standin(n_classes, c_in, seed, filters, side, **kw)


def conv_flops_per_frame(model_config: dict) -> float:
    """Algorithmic FLOPs/frame of a Keras graph: sum 2*D*H*W*k^3*Cin*Cout + sum 2*in*out
    (SURVEY.md 8(d)); computed from the config alone by shape propagation."""
    from .keras_graph import parse_model_config
    g = parse_model_config(model_config)
    return g.flops_per_frame()


def synthetic_frames(n: int, side: int = 21, c: int = 6, seed: int = 1234,
                     first_index: int = 0, dtype=np.float32) -> np.ndarray:
    """Deterministic synthetic gaussian-blob frames, counter-based on the GLOBAL frame index
    (so sharding across ranks does not change any frame): ~40-120 'atoms' per frame,
    sigma ~0.6 voxel, channel uniform over ``c``."""
    out = np.zeros((n, side, side, side, c), dtype=dtype)
    ax = np.arange(side, dtype=np.float64)
    sig = 0.6
    for i in range(n):
        r = np.random.default_rng([seed, first_index + i])
        na = int(r.integers(40, 121))
        pos = r.uniform(0, side - 1, size=(na, 3))
        # per-frame channel mix (Dirichlet) so frames differ in composition, not just position
        mix = r.dirichlet(np.full(c, 0.7))
        ch = r.choice(c, size=na, p=mix)
        g = np.exp(-0.5 * ((ax[None, None, :] - pos.T[:, :, None]) / sig) ** 2)   # (3, na, side)
        for k in range(c):
            m = ch == k
            if m.any():
                out[i, :, :, :, k] = np.einsum("ax,ay,az->xyz", g[0, m], g[1, m], g[2, m],
                                               optimize=True).astype(dtype)
    return out
