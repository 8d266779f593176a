"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of what ``frame_model.predict(X_batch)`` computes in the reference
(``/root/reference/predict.py:121,142``).  The arithmetic lives in third-party
tensorflow==2.13.0 (``/root/reference/requirements.txt:8``), which is absent from this image
and from ``/root/reference`` -> **parity unpinned** for the conv stack: no reference test or
golden vector covers it.  The restatement follows the published Keras 2.13
inference semantics of each layer and is guarded by a second, independent
restatement (``forward_torch`` below, oneDNN ``conv3d``) -- the two must agree to ~1e-6.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.

``forward_numpy``  : layer-by-layer, un-fused, float64 by default; convolution is a loop
                     over kernel taps of shifted-view matmuls (obviously-correct form).
``forward_torch``  : same graph through ``torch.nn.functional`` on CPU in fp32 -- the class
                     of backend TF-CPU uses; this is the timed CPU baseline (kind "port").
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- graph walk
def _layers_in_order(model_config: dict):
    """Yield (name, class_name, config, [input names]) for Sequential and Functional configs."""
    cls = model_config["class_name"]
    cfg = model_config["config"]
    layers = cfg["layers"] if isinstance(cfg, dict) else cfg   # very old Sequential: list
    if cls == "Sequential":
        prev = None
        for i, layer in enumerate(layers):
            name = layer["config"]["name"]
            if layer["class_name"] == "InputLayer":
                yield name, "InputLayer", layer["config"], []
            else:
                if prev is None:          # implicit input
                    yield "__input__", "InputLayer", {
                        "batch_input_shape": layer["config"].get("batch_input_shape")}, []
                    prev = "__input__"
                yield name, layer["class_name"], layer["config"], [prev]
            prev = name
        return
    for layer in layers:
        inb = layer.get("inbound_nodes", [])
        ins = []
        if inb:
            node = inb[0]
            for item in node:
                ins.append(item[0])
        yield layer["name"], layer["class_name"], layer["config"], ins


def _output_name(model_config: dict, last: str) -> str:
    cfg = model_config["config"]
    if model_config["class_name"] != "Sequential" and "output_layers" in cfg:
        return cfg["output_layers"][0][0]
    return last


def _w(weights: dict, layer: str, key: str):
    for k, v in weights[layer].items():
        if k.split("/")[-1] == key:
            return np.asarray(v)
    raise KeyError(f"{layer}: no weight {key}")


def _has_w(weights: dict, layer: str, key: str) -> bool:
    return layer in weights and any(k.split("/")[-1] == key for k in weights[layer])


# ----------------------------------------------------------------------------- numpy ops
def _same_pads(n: int, k: int, s: int):
    """Keras/TF 'same': out = ceil(n/s); total = max((out-1)*s + k - n, 0); before = total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def np_activation(x, fn: str, alpha: float = 1.0):
    if fn in (None, "linear"):
        return x
    if fn == "relu":
        return np.maximum(x, 0)
    if fn == "elu":
        return np.where(x > 0, x, alpha * np.expm1(np.minimum(x, 0)))
    if fn == "softmax":
        z = x - x.max(axis=-1, keepdims=True)
        e = np.exp(z)
        return e / e.sum(axis=-1, keepdims=True)
    if fn == "sigmoid":
        return 1.0 / (1.0 + np.exp(-x))
    if fn == "tanh":
        return np.tanh(x)
    raise NotImplementedError(f"activation {fn}")


def np_conv3d(x, kernel, bias, padding: str, strides=(1, 1, 1)):
    """Cross-correlation, NDHWC x DHWIO, stride s, 'same'/'valid' (Keras Conv3D)."""
    kd, kh, kw, ci, co = kernel.shape
    n, d, h, w, c = x.shape
    assert c == ci
    sd, sh, sw = strides
    if padding == "same":
        od, pd0, pd1 = _same_pads(d, kd, sd)
        oh, ph0, ph1 = _same_pads(h, kh, sh)
        ow, pw0, pw1 = _same_pads(w, kw, sw)
        x = np.pad(x, ((0, 0), (pd0, pd1), (ph0, ph1), (pw0, pw1), (0, 0)))
    else:
        od, oh, ow = (d - kd) // sd + 1, (h - kh) // sh + 1, (w - kw) // sw + 1
    out = np.zeros((n, od, oh, ow, co), dtype=x.dtype)
    for a in range(kd):
        for b in range(kh):
            for cidx in range(kw):
                view = x[:, a:a + (od - 1) * sd + 1:sd, b:b + (oh - 1) * sh + 1:sh,
                         cidx:cidx + (ow - 1) * sw + 1:sw, :]
                out += view @ kernel[a, b, cidx].astype(x.dtype)
    if bias is not None:
        out += bias.astype(x.dtype)
    return out


def np_pool3d(x, size, strides, padding: str, kind: str):
    n, d, h, w, c = x.shape
    dims = (d, h, w)
    outs, pads = [], []
    for i in range(3):
        if padding == "same":
            o, p0, p1 = _same_pads(dims[i], size[i], strides[i])
        else:
            o, p0, p1 = (dims[i] - size[i]) // strides[i] + 1, 0, 0
        outs.append(o)
        pads.append((p0, p1))
    fill = -np.inf if kind == "max" else 0.0
    xp = np.pad(x, ((0, 0), *pads, (0, 0)), constant_values=fill)
    cnt = np.pad(np.ones((1, d, h, w, 1), dtype=x.dtype), ((0, 0), *pads, (0, 0)))
    acc = None
    num = None
    for a in range(size[0]):
        for b in range(size[1]):
            for cc in range(size[2]):
                sl = (slice(None),
                      slice(a, a + (outs[0] - 1) * strides[0] + 1, strides[0]),
                      slice(b, b + (outs[1] - 1) * strides[1] + 1, strides[1]),
                      slice(cc, cc + (outs[2] - 1) * strides[2] + 1, strides[2]),
                      slice(None))
                v = xp[sl]
                if kind == "max":
                    acc = v.copy() if acc is None else np.maximum(acc, v)
                else:
                    acc = v.copy() if acc is None else acc + v
                    num = cnt[sl].copy() if num is None else num + cnt[sl]
    if kind == "avg":
        acc = acc / num          # Keras/TF average excludes padded elements
    return acc


def forward_numpy(model_config: dict, weights: dict, X: np.ndarray, dtype=np.float64,
                  return_all: bool = False):
    """Un-fused, layer-by-layer evaluation.  ``X`` (B,D,H,W,C) any real/bool dtype; it is
    first cast to float32 as Keras does for a float32 InputLayer, then to ``dtype``."""
    vals = {}
    last = None
    x_in = np.asarray(X).astype(np.float32).astype(dtype)
    for name, cls, cfg, ins in _layers_in_order(model_config):
        a = [vals[i] for i in ins]
        if cls == "InputLayer":
            y = x_in
        elif cls == "Conv3D":
            k = _w(weights, name, "kernel:0").astype(dtype)
            b = _w(weights, name, "bias:0").astype(dtype) if cfg.get("use_bias", True) else None
            assert tuple(cfg.get("dilation_rate", (1, 1, 1))) == (1, 1, 1)
            y = np_conv3d(a[0], k, b, cfg["padding"], tuple(cfg.get("strides", (1, 1, 1))))
            y = np_activation(y, cfg.get("activation", "linear"))
        elif cls == "Dense":
            k = _w(weights, name, "kernel:0").astype(dtype)
            y = a[0] @ k
            if cfg.get("use_bias", True):
                y = y + _w(weights, name, "bias:0").astype(dtype)
            y = np_activation(y, cfg.get("activation", "linear"))
        elif cls == "BatchNormalization":
            c = a[0].shape[-1]
            eps = cfg.get("epsilon", 1e-3)
            gamma = _w(weights, name, "gamma:0").astype(dtype) if cfg.get("scale", True) else np.ones(c, dtype)
            beta = _w(weights, name, "beta:0").astype(dtype) if cfg.get("center", True) else np.zeros(c, dtype)
            mean = _w(weights, name, "moving_mean:0").astype(dtype)
            var = _w(weights, name, "moving_variance:0").astype(dtype)
            y = gamma * (a[0] - mean) / np.sqrt(var + dtype(eps)) + beta
        elif cls == "ELU":
            y = np_activation(a[0], "elu", cfg.get("alpha", 1.0))
        elif cls == "ReLU":
            assert not cfg.get("max_value") and not cfg.get("negative_slope") and not cfg.get("threshold")
            y = np_activation(a[0], "relu")
        elif cls == "Softmax":
            y = np_activation(a[0], "softmax")
        elif cls == "Activation":
            y = np_activation(a[0], cfg["activation"])
        elif cls in ("Dropout", "SpatialDropout3D", "GaussianNoise", "GaussianDropout"):
            y = a[0]
        elif cls in ("MaxPooling3D", "AveragePooling3D"):
            size = tuple(cfg["pool_size"])
            st = tuple(cfg.get("strides") or size)
            y = np_pool3d(a[0], size, st, cfg["padding"], "max" if cls.startswith("Max") else "avg")
        elif cls == "GlobalAveragePooling3D":
            y = a[0].mean(axis=(1, 2, 3))
        elif cls == "GlobalMaxPooling3D":
            y = a[0].max(axis=(1, 2, 3))
        elif cls == "Flatten":
            y = a[0].reshape(a[0].shape[0], -1)
        elif cls == "Concatenate":
            y = np.concatenate(a, axis=-1)
        elif cls == "Add":
            y = a[0]
            for t in a[1:]:
                y = y + t
        else:
            raise NotImplementedError(f"oracle: layer class {cls}")
        vals[name] = y
        last = name
    out = vals[_output_name(model_config, last)]
    return (out, vals) if return_all else out


# ----------------------------------------------------------------------------- torch restatement
def forward_torch(model_config: dict, weights: dict, X: np.ndarray, threads: int | None = None,
                  dtype: str = "float32"):
    """Independent restatement on torch-CPU (NCDHW, oneDNN conv3d) -- also the timed CPU
    baseline.  Explicit asymmetric 'same' padding, TF pooling semantics."""
    import torch
    import torch.nn.functional as F
    if threads:
        torch.set_num_threads(threads)
    td = getattr(torch, dtype)
    vals = {}
    last = None
    with torch.no_grad():
        x0 = torch.from_numpy(np.ascontiguousarray(np.asarray(X).astype(np.float32))).to(td)
        x0 = x0.permute(0, 4, 1, 2, 3).contiguous()
        for name, cls, cfg, ins in _layers_in_order(model_config):
            a = [vals[i] for i in ins]

            def act(t, fn, alpha=1.0):
                if fn in (None, "linear"):
                    return t
                if fn == "relu":
                    return F.relu(t)
                if fn == "elu":
                    return F.elu(t, alpha)
                if fn == "softmax":
                    return F.softmax(t, dim=1 if t.dim() > 2 else -1)
                if fn == "sigmoid":
                    return torch.sigmoid(t)
                if fn == "tanh":
                    return torch.tanh(t)
                raise NotImplementedError(fn)

            if cls == "InputLayer":
                y = x0
            elif cls == "Conv3D":
                k = torch.from_numpy(_w(weights, name, "kernel:0")).to(td).permute(4, 3, 0, 1, 2).contiguous()
                b = torch.from_numpy(_w(weights, name, "bias:0")).to(td) if cfg.get("use_bias", True) else None
                t = a[0]
                st = tuple(cfg.get("strides", (1, 1, 1)))
                if cfg["padding"] == "same":
                    pads = []
                    for dim, kk, ss in zip(t.shape[2:], k.shape[2:], st):
                        _, p0, p1 = _same_pads(dim, kk, ss)
                        pads.append((p0, p1))
                    t = F.pad(t, (*pads[2], *pads[1], *pads[0]))
                y = act(F.conv3d(t, k, b, stride=st), cfg.get("activation", "linear"))
            elif cls == "Dense":
                k = torch.from_numpy(_w(weights, name, "kernel:0")).to(td)
                y = a[0] @ k
                if cfg.get("use_bias", True):
                    y = y + torch.from_numpy(_w(weights, name, "bias:0")).to(td)
                y = act(y, cfg.get("activation", "linear"))
            elif cls == "BatchNormalization":
                c = a[0].shape[1]
                eps = cfg.get("epsilon", 1e-3)
                g = torch.from_numpy(_w(weights, name, "gamma:0")).to(td) if cfg.get("scale", True) else torch.ones(c, dtype=td)
                be = torch.from_numpy(_w(weights, name, "beta:0")).to(td) if cfg.get("center", True) else torch.zeros(c, dtype=td)
                mu = torch.from_numpy(_w(weights, name, "moving_mean:0")).to(td)
                var = torch.from_numpy(_w(weights, name, "moving_variance:0")).to(td)
                shp = [1, c] + [1] * (a[0].dim() - 2)
                y = (a[0] - mu.view(shp)) * (g / torch.sqrt(var + eps)).view(shp) + be.view(shp)
            elif cls == "ELU":
                y = F.elu(a[0], cfg.get("alpha", 1.0))
            elif cls == "ReLU":
                y = F.relu(a[0])
            elif cls == "Softmax":
                y = F.softmax(a[0], dim=-1 if a[0].dim() == 2 else 1)
            elif cls == "Activation":
                y = act(a[0], cfg["activation"])
            elif cls in ("Dropout", "SpatialDropout3D", "GaussianNoise", "GaussianDropout"):
                y = a[0]
            elif cls in ("MaxPooling3D", "AveragePooling3D"):
                size = tuple(cfg["pool_size"])
                st = tuple(cfg.get("strides") or size)
                t = a[0]
                is_max = cls.startswith("Max")
                if cfg["padding"] == "same":
                    pads = []
                    for dim, kk, ss in zip(t.shape[2:], size, st):
                        _, p0, p1 = _same_pads(dim, kk, ss)
                        pads.append((p0, p1))
                    if is_max:
                        t = F.pad(t, (*pads[2], *pads[1], *pads[0]), value=float("-inf"))
                        y = F.max_pool3d(t, size, st)
                    else:
                        ones = torch.ones((1, 1, *t.shape[2:]), dtype=td)
                        tp = F.pad(t, (*pads[2], *pads[1], *pads[0]))
                        op = F.pad(ones, (*pads[2], *pads[1], *pads[0]))
                        y = F.avg_pool3d(tp, size, st) / F.avg_pool3d(op, size, st)
                else:
                    y = F.max_pool3d(t, size, st) if is_max else F.avg_pool3d(t, size, st)
            elif cls == "GlobalAveragePooling3D":
                y = a[0].mean(dim=(2, 3, 4))
            elif cls == "GlobalMaxPooling3D":
                y = a[0].amax(dim=(2, 3, 4))
            elif cls == "Flatten":
                y = a[0].permute(0, 2, 3, 4, 1).reshape(a[0].shape[0], -1)   # NDHWC row-major
            elif cls == "Concatenate":
                y = torch.cat(a, dim=1)
            elif cls == "Add":
                y = a[0]
                for t in a[1:]:
                    y = y + t
            else:
                raise NotImplementedError(f"oracle(torch): layer class {cls}")
            vals[name] = y
            last = name
        out = vals[_output_name(model_config, last)]
        return out.to(torch.float32).numpy() if out.dtype != torch.float64 else out.numpy()


def fp16_argmax(probs: np.ndarray) -> np.ndarray:
    """argmax the way the reference does: cast to float16 first (design_utils/utils.py:768,
    predict.py:163), then np.argmax with first-index tie-break (utils.py:659)."""
    return np.argmax(np.asarray(probs).astype(np.float16), axis=1)


def near_tie_rows(probs: np.ndarray, ulps: int = 1) -> np.ndarray:
    """Rows whose top-2 fp16 values are equal or within ``ulps`` fp16 ulps: a 1e-4 deviation
    can flip these, so bit-exact argmax is only meaningful outside this set."""
    p16 = np.asarray(probs).astype(np.float16)
    srt = np.sort(p16, axis=1)
    top, second = srt[:, -1], srt[:, -2]
    gap = top.view(np.uint16).astype(np.int32) - second.view(np.uint16).astype(np.int32)
    return gap <= ulps
