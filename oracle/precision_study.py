"""ORACLE-side study (test infrastructure): which tensor-core operand scheme holds the
north-star tolerance (|dp| <= 1e-4 vs the fp32 CPU path, fp16-argmax identical)?

Emulates on CPU what each scheme would compute: operands rounded to the MMA input type,
products accumulated exactly (fp64) -- isolating operand-rounding error, which dominates --
on the TIMED stand-in.  Run:  python -m oracle.precision_study [n_frames]
Results are quoted in DESIGN.md ("Numerics").
"""
from __future__ import annotations

import sys

import numpy as np
import torch
import torch.nn.functional as F

from oracle import keras_oracle as ko
from timed_design_b200 import standins


def _round(t: torch.Tensor, kind: str) -> torch.Tensor:
    if kind == "bf16":
        return t.float().to(torch.bfloat16).double()
    if kind == "fp16":
        return t.float().to(torch.float16).double()
    if kind == "tf32":
        i = t.float().view(torch.int32)
        i = (i + 0x00001000) & ~0x00001FFF   # RN to 10-bit mantissa (ties away; fine here)
        return i.view(torch.float32).double()
    raise ValueError(kind)


def _split_conv(a, w, b, scheme):
    """a: NCDHW float64 (already padded); w: OIDHW float64."""
    if scheme == "fp32":
        return F.conv3d(a.float(), w.float(), None if b is None else b.float()).double()
    kind, terms = scheme.split("x")
    terms = int(terms)
    a_hi = _round(a, kind)
    w_hi = _round(w, kind)
    y = F.conv3d(a_hi, w_hi)
    if terms >= 3:
        a_lo = _round(a.float().double() - a_hi, kind)
        w_lo = _round(w.float().double() - w_hi, kind)
        y = y + F.conv3d(a_lo, w_hi) + F.conv3d(a_hi, w_lo)
        if terms >= 4:
            y = y + F.conv3d(a_lo, w_lo)
    if b is not None:
        y = y + b.view(1, -1, 1, 1, 1)
    return y


def forward_emulated(cfg, weights, X, scheme):
    """TIMED-shaped graphs only (Conv3D/ELU/BN/MaxPool/Dropout/GAP/Softmax)."""
    vals = {}
    last = None
    x0 = torch.from_numpy(np.asarray(X, dtype=np.float32)).double().permute(0, 4, 1, 2, 3)
    for name, cls, c, ins in ko._layers_in_order(cfg):
        a = [vals[i] for i in ins]
        if cls == "InputLayer":
            y = x0
        elif cls == "Conv3D":
            k = torch.from_numpy(ko._w(weights, name, "kernel:0")).double().permute(4, 3, 0, 1, 2)
            b = torch.from_numpy(ko._w(weights, name, "bias:0")).double() if c.get("use_bias", True) else None
            t = a[0]
            pads = []
            for dim, kk in zip(t.shape[2:], k.shape[2:]):
                _, p0, p1 = ko._same_pads(dim, kk, 1)
                pads.append((p0, p1))
            t = F.pad(t, (*pads[2], *pads[1], *pads[0]))
            y = _split_conv(t, k, b, scheme)
            # intermediate activations are kept in fp32 between layers
            y = y.float().double()
        elif cls == "ELU":
            y = F.elu(a[0].float()).double()
        elif cls == "BatchNormalization":
            n_c = a[0].shape[1]
            g = torch.from_numpy(ko._w(weights, name, "gamma:0")).double()
            be = torch.from_numpy(ko._w(weights, name, "beta:0")).double()
            mu = torch.from_numpy(ko._w(weights, name, "moving_mean:0")).double()
            var = torch.from_numpy(ko._w(weights, name, "moving_variance:0")).double()
            sc = (g / torch.sqrt(var + c.get("epsilon", 1e-3))).float()
            sh = (be - mu * sc.double()).float()
            shp = [1, n_c, 1, 1, 1]
            y = (a[0].float() * sc.view(shp) + sh.view(shp)).double()
        elif cls == "MaxPooling3D":
            t = a[0]
            pads = []
            for dim in t.shape[2:]:
                _, p0, p1 = ko._same_pads(dim, 2, 2)
                pads.append((p0, p1))
            t = F.pad(t, (*pads[2], *pads[1], *pads[0]), value=float("-inf"))
            y = F.max_pool3d(t, 2, 2)
        elif cls in ("SpatialDropout3D", "Dropout"):
            y = a[0]
        elif cls == "GlobalAveragePooling3D":
            y = a[0].mean(dim=(2, 3, 4))
        elif cls == "Softmax":
            y = F.softmax(a[0].float(), dim=-1).double()
        else:
            raise NotImplementedError(cls)
        vals[name] = y
        last = name
    return vals[last].numpy()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ncls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    cfg, w = standins.timed_standin(ncls)
    X = standins.synthetic_frames(n)
    ref64 = ko.forward_numpy(cfg, w, X, np.float64)
    ref32 = ko.forward_torch(cfg, w, X)
    print(f"frames={n} classes={ncls}")
    print(f"  torch-fp32 vs numpy-fp64 : max|dp| = {np.abs(ref32 - ref64).max():.3e}")
    for scheme in ("bf16x1", "fp16x1", "tf32x1", "bf16x3", "bf16x4", "fp16x3"):
        p = forward_emulated(cfg, w, X, scheme)
        dp = np.abs(p - ref64).max()
        flips = int((ko.fp16_argmax(p) != ko.fp16_argmax(ref64)).sum())
        print(f"  {scheme:7s} vs numpy-fp64 : max|dp| = {dp:.3e}   fp16-argmax flips = {flips}/{n}")


if __name__ == "__main__":
    main()
