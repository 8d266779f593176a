"""Shared helpers for the parity tests (tests may import oracle/, the product may not)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from timed_design_b200 import _lib
from timed_design_b200.keras_graph import ACT_CODES, OP_CONV3D


def conv_desc(kernel, bias=None, scale=None, shift=None, padding="same", act1=None, act2=None):
    """tb_op_desc for timed_b200_conv3d_fwd + keep-alive list."""
    d = _lib.tb_op_desc()
    keep = []
    d.op = OP_CONV3D
    kd, kh, kw, ci, co = kernel.shape
    d.kernel[0], d.kernel[1], d.kernel[2] = kd, kh, kw
    d.stride[0] = d.stride[1] = d.stride[2] = 1
    d.pad_same = 1 if padding == "same" else 0
    d.c_out = co
    d.act1, d.act2 = ACT_CODES[act1], ACT_CODES[act2]
    d.alpha1 = d.alpha2 = 1.0
    for name, arr in (("kernel_w", kernel), ("bias", bias), ("scale", scale), ("shift", shift)):
        if arr is not None:
            a = np.ascontiguousarray(arr, dtype=np.float32)
            keep.append(a)
            setattr(d, name, _lib.fptr(a))
    return d, keep


def run_conv_gpu(x, kernel, bias=None, scale=None, shift=None, padding="same", act1=None, act2=None):
    """y = act2(scale*act1(conv3d(x)+bias)+shift) through the C ABI on cuda:0 (torch = container)."""
    import torch
    lib = _lib.load()
    n, D, H, W, ci = x.shape
    kd, kh, kw, _, co = kernel.shape
    if padding == "same":
        od, oh, ow = D, H, W
    else:
        od, oh, ow = D - kd + 1, H - kh + 1, W - kw + 1
    dx = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
    dy = torch.full((n, od, oh, ow, co), float("nan"), dtype=torch.float32, device="cuda")
    d, keep = conv_desc(kernel, bias, scale, shift, padding, act1, act2)
    _lib.check(lib.timed_b200_conv3d_fwd(C.c_void_p(dx.data_ptr()), n, D, H, W, ci, C.byref(d), 0,
                                         C.c_void_p(dy.data_ptr())))
    torch.cuda.synchronize()
    return dy.cpu().numpy()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
