"""Isolate tensor-core ACCUMULATION error: operands that are exactly representable in bf16 (so the
hi/lo split is exact and lo planes are zero) through convs of growing K, against the fp64 oracle.
Run under gpurun."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import keras_oracle as ko  # noqa: E402
from tests.helpers import run_conv_gpu  # noqa: E402


def bf16_exact(a):
    return torch.from_numpy(a.astype(np.float32)).to(torch.bfloat16).float().numpy()


rng = np.random.default_rng(0)
for ci, co, k in ((64, 128, 3), (256, 512, 3), (512, 32, 3), (512, 32, 1), (128, 256, 3)):
    x = rng.standard_normal((2, 6, 6, 6, ci)).astype(np.float32)
    w = (rng.standard_normal((k, k, k, ci, co)) * np.sqrt(2.0 / (k ** 3 * ci))).astype(np.float32)
    for exact in (True, False):
        xx, ww = (bf16_exact(x), bf16_exact(w)) if exact else (x, w)
        y = run_conv_gpu(xx, ww, padding="same")
        ref = ko.np_conv3d(xx.astype(np.float64), ww.astype(np.float64), None, "same")
        err = y - ref
        rms = np.sqrt((ref ** 2).mean())
        print(json.dumps({"cin": ci, "cout": co, "k": k, "K": k ** 3 * ci, "bf16_exact_operands": exact,
                          "err_rms_rel": float(np.sqrt((err ** 2).mean()) / rms),
                          "mean_signed_rel": float((err * np.sign(ref)).mean() / rms),
                          "max_rel_to_rms": float(np.abs(err).max() / rms)}))
