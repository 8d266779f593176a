"""Isolate tensor-core ACCUMULATION error: operands exactly representable in bf16 (the hi/lo split is exact, lo planes
are zero) through convs of growing K, against the fp64 oracle.  Reports, per shape, the rms relative error, the
SYSTEMATIC part (least-squares slope of err on ref: the accumulator's relative shrink, split by the sign of the result
to tell truncation toward zero from truncation toward -inf) and the rms that remains once the slope is removed.
Run under gpurun:  python tools/accum_error.py [--out file]"""
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


def torch_ref64(x, w, padding):
    import torch.nn.functional as F
    t = torch.from_numpy(x.astype(np.float64)).permute(0, 4, 1, 2, 3)
    k = torch.from_numpy(w.astype(np.float64)).permute(4, 3, 0, 1, 2)
    if padding == "same":
        p = [(kk - 1) // 2 for kk in w.shape[:3]]
        t = F.pad(t, (p[2], w.shape[2] - 1 - p[2], p[1], w.shape[1] - 1 - p[1], p[0], w.shape[0] - 1 - p[0]))
    return F.conv3d(t, k).permute(0, 2, 3, 4, 1).numpy()


out = [a for a in sys.argv[1:] if not a.startswith("--")]
rng = np.random.default_rng(0)
rows = []
# (cin, cout, k, frames, positive inputs?): K = k^3 * cin; 256 -> 512 at 6^3 is TIMED block 5
for ci, co, k, n, positive in ((32, 64, 3, 4, False), (64, 128, 3, 4, False), (128, 256, 3, 4, False), (256, 512, 3, 4, False),
                               (512, 256, 3, 2, False), (256, 512, 3, 700, False), (256, 512, 3, 4, True), (512, 192, 1, 8, False)):
    x = rng.standard_normal((n, 6, 6, 6, ci)).astype(np.float32)
    if positive:
        x = np.abs(x)
    w = (rng.standard_normal((k, k, k, ci, co)) * np.sqrt(2.0 / (k ** 3 * ci))).astype(np.float32)
    xx, ww = bf16_exact(x), bf16_exact(w)
    y = run_conv_gpu(xx, ww, padding="same").astype(np.float64)
    sel = slice(0, 4)
    ref = torch_ref64(xx[sel], ww, "same")
    y = y[sel]
    err = y - ref
    rms = np.sqrt((ref ** 2).mean())
    slope = float((err * ref).sum() / (ref * ref).sum())
    pos, neg = ref > 0, ref < 0
    slope_pos = float((err[pos] * ref[pos]).sum() / (ref[pos] ** 2).sum())
    slope_neg = float((err[neg] * ref[neg]).sum() / (ref[neg] ** 2).sum())
    resid = err - slope * ref
    n_mma = k ** 3 * ci // 16
    row = {"cin": ci, "cout": co, "k": k, "K": k ** 3 * ci, "frames": n, "positive_inputs": positive, "mma_per_accumulator": n_mma,
           "err_rms_rel": float(np.sqrt((err ** 2).mean()) / rms), "slope": slope, "slope_per_mma": slope / n_mma,
           "slope_pos": slope_pos, "slope_neg": slope_neg, "mean_err_over_rms": float(err.mean() / rms),
           "resid_rms_rel": float(np.sqrt((resid ** 2).mean()) / rms)}
    rows.append(row)
    print(json.dumps(row), flush=True)
if "--out" in sys.argv:
    Path(sys.argv[sys.argv.index("--out") + 1]).write_text("\n".join(json.dumps(r) for r in rows) + "\n")
