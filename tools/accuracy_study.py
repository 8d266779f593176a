"""GPU vs oracle accuracy over many frames of a stand-in (run under gpurun).
    python tools/accuracy_study.py [n_frames] [classes]"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import keras_oracle as ko  # noqa: E402
from timed_design_b200 import standins  # noqa: E402
from timed_design_b200.model import Model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ncls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg, w = standins.timed_standin(ncls)
X = standins.synthetic_frames(n, seed=77)
p = Model(cfg, w).predict(X)
ref32 = ko.forward_torch(cfg, w, X)
ref64 = ko.forward_torch(cfg, w, X, dtype="float64")
d32 = np.abs(p - ref32).max(1)
d64 = np.abs(p - ref64).max(1)
r3264 = np.abs(ref32 - ref64).max(1)
out = {"frames": n, "classes": ncls,
       "gpu_vs_fp32": {"max": float(d32.max()), "p99": float(np.percentile(d32, 99)), "median": float(np.median(d32)),
                       "n_over_1e-4": int((d32 > 1e-4).sum())},
       "gpu_vs_fp64": {"max": float(d64.max()), "p99": float(np.percentile(d64, 99)), "median": float(np.median(d64)),
                       "n_over_1e-4": int((d64 > 1e-4).sum())},
       "fp32_vs_fp64": {"max": float(r3264.max()), "median": float(np.median(r3264))},
       "argmax_flips_vs_fp32": int((ko.fp16_argmax(p) != ko.fp16_argmax(ref32)).sum()),
       "near_ties": int(ko.near_tie_rows(ref32).sum()),
       "worst_frames": [int(i) for i in np.argsort(-d64)[:5]],
       "worst_rows_pmax": [float(ref64[i].max()) for i in np.argsort(-d64)[:5]]}
print(json.dumps(out))
