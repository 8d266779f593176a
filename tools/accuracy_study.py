"""GPU vs oracle accuracy of the stand-in graphs at their BENCHMARK sizes (run under gpurun).

    python tools/accuracy_study.py --model timed20 --frames 512 [--gain 16] [--fast-accum] [--out profiles/x.json]

Models: timed20, timed338, densecpd, prodconn (all at 21^3 x 6).  --gain sets the logit gain of the last BatchNorm of the
TIMED stand-ins (default 8; probability errors of any finite-precision evaluation scale with it).
The oracle is the torch-CPU fp32 restatement (what the 1e-4 contract is stated against) plus its fp64 evaluation.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import keras_oracle as ko  # noqa: E402
from timed_design_b200 import standins  # noqa: E402
from timed_design_b200.model import Model  # noqa: E402


def build(model: str, gain: float):
    if model == "timed20":
        return standins.timed_standin(20, logit_gain=gain)
    if model == "timed338":
        return standins.timed_standin(338, seed=8, logit_gain=gain)
    if model == "densecpd":
        return standins.densecpd_standin(20)
    if model == "prodconn":
        return standins.prodconn_standin(20)
    raise SystemExit(f"unknown model {model}")


def study(model: str, n: int, gain: float, fast: bool, fp64: bool = True) -> dict:
    cfg, w = build(model, gain)
    X = standins.synthetic_frames(n, seed=77)
    m = Model(cfg, w, precise=not fast)
    p = m.predict(X, batch_size=4096)
    t0 = time.perf_counter()
    ref32 = ko.forward_torch(cfg, w, X)
    t_cpu = time.perf_counter() - t0
    d32 = np.abs(p - ref32).max(1)
    out = {"model": model, "frames": n, "classes": int(p.shape[1]), "logit_gain": gain,
           "accumulation": "corrections in the main accumulator (A/B switch)" if fast else "separate correction accumulator (default)",
           "gpu_vs_fp32": {"max": float(d32.max()), "p99": float(np.percentile(d32, 99)), "median": float(np.median(d32)),
                           "n_over_1e-4": int((d32 > 1e-4).sum()), "n_over_5e-5": int((d32 > 5e-5).sum())},
           "argmax_flips_vs_fp32": int((ko.fp16_argmax(p) != ko.fp16_argmax(ref32)).sum()),
           "argmax_flips_outside_near_ties": int(((ko.fp16_argmax(p) != ko.fp16_argmax(ref32)) & ~ko.near_tie_rows(ref32)).sum()),
           "near_ties": int(ko.near_tie_rows(ref32).sum()),
           "pmax_median": float(np.median(ref32.max(1))), "distinct_argmax": int(len(set(ko.fp16_argmax(ref32)))),
           "cpu_oracle_seconds": round(t_cpu, 2),
           "kernels": sorted({m.op_kernel(i, n) for i, op in enumerate(m.graph.ops) if op.kind == 1})}
    if fp64:
        ref64 = ko.forward_torch(cfg, w, X, dtype="float64")
        d64 = np.abs(p - ref64).max(1)
        r = np.abs(ref32 - ref64).max(1)
        out["gpu_vs_fp64"] = {"max": float(d64.max()), "p99": float(np.percentile(d64, 99)), "median": float(np.median(d64))}
        out["fp32_vs_fp64"] = {"max": float(r.max()), "median": float(np.median(r))}
    m.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="timed20")
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--gain", type=float, default=8.0)
    ap.add_argument("--fast-accum", action="store_true")
    ap.add_argument("--no-fp64", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = study(a.model, a.frames, a.gain, a.fast_accum, not a.no_fp64)
    line = json.dumps(res)
    print(line, flush=True)
    if a.out:
        with open(a.out, "a") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
