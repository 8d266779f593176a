"""GPU bring-up probe: runs single-conv cases through the C ABI, each in its own subprocess
(a trap or hang in one case must not take the others down), compares with the numpy oracle
and writes one JSON line per case to gpurun_out/probe.jsonl.

    python tools/gpu_probe.py            # all cases
    python tools/gpu_probe.py --case 3   # one case, in-process (used by the parent)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

# (name, n, side, c_in, c_out, k, padding, onehot)
CASES = [
    ("k1_c64_n32", 2, 4, 64, 32, 1, "same", False),
    ("k1_c16_n32", 2, 4, 16, 32, 1, "same", False),
    ("k1_c32_n32", 2, 4, 32, 32, 1, "same", False),
    ("k1_c128_n64", 2, 4, 128, 64, 1, "same", False),
    ("k1_c6_n20", 3, 5, 6, 20, 1, "same", False),
    ("k3_c64_n32_onehot", 2, 6, 64, 32, 3, "same", True),
    ("k3_c64_n128", 2, 6, 64, 128, 3, "same", False),
    ("k3_c6_n32_21", 1, 21, 6, 32, 3, "same", False),
    ("k3_c32_n64_11", 2, 11, 32, 64, 3, "same", False),
    ("k3_c256_n512", 4, 6, 256, 512, 3, "same", False),
    ("k3_c512_n20", 4, 6, 512, 20, 3, "same", False),
    ("k3_c512_n338", 2, 6, 512, 338, 3, "same", False),
    ("k3_valid_c16", 2, 8, 16, 48, 3, "valid", False),
    ("k5_c6_n16", 1, 9, 6, 16, 5, "same", False),
    ("k7_c6_n16", 1, 9, 6, 16, 7, "same", False),
    ("k4_valid_dense", 5, 4, 64, 128, 4, "valid", False),
    ("k3_c128_n256_mt2", 700, 6, 128, 256, 3, "same", False),
    ("k3_c64_n128_mt2", 700, 6, 64, 128, 3, "same", False),
]


def run_case(idx: int) -> dict:
    from oracle import keras_oracle as ko
    from tests.helpers import run_conv_gpu
    name, n, side, ci, co, k, padding, onehot = CASES[idx]
    rng = np.random.default_rng(100 + idx)
    x = rng.standard_normal((n, side, side, side, ci)).astype(np.float32)
    if onehot:
        w = np.zeros((k, k, k, ci, co), dtype=np.float32)
        for o in range(co):
            t = o % (k * k * k)
            w[t // (k * k), (t // k) % k, t % k, (o * 7) % ci, o] = 1.0
    else:
        w = (rng.standard_normal((k, k, k, ci, co)) * np.sqrt(2.0 / (k ** 3 * ci))).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    t0 = time.time()
    y = run_conv_gpu(x, w, bias=b, padding=padding)
    dt = time.time() - t0
    nref = min(n, 4)
    ref = ko.np_conv3d(x[:nref].astype(np.float64), w.astype(np.float64), b.astype(np.float64), padding)
    err = np.abs(y[:nref] - ref)
    scale = float(np.abs(ref).max())
    out = {"case": name, "idx": idx, "max_abs_err": float(err.max()), "ref_max": scale,
           "rel_err": float(err.max() / scale), "nan": int(np.isnan(y).sum()), "seconds": round(dt, 3)}
    if n > nref:   # large-batch cases: the remaining frames against a torch conv
        import torch
        import torch.nn.functional as F
        xt = torch.from_numpy(x[nref:nref + 8]).permute(0, 4, 1, 2, 3).double()
        wt = torch.from_numpy(w).permute(4, 3, 0, 1, 2).double()
        p = (k - 1) // 2 if padding == "same" else 0
        r2 = F.conv3d(xt, wt, torch.from_numpy(b).double(), padding=p).permute(0, 2, 3, 4, 1).numpy()
        out["rel_err_mid"] = float(np.abs(y[nref:nref + 8] - r2).max() / np.abs(r2).max())
        xt = torch.from_numpy(x[-3:]).permute(0, 4, 1, 2, 3).double()
        r3 = F.conv3d(xt, wt, torch.from_numpy(b).double(), padding=p).permute(0, 2, 3, 4, 1).numpy()
        out["rel_err_tail"] = float(np.abs(y[-3:] - r3).max() / np.abs(r3).max())
    if out["rel_err"] > 1e-4 or out["nan"]:
        bad = np.argwhere(~(err <= 1e-4 * scale))
        out["n_bad"] = int(len(bad))
        out["first_bad"] = [[int(v) for v in row] for row in bad[:12]]
        out["bad_vals"] = [[float(y[tuple(row)]), float(ref[tuple(row)])] for row in bad[:12]]
        # which output channels / positions are wrong?
        out["bad_channels"] = sorted({int(r[4]) for r in bad})[:40]
        out["bad_w"] = sorted({int(r[3]) for r in bad})[:40]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=None)
    ap.add_argument("--only", type=str, default=None, help="comma-separated case indices")
    args = ap.parse_args()
    if args.case is not None:
        print("RESULT " + json.dumps(run_case(args.case)))
        return
    outdir = ROOT / "gpurun_out"
    outdir.mkdir(exist_ok=True)
    idxs = [int(v) for v in args.only.split(",")] if args.only else range(len(CASES))
    with open(outdir / "probe.jsonl", "w") as f:
        for i in idxs:
            try:
                res = subprocess.run([sys.executable, __file__, "--case", str(i)], capture_output=True,
                                     text=True, timeout=240)
                line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")]
                rec = json.loads(line[-1][7:]) if line else {
                    "case": CASES[i][0], "idx": i, "error": (res.stdout[-1500:] + res.stderr[-2500:])}
            except subprocess.TimeoutExpired:
                rec = {"case": CASES[i][0], "idx": i, "error": "timeout"}
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec)[:600])


if __name__ == "__main__":
    main()
