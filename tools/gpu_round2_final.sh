#!/bin/bash
# Round-2 measurement set on one GPU: bench lines of every BASELINE config, the 1M-frame long run, accuracy at the
# benchmark sizes, the real-input routes.  Outputs under gpurun_out/ (copied to profiles/r2f_* by hand).
mkdir -p gpurun_out
T=${TAG:-r2f}
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/${T}_gpu.txt
timeout 400 python bench.py > gpurun_out/${T}_bench_timed20.json 2>> gpurun_out/${T}_err.log
timeout 400 python bench.py --frames 1000000 --no-cpu-baseline > gpurun_out/${T}_bench_timed20_1M.json 2>> gpurun_out/${T}_err.log
timeout 400 python bench.py --config timed338 > gpurun_out/${T}_bench_timed338.json 2>> gpurun_out/${T}_err.log
timeout 600 python bench.py --config densecpd --steps 3 > gpurun_out/${T}_bench_densecpd.json 2>> gpurun_out/${T}_err.log
timeout 400 python bench.py --config sampler > gpurun_out/${T}_bench_sampler20.json 2>> gpurun_out/${T}_err.log
timeout 400 python bench.py --config sampler --classes 338 > gpurun_out/${T}_bench_sampler338.json 2>> gpurun_out/${T}_err.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_err.log
timeout 400 python tools/bench_structures.py 64 > gpurun_out/${T}_structures.json 2>> gpurun_out/${T}_err.log
A=gpurun_out/${T}_accuracy.jsonl; rm -f $A
for g in 8 16 32; do timeout 300 python tools/accuracy_study.py --model timed20 --frames 512 --gain $g --out $A > /dev/null 2>> gpurun_out/${T}_err.log; done
for g in 8 16; do timeout 300 python tools/accuracy_study.py --model timed338 --frames 256 --gain $g --out $A > /dev/null 2>> gpurun_out/${T}_err.log; done
timeout 300 python tools/accuracy_study.py --model densecpd --frames 16 --out $A > /dev/null 2>> gpurun_out/${T}_err.log
timeout 300 python tools/accuracy_study.py --model prodconn --frames 16 --out $A > /dev/null 2>> gpurun_out/${T}_err.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1]); r=l.get("roofline") or {}
        print(f.split("bench_")[1], round(l["value"],1), l["unit"], "ms", round(l["ms_per_step"],3), "e2e", round((l.get("e2e") or {}).get("value") or 0,1), "frac", round(r.get("frac") or 0,4), "wg", round((r.get("whole_graph") or {}).get("frac") or 0,4), "clk", (l.get("clocks") or {}).get("sm_mhz"), "cpu", round((l.get("cpu_baseline") or {}).get("value") or 0,1))
    except Exception as e: print(f, "ERR", e)
print(open("gpurun_out/${T}_structures.json").read()[:700])
for l in open("$A"):
    r=json.loads(l); print(r["model"], r["logit_gain"], r["gpu_vs_fp32"]["max"], "flips", r["argmax_flips_outside_near_ties"])
PY
tail -5 gpurun_out/${T}_err.log
