#!/bin/bash
# Per-role timing of the conv kernels: runs bench.py with TMA / MMA / epilogue disabled in turn
# (TIMED_B200_DBG) and prints per-op milliseconds.  Output values are garbage by construction.
# The switches exist only in the bring-up build: python -m timed_design_b200.build --debug (run it before gpurun).
export TIMED_B200_LIB=$PWD/timed_design_b200/libtimed_b200_dbg.so
for d in 0 1 2 4 3 5 6 7; do
  echo "== TIMED_B200_DBG=$d"
  TIMED_B200_DBG=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
l=json.loads(sys.stdin.readline()); print(round(l['ms_per_step'],2), {k.split(':')[1]: round(v,2) for k,v in l['roofline']['per_op_ms'].items() if 'conv' in k})"
done
