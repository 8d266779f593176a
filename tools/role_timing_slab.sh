#!/bin/bash
# Role timing of the slab / thin-z conv kernels (bring-up build): DBG 0 full, 1 no copies, 2 no MMAs, 4 no epilogue,
# 5 MMAs only, 6 copies only.  python -m timed_design_b200.build --debug first.
export TIMED_B200_LIB=$PWD/timed_design_b200/libtimed_b200_dbg.so
for d in ${DBGS:-0 5 2}; do
  echo "== DBG=$d"
  TIMED_B200_DBG=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
l=json.loads(sys.stdin.readline()); print(round(l['ms_per_step'],2), {k.split(':')[1]: round(v,2) for k,v in l['roofline']['per_op_ms'].items() if 'conv' in k})"
done
