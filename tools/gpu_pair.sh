#!/bin/bash
# pair-mode bring-up: parity tests first (bounded), then A/B bench lines
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -m gpu ) > gpurun_out/pair_pytest_conv.log 2>&1
tail -15 gpurun_out/pair_pytest_conv.log
if ! grep -q " passed" gpurun_out/pair_pytest_conv.log || grep -q "failed" gpurun_out/pair_pytest_conv.log; then echo "CONV TESTS FAILED"; fi
( time timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_conv_gpu.py ) > gpurun_out/pair_pytest_rest.log 2>&1
tail -5 gpurun_out/pair_pytest_rest.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
summ() { python -c "
import sys, json
l=json.loads(sys.stdin.readlines()[-1]); print(round(l['value']), round(l['ms_per_step'],2), l['clocks'].get('sm_mhz'), {k.split(':')[1]: round(v,2) for k,v in l['roofline']['per_op_ms'].items()})"; }
echo "== pair kc64";            $B 2>&1 | summ
echo "== no pair";   TIMED_B200_NO_PAIR=1 $B 2>&1 | summ
echo "== pair kc32"; TIMED_B200_PAIR_KC=32 $B 2>&1 | summ
echo "== pair kc64 again";      $B 2>&1 | summ
for d in 1 2 4 3 5 6 7; do echo "== DBG=$d"; TIMED_B200_DBG=$d $B 2>&1 | summ; done
