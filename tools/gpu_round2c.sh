#!/bin/bash
# Round-2 GPU call C: accumulator-truncation compensation -- residual slope, accuracy across logit gains, tests.
mkdir -p gpurun_out
timeout 600 python tools/accum_error.py --out gpurun_out/c_accum_error.jsonl 2> gpurun_out/c_err.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['K'], r['frames'], 'err', '%.2e' % r['err_rms_rel'], 'slope', '%.2e' % r['slope'], 'resid', '%.2e' % r['resid_rms_rel'])"
TIMED_B200_NO_ACC_COMP=1 timeout 300 python tools/layer_error.py 8 2>> gpurun_out/c_err.log | cut -c1-260
timeout 300 python tools/layer_error.py 8 2>> gpurun_out/c_err.log | tee gpurun_out/c_layer_error.jsonl | cut -c1-260
A=gpurun_out/c_accuracy.jsonl; rm -f $A
for g in 8 16 32; do timeout 300 python tools/accuracy_study.py --model timed20 --frames 512 --gain $g --out $A > /dev/null 2>> gpurun_out/c_err.log; done
for g in 8 16 32; do timeout 300 python tools/accuracy_study.py --model timed338 --frames 256 --gain $g --out $A > /dev/null 2>> gpurun_out/c_err.log; done
timeout 300 python tools/accuracy_study.py --model densecpd --frames 16 --out $A > /dev/null 2>> gpurun_out/c_err.log
timeout 300 python tools/accuracy_study.py --model prodconn --frames 16 --out $A > /dev/null 2>> gpurun_out/c_err.log
python -c "
import json
for l in open('$A'):
    r = json.loads(l); print(r['model'], r['logit_gain'], r['gpu_vs_fp32'], 'fp64', r['gpu_vs_fp64']['max'], 'flips', r['argmax_flips_outside_near_ties'])"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c_pytest.log 2>&1
grep -E "^FAILED|passed|failed" gpurun_out/c_pytest.log | cut -c1-200
tail -5 gpurun_out/c_err.log
