#!/bin/bash
# Round-2 GPU call A: parity tests, accuracy study at benchmark sizes (default vs fast accumulation, logit gains), bench lines of all configs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
A=gpurun_out/a_accuracy.jsonl; rm -f $A
for g in 8 16 32; do timeout 300 python tools/accuracy_study.py --model timed20 --frames 512 --gain $g --out $A > /dev/null 2>> gpurun_out/a_err.log; done
for g in 8 16; do timeout 300 python tools/accuracy_study.py --model timed20 --frames 512 --gain $g --fast-accum --out $A > /dev/null 2>> gpurun_out/a_err.log; done
timeout 300 python tools/accuracy_study.py --model timed338 --frames 256 --out $A > /dev/null 2>> gpurun_out/a_err.log
timeout 300 python tools/accuracy_study.py --model timed338 --frames 256 --gain 16 --out $A > /dev/null 2>> gpurun_out/a_err.log
timeout 300 python tools/accuracy_study.py --model densecpd --frames 16 --out $A > /dev/null 2>> gpurun_out/a_err.log
timeout 300 python tools/accuracy_study.py --model prodconn --frames 16 --out $A > /dev/null 2>> gpurun_out/a_err.log
cat $A | cut -c1-420
timeout 300 python bench.py > gpurun_out/a_bench_timed20.json 2>> gpurun_out/a_err.log
timeout 300 python bench.py --fast-accum --no-cpu-baseline > gpurun_out/a_bench_timed20_fast.json 2>> gpurun_out/a_err.log
timeout 300 python bench.py --config timed338 > gpurun_out/a_bench_timed338.json 2>> gpurun_out/a_err.log
timeout 400 python bench.py --config densecpd --steps 3 > gpurun_out/a_bench_densecpd.json 2>> gpurun_out/a_err.log
timeout 300 python bench.py --config sampler > gpurun_out/a_bench_sampler.json 2>> gpurun_out/a_err.log
timeout 300 python bench.py --config sampler --classes 338 --no-cpu-baseline > gpurun_out/a_bench_sampler338.json 2>> gpurun_out/a_err.log
for f in gpurun_out/a_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=l.get('roofline') or {}
    print(round(l['value'],1), l['unit'], 'ms/step', round(l['ms_per_step'],3), 'e2e', (l.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'wg', (r.get('whole_graph') or {}).get('frac'), 'clk', (l.get('clocks') or {}).get('sm_mhz'))
    po=r.get('per_op_ms') or {}
    if len(po) < 30: print({k.split(':')[1]: v for k,v in po.items()})
except Exception as e: print('ERR', e)
PY
done
tail -20 gpurun_out/a_err.log
