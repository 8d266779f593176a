#!/bin/bash
# Round-2 GPU call B: rerun parity tests after the config-score fix, accumulation-error model, default bench lines.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
grep -E "^FAILED|passed|failed" gpurun_out/b_pytest.log | cut -c1-200
timeout 600 python tools/accum_error.py --out gpurun_out/b_accum_error.jsonl 2> gpurun_out/b_err.log | cut -c1-500
timeout 300 python tools/layer_error.py 8 2>> gpurun_out/b_err.log | tee gpurun_out/b_layer_error.jsonl | cut -c1-300
timeout 300 python bench.py > gpurun_out/b_bench_timed20.json 2>> gpurun_out/b_err.log
timeout 300 python bench.py --config timed338 > gpurun_out/b_bench_timed338.json 2>> gpurun_out/b_err.log
for f in gpurun_out/b_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
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
tail -20 gpurun_out/b_err.log
