#!/bin/bash
# Quick GPU check: parity tests + the default bench line (+ any extra bench.py argument sets given as "a b c" strings).
mkdir -p gpurun_out
TAG=${TAG:-q}
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${TAG}_pytest.log | cut -c1-220
grep -E "Error|error:" gpurun_out/${TAG}_pytest.log | sort | uniq -c | head -5 | cut -c1-300
i=0
for args in "" "$@"; do
  timeout 400 python bench.py $args > gpurun_out/${TAG}_bench_$i.json 2>> gpurun_out/${TAG}_err.log
  echo "== bench.py $args"
  python - gpurun_out/${TAG}_bench_$i.json <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=l.get('roofline') or {}
    print(round(l['value'],1), l['unit'], 'ms/step', round(l['ms_per_step'],3), 'e2e', round((l.get('e2e') or {}).get('value') or 0,1), 'frac', round(r.get('frac') or 0,4), 'wg', round((r.get('whole_graph') or {}).get('frac') or 0,4), 'clk', (l.get('clocks') or {}).get('sm_mhz'), 'launches', l.get('gpu_launches'))
    po=r.get('per_op_ms') or {}
    if len(po) < 30: print({k.split(':')[1]: v for k,v in po.items()})
except Exception as e: print('ERR', e)
PY
  i=$((i+1))
done
tail -5 gpurun_out/${TAG}_err.log 2>/dev/null
