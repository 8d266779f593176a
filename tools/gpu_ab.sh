#!/bin/bash
# GPU parity tests, then A/B bench lines.  usage: tools/gpu_ab.sh tag "label|ENV=1 ENV2=x" "label2|" ...
tag=$1; shift
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/${tag}_pytest.log 2>&1
tail -6 gpurun_out/${tag}_pytest.log
summ() { python -c "
import sys, json
try:
    l=json.loads(sys.stdin.readlines()[-1]); print(round(l['value']), round(l['ms_per_step'],2), l['clocks'].get('sm_mhz'), {k.split(':')[1]: round(v,2) for k,v in l['roofline']['per_op_ms'].items()})
except Exception as e: print('FAILED', e)"; }
for v in "$@"; do
  label=${v%%|*}; envs=${v#*|}
  echo "== $label [$envs]"
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/${tag}_${label}.err | tee gpurun_out/${tag}_${label}.json | summ
done
