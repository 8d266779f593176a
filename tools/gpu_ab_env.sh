#!/bin/bash
# A/B of an environment switch on the default bench: tools/gpu_ab_env.sh VAR [bench args]; prints frames/s and per-conv ms.
V=$1; shift
run() { python bench.py --no-e2e --no-cpu-baseline --steps 20 "$@" 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); po=l['roofline']['per_op_ms']; print(round(l['value'],0), {k.split(':')[1]: round(v,2) for k,v in po.items() if 'conv' in k}, l['clocks']['sm_mhz'])"; }
for i in 1 2; do echo "$V=1"; env $V=1 bash -c "$(declare -f run); run $*"; echo "default"; run "$@"; done
