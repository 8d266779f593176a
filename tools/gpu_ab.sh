#!/bin/bash
# A/B of two builds of the library on the default bench: tools/gpu_ab.sh [lib_a lib_b]; prints frames/s and per-conv ms.
A=${1:-libtimed_b200_base.so}; B=${2:-libtimed_b200.so}
for i in 1 2; do for lib in $A $B; do echo "$lib"; TIMED_B200_LIB=$PWD/timed_design_b200/$lib timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); po=l['roofline']['per_op_ms']; print(round(l['value'],0), {k.split(':')[1]: round(v,2) for k,v in po.items() if 'conv' in k}, l['clocks']['sm_mhz'])"; done; done
