#!/bin/bash
# End-to-end (Model.predict from pinned host frames) against the pass-size policy: ramp (default) vs uniform passes.
run() { python bench.py --steps 5 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); e=l['e2e']; print('device', round(l['value']), 'e2e f32', round(e['value']), {k: round(v['value']) for k,v in e['by_host_dtype'].items()})"; }
for c in 1024 2048 4096; do echo "ramp, chunk $c"; run --e2e-chunk $c; done
for c in 256 512 768 1024; do echo "uniform $c"; TIMED_B200_NO_RAMP=1 run --e2e-chunk $c; done
