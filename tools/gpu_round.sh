#!/bin/bash
# One gpurun call: GPU tests, bench line, ncu launch list, ncu full capture of the conv kernels.
# usage: tools/gpu_round.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 3000 gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
# 5 forwards precede (Model warm-up is inside): capture one forward's worth of conv-side kernels of a late step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv|col2im' -s 28 -c 7 \
   -o gpurun_out/${tag}_conv -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out
