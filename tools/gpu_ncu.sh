#!/bin/bash
# ncu full capture of one forward's kernels (a late step) + launch list. usage: tools/gpu_ncu.sh tag [regex] [bench args]
tag=$1; rx=${2:-"conv_|head_|pool3d|input_convert|cpv_"}; shift; shift
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/${tag}_ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 30 -c 10 \
   -o gpurun_out/${tag}_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
