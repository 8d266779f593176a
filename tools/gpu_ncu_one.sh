#!/bin/bash
# ncu --set full of a few launches of one kernel in a DenseCPD forward at batch 512: tools/gpu_ncu_one.sh tag regex [skip] [count]
tag=$1; rx=$2; skip=${3:-40}; cnt=${4:-2}
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt \
   -o gpurun_out/${tag}_full -f python bench.py --config densecpd --batch 512 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
python tools/ncu_extract.py gpurun_out/${tag}_full.ncu-rep gpurun_out/${tag}_kernels.csv; cut -c1-330 gpurun_out/${tag}_kernels.csv
tail -3 gpurun_out/${tag}_ncu.log
