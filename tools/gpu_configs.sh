#!/bin/bash
# All BASELINE.json configurations on one GPU: tests, default bench line (config 2), reference arm, 338-class (3),
# DenseCPD (4), sampler sweep (5), plumbing CLI run (1).  usage: tools/gpu_configs.sh tag
tag=$1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 2500 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>&1; tail -c 600 gpurun_out/${tag}_bench_ref.json
timeout 600 python bench.py --classes 338 --no-cpu-baseline > gpurun_out/${tag}_bench_338.json 2> gpurun_out/${tag}_bench_338.err; tail -c 1800 gpurun_out/${tag}_bench_338.json
timeout 900 python bench.py --model densecpd --batch 512 --steps 5 --no-cpu-baseline > gpurun_out/${tag}_bench_densecpd.json 2> gpurun_out/${tag}_bench_densecpd.err; tail -c 2500 gpurun_out/${tag}_bench_densecpd.json
timeout 600 python tools/bench_sampler.py --classes 20 > gpurun_out/${tag}_sampler20.json 2>&1; cat gpurun_out/${tag}_sampler20.json
timeout 600 python tools/bench_sampler.py --classes 338 > gpurun_out/${tag}_sampler338.json 2>&1; cat gpurun_out/${tag}_sampler338.json
