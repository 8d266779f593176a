#!/bin/bash
# sampler parity tests + bench lines (20 and 338 categories)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_device_post_gpu.py tests/test_cli_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -5
for a in "--config sampler" "--config sampler --classes 338 --no-cpu-baseline"; do
  timeout 300 python bench.py $a 2>>gpurun_out/j_err.log | tee -a gpurun_out/j_sampler.jsonl | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); print(round(l['value']/1e9,2),'G residues/s', round(l['ms_per_step'],2), 'ms/sweep; e2e', round(l['e2e']['value']/1e9,2), 'draw kernel ms', l['roofline']['kernel_ms'], 'GB/s', round(l['roofline']['achieved'],1), 'cpu', (l.get('cpu_baseline') or {}).get('value'))"
done
tail -3 gpurun_out/j_err.log
