#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "slab or k3_c32_n64_11 or valid_c16" ) > gpurun_out/slab_pytest1.log 2>&1
tail -30 gpurun_out/slab_pytest1.log
