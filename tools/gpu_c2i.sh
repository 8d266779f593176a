#!/bin/bash
# Fused col2im (kw in the GEMM epilogue) check: conv parity tests, DenseCPD parity, A/B bench lines, one ncu --set full of
# a growth conv + a 1x1 conv of DenseCPD at batch 512.  Outputs under gpurun_out/${TAG}_*.
mkdir -p gpurun_out
T=${TAG:-r2c}
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_pytest_conv.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${T}_pytest_conv.log | cut -c1-240
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "densecpd" > gpurun_out/${T}_pytest_dense.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${T}_pytest_dense.log | cut -c1-240
grep -E "Error|error:|assert" gpurun_out/${T}_pytest_*.log | sort | uniq -c | head -12 | cut -c1-300
timeout 300 python bench.py --config densecpd --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_densecpd_fused.json 2>> gpurun_out/${T}_err.log
TIMED_B200_NO_C2I_FUSE=1 timeout 300 python bench.py --config densecpd --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_densecpd_zmatrix.json 2>> gpurun_out/${T}_err.log
python - <<PY
import json
for n in ("fused", "zmatrix"):
    try:
        l = json.loads(open("gpurun_out/${T}_bench_densecpd_%s.json" % n).read().strip().splitlines()[-1]); r = l["roofline"]
        po, pk = r["per_op_ms"], r["per_op_kernel"]
        agg = {}
        for k, v in po.items(): agg[pk.get(k, "?")] = agg.get(pk.get(k, "?"), 0) + v
        print(n, round(l["value"], 1), l["unit"], "ms", round(l["ms_per_step"], 2), "chunk", l["config"].get("forward_chunk_frames"), "wg", round(r["whole_graph"]["frac"], 4), "clk", l["clocks"].get("sm_mhz"))
        print("   ", {k[:48]: round(v, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]})
        print("   ", {k: po[k] for k in list(po)[3:9]})
    except Exception as e: print(n, "ERR", e)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_umma" -s 115 -c 3 \
   -o gpurun_out/${T}_dense_full -f python bench.py --config densecpd --batch 512 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
python tools/ncu_extract.py gpurun_out/${T}_dense_full.ncu-rep gpurun_out/${T}_dense_ncu_kernels.csv; cut -c1-330 gpurun_out/${T}_dense_ncu_kernels.csv
tail -5 gpurun_out/${T}_err.log
