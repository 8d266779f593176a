#!/bin/bash
# Fused DenseNet pre-activation (bnrelu_conv1x1_kernel) check: parity tests, A/B bench lines.  Outputs: gpurun_out/${TAG}_*.
mkdir -p gpurun_out
T=${TAG:-r2x}
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "dense" > gpurun_out/${T}_pytest_dense.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${T}_pytest_dense.log | cut -c1-240
grep -E "Error|error:|assert|timed out" gpurun_out/${T}_pytest_dense.log | sort | uniq -c | head -12 | cut -c1-300
timeout 300 python bench.py --config densecpd --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_densecpd_fused.json 2>> gpurun_out/${T}_err.log
TIMED_B200_NO_XFORM=1 timeout 300 python bench.py --config densecpd --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_densecpd_noxform.json 2>> gpurun_out/${T}_err.log
python - <<PY
import json
for n in ("fused", "noxform"):
    try:
        l = json.loads(open("gpurun_out/${T}_bench_densecpd_%s.json" % n).read().strip().splitlines()[-1]); r = l["roofline"]
        po, pk = r["per_op_ms"], r["per_op_kernel"]
        agg = {}
        for k, v in po.items(): agg[pk.get(k, "?")] = agg.get(pk.get(k, "?"), 0) + v
        print(n, round(l["value"], 1), l["unit"], "ms", round(l["ms_per_step"], 2), "chunk", l["config"].get("forward_chunk_frames"), "wg", round(r["whole_graph"]["frac"], 4), "clk", l["clocks"].get("sm_mhz"), "launches", l.get("gpu_launches"))
        print("   ", {k[:40]: round(v, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]})
        print("   ", {k: po[k] for k in list(po)[2:12]})
    except Exception as e: print(n, "ERR", e)
PY
tail -5 gpurun_out/${T}_err.log
