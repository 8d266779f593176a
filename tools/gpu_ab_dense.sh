#!/bin/bash
# A/B of environment switches on the DenseCPD bench line: tools/gpu_ab_dense.sh "ENV1=.. ENV2=.." "ENV=.." ...  ("" = default)
mkdir -p gpurun_out
i=0
for envs in "" "$@"; do
  env $envs timeout 300 python bench.py --config densecpd --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_dense_$i.json 2>> gpurun_out/ab_dense_err.log
  python - gpurun_out/ab_dense_$i.json "$envs" <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = l["roofline"]
po, pk = r["per_op_ms"], r["per_op_kernel"]
agg = {}
for k, v in po.items(): agg[pk.get(k, "?")[:28]] = agg.get(pk.get(k, "?")[:28], 0) + v
print(repr(sys.argv[2]), round(l["value"], 1), "frames/s", round(l["ms_per_step"], 2), "ms clk", l["clocks"].get("sm_mhz"), {k: round(v, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:3]}, [po[k] for k in list(po)[3:12:4]])
PY
  i=$((i+1))
done
tail -3 gpurun_out/ab_dense_err.log
