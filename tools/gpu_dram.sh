#!/bin/bash
# DRAM bytes and duration per launch of the conv kernels (light ncu pass; a late forward): tools/gpu_dram.sh [regex]
rx=${1:-"conv_"}; skip=${2:-24}; cnt=${3:-6}
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
  -k regex:"$rx" -s $skip -c $cnt --csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
cur={}
for r in rows:
    k=(r[0], r[4][:40]); cur.setdefault(k,{})[r[-3]]=r[-1]+' '+r[-2]
for k,v in cur.items(): print(k[1], v)
"
