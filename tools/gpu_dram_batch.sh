for b in 256 1024; do echo "batch $b"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_pair -s 12 -c 3 --csv python bench.py --batch $b --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
cur={}
for r in rows:
    k=(r[0], r[4][:40]); cur.setdefault(k,{})[r[-3]]=r[-1]+' '+r[-2]
for k,v in cur.items(): print(k[1], v)
"; done
