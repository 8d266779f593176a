run() { python bench.py --no-e2e --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); po=l['roofline']['per_op_ms']; print(round(l['value'],0), {k.split(':')[1]: round(v,2) for k,v in po.items() if 'conv' in k}, l['clocks']['sm_mhz'])"; }
for i in 1 2; do
echo base; TIMED_B200_LIB=$PWD/timed_design_b200/libtimed_b200_base.so run
echo "new issuers=1"; TIMED_B200_THINZ_ISSUERS=1 run
echo "new issuers=2"; run
done
