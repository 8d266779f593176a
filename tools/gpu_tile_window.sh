for w in 0 148 296 74; do echo "window $w"; TIMED_B200_TILE_WINDOW=$w bash tools/gpu_dram.sh conv_pair 9 3 2>&1 | grep "conv_pair" | sed -n 2p | cut -c1-260; done
for w in 0 148 296; do echo "bench window $w"; TIMED_B200_TILE_WINDOW=$w python bench.py --no-e2e --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); po=l['roofline']['per_op_ms']; print(round(l['value'],0), {k.split(':')[1]: round(v,2) for k,v in po.items() if 'conv' in k}, l['clocks']['sm_mhz'])"; done
