"""Compact per-launch table from an .ncu-rep (raw page): duration, tensor-pipe activity, L2 / L1 / DRAM
throughput, DRAM bytes, achieved clocks.

    python tools/ncu_extract.py gpurun_out/x.ncu-rep [out.csv]
"""
import csv
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"),
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clock"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("launch__registers_per_thread", "regs"),
]

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
out = [[n + (f" [{units[hdr.index(c)]}]" if units[hdr.index(c)] else "") for c, n in COLS if c in hdr]]
for r in rows[2:]:
    out.append([r[hdr.index(c)][:48] for c, n in COLS if c in hdr])
w = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
w.writerows(out)
