"""Summarise an ncu capture of conv_umma_kernel: for launch <k>, list every mbarrier TRYWAIT spin loop,
UTCHMMA / UTMALDG / LDTM site with its execution count and sampling share, plus the stall-reason totals.

    python tools/ncu_roles.py gpurun_out/prof.ncu-rep 4
"""
import collections
import csv
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{int(kid) + 1}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    if r[isamp] == "# Samples":              # the CSV page repeats the listing: keep the first copy
        break
    data.append(r)
tot = sum(int(r[isamp]) for r in data)
print(f"kernel {kid}: {len(data)} SASS instructions, {tot} samples")
for k, r in enumerate(data):
    s = int(r[isamp])
    if any(t in r[isrc] for t in ("TRYWAIT", "UTCBAR", "LDTM", "UTMALDG.2D")) or \
            ("UTCHMMA" in r[isrc] and int(r[iex]) > 0) or ("IM2COL" in r[isrc] and int(r[iex]) > 0) or s > tot * 0.01:
        print(f"{k:6d} samples {s:8d} {100 * s / tot:5.1f}%  executed {int(r[iex]):>11d}  {r[isrc].strip()[:80]}")
st = collections.Counter()
for i, h in enumerate(hdr):
    if h.startswith("stall_") and "Not Issued" not in h:
        st[h] = sum(int(r[i]) for r in data if r[i].isdigit())
print(st.most_common(10))
