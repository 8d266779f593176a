"""Instruction evidence per kernel instantiation from the built library's SASS (no GPU needed):
    python tools/sass_evidence.py > profiles/r2_sass_evidence.md
Counts the mnemonics that prove the Blackwell paths: UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), UTMALDG (TMA tensor loads;
IM2COL = im2col mode), UBLKCP (1-D bulk copies), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), STG.*.256 (256-bit stores),
USETMAXREG (setmaxnreg), cache-hinted accesses (EF = evict-first)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "timed_design_b200" / "libtimed_b200.so"
KEEP = re.compile(r"^(UTCHMMA|UTMALDG|UTMASTG|FENCE\.VIEW\.ASYNC|UBLKCP|LDTM|UTCBAR|UTCATOMSWS|USETMAXREG|STG\.E[A-Z0-9.]*\.256|BAR\.SYNC|REDG|REDUX|DFMA|SHFL|WARPSYNC|SYNCS\.ARRIVE\.TRANS64\.RED)")
WANT = ("conv_umma_kernelILi2ELi0ELi1", "conv_pair_kernelILi2ELi0ELi1", "thin_conv_kernelILi0ELi0ELi0", "thinz_conv_kernelILi2ELi0ELi1ELi1",
        "slab_conv_kernelILi2ELi0ELi0", "sample_tiled_kernelILi32", "voxelise_kernelE", "head_col2im_pool_softmax_kernel",
        "bnrelu_conv1x1_kernelILi0ELi1ELi1", "conv_umma_kernelILi0ELi0ELi0", "inflate_streams_kernel"
        )

out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1) if any(w in m.group(1) for w in WANT) else None
        if cur:
            counts[cur] = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
        if m and KEEP.match(m.group(1)):
            counts[cur][m.group(1)] += 1
print("# SASS evidence, round 2 (cuobjdump -sass timed_design_b200/libtimed_b200.so; instruction counts per kernel instantiation; "
      "`python tools/sass_evidence.py`)\n")
print("`UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `UTMALDG.*` = cp.async.bulk.tensor (TMA; `IM2COL` = im2col mode, plain = tiled: "
      "the voxel-stationary boxes), `UTMASTG` = cp.async.bulk.tensor store (shared -> global), `FENCE.VIEW.ASYNC` = fence.proxy.async, `UBLKCP` = 1-D cp.async.bulk, `LDTM` = tcgen05.ld, `UTCBAR` = tcgen05.commit, `USETMAXREG` = "
      "setmaxnreg, `STG.E.EF*.256` = 256-bit store with the L2 evict-first hint.\n")
for k, c in counts.items():
    print(f"* `{k}`: " + ", ".join(f"{n} × {v}" for n, v in sorted(c.items())))
