"""Where the time of frames.load_batch_device goes (run under gpurun): python tools/inflate_breakdown.py [n_structures]"""
import ctypes as C
import json
import shutil
import sys
import tempfile
import time
import warnings
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from timed_design_b200 import _lib, frames, voxelise  # noqa: E402

warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src = Path(__file__).resolve().parents[1] / "tests" / "golden" / "1ubq.pdb1.gz"
tmp = Path(tempfile.mkdtemp())
files = []
for i in range(n):
    f = tmp / f"s{i:04d}.pdb1.gz"
    shutil.copy(src, f)
    files.append(f)
data = voxelise.make_frame_dataset(files, tmp, "data", codec="CNOCBCA")
flat, _ = frames.create_flat_dataset_map(data)
fobj = frames._open(data)
t0 = time.perf_counter()
offs, sizes, objs = [], [], []
for row in flat:
    ds = fobj[str(row[0])][str(row[1])][str(row[2])]
    info = ds.chunk_table()
    offs.append(info[1][0][1]); sizes.append(info[1][0][2]); objs.append(ds)
t1 = time.perf_counter()
y = [ds.attrs["encoded_residue"] for ds in objs]
t2 = time.perf_counter()
offs, sizes = np.asarray(offs, np.int64), np.asarray(sizes, np.int64)
base = np.frombuffer(fobj.buf, dtype=np.uint8)
lo, hi = int(offs.min()), int((offs + sizes).max())
torch.cuda.synchronize()
t3 = time.perf_counter()
comp = torch.from_numpy(base[lo:hi]).cuda()
torch.cuda.synchronize()
t4 = time.perf_counter()
d_off, d_size = torch.from_numpy(offs - lo).cuda(), torch.from_numpy(sizes).cuda()
out = torch.empty((len(flat), 21, 21, 21, 5), dtype=torch.float32, device="cuda")
st = torch.empty(len(flat), dtype=torch.int32, device="cuda")
lib = _lib.load()
ms = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.timed_b200_inflate_device(C.c_void_p(comp.data_ptr()), len(flat), C.c_void_p(d_off.data_ptr()),
                                             C.c_void_p(d_size.data_ptr()), out[0].numel() * 4, C.c_void_p(out.data_ptr()),
                                             C.c_void_p(st.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print(json.dumps({"frames": len(flat), "chunk_table_walk_s": t1 - t0, "label_attrs_s": t2 - t1, "h2d_s": t4 - t3,
                  "h2d_MB": (hi - lo) / 1e6, "stored_MB": int(sizes.sum()) / 1e6, "inflate_kernel_ms": ms,
                  "inflated_GB": out.numel() * 4 / 1e9, "status_ok": bool((st == 0).all().item())}))
shutil.rmtree(tmp)
