"""End-to-end frames/s of the two real-input routes of predict.py on one GPU (run under gpurun):
  (a) structure files -> GPU voxeliser -> network (no dataset file, frames never leave the device);
  (b) gzip .hdf5 frame dataset -> load_batch (native inflater on the host threads) -> Model.predict.
    python tools/bench_structures.py [n_structures]"""
import json
import shutil
import sys
import tempfile
import time
import warnings
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from timed_design_b200 import frames, predict, standins, voxelise  # noqa: E402
from timed_design_b200.model import Model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src = Path(__file__).resolve().parents[1] / "tests" / "golden" / "1ubq.pdb1.gz"
tmp = Path(tempfile.mkdtemp())
files = []
for i in range(n):
    f = tmp / f"s{i:04d}.pdb1.gz"
    shutil.copy(src, f)
    files.append(f)
cfg, w = standins.timed_standin(20, c_in=5)
m = Model(cfg, w, max_chunk_frames=4096)
warnings.simplefilter("ignore")
# warm-up
fr, flat = voxelise.voxelise_structures(files[:2], "CNOCBCA", return_device=True)
predict._forward_device_rows(m, fr)
t0 = time.perf_counter()
fr, flat = voxelise.voxelise_structures(files, "CNOCBCA", return_device=True)
t1 = time.perf_counter()
p = predict._forward_device_rows(m, fr)
t2 = time.perf_counter()
out = {"structures": n, "frames": int(fr.shape[0]),
       "structure_route": {"voxelise_s": t1 - t0, "network_s": t2 - t1, "frames_per_s": fr.shape[0] / (t2 - t0),
                           "note": "timed_b200_pdb_parse (host threads) + timed_b200_voxelise + graph_forward, frames stay on the device"}}
t0 = time.perf_counter()
states = voxelise.load_states(files, "CNOCBCA")
out["structure_route"]["parse_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
parsed = [voxelise.fast_tables(f, "CNOCBCA", 1.0) for f in files]
out["structure_route"]["python_parser_s"] = time.perf_counter() - t0
# (b) the dataset route on the same frames
data = voxelise.make_frame_dataset(files, tmp, "data", codec="CNOCBCA")
flat2, _ = frames.create_flat_dataset_map(data)
t0 = time.perf_counter()
X, y = frames.load_batch(data, flat2)
t1 = time.perf_counter()
p2 = m.predict(X, batch_size=4096)
t2 = time.perf_counter()
out["hdf5_route"] = {"load_batch_s": t1 - t0, "predict_s": t2 - t1, "frames_per_s": len(flat2) / (t2 - t0),
                     "file_MB": data.stat().st_size / 1e6, "note": "gzip float32 frames, native inflater on the host threads"}
out["max_abs_diff_between_routes"] = float(np.abs(p - p2).max())
# (c) the same file with the STORED chunks inflated on the device (frames.load_batch_device, csrc/inflate.cuh)
import torch  # noqa: E402
frames.load_batch_device(data, flat2[:64])                       # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
dev = frames.load_batch_device(data, flat2)
torch.cuda.synchronize()
t1 = time.perf_counter()
if dev is not None:
    p3 = predict._forward_device_rows(m, dev[0])
    t2 = time.perf_counter()
    X3 = dev[0].cpu().numpy()
    out["hdf5_device_inflate_route"] = {"load_s": t1 - t0, "predict_s": t2 - t1, "frames_per_s": len(flat2) / (t2 - t0),
                                        "frames_equal_host_inflate": bool(np.array_equal(X3, X)) and bool(np.array_equal(dev[1], y)),
                                        "max_abs_diff_vs_host_route": float(np.abs(p3 - p2).max()),
                                        "note": "stored chunks H2D + one warp per chunk inflating on the device; load_s is the Python object walk + copy + inflate"}

print(json.dumps(out))
shutil.rmtree(tmp)
