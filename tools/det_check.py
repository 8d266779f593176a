"""Results must not depend on what the caller's workspace contains: forward on a workspace filled with 0xFF bytes
(bf16 / fp32 NaN patterns) and compare with the host path.  PYTHONPATH=. python tools/det_check.py [frames] [model]"""
import sys
import numpy as np, torch
from timed_design_b200 import standins
from timed_design_b200.model import Model
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
name = sys.argv[2] if len(sys.argv) > 2 else "timed20"
cfg, w = {"timed20": lambda: standins.timed_standin(20), "timed338": lambda: standins.timed_standin(338),
          "densecpd": standins.densecpd_standin, "prodconn": standins.prodconn_standin, "tiny": standins.tiny_standin}[name]()
m = Model(cfg, w)
side = 9 if name == "tiny" else 21
X = standins.synthetic_frames(n, side=side, seed=21)
a = m.predict(X)
d = torch.from_numpy(X).cuda()
probs = torch.empty(a.shape, dtype=torch.float32, device="cuda")
ws = torch.empty(m.workspace_bytes(n), dtype=torch.uint8, device="cuda")
for fill in (255, 0, 127):
    ws.fill_(fill)
    m.forward_device(d, probs, ws, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
    p = probs.cpu().numpy()
    print(f"{name} n={n} ws filled with {fill:#x}: max|dev - host| =", np.abs(p - a).max(), "nan rows", int(np.isnan(p).any(1).sum()))
