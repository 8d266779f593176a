"""Per-layer error of the GPU conv (fused bias+ELU+BN) against the fp64 oracle, each layer fed the
oracle's own fp32 input activation, so errors do not compound.  Run under gpurun."""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import keras_oracle as ko  # noqa: E402
from tests.helpers import run_conv_gpu  # noqa: E402
from timed_design_b200 import standins  # noqa: E402
from timed_design_b200.keras_graph import OP_CONV3D, parse_model_config  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg, w = standins.timed_standin(20)
X = standins.synthetic_frames(n, seed=77)
_, vals = ko.forward_numpy(cfg, w, X, np.float64, return_all=True)
g = parse_model_config(cfg, w)
names = [l["name"] for l in cfg["config"]["layers"]]
for op in g.ops:
    if op.kind != OP_CONV3D:
        continue
    idx = names.index(op.name)
    src = cfg["config"]["layers"][idx]["inbound_nodes"][0][0][0]
    x = vals[src].astype(np.float32)
    # the fused op ends at the BatchNormalization that follows conv -> elu -> bn
    bn_name = names[idx + 2]
    ref = vals[bn_name]
    y = run_conv_gpu(x, op.kernel_w, bias=op.bias, scale=op.scale, shift=op.shift, padding="same", act1="elu")
    err = y.astype(np.float64) - ref
    rms = float(np.sqrt((ref ** 2).mean()))
    print(json.dumps({"layer": op.name, "K": int(np.prod(op.kernel_w.shape[:4])), "N": int(op.c_out),
                      "ref_rms": rms, "err_rms_rel": float(np.sqrt((err ** 2).mean()) / rms),
                      "err_max_rel_to_rms": float(np.abs(err).max() / rms),
                      "mean_signed_err_rel": float((err * np.sign(ref)).mean() / rms)}))
