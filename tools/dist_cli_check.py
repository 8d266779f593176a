"""Two-GPU check of the drop-in drivers (run under `gpurun --gpus 2`): predict.py and sample.py launched with torchrun
(NCCL) must leave exactly the files a single-GPU run leaves.

    python tools/dist_cli_check.py
"""
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from timed_design_b200 import standins  # noqa: E402
from timed_design_b200.hdf5 import write_frame_dataset, write_keras_h5  # noqa: E402


def run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-4000:])
        raise SystemExit(f"FAILED: {' '.join(map(str, cmd))}")


def main():
    d = Path(tempfile.mkdtemp())
    ubq = json.loads((ROOT / "tests" / "golden" / "1ubq_chainA.json").read_text())
    frames = standins.synthetic_frames(76, seed=0)
    res = {str(rid): (frames[i], ubq["labels"][i]) for i, rid in enumerate(ubq["residue_ids"])}
    resb = {str(i + 1): (frames[i + 5], ubq["labels"][i]) for i in range(31)}
    write_frame_dataset(d / "data.hdf5", {"1ubq": {"A": res, "B": resb}}, (21, 21, 21, 6))
    cfg, w = standins.timed_standin(20, filters=(8, 16, 16, 24, 32), calib_frames=4)
    write_keras_h5(d / "TIMED.h5", cfg, w)
    torchrun = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                "--master-addr", "127.0.0.1", "--master-port", "29533"]
    for name, launcher in (("one", [sys.executable]), ("two", torchrun)):
        wd = d / name
        wd.mkdir()
        run(launcher + [str(ROOT / "predict.py"), "--path_to_dataset", str(d / "data.hdf5"), "--path_to_model",
                        str(d / "TIMED.h5"), "--path_to_output", str(wd / "out"), "--path_to_datasetmap",
                        str(wd / "out" / "datasetmap.txt"), "--batch_size", "16", "--yes", "--binary_outputs"], wd)
        run(launcher + [str(ROOT / "sample.py"), "--path_to_pred_matrix", str(wd / "out" / "TIMED.csv"),
                        "--path_to_datasetmap", str(wd / "out" / "TIMED.txt"), "--sample_n", "37", "--temperature", "0.7",
                        "--seed", "11"], wd)
    ok = True
    for sub in ("out", "."):
        a = sorted(p.name for p in (d / "one" / sub).iterdir() if p.is_file())
        b = sorted(p.name for p in (d / "two" / sub).iterdir() if p.is_file())
        if a != b:
            print("file sets differ:", sub, a, b)
            ok = False
        for n in a:
            if n in b and (d / "one" / sub / n).read_bytes() != (d / "two" / sub / n).read_bytes():
                print("DIFFERENT:", sub, n)
                ok = False
    m = np.load(d / "two" / "out" / "TIMED.npy")
    print("files:", sorted(p.name for p in (d / "two" / "out").iterdir()), sorted(p.name for p in (d / "two").iterdir() if p.is_file()))
    print("matrix", m.shape, "OK" if ok else "MISMATCH")
    raise SystemExit(0 if ok else 1)


if __name__ == "__main__":
    main()
