"""Build libtimed_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m timed_design_b200.build [--force]
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libtimed_b200.so"
SOURCES = [CSRC / "api.cu"]
DEPS = sorted(CSRC.glob("*.cu*")) + [HERE.parent / "include" / "timed_b200.h"]
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-std=c++17"]


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    return any(d.stat().st_mtime > t for d in DEPS)


DEBUG_OUT = HERE / "libtimed_b200_dbg.so"


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> Path:
    """``debug=True`` builds libtimed_b200_dbg.so with -DTIMED_B200_DEBUG: the role-timing switches
    (TIMED_B200_DBG) exist only there; select it with TIMED_B200_LIB (tools/role_timing.sh)."""
    if debug:
        cmd = ["nvcc", *NVCC_FLAGS, "-DTIMED_B200_DEBUG", *map(str, SOURCES), "-lz", "-o", str(DEBUG_OUT)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
        return DEBUG_OUT
    if not force and not needs_build():
        return OUT
    cmd = ["nvcc", *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []),
           *map(str, SOURCES), "-lz", "-o", str(OUT)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
