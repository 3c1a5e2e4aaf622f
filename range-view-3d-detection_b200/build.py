"""Build librv3d.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

    python range-view-3d-detection_b200/build.py [--force]

Output: range-view-3d-detection_b200/rv3d/_lib/librv3d.so (git-ignored, travels to the GPU box).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
INCLUDE = HERE.parent / "include"
OUT_DIR = HERE / "rv3d" / "_lib"
OBJ_DIR = HERE / "build"
LIB = OUT_DIR / "librv3d.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", str(INCLUDE), "-I", str(CSRC),
    # Every float op on this path must round exactly once so results are bit-identical to the
    # CPU oracle (gcc -ffp-contract=off): no FMA contraction, IEEE div/sqrt, no flush-to-zero.
    "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
    *os.environ.get("RV3D_NVCC_DEFS", "").split(),      # experiments only, e.g. -DRV3D_NMS_THREADS=1024
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def _stale(out: Path, deps) -> bool:
    return not out.exists() or any(out.stat().st_mtime < d.stat().st_mtime for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    objs, jobs = [], []
    for src in sources():
        obj = OBJ_DIR / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, *headers]):
            cmd = [NVCC, *COMMON, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for r in ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed: " + " ".join(r.args))
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
               "-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
