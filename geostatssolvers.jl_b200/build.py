"""Builds libgskrige.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m gskrige.build        (or __graft_entry__.build())

One object per .cu, compiled in parallel; the shared library lands next to the sources
(geostatssolvers.jl_b200/csrc/libgskrige.so) so that it travels with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "libgskrige.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v", "-ccbin", "g++"]
# development builds only (scripts/dev): e.g. GSK_NVCC_EXTRA="-DGSK_DEV_TUNABLES"
FLAGS += os.environ.get("GSK_NVCC_EXTRA", "").split()


def _newer(src: Path, dst: Path, deps) -> bool:
    if not dst.exists():
        return True
    t = dst.stat().st_mtime
    return any(d.stat().st_mtime > t for d in [src, *deps])


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(CSRC.glob("*.cuh")) + [CSRC.parent.parent / "include" / "gskrige.h"]
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)
    jobs = []
    for s in srcs:
        o = objdir / (s.stem + ".o")
        if force or _newer(s, o, hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC, *FLAGS, "-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (objdir / (s.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s.name}:\n{r.stderr[-4000:]}")
        return s.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(compile_one, jobs):
                if verbose:
                    print("compiled", name, flush=True)
    objs = [objdir / (s.stem + ".o") for s in srcs]
    if jobs or not LIB.exists():
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-ccbin", "g++"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
