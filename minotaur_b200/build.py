"""In-tree build of libmntr_gpu.so for sm_100a (nvcc cross-compiles without a GPU).

``python -m minotaur_b200.build`` or ``minotaur_b200.build.build()``.  The library is a plain
C-ABI shared object (include/mntr_gpu.h): no torch, no Python in it.  The same sources are
built by minotaur_b200/CMakeLists.txt when the module is dropped into the Minotaur tree as
``src/gpu``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmntr_gpu.so")
OBJ = os.path.join(HERE, "build")

SOURCES = ["linear_single.cu", "linear_batch.cu", "linear_rounds.cu", "root_rows.cu", "quad_relations.cu", "mntr_gpu.cu", "mntr_group.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--fmad=false",        # never contract a*b+c: the reference builds without FMA
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v"]


def nvcc() -> str:
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found: the CUDA engine cannot be built")
    return path


def _stale(target: str, deps: List[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "mntr_gpu.h"))
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc()] + NVCC_FLAGS + ["-c", s, "-o", o]
            res = subprocess.run(cmd, capture_output=True, text=True)
            with open(o + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError(f"nvcc failed on {src}")
            if verbose:
                sys.stderr.write(res.stderr)
    if force or _stale(LIB, objs):
        # cudart is linked statically (nvcc default); libstdc++ dynamically and explicitly,
        # because this image's g++ wrapper otherwise links the static archive.
        cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + \
              ["-Xlinker", "-soname=libmntr_gpu.so", "-Xlinker", "--no-as-needed", "-lstdc++", "-lm", "-ldl", "-lpthread", "-lrt"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link of libmntr_gpu.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
