"""Build libtetsim_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m tetsim_b200.build [--force]

kernels_exact.cu is compiled with -fmad=false: the BITEXACT flavour restates the reference's
arithmetic (JS doubles / separately rounded f32), which never fuses a multiply with an add.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libtetsim_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"]
UNITS = {
    "kernels_exact.cu": ["-fmad=false"],
    "kernels_fast.cu": [],
    "tetsim_capi.cu": [],
    "mesh_prep.cpp": [],
}
HEADERS = ["device_math.cuh", "kernels.cuh", "table.cuh", "launch.h", "mesh_prep.h",
           os.path.join("..", "..", "include", "tetsim_b200.h")]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src, extra in UNITS.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append([NVCC] + COMMON + extra + ["-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            for warn in ex.map(run, jobs):
                if verbose and warn.strip():
                    print(warn)
    if jobs or force or _stale(LIB, objs):
        run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
