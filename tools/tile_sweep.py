#!/usr/bin/env python3
"""Sweep the experimental variants of the clustered-Jacobi tile kernel on one GPU.

    python tools/tile_sweep.py [--cells 407,64,64] [--reps 30] [--out gpurun_out/tile_sweep.txt]

For every tile size T in --sizes a body is created once (the host pre-pass depends on T only); the
variant switches (TETSIM_TILE_TPT, TETSIM_TILE_STAGES, ...) are read by the library at every launch, so
the same body is timed under each of them with tetsim_time_kernel (CUDA events around `reps` launches).
Correctness: every variant first runs 6 substeps on a small jittered beam and is compared with the
default variant of the same T (must agree bit for bit: same tiles, same per-tet math, same sum order);
the default variant itself is what tests/test_parity_gpu.py checks against the oracle.
"""
import argparse
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetsim_b200 as ts  # noqa: E402
from tetsim_b200 import mesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", default="407,64,64")
ap.add_argument("--check-cells", default="24,12,12")
ap.add_argument("--sizes", default="256,512,128")
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--out", default="")
a = ap.parse_args()

def V(tpt=0, stages=0, minb=0):
    e = {}
    if tpt: e["TETSIM_TILE_TPT"] = str(tpt)
    if stages: e["TETSIM_TILE_STAGES"] = str(stages)
    if minb: e["TETSIM_TILE_MINB"] = str(minb)
    return e


VARIANTS = {   # T -> list of env dicts (the first is the default kernel of that tile size: 2 tets per thread, 2 stages)
    128: [V(), V(1, 3), V(2, 3)],
    256: [V(), V(1, 3), V(2, 3), V(2, 2, 6), V(4, 2)],
    512: [V(), V(1), V(2, 2, 3), V(2, 3, 4), V(4, 2)],
}
KEYS = ("TETSIM_TILE_TPT", "TETSIM_TILE_STAGES", "TETSIM_TILE_MINB")


def set_env(env):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)


def make(cells, T):
    v, t = mesh.make_beam(cells, jitter=0.2)
    pp = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=20, worldBounds=list(mesh.wide_bounds(64.0)))
    return ts.SoftBody(v, t, None, pp, solver="jacobi", arithmetic="fast", cluster_size=T), pp


lines = []


def say(s):
    print(s, flush=True)
    lines.append(s)


cells = tuple(int(c) for c in a.cells.split(","))
ccells = tuple(int(c) for c in a.check_cells.split(","))
for T in (int(s) for s in a.sizes.split(",")):
    # ---- correctness of every variant on a small mesh ----
    ref = None
    ok = {}
    for env in VARIANTS[T]:
        set_env(env)
        b, pp = make(ccells, T)
        for _ in range(6):
            b.simulate(1.0 / 1200.0, pp)
        p = b.pos.copy()
        b.close()
        if ref is None:
            ref = p
        finite = bool(np.isfinite(p).all())
        d = float(np.abs(p - ref).max()) if finite else float("nan")
        ok[tuple(sorted(env.items()))] = (finite, d)
    # ---- timing on the full mesh ----
    set_env({})
    b, pp = make(cells, T)
    b.step(pp)
    b.synchronize()
    info = b.info()
    say("T=%d: tets %d, tiles %d, tile verts/tet %.3f (max %d per tile)" % (
        T, info["localTets"], info["numClusters"], info["sumLocalVerts"] / max(info["localTets"], 1), info["maxTileVerts"]))
    for env in VARIANTS[T]:
        set_env(env)
        best = 1e9
        for _ in range(3):
            ms, nbytes = b.time_kernel(a.reps)
            best = min(best, ms)
        fin, d = ok[tuple(sorted(env.items()))]
        say("  T=%-4d %-58s %.4f ms/launch  %.0f GB/s algorithmic  (check: finite=%s, max |dpos| vs default %.3g)" % (
            T, " ".join("%s=%s" % (k.replace("TETSIM_TILE_", ""), v) for k, v in env.items()) or "default", best,
            nbytes / best / 1e6, fin, d))
    set_env({})
    b.close()

if a.out:
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")
