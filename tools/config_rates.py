#!/usr/bin/env python3
"""Throughput of every BASELINE.json config on one B200 (the bench line covers config 4 only).

    python tools/config_rates.py [--frames 30]

Prints substeps/s and tet-projections/s per config, CUDA-event timed through tetsim_step
(one CUDA-graph launch per frame), plus the CPU oracle's rate on the Dragon for reference.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetsim_b200 as ts  # noqa: E402
from tetsim_b200 import mesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=30)
ap.add_argument("--big", action="store_true", help="also run the 784-Dragon (3.0M-tet) scene")
a = ap.parse_args()
m = mesh.load_dragon()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)


def rate(name, body, pp, frames):
    for _ in range(3):
        body.step(pp)
    body.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(frames):
        body.step(pp)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sub = frames * pp["numSubsteps"]
    info = body.info()
    print("%-58s %9.0f substeps/s  %10.1f Mtet/s  (%d tets, %d launches/substep, finite=%s)" % (
        name, sub / ms * 1e3, info["numTets"] * info["iters"] * sub / ms / 1e3, info["numTets"],
        info["launchesPerSubstep"], bool(np.isfinite(body.pos).all())))
    body.close()


def dragon(pp, **kw):
    return ts.SoftBody(m["tet_verts"], m["tet_ids"], None, pp, stream=stream.cuda_stream, **kw)


p10 = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=10)
p20 = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=20)
print("== one B200, %d frames per config ==" % a.frames)
rate("C1/C3(i) Dragon NH GS exact order, BITEXACT (parity mode)", dragon(p10, solver="gs_exact", arithmetic="bitexact"), p10, a.frames)
rate("C1/C3(i) Dragon NH GS exact order, FAST_F32", dragon(p10, solver="gs_exact", arithmetic="fast"), p10, a.frames)
rate("C3(ii)   Dragon NH GS 32 colours, FAST_F32", dragon(p10, solver="gs_color", arithmetic="fast"), p10, a.frames)
rate("C2       Dragon polar-decomposition Jacobi, FAST_F32", ts.SoftBodyGPU(m["tet_verts"], m["tet_ids"], None, dict(p20), stream=stream.cuda_stream), p20, a.frames)
rate("C2       Dragon polar-decomposition Jacobi, BITEXACT", ts.SoftBodyGPU(m["tet_verts"], m["tet_ids"], None, dict(p20), arithmetic="bitexact", stream=stream.cuda_stream), p20, a.frames)
rate("         Dragon NH Jacobi (tile kernel), FAST_F32", dragon(p20, solver="jacobi"), p20, a.frames)
wb = list(mesh.wide_bounds(64.0))
for n in ([8, 28] if a.big else [8]):
    v, t = mesh.tile_bodies(m["tet_verts"], m["tet_ids"], n, n, y_shift=-0.40)
    pw = dict(p10, worldBounds=wb)
    for arith in ("bitexact", "fast"):
        rate("C5       %dx tiled Dragon + ground, NH GS exact, %s" % (n * n, arith),
             ts.SoftBody(v, t, None, pw, solver="gs_exact", arithmetic=arith, stream=stream.cuda_stream), pw, max(3, a.frames // 3))
    pj = dict(p20, worldBounds=wb)
    rate("C5       %dx tiled Dragon + ground, NH Jacobi tile kernel" % (n * n),
         ts.SoftBody(v, t, None, pj, solver="jacobi", stream=stream.cuda_stream), pj, a.frames)
# The CPU baseline of config 1 (the reference's algorithm on the Dragon, substeps/s) is measured by bench.py
# (cpu_baseline.dragon_substeps_per_s): tools never load oracle/.
