#!/usr/bin/env python3
"""Ablations of the tile kernel (TETSIM_TILE_DEBUG bits, compiled only into the one-tet-per-thread T = 256 instantiation that
tetsim_time_kernel selects when the variable is set; results are WRONG by construction, only the time means something):
  1 no corner sums   2 no solve (zeros scattered)   4 no vertex gather   8 synthetic records (no HBM tet stream)   16 no scatter

    python tools/tile_ablation.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetsim_b200 as ts  # noqa: E402
from tetsim_b200 import mesh  # noqa: E402

v, t = mesh.make_beam((407, 64, 64))
pp = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=20, worldBounds=list(mesh.wide_bounds(64.0)))
b = ts.SoftBody(v, t, None, pp, solver="jacobi", arithmetic="fast", cluster_size=256)
rows = [(0, "everything (DBG instantiation, one tet per thread, 3 stages)"), (2, "no solve"), (1, "no corner sums"), (16, "no scatter"),
        (4, "no vertex gather"), (8, "no HBM tet stream (synthetic records)"), (2 + 16 + 1, "data movement only: stream + gather, no solve / scatter / sums"),
        (8 + 4, "compute only: solve + scatter + sums on whatever is in shared memory"), (8 + 4 + 2 + 16 + 1, "empty pipeline (barriers, metadata, loop)")]
for bits, what in rows:
    os.environ["TETSIM_TILE_DEBUG"] = str(bits)
    ms = min(b.time_kernel(20)[0] for _ in range(3))
    print("debug %2d  %.4f ms/launch   %s" % (bits, ms, what), flush=True)
os.environ.pop("TETSIM_TILE_DEBUG")
print("production kernel at T=256 (two tets per thread): %.4f ms/launch" % min(b.time_kernel(20)[0] for _ in range(3)))
