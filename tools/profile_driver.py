#!/usr/bin/env python3
"""Small driver for ncu: builds the bench workload, runs a few steps and the tile kernel alone.

    ncu --set full --clock-control none --import-source on -k regex:k_jacobi_tiles -s 2 -c 1 \
        -o gpurun_out/prof python tools/profile_driver.py [--cells x,y,z] [--cluster-size T]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetsim_b200 as ts  # noqa: E402
from tetsim_b200 import mesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", default="407,64,64")
ap.add_argument("--cluster-size", type=int, default=256)
ap.add_argument("--substeps", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--atomic", action="store_true")
ap.add_argument("--no-reorder", action="store_true")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
cells = tuple(int(c) for c in a.cells.split(","))
v, t = mesh.make_beam(cells)
pp = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=a.substeps, worldBounds=list(mesh.wide_bounds(64.0)))
b = ts.SoftBody(v, t, None, pp, solver="jacobi", arithmetic="fast", cluster_size=a.cluster_size,
                deterministic=not a.atomic, reorder=not a.no_reorder)
for _ in range(a.steps):
    b.step(pp)
b.synchronize()
ms, nbytes = b.time_kernel(a.reps)
info = b.info()
print("tile kernel %.4f ms/launch, %.1f GB/s algorithmic, %.3e tets/s; T=%d tiles %d, tile verts %.2f per tet (max %d per tile), "
      "meta %.1f B/tet" % (ms, nbytes / ms / 1e6, info["localTets"] / ms * 1e3, info["clusterSize"], info["numClusters"],
                          info["sumLocalVerts"] / max(info["localTets"], 1), info["maxTileVerts"],
                          info["tileMetaBytes"] / max(info["localTets"], 1)))
