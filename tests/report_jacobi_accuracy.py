#!/usr/bin/env python3
"""(Test-side report, not collected by pytest: the oracle is the checker here.)  Accuracy of the clustered-Jacobi tile kernel (FAST_F32) against the oracle's Jacobi on Dragon:
vector-relative position error, max velocity difference and volError difference after 100 substeps at
dt = 1/1200, for every tile size (and whatever TETSIM_TILE_* variant switches are set)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle  # noqa: E402
import tetsim_b200 as ts  # noqa: E402
from tetsim_b200 import mesh  # noqa: E402
from util import vec_rel_err  # noqa: E402

m = mesh.load_dragon()
dt = 1.0 / 1200.0
ref = oracle.SoftBodyOracle(m["tet_verts"], m["tet_ids"])
hist = {}
for s in range(1, 101):
    ref.simulate_jacobi(dt, 1)
    if s in (10, 50, 100):
        hist[s] = (ref.pos.copy(), ref.vel.copy(), ref.volError)
for T in (32, 64, 128, 256, 512):
    sb = ts.SoftBody(m["tet_verts"], m["tet_ids"], None, None, solver="jacobi", arithmetic="fast", cluster_size=T,
                     track_vol_error=True)
    out = []
    for s in range(1, 101):
        sb.simulate(dt)
        if s in hist:
            p, v, ve = hist[s]
            out.append("@%d pos %.2e vel %.2e volErr %.1e" % (s, vec_rel_err(sb.pos, p), float(np.max(np.abs(sb.vel - v))),
                                                              abs(sb.volError - ve)))
    print("T=%-3d %s" % (T, " | ".join(out)), flush=True)
    sb.close()
