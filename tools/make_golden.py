#!/usr/bin/env python3
"""Freeze regression vectors of the CPU oracle on the Dragon mesh -> tests/golden/softbody_golden.npz.

PARITY UNPINNED: the reference has no golden vectors and cannot be executed here (no JS engine), so
these come from oracle/ (the C restatement).  What pins them:
  * tests/test_oracle.py re-derives the Gauss-Seidel vectors with the independent numpy restatement
    (oracle/oracle_np.py) and requires bit equality;
  * the SURVEY-time emulation anchors (volError, sum of positions, pos[0],
    invRestPose[0..8], invMass[0..3]), produced by a third, separately written numba emulation, are
    stored below as literals and checked against the oracle by the same test.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

DT600 = (1.0 * (1.0 / 60.0)) / 10
DT1200 = (1.0 * (1.0 / 60.0)) / 20

m = np.load(os.path.join(ROOT, "tetsim_b200", "assets", "dragon_mesh.npz"))
V, T = m["tet_verts"], m["tet_ids"]
out = {}

sb = oracle.SoftBodyOracle(V, T)
out["invRestPose"] = sb.invRestPose.copy()
out["invRestVolume"] = sb.invRestVolume.copy()
out["invMass"] = sb.invMass.copy()
for s in range(1, 101):
    sb.simulate(DT600)
    if s in (1, 10, 100):
        out["gs_pos_%d" % s] = sb.pos.copy()
        out["gs_vel_%d" % s] = sb.vel.copy()
        out["gs_vol_error_%d" % s] = np.float64(sb.volError)

# greedy colour order (tet order, smallest free colour): computed here independently of the library
ids = T.reshape(-1, 4)
used = [set() for _ in range(V.size // 3)]
color = np.zeros(len(ids), np.int32)
for e, t in enumerate(ids):
    taken = used[t[0]] | used[t[1]] | used[t[2]] | used[t[3]]
    c = 0
    while c in taken:
        c += 1
    color[e] = c
    for k in t:
        used[k].add(c)
out["greedy_color"] = color
order = np.argsort(color, kind="stable").astype(np.int32)
sb = oracle.SoftBodyOracle(V, T)
for s in range(100):
    sb.simulate(DT600, order=order)
out["gs_color_pos_100"] = sb.pos.copy()

for iters in (1, 4):
    sb = oracle.SoftBodyOracle(V, T)
    for s in range(1, 101):
        sb.simulate_jacobi(DT1200, iters)
        if s in (10, 100):
            out["jacobi%d_pos_%d" % (iters, s)] = sb.pos.copy()

for bug in (1, 0):
    po = oracle.PolarOracle(V, T, reference_table_bug=bool(bug))
    for s in range(1, 101):
        po.simulate(DT1200)
        if s in (10, 100):
            out["polar_bug%d_pos_%d" % (bug, s)] = po.pos.copy()
    out["polar_bug%d_quat_100" % bug] = po.quat.copy()

path = os.path.join(ROOT, "tests", "golden", "softbody_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", sorted(out))
