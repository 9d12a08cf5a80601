#!/usr/bin/env python3
"""Golden vectors produced by the REFERENCE'S OWN TEXT -> tests/golden/ref_golden.npz.

    python tools/make_ref_golden.py          (needs /root/reference; about two minutes of pure-Python execution)

1. tools/transpile_reference.py re-emits src/Softbody.js (class SoftBody, every method) and SoftBodyGPU.initPhysics
   mechanically as Python under JS number semantics (oracle/jsrt.py) into oracle/_ref/ (git-ignored).
2. The scenarios of oracle/ref_scenarios.py are run on that code; positions / prevPos / velocities / volError / grabId
   at the listed substeps, the initPhysics arrays, the skinned surface mesh and the WebGL variant's init textures
   (reverse tables with the `<= 0.0` slot rule, elems0, quats0, invMass, invRestVolume) are stored.
The file is the fixture that travels to the GPU box: the C oracle (CPU tests) and the CUDA BITEXACT path (GPU tests)
must reproduce every array bit for bit.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_runner, ref_scenarios  # noqa: E402


def main():
    if not ref_runner.reference_present():
        sys.exit("make_ref_golden: /root/reference is not present")
    ref_runner.ensure()
    m = np.load(os.path.join(ROOT, "tetsim_b200", "assets", "dragon_mesh.npz"))
    V, T = m["tet_verts"], m["tet_ids"]
    out = {}
    t0 = time.time()
    for sc in ref_scenarios.SCENARIOS:
        first = sc["name"] == "free100"
        body = ref_runner.RefSoftBody(ref_scenarios.shifted(V, sc["shift"]), T, sc["params"],
                                      m["vis_verts"] if first else None, m["vis_tri_ids"] if first else None,
                                      m["tet_edge_ids"] if first else None)
        if first:
            out["invRestPose"], out["invRestVolume"], out["invMass"] = body.invRestPose, body.invRestVolume, body.invMass
            out["vis_pos_0"] = body.visPositions          # updateVisMesh() at the end of the constructor, :57

        def save(step, d, name=sc["name"]):
            for k, v in d.items():
                out["%s_%s_%d" % (name, k, step)] = v

        ref_scenarios.run(sc, body, lambda b: dict(pos=b.pos, prev=b.prevPos, vel=b.vel, volError=np.float64(b.volError),
                                                   grabId=np.int32(b.grabId)), save)
        if first:
            body.endFrame()                               # updateEdgeMesh + updateVisMesh, :244-277
            out["vis_pos_100"] = body.visPositions
            out["edge_pos_100"] = body.edgePositions
            out["caller_vertices_100"] = ref_runner._np(body.vertices)   # the aliasing quirk: the caller's array is overwritten
        print("%-12s %3d substeps  %.1f s" % (sc["name"], sc["steps"], time.time() - t0), flush=True)

    g = ref_runner.RefSoftBodyGPUInit(V, T)
    out["gpu_texDim"] = np.int32(g.texDim)
    out["gpu_biggestT"] = np.int32(int(g.biggestT))
    for k in range(9):
        out["gpu_table_%d" % k] = g.tex("particleToElemVertsTable", k)
    for k in range(4):
        out["gpu_elems0_%d" % k] = g.tex("elems0", k)
    for name in ("pos0", "vel0", "invMass", "invRestVolumeAndColor", "elemToParticlesTable", "quats0"):
        out["gpu_" + name] = g.tex(name)
    # ---- the WebGL solver, whole substeps: initPhysics + simulate + the GPGPU runtime's compute() transpiled from JavaScript,
    # the seven passes transpiled from GLSL (tools/transpile_shaders.py), executed texel by texel (oracle/ref_runner.py) ----
    from tetsim_b200 import mesh as tmesh
    for name, (pv, pt), params, steps, save in ref_scenarios.polar_scenarios(V, T, tmesh):
        body = ref_runner.RefSoftBodyGPU(pv, pt, params)
        for s in range(1, steps + 1):
            body.simulate(ref_scenarios.FRAME_DT / 20, params)
            if s in save:
                for k, a in dict(pos=body.pos, prev=body.prevPos, vel=body.vel, quat=body.quat, rest=body.rest).items():
                    out["%s_%s_%d" % (name, k, s)] = a
        print("%-12s %3d substeps  %.1f s" % (name, steps, time.time() - t0), flush=True)
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
