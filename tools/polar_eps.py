#!/usr/bin/env python3
"""Exit rule of the FAST rotation extraction (device_math.cuh, pl_extract_rotation_fast): error against the CUDA BITEXACT path
(= the shader's nine iterations) and tile-kernel time, per noise multiplier k (TETSIM_POLAR_NOISE_K; 0 = fixed 1e-6 floor).

    python tools/polar_eps.py 0 2 4 8
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tetsim_b200 as ts  # noqa: E402
from tetsim_b200 import mesh  # noqa: E402


def run(verts, tets, arithmetic, substeps, pp):
    b = ts.SoftBodyGPU(verts, tets, None, dict(pp), arithmetic=arithmetic)
    out = []
    for s in range(substeps // 20):
        b.step(pp)
        out.append(b.pos.copy())
    return b, out


def main():
    ks = [float(a) for a in sys.argv[1:]] or [0.0, 4.0]
    pp = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=20, worldBounds=list(mesh.wide_bounds(64.0)))
    cases = {
        "beam 48x12x12 h=0.01 jitter 0.2, y0=0.01 (contact), 400 substeps": mesh.make_beam((48, 12, 12), h=0.01, y0=0.01, jitter=0.2),
        "dragon, 400 substeps": None,
    }
    d = mesh.load_dragon()
    cases["dragon, 400 substeps"] = (d["tet_verts"], d["tet_ids"])
    big = mesh.make_beam((407, 64, 64))
    for k in ks:
        os.environ["TETSIM_POLAR_NOISE_K"] = repr(k)
        print("k = %g" % k, flush=True)
        for name, (v, t) in cases.items():
            os.environ["TETSIM_POLAR_NOISE_K"] = repr(k)
            _, fast = run(v, t, "fast", 400, pp)
            _, ex = run(v, t, "bitexact", 400, pp)
            errs = [float(np.abs(a - b).max()) for a, b in zip(fast, ex)]
            print("   %-66s max|fast - bitexact| after 20/100/200/400 substeps: %.2e %.2e %.2e %.2e" % (name, errs[0], errs[4], errs[9], errs[19]), flush=True)
        b = ts.SoftBodyGPU(big[0], big[1], None, dict(pp), arithmetic="fast", cluster_size=128)
        done = 0
        for frames in (2, 30, 40):   # in free fall / after the impact on the floor (t = 0.45 s) / later
            for _ in range(frames):
                b.step(pp)
            done += 20 * frames
            ms, nbytes = b.time_kernel(10)
            print("   10M-tet beam after %4d substeps: k_polar_tiles %.4f ms  (%.3f of 6454.9 GB/s)" % (done, ms, nbytes / ms / 1e6 / 6454.9), flush=True)
        b.close()


if __name__ == "__main__":
    main()
