"""Scenarios run identically by the transpiled reference (oracle/ref_runner.py), the C oracle and the CUDA library --
TEST INFRASTRUCTURE.  tools/make_ref_golden.py runs them on the transpiled reference and commits the results as
tests/golden/ref_golden.npz; the tests then demand bit equality from the C oracle (CPU) and from the CUDA BITEXACT
path (GPU box, where /root/reference does not exist).

Every scenario is the Dragon mesh (src/Dragon.js) under the reference's own call pattern: simulate(dt, physicsParams)
per substep (src/main.js:80-84), startGrab/moveGrabbed/endGrab from the pointer handlers (src/Softbody.js:451-469).
"""
from __future__ import annotations

import numpy as np

DEFAULTS = dict(gravity=-9.81, friction=1000.0, density=1000.0, devCompliance=1.0 / 100000.0, volCompliance=0.0,
                worldBounds=(-2.5, -1.0, -2.5, 2.5, 10.0, 2.5))
FRAME_DT = 1.0 * (1.0 / 60.0)   # timeScale * timeStep, src/main.js:79


def _p(**kw):
    d = dict(DEFAULTS)
    d.update(kw)
    return d


SCENARIOS = [
    # BASELINE config 1: defaults, dt = 1/600 (10 substeps per frame), free fall, 100 substeps
    dict(name="free100", shift=(0.0, 0.0, 0.0), params=_p(), dt=FRAME_DT / 10, steps=100, save=(1, 10, 50, 100), events={}),
    # floor contact from the first substep, friction below the min(1, dt*friction) knee, bounds that clamp in x and z
    dict(name="contact40", shift=(0.0, -0.46, 0.0), params=_p(friction=100.0, worldBounds=(-0.9, -1.0, -0.4, 0.95, 10.0, 0.35)),
         dt=FRAME_DT / 10, steps=40, save=(1, 20, 40), events={}),
    # non-zero volume compliance (alpha path of the hydrostatic constraint), other gravity, the demo's CPU dt = 1/300
    dict(name="compliant20", shift=(0.0, 0.0, 0.0), params=_p(gravity=-5.0, devCompliance=2.0e-5, volCompliance=1.0e-6),
         dt=FRAME_DT / 5, steps=20, save=(1, 20), events={}),
    # grab: nearest-vertex pick, pinned vertex dragged, released (src/Softbody.js:233-235, 279-298)
    dict(name="grab30", shift=(0.0, 0.0, 0.0), params=_p(), dt=FRAME_DT / 10, steps=30, save=(5, 15, 30),
         events={0: ("start", (0.3, 1.6, 0.05)), 5: ("move", (0.35, 1.75, 0.1)), 10: ("move", (0.4, 1.9, 0.0)), 20: ("end", None)}),
]


def shifted(verts, shift):
    v = np.asarray(verts, np.float32).reshape(-1, 3).copy()
    v += np.asarray(shift, np.float32)   # one f32 rounding per coordinate; every implementation receives these f32 values
    return v.reshape(-1)


def run(sc, body, read, on_save):
    """Drive `body` (simulate/startGrab/moveGrabbed/endGrab) through scenario `sc`; `read(body)` returns a dict of arrays."""
    for s in range(sc["steps"]):
        ev = sc["events"].get(s)
        if ev:
            kind, p = ev
            if kind == "start":
                body.startGrab(p)
            elif kind == "move":
                body.moveGrabbed(p)
            else:
                body.endGrab()
        body.simulate(sc["dt"], sc["params"])
        if s + 1 in sc["save"]:
            on_save(s + 1, read(body))


def polar_scenarios(dragon_verts, dragon_tets, mesh):
    """(name, (verts, tets), physicsParams, substeps, checkpoints) for the WebGL solver (SoftBodyGPU), dt = 1/1200 (its default
    20 substeps per frame).  No grab: the shader's indexFromUV decode pins the wrong particle (src/SoftbodyGPU.js:336-338, its
    own comment says so) and is deliberately not replicated; bounds are the shader's hard-coded ones (:347) = the defaults."""
    beam = mesh.make_beam((3, 2, 2), h=0.2, y0=0.0008, jitter=0.2)
    return [
        ("polar_dragon", (dragon_verts, dragon_tets), _p(), 10, (1, 5, 10)),
        # floor contact within the run, friction below the min(1, dt * friction) knee, other gravity
        ("polar_beam", beam, _p(friction=300.0, gravity=-12.0), 40, (1, 20, 40)),
    ]
