"""CPU oracle for the TetSim substep path -- TEST INFRASTRUCTURE, not product code.

ctypes bindings over ``oracle/liboracle.so`` (built from ``softbody_oracle.c`` and
``polar_oracle.c`` by ``oracle/Makefile``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; nothing
under ``tetsim_b200/`` does.

PARITY PINNED TO THE REFERENCE'S OWN TEXT: the reference (zalo/TetSim) has no tests or golden vectors and is
JavaScript + GLSL, which this image cannot execute; ``tools/transpile_reference.py`` and ``tools/transpile_shaders.py``
re-emit it mechanically as Python (``oracle/_ref/``, run under ``oracle/jsrt.py`` / ``oracle/glslrt.py``), and the vectors
that produces (``tests/golden/ref_golden.npz``) are what the C restatements here must reproduce bit for bit
(``tests/test_reference_pin.py``).  See the headers of ``softbody_oracle.c`` and ``polar_oracle.c``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

# Defaults of the reference's physicsParams object, src/main.js:22-36.
DEFAULT_PARAMS = dict(
    gravity=-9.81,
    friction=1000.0,
    density=1000.0,
    devCompliance=1.0 / 100000.0,
    volCompliance=0.0,
    worldBounds=(-2.5, -1.0, -2.5, 2.5, 10.0, 2.5),
)


class OracleParams(C.Structure):
    _fields_ = [
        ("gravity", C.c_double),
        ("friction", C.c_double),
        ("density", C.c_double),
        ("devCompliance", C.c_double),
        ("volCompliance", C.c_double),
        ("worldBounds", C.c_double * 6),
    ]


class PolarParams(C.Structure):
    _fields_ = [("gravity", C.c_float), ("friction", C.c_float), ("worldBounds", C.c_float * 6)]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("softbody_oracle.c", "polar_oracle.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_nearest_vertex.restype = C.c_int
        _lib.oracle_polar_build_table.restype = C.c_int
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def make_params(**kw) -> OracleParams:
    d = dict(DEFAULT_PARAMS)
    d.update(kw)
    p = OracleParams()
    p.gravity, p.friction, p.density = d["gravity"], d["friction"], d["density"]
    p.devCompliance, p.volCompliance = d["devCompliance"], d["volCompliance"]
    for i, v in enumerate(d["worldBounds"]):
        p.worldBounds[i] = v
    return p


def valence_of(num_verts: int, tet_ids) -> np.ndarray:
    return np.bincount(np.asarray(tet_ids).reshape(-1), minlength=num_verts).astype(np.int32)


class SoftBodyOracle:
    """State + methods of the reference's ``SoftBody`` (src/Softbody.js:3-298), CPU restatement."""

    def __init__(self, verts, tet_ids, **params):
        self.params = dict(DEFAULT_PARAMS)
        self.params.update(params)
        self.pos = _f32(verts).reshape(-1).copy()
        self.prevPos = self.pos.copy()
        self.numParticles = self.pos.size // 3
        self.tetIds = _i32(tet_ids).reshape(-1).copy()
        self.numElems = self.tetIds.size // 4
        self.vel = np.zeros(3 * self.numParticles, np.float32)
        self.invMass = np.zeros(self.numParticles, np.float32)
        self.invRestPose = np.zeros(9 * self.numElems, np.float32)
        self.invRestVolume = np.zeros(self.numElems, np.float32)
        self.volError = 0.0
        self.grabId = -1
        self.grabPos = np.zeros(3, np.float32)
        self.valence = valence_of(self.numParticles, self.tetIds)
        self._acc = np.zeros(3 * self.numParticles, np.float32)
        lib().oracle_init_physics(
            self.numParticles, self.numElems, _p(self.pos, C.c_float), _p(self.tetIds, C.c_int),
            C.c_double(self.params["density"]), _p(self.invRestPose, C.c_float),
            _p(self.invRestVolume, C.c_float), _p(self.invMass, C.c_float))

    def _cparams(self, override):
        d = dict(self.params)
        if override:
            d.update(override)
        return make_params(**d)

    def simulate(self, dt: float, params: dict | None = None, order=None):
        """One substep, Gauss-Seidel in ``order`` (None = tet index order, the reference's)."""
        p = self._cparams(params)
        ve = C.c_double(0.0)
        o = _i32(order) if order is not None else None
        lib().oracle_simulate(
            self.numParticles, self.numElems, _p(self.pos, C.c_float), _p(self.prevPos, C.c_float),
            _p(self.vel, C.c_float), _p(self.invMass, C.c_float), _p(self.invRestPose, C.c_float),
            _p(self.invRestVolume, C.c_float), _p(self.tetIds, C.c_int), _p(o, C.c_int), C.c_double(dt),
            C.byref(p), self.grabId, _p(self.grabPos, C.c_float), C.byref(ve))
        self.volError = ve.value

    def simulate_jacobi(self, dt: float, iters: int = 1, params: dict | None = None):
        p = self._cparams(params)
        ve = C.c_double(0.0)
        lib().oracle_simulate_jacobi(
            self.numParticles, self.numElems, _p(self.pos, C.c_float), _p(self.prevPos, C.c_float),
            _p(self.vel, C.c_float), _p(self.invMass, C.c_float), _p(self.invRestPose, C.c_float),
            _p(self.invRestVolume, C.c_float), _p(self.tetIds, C.c_int), _p(self.valence, C.c_int),
            _p(self._acc, C.c_float), int(iters), C.c_double(dt), C.byref(p), self.grabId,
            _p(self.grabPos, C.c_float), C.byref(ve))
        self.volError = ve.value

    def jacobi_accumulate(self, tet_list, dt: float) -> np.ndarray:
        """dx sums (3 per vertex) contributed by the tets in ``tet_list`` from the current positions."""
        tl = _i32(tet_list)
        acc = np.zeros(3 * self.numParticles, np.float32)
        p = self._cparams(None)
        lib().oracle_jacobi_accumulate(
            tl.size, _p(tl, C.c_int), _p(self.pos, C.c_float), _p(self.invMass, C.c_float),
            _p(self.invRestPose, C.c_float), _p(self.invRestVolume, C.c_float), _p(self.tetIds, C.c_int),
            _p(acc, C.c_float), C.c_double(dt), C.byref(p))
        return acc

    # grab API, src/Softbody.js:279-298
    def startGrab(self, pos):
        p = np.asarray(pos, np.float64)
        self.grabId = lib().oracle_nearest_vertex(self.numParticles, _p(self.pos, C.c_float), _p(p, C.c_double))
        self.grabPos[:] = p

    def moveGrabbed(self, pos):
        self.grabPos[:] = np.asarray(pos, np.float64)

    def endGrab(self):
        self.grabId = -1


def skin(vis_verts, tet_ids, pos) -> np.ndarray:
    vv = _f32(vis_verts).reshape(-1)
    n = vv.size // 4
    out = np.zeros(3 * n, np.float32)
    ids, p = _i32(tet_ids).reshape(-1), _f32(pos).reshape(-1)
    lib().oracle_skin(n, _p(vv, C.c_float), _p(ids, C.c_int), _p(p, C.c_float), _p(out, C.c_float))
    return out


def vertex_normals(pos, tri_ids) -> np.ndarray:
    p, t = _f32(pos).reshape(-1), _i32(tri_ids).reshape(-1)
    out = np.zeros_like(p)
    lib().oracle_vertex_normals(p.size // 3, t.size // 3, _p(p, C.c_float), _p(t, C.c_int), _p(out, C.c_float))
    return out


def polar_skin(vis_verts, tet_ids, pos, quat, rest_normals=None):
    """The WebGL variant's vertex-shader skinning (src/SoftbodyGPU.js:424-448): (positions, normals or None)."""
    vv = _f32(vis_verts).reshape(-1)
    n = vv.size // 4
    out = np.zeros(3 * n, np.float32)
    nrm = np.zeros(3 * n, np.float32) if rest_normals is not None else None
    rn = _f32(rest_normals).reshape(-1) if rest_normals is not None else None
    lib().oracle_polar_skin(n, _p(vv, C.c_float), _p(_i32(tet_ids).reshape(-1), C.c_int), _p(_f32(pos).reshape(-1), C.c_float),
                            _p(_f32(quat).reshape(-1), C.c_float), _p(rn, C.c_float), _p(out, C.c_float), _p(nrm, C.c_float))
    return out, nrm


class PolarOracle:
    """State + substep of the reference's ``SoftBodyGPU`` (src/SoftbodyGPU.js), CPU restatement in f32."""

    def __init__(self, verts, tet_ids, reference_table_bug: bool = True, **params):
        self.params = dict(DEFAULT_PARAMS)
        self.params.update(params)
        v = _f32(verts).reshape(-1)
        self.pos = v.copy()
        self.prevPos = v.copy()
        self.numParticles = v.size // 3
        self.tetIds = _i32(tet_ids).reshape(-1).copy()
        self.numElems = self.tetIds.size // 4
        self.vel = np.zeros_like(v)
        self.rest = np.zeros(12 * self.numElems, np.float32)
        self.quat = np.zeros(4 * self.numElems, np.float32)
        self.invRestVolume = np.zeros(self.numElems, np.float32)
        self.grabId = -1
        self.grabPos = np.zeros(3, np.float32)
        L = lib()
        L.oracle_polar_init(self.numParticles, self.numElems, _p(v, C.c_float), _p(self.tetIds, C.c_int),
                            _p(self.rest, C.c_float), _p(self.quat, C.c_float), _p(self.invRestVolume, C.c_float))
        self.tblStart = np.zeros(self.numParticles + 1, np.int32)
        self.tblEntries = np.zeros(4 * self.numElems, np.int32)
        n = L.oracle_polar_build_table(self.numParticles, self.numElems, _p(self.tetIds, C.c_int),
                                       int(bool(reference_table_bug)), 36, _p(self.tblStart, C.c_int),
                                       _p(self.tblEntries, C.c_int))
        self.tblEntries = self.tblEntries[:n].copy()

    def simulate(self, dt: float, params: dict | None = None):
        d = dict(self.params)
        if params:
            d.update(params)
        p = PolarParams()
        p.gravity, p.friction = d["gravity"], d["friction"]
        for i, b in enumerate(d["worldBounds"]):
            p.worldBounds[i] = b
        lib().oracle_polar_simulate(
            self.numParticles, self.numElems, _p(self.pos, C.c_float), _p(self.prevPos, C.c_float),
            _p(self.vel, C.c_float), _p(self.rest, C.c_float), _p(self.quat, C.c_float),
            _p(self.invRestVolume, C.c_float), _p(self.tetIds, C.c_int), _p(self.tblStart, C.c_int),
            _p(self.tblEntries, C.c_int), C.c_float(dt), C.byref(p), self.grabId, _p(self.grabPos, C.c_float))
