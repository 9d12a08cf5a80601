"""Host-side mirror of the reference's solver classes, over the C ABI.

    SoftBody     <- src/Softbody.js:3-298      Neo-Hookean XPBD (Gauss-Seidel in the reference)
    SoftBodyGPU  <- src/SoftbodyGPU.js:4-712   polar-decomposition shape matching (Jacobi)

Same constructor argument order, same method and field names, same call pattern as Main.update
(src/main.js:74-96): ``simulate(dt, physicsParams)`` once per substep, ``endFrame()`` once per frame,
``startGrab/moveGrabbed/endGrab`` from the pointer handlers.  ``physicsParams`` is the reference's
plain object (a dict here, keys of src/main.js:22-36) and is re-read on every call.  The JavaScript
versions of these two classes (tetsim_b200/js/softbody.mjs) are line-for-line the same thin
wrappers over the N-API shim; this image has no JS engine, so the Python ones are what runs.

There is no CPU path: constructing a body without a usable B200 raises TetSimError.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np

from . import _capi
from ._capi import (ARITH_BITEXACT, ARITH_FAST_F32, NH_GS_COLOR, NH_GS_EXACT, NH_JACOBI, POLAR_JACOBI, TetSimError,
                    check, default_options, default_params, ptr)

def default_physics_params(cpu_sim: bool = True) -> dict:
    """The physicsParams object of src/main.js:22-36: numSubsteps is 5 with ?cpu=true (SoftBody) and 20 otherwise (SoftBodyGPU)."""
    n = 5 if cpu_sim else 20
    return {
        "gravity": -9.81, "timeScale": 1.0, "timeStep": 1.0 / 60.0, "numSubsteps": n, "dt": 1.0 / (60.0 * n),
        "friction": 1000.0, "density": 1000.0, "devCompliance": 1.0 / 100000.0, "volCompliance": 0.0,
        "worldBounds": [-2.5, -1.0, -2.5, 2.5, 10.0, 2.5], "computeNormals": True, "ShowTetMesh": False, "cpuSim": cpu_sim,
    }


DEFAULT_PHYSICS_PARAMS = default_physics_params(True)

_SOLVERS = {"gs_exact": NH_GS_EXACT, "gs_color": NH_GS_COLOR, "jacobi": NH_JACOBI, "polar": POLAR_JACOBI}
_ARITH = {"fast": ARITH_FAST_F32, "bitexact": ARITH_BITEXACT}


def _params_struct(pp: dict):
    return default_params(**{k: pp[k] for k in ("gravity", "friction", "density", "devCompliance", "volCompliance",
                                                "worldBounds") if k in pp})


class _Body:
    """Shared plumbing of the two reference classes."""

    _default_solver = "gs_exact"

    def __init__(self, vertices, tetIds, tetEdgeIds, physicsParams, visVerts=None, visTriIds=None, visMaterial=None,
                 world=None, *, solver=None, arithmetic="fast", iters=1, deterministic=True, reference_table_bug=True,
                 reorder=True, cluster_size=256, track_vol_error=None, device=-1, stream=0, rank=0, world_size=1,
                 nccl_unique_id=None, exchange="allreduce"):
        self.physicsParams = physicsParams if physicsParams is not None else default_physics_params(self._default_solver != "polar")
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1)
        t = np.ascontiguousarray(tetIds, np.int32).reshape(-1)
        if v.size % 3 or t.size % 4:
            raise ValueError("vertices must hold 3 floats per particle and tetIds 4 ints per element")
        self.numParticles = v.size // 3           # src/Softbody.js:9
        self.numElems = t.size // 4               # :10
        self.tetIds = t                           # :19 (a copy: the library never aliases caller memory)
        self.tetEdgeIds = None if tetEdgeIds is None else np.asarray(tetEdgeIds, np.int32).reshape(-1)
        self.grabId = -1                          # :23
        self.grabPos = np.zeros(3, np.float32)    # :22
        self.visVerts = None if visVerts is None else np.ascontiguousarray(visVerts, np.float32).reshape(-1)
        self.visTriIds = None if visTriIds is None else np.ascontiguousarray(visTriIds, np.int32).reshape(-1)
        self.numVisVerts = 0 if self.visVerts is None else self.visVerts.size // 4
        self.visMaterial = visMaterial
        self.world = world
        self._nccl_id = None
        opt = default_options(
            solver=_SOLVERS[solver or self._default_solver], arithmetic=_ARITH[arithmetic], iters=int(iters),
            deterministic=int(bool(deterministic)), referenceTableBug=int(bool(reference_table_bug)),
            reorder=int(bool(reorder)), clusterSize=int(cluster_size),
            trackVolError=-1 if track_vol_error is None else int(bool(track_vol_error)),
            device=int(device), rank=int(rank), worldSize=int(world_size), stream=int(stream) or None,
            exchange={"allreduce": 0, "halo": 1, "peer": 2}[exchange])
        if nccl_unique_id is not None:
            self._nccl_id = C.create_string_buffer(bytes(nccl_unique_id), 128)
            opt.ncclUniqueId = C.cast(self._nccl_id, C.c_void_p)
        self._h = C.c_void_p()
        prm = _params_struct(self.physicsParams)
        check(_capi.lib().tetsim_create(
            v.ctypes.data_as(C.POINTER(C.c_float)), self.numParticles, t.ctypes.data_as(C.POINTER(C.c_int32)),
            self.numElems, C.byref(prm), C.byref(opt), C.byref(self._h)))
        self._cache = {}
        # the two scene objects Main adds (src/main.js:67-68); plain holders of the render buffers
        self.edgeMesh = SimpleNamespace(positions=v.copy(), index=self.tetEdgeIds, visible=True, userData=self)
        self.visMesh = SimpleNamespace(positions=np.zeros(3 * self.numVisVerts, np.float32),
                                       normals=np.zeros(3 * self.numVisVerts, np.float32), index=self.visTriIds,
                                       material=visMaterial, userData=self)
        if self.numVisVerts:
            self.updateVisMesh()

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _capi.lib().tetsim_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the hot path ----
    def simulate(self, dt, physicsParams=None):
        """One substep: src/Softbody.js:195-240 / src/SoftbodyGPU.js:610-641."""
        if physicsParams is not None:
            self.physicsParams = physicsParams
        prm = _params_struct(self.physicsParams)
        self._cache.clear()
        check(_capi.lib().tetsim_simulate(self._h, float(dt), C.byref(prm)))

    def step(self, physicsParams=None):
        """The substep loop of Main.update (src/main.js:79-84) as one CUDA-graph launch."""
        if physicsParams is not None:
            self.physicsParams = physicsParams
        pp = self.physicsParams
        prm = _params_struct(pp)
        self._cache.clear()
        frame_dt = pp.get("timeScale", 1.0) * pp.get("timeStep", 1.0 / 60.0)
        check(_capi.lib().tetsim_step(self._h, float(frame_dt), int(pp.get("numSubsteps", 1)), C.byref(prm)))

    def synchronize(self):
        check(_capi.lib().tetsim_synchronize(self._h))

    # ---- peer-memory exchange (exchange="peer"): hand-shake of the exchange buffers ----
    def ipc_handle(self) -> bytes:
        """This rank's blob for tetsim_set_peers (all-gather them in rank order)."""
        buf = C.create_string_buffer(_capi.PEER_BLOB_BYTES)
        check(_capi.lib().tetsim_get_ipc_handle(self._h, buf))
        return buf.raw

    def set_peers(self, blobs):
        """blobs: the ipc_handle() of every rank, in rank order."""
        raw = b"".join(bytes(b) for b in blobs)
        assert len(raw) % _capi.PEER_BLOB_BYTES == 0
        check(_capi.lib().tetsim_set_peers(self._h, C.create_string_buffer(raw, len(raw))))

    # ---- readable state (src/Softbody.js:12-20) ----
    def _fetch3(self, name, fn):
        if name not in self._cache:
            out = np.empty(3 * self.numParticles, np.float32)
            check(fn(self._h, ptr(out)))
            self._cache[name] = out
        return self._cache[name]

    @property
    def pos(self):
        return self._fetch3("pos", _capi.lib().tetsim_get_positions)

    @property
    def prevPos(self):
        return self._fetch3("prev", _capi.lib().tetsim_get_prev_positions)

    @property
    def vel(self):
        return self._fetch3("vel", _capi.lib().tetsim_get_velocities)

    def _rest(self):
        if getattr(self, "_rest_cache", None) is None:  # static after construction (initPhysics)
            q = np.empty(9 * self.numElems, np.float32)
            r = np.empty(self.numElems, np.float32)
            m = np.empty(self.numParticles, np.float32)
            check(_capi.lib().tetsim_get_rest(self._h, ptr(q), ptr(r), ptr(m)))
            self._rest_cache = (q, r, m)
        return self._rest_cache

    @property
    def invRestPose(self):
        return self._rest()[0]

    @property
    def invRestVolume(self):
        return self._rest()[1]

    @property
    def invMass(self):
        return self._rest()[2]

    @property
    def volError(self):
        out = C.c_double(0.0)
        check(_capi.lib().tetsim_get_vol_error(self._h, C.byref(out)))
        return out.value

    @property
    def resident(self):
        out = np.zeros(self.numParticles, np.uint8)
        check(_capi.lib().tetsim_get_resident(self._h, ptr(out)))
        return out.astype(bool)

    def set_state(self, pos=None, prevPos=None, vel=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32).reshape(-1) for a in (pos, prevPos, vel)]
        for a in arrs:
            if a is not None and a.size != 3 * self.numParticles:
                raise ValueError("state arrays must hold 3 floats per particle")
        self._cache.clear()
        check(_capi.lib().tetsim_set_state(self._h, *[ptr(a) for a in arrs]))
        self.synchronize()

    # ---- rank-local state access (multi-GPU hosts): arrays in the handle's own vertex order ----
    @property
    def resident_ids(self):
        """Caller vertex id of every vertex this handle holds (-1: a replica this rank does not maintain)."""
        out = np.empty(self.info()["localVerts"], np.int32)
        check(_capi.lib().tetsim_get_resident_ids(self._h, ptr(out)))
        return out

    def set_state_resident(self, pos=None, prevPos=None, vel=None):
        n = 3 * self.info()["localVerts"]
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32).reshape(-1) for a in (pos, prevPos, vel)]
        for a in arrs:
            if a is not None and a.size != n:
                raise ValueError("resident state arrays must hold 3 floats per resident vertex")
        self._cache.clear()
        check(_capi.lib().tetsim_set_state_resident(self._h, *[ptr(a) for a in arrs]))
        self.synchronize()

    @property
    def pos_resident(self):
        out = np.empty(3 * self.info()["localVerts"], np.float32)
        check(_capi.lib().tetsim_get_positions_resident(self._h, ptr(out)))
        return out

    def nearest_vertex(self, pos):
        """(caller vertex id, f64 squared distance) of the nearest vertex this rank maintains; (-1, inf) if none."""
        p = self._xyz(pos)
        gid, d2 = C.c_int32(-1), C.c_double(0.0)
        check(_capi.lib().tetsim_nearest_vertex(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), C.byref(gid), C.byref(d2)))
        return gid.value, d2.value

    def set_grab(self, grab_id, pos):
        p = self._xyz(pos)
        check(_capi.lib().tetsim_set_grab(self._h, int(grab_id), p.ctypes.data_as(C.POINTER(C.c_double))))
        self.grabId = int(grab_id)
        self.grabPos[:] = p

    def info(self) -> dict:
        i = _capi.TetSimInfo()
        check(_capi.lib().tetsim_get_info(self._h, C.byref(i)))
        return i.as_dict()

    def time_kernel(self, reps: int = 10):
        """(mean ms per launch, algorithmic bytes per launch) of the dominant kernel; see tetsim_time_kernel."""
        ms, nbytes = C.c_double(0.0), C.c_int64(0)
        check(_capi.lib().tetsim_time_kernel(self._h, int(reps), C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    # ---- frame end / render buffers (src/Softbody.js:244-277) ----
    def endFrame(self):
        self.updateEdgeMesh()
        self.updateVisMesh()

    def updateEdgeMesh(self):
        self.edgeMesh.positions[:] = self.pos

    def updateVisMesh(self):
        if not self.numVisVerts:
            return
        want_normals = self.visTriIds is not None and self.physicsParams.get("computeNormals", True)
        check(_capi.lib().tetsim_skin(
            self._h, ptr(self.visVerts), self.numVisVerts, ptr(self.visTriIds) if want_normals else None,
            self.visTriIds.size // 3 if want_normals else 0, ptr(self.visMesh.positions),
            ptr(self.visMesh.normals) if want_normals else None))

    # ---- grab (src/Softbody.js:279-298) ----
    @staticmethod
    def _xyz(pos):
        if hasattr(pos, "x"):
            return np.array([pos.x, pos.y, pos.z], np.float64)
        if isinstance(pos, dict):
            return np.array([pos["x"], pos["y"], pos["z"]], np.float64)
        return np.asarray(pos, np.float64).reshape(3)

    def startGrab(self, pos):
        p = self._xyz(pos)
        gid = C.c_int32(-1)
        check(_capi.lib().tetsim_start_grab(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), C.byref(gid)))
        self.grabId = gid.value
        self.grabPos[:] = p

    def moveGrabbed(self, pos):
        p = self._xyz(pos)
        check(_capi.lib().tetsim_move_grabbed(self._h, p.ctypes.data_as(C.POINTER(C.c_double))))
        self.grabPos[:] = p

    def endGrab(self):
        check(_capi.lib().tetsim_end_grab(self._h))
        self.grabId = -1


class SoftBody(_Body):
    """Drop-in for ``new SoftBody(vertices, tetIds, tetEdgeIds, physicsParams, visVerts, visTriIds, visMaterial)``.

    solver: "gs_exact" (default; the reference's in-order Gauss-Seidel as a dependency-level schedule),
    "gs_color" (greedy graph colouring), "jacobi" (clustered Jacobi, the throughput mode).
    arithmetic: "fast" (f32/FMA) or "bitexact" (f64 expressions + f32 stores, the reference's JS arithmetic).
    """
    _default_solver = "gs_exact"


class SoftBodyGPU(_Body):
    """Drop-in for ``new SoftBodyGPU(..., visMaterial, world)``: the polar-decomposition Jacobi solver."""
    _default_solver = "polar"

    def __init__(self, vertices, tetIds, tetEdgeIds, physicsParams, visVerts=None, visTriIds=None, visMaterial=None,
                 world=None, **kw):
        kw.setdefault("solver", "polar")
        super().__init__(vertices, tetIds, tetEdgeIds, physicsParams, visVerts, visTriIds, visMaterial, world, **kw)
        # the constructor ends with computeVertexNormals() + updateVisMesh() on the rest pose (src/SoftbodyGPU.js:484-485):
        # those normals stay in the geometry as `objectNormal` and the vertex shader rotates them by the tets' quaternions
        self.restNormals = self.visMesh.normals.copy()

    def simulate(self, dt, physicsParams=None):
        pp = physicsParams if physicsParams is not None else self.physicsParams
        pp["dt"] = dt  # src/SoftbodyGPU.js:611 mutates the shared params object
        super().simulate(dt, pp)

    def endFrame(self):  # src/SoftbodyGPU.js:643-647: only toggles the edge mesh; the vis mesh is skinned at render time
        self.edgeMesh.visible = bool(self.physicsParams.get("ShowTetMesh", False))

    def renderVisMesh(self):
        """What the vis material's patched vertex shader does at render time (src/SoftbodyGPU.js:424-448): barycentric
        skinning in f32 and normal = Rotate(rest normal, tet quaternion); fills visMesh.positions / visMesh.normals."""
        if not self.numVisVerts:
            return
        want_n = self.visTriIds is not None
        check(_capi.lib().tetsim_skin_gpu(self._h, ptr(self.visVerts), self.numVisVerts, ptr(self.restNormals) if want_n else None,
                                          ptr(self.visMesh.positions), ptr(self.visMesh.normals) if want_n else None))

    def readToCPU(self, variable="pos", buffer=None):
        """src/SoftbodyGPU.js:649-653; returns RGBA-strided floats like readRenderTargetPixels."""
        src = {"pos": self.pos, "prevPos": self.prevPos, "vel": self.vel}[variable]
        out = buffer if buffer is not None else np.zeros(4 * self.numParticles, np.float32)
        out.reshape(-1, 4)[: self.numParticles, :3] = src.reshape(-1, 3)
        return out

    @property
    def quats(self):
        q = np.empty(4 * self.numElems, np.float32)
        check(_capi.lib().tetsim_get_polar_state(self._h, None, ptr(q)))
        return q

    @property
    def elems(self):
        r = np.empty(12 * self.numElems, np.float32)
        check(_capi.lib().tetsim_get_polar_state(self._h, ptr(r), None))
        return r
