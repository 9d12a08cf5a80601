"""ctypes binding of libtetsim_b200.so -- one Python function per entry point of include/tetsim_b200.h.

This is the stand-in for the N-API shim a JavaScript host would use (tetsim_b200/js/):
no JS engine exists in this image, so the tests and the benchmark drive the same C ABI from Python.
There is no fallback: if the shared library is missing or no B200 is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtetsim_b200.so")

# enum TetSimSolver / TetSimArithmetic
NH_GS_EXACT, NH_GS_COLOR, NH_JACOBI, POLAR_JACOBI = 0, 1, 2, 3
ARITH_FAST_F32, ARITH_BITEXACT = 0, 1

E_INVALID, E_CUDA, E_NCCL, E_STATE, E_NOMEM = -1, -2, -3, -4, -5
PEER_BLOB_BYTES = 128  # TETSIM_PEER_BLOB_BYTES

# every symbol include/tetsim_b200.h declares (tests/test_capi_symbols.py checks the header against this)
SYMBOLS = [
    "tetsim_last_error", "tetsim_version", "tetsim_device_count", "tetsim_default_params",
    "tetsim_default_options", "tetsim_create", "tetsim_destroy", "tetsim_simulate", "tetsim_step",
    "tetsim_synchronize", "tetsim_get_positions", "tetsim_get_prev_positions", "tetsim_get_velocities",
    "tetsim_get_resident", "tetsim_set_state", "tetsim_get_rest", "tetsim_get_vol_error",
    "tetsim_get_polar_state", "tetsim_start_grab", "tetsim_move_grabbed", "tetsim_end_grab", "tetsim_skin",
    "tetsim_get_info", "tetsim_time_kernel", "tetsim_nccl_unique_id", "tetsim_get_ipc_handle", "tetsim_set_peers",
    "tetsim_level_schedule", "tetsim_greedy_colors", "tetsim_plan_partition", "tetsim_plan_halo",
    "tetsim_get_positions_async", "tetsim_wait_positions", "tetsim_get_resident_ids", "tetsim_set_state_resident",
    "tetsim_get_positions_resident", "tetsim_get_positions_resident_async", "tetsim_nearest_vertex", "tetsim_set_grab",
    "tetsim_connected_components", "tetsim_skin_gpu",
]


class TetSimError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("tetsim error %d: %s" % (code, message))
        self.code = code


class TetSimParams(C.Structure):
    _fields_ = [
        ("gravity", C.c_double),
        ("friction", C.c_double),
        ("density", C.c_double),
        ("devCompliance", C.c_double),
        ("volCompliance", C.c_double),
        ("worldBounds", C.c_double * 6),
    ]


class TetSimOptions(C.Structure):
    _fields_ = [
        ("solver", C.c_int32),
        ("arithmetic", C.c_int32),
        ("iters", C.c_int32),
        ("deterministic", C.c_int32),
        ("referenceTableBug", C.c_int32),
        ("reorder", C.c_int32),
        ("clusterSize", C.c_int32),
        ("trackVolError", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("worldSize", C.c_int32),
        ("exchange", C.c_int32),
        ("stream", C.c_void_p),
        ("ncclUniqueId", C.c_void_p),
    ]


class TetSimInfo(C.Structure):
    _fields_ = [
        ("numVerts", C.c_int32), ("numTets", C.c_int32),
        ("solver", C.c_int32), ("arithmetic", C.c_int32), ("iters", C.c_int32),
        ("numLevels", C.c_int32), ("maxLevelSize", C.c_int32), ("numComponents", C.c_int32),
        ("bodyKernel", C.c_int32), ("numClusters", C.c_int32), ("clusterSize", C.c_int32),
        ("localTets", C.c_int32), ("localVerts", C.c_int32), ("boundaryVerts", C.c_int32),
        ("maxValence", C.c_int32), ("launchesPerSubstep", C.c_int32),
        ("deviceBytes", C.c_int64), ("sumLocalVerts", C.c_int64), ("kernelLaunches", C.c_int64),
        ("tileMetaBytes", C.c_int64), ("maxTileVerts", C.c_int32), ("boundaryTiles", C.c_int32),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_lib = None


def lib() -> C.CDLL:
    """Load the shared library; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libtetsim_b200.so is not built: run `python -m tetsim_b200.build` (nvcc, sm_100a). "
            "tetsim_b200 has no CPU or PyTorch fallback path.")
    L = C.CDLL(LIB_PATH)
    vp, f32p, i32p, dblp = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.tetsim_last_error.restype = C.c_char_p
    L.tetsim_version.restype = C.c_int
    L.tetsim_device_count.restype = C.c_int
    L.tetsim_default_params.argtypes = [C.POINTER(TetSimParams)]
    L.tetsim_default_params.restype = None
    L.tetsim_default_options.argtypes = [C.POINTER(TetSimOptions)]
    L.tetsim_default_options.restype = None
    L.tetsim_create.argtypes = [f32p, C.c_int32, i32p, C.c_int32, C.POINTER(TetSimParams), C.POINTER(TetSimOptions),
                                C.POINTER(vp)]
    L.tetsim_destroy.argtypes = [vp]
    L.tetsim_destroy.restype = None
    L.tetsim_simulate.argtypes = [vp, C.c_double, C.POINTER(TetSimParams)]
    L.tetsim_step.argtypes = [vp, C.c_double, C.c_int32, C.POINTER(TetSimParams)]
    L.tetsim_synchronize.argtypes = [vp]
    for name in ("tetsim_get_positions", "tetsim_get_prev_positions", "tetsim_get_velocities"):
        getattr(L, name).argtypes = [vp, vp]
    L.tetsim_get_resident.argtypes = [vp, vp]
    for name in ("tetsim_get_positions_async", "tetsim_get_positions_resident", "tetsim_get_positions_resident_async",
                 "tetsim_get_resident_ids"):
        getattr(L, name).argtypes = [vp, vp]
    L.tetsim_wait_positions.argtypes = [vp]
    L.tetsim_set_state_resident.argtypes = [vp, vp, vp, vp]
    L.tetsim_nearest_vertex.argtypes = [vp, dblp, i32p, dblp]
    L.tetsim_set_grab.argtypes = [vp, C.c_int32, dblp]
    L.tetsim_set_state.argtypes = [vp, vp, vp, vp]
    L.tetsim_get_rest.argtypes = [vp, vp, vp, vp]
    L.tetsim_get_vol_error.argtypes = [vp, dblp]
    L.tetsim_get_polar_state.argtypes = [vp, vp, vp]
    L.tetsim_start_grab.argtypes = [vp, dblp, i32p]
    L.tetsim_move_grabbed.argtypes = [vp, dblp]
    L.tetsim_end_grab.argtypes = [vp]
    L.tetsim_skin.argtypes = [vp, vp, C.c_int32, vp, C.c_int32, vp, vp]
    L.tetsim_skin_gpu.argtypes = [vp, vp, C.c_int32, vp, vp, vp]
    L.tetsim_get_info.argtypes = [vp, C.POINTER(TetSimInfo)]
    L.tetsim_time_kernel.argtypes = [vp, C.c_int32, dblp, C.POINTER(C.c_int64)]
    L.tetsim_nccl_unique_id.argtypes = [vp]
    L.tetsim_get_ipc_handle.argtypes = [vp, vp]
    L.tetsim_set_peers.argtypes = [vp, vp]
    L.tetsim_level_schedule.argtypes = [i32p, C.c_int32, C.c_int32, i32p]
    L.tetsim_greedy_colors.argtypes = [i32p, C.c_int32, C.c_int32, i32p]
    L.tetsim_connected_components.argtypes = [i32p, C.c_int32, C.c_int32, i32p]
    L.tetsim_plan_partition.argtypes = [f32p, C.c_int32, i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        i32p, i32p, i32p]
    L.tetsim_plan_halo.argtypes = [f32p, C.c_int32, i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, i32p, i32p, i32p, i32p, i32p]
    _lib = L
    return L


def check(rc: int) -> int:
    if rc < 0:
        raise TetSimError(rc, lib().tetsim_last_error().decode("utf-8", "replace"))
    return rc


def default_params(**kw) -> TetSimParams:
    p = TetSimParams()
    lib().tetsim_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "worldBounds":
            for i, b in enumerate(v):
                p.worldBounds[i] = float(b)
        elif k in ("gravity", "friction", "density", "devCompliance", "volCompliance"):
            setattr(p, k, float(v))
    return p


def default_options(**kw) -> TetSimOptions:
    o = TetSimOptions()
    lib().tetsim_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def level_schedule(tet_ids, num_verts: int):
    ids = np.ascontiguousarray(tet_ids, np.int32).reshape(-1)
    level = np.zeros(ids.size // 4, np.int32)
    n = check(lib().tetsim_level_schedule(ids.ctypes.data_as(C.POINTER(C.c_int32)), ids.size // 4, num_verts,
                                          level.ctypes.data_as(C.POINTER(C.c_int32))))
    return level, n


def greedy_colors(tet_ids, num_verts: int):
    ids = np.ascontiguousarray(tet_ids, np.int32).reshape(-1)
    color = np.zeros(ids.size // 4, np.int32)
    n = check(lib().tetsim_greedy_colors(ids.ctypes.data_as(C.POINTER(C.c_int32)), ids.size // 4, num_verts,
                                         color.ctypes.data_as(C.POINTER(C.c_int32))))
    return color, n


def connected_components(tet_ids, num_verts: int):
    """(body of every vertex, number of bodies): connected components over shared vertices (host only, no GPU)."""
    ids = np.ascontiguousarray(tet_ids, np.int32).reshape(-1)
    comp = np.zeros(max(num_verts, 1), np.int32)
    n = check(lib().tetsim_connected_components(ids.ctypes.data_as(C.POINTER(C.c_int32)), ids.size // 4, num_verts,
                                                comp.ctypes.data_as(C.POINTER(C.c_int32))))
    return comp[:num_verts], n


def plan_partition(verts, tet_ids, cluster_size: int, reorder: bool, rank: int, world_size: int) -> dict:
    """Host-only view of the tet partition a multi-GPU Jacobi handle would use (no GPU needed)."""
    v = np.ascontiguousarray(verts, np.float32).reshape(-1)
    t = np.ascontiguousarray(tet_ids, np.int32).reshape(-1)
    n, m = v.size // 3, t.size // 4
    counts = np.zeros(4, np.int32)
    l2c = np.zeros(max(n, 1), np.int32)
    lt = np.zeros(max(m, 1), np.int32)
    i32p = C.POINTER(C.c_int32)
    check(lib().tetsim_plan_partition(v.ctypes.data_as(C.POINTER(C.c_float)), n, t.ctypes.data_as(i32p), m,
                                      int(cluster_size), int(bool(reorder)), int(rank), int(world_size),
                                      counts.ctypes.data_as(i32p), l2c.ctypes.data_as(i32p), lt.ctypes.data_as(i32p)))
    nloc = int(counts[1] + counts[2])
    return dict(localTets=lt[: counts[0]].copy(), numInterior=int(counts[1]), numBoundary=int(counts[2]),
                numClusters=int(counts[3]), localToCaller=l2c[:nloc].copy())


def plan_halo(verts, tet_ids, cluster_size: int, reorder: bool, rank: int, world_size: int) -> dict:
    """Host-only view of the neighbour lists of the halo / peer-memory exchange (no GPU needed)."""
    v = np.ascontiguousarray(verts, np.float32).reshape(-1)
    t = np.ascontiguousarray(tet_ids, np.int32).reshape(-1)
    cap = max(world_size, 1)
    i32p = C.POINTER(C.c_int32)
    arrs = [np.zeros(cap + 1, np.int32) for _ in range(5)]
    n = check(lib().tetsim_plan_halo(v.ctypes.data_as(C.POINTER(C.c_float)), v.size // 3, t.ctypes.data_as(i32p),
                                     t.size // 4, int(cluster_size), int(bool(reorder)), int(rank), int(world_size), cap,
                                     *[x.ctypes.data_as(i32p) for x in arrs]))
    peers, seg, off, tot, slot = arrs
    return dict(peers=peers[:n].copy(), segStart=seg[: n + 1].copy(), remoteOff=off[:n].copy(),
                remoteTotal=tot[:n].copy(), remoteSlot=slot[:n].copy())
