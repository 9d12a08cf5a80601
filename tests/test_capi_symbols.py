"""No-GPU checks of the drop-in boundary: the library loads, exports every symbol the header declares,
its host-only tools work, and anything that needs a device fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import tetsim_b200 as ts
from tetsim_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tetsim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tetsim_[a-z0-9_]+)\s*\(", text)))


def test_header_python_and_library_agree():
    hdr = header_symbols()
    assert hdr == sorted(_capi.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (tetsim_[a-z0-9_]+)", out)))
    assert exported == hdr
    L = _capi.lib()
    for s in hdr:
        assert getattr(L, s) is not None
    assert L.tetsim_version() == 100


def test_struct_layouts_match_header():
    assert C.sizeof(_capi.TetSimParams) == 11 * 8
    assert C.sizeof(_capi.TetSimOptions) == 12 * 4 + 2 * 8
    assert C.sizeof(_capi.TetSimInfo) == 16 * 4 + 4 * 8 + 2 * 4
    p = _capi.default_params()
    assert (p.gravity, p.friction, p.density, p.devCompliance, p.volCompliance) == (-9.81, 1000.0, 1000.0, 1e-5, 0.0)
    assert list(p.worldBounds) == [-2.5, -1.0, -2.5, 2.5, 10.0, 2.5]      # src/main.js:32
    o = _capi.default_options()
    assert (o.solver, o.arithmetic, o.iters, o.deterministic, o.referenceTableBug, o.clusterSize, o.worldSize) == (
        0, 0, 1, 1, 1, 256, 1)


def test_sass_is_sm100a_without_tensor_or_cas_loops():
    """The hot kernels are plain FP32 + 128-bit memory ops: no tensor-core MMA (there is no dense
    contraction on this path) and no shared-memory CAS spin loops in the clustered Jacobi kernel."""
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    # the default tile kernels: T = 512 (64 registers, 4 CTAs of 256 threads per SM) and T = 256, two tets per thread, two stages
    res = subprocess.run(["cuobjdump", "-res-usage", _capi.LIB_PATH], capture_output=True, text=True).stdout
    names = re.findall(r"Function (\S+):", res)
    for prefix in ("_ZN4tsim15k_jacobi_tilesNILi512ELi2ELi2ELi4ELb0E", "_ZN4tsim15k_jacobi_tilesNILi256ELi2ELi2ELi0ELb0E"):
        fun = [n for n in names if n.startswith(prefix)]
        assert len(fun) == 1, (prefix, fun)
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun[0], _capi.LIB_PATH], capture_output=True, text=True).stdout
        # TMA bulk copies on mbarriers, bulk L2 prefetch, cp.async gathers, 128-bit shared-memory traffic
        assert "UBLKCP" in sass and "UBLKPF" in sass and "SYNCS" in sass and "LDGSTS.E.BYPASS.128" in sass, fun
        assert "STS.128" in sass and "LDS.128" in sass and "STG.E.128" in sass, fun
        assert "ATOMS.CAST" not in sass and "HMMA" not in sass and "UTCHMMA" not in sass and "CALL" not in sass, fun
    regs = int(re.search(r"REG:(\d+)", res[res.index("k_jacobi_tilesNILi512ELi2ELi2ELi4ELb0E"):][:200]).group(1))
    assert regs <= 64 and "STACK:0" in res[res.index("k_jacobi_tilesNILi512ELi2ELi2ELi4ELb0E"):][:200], regs   # 4 x 256 threads resident, no spills
    # the single-GPU kernels carry no peer-exchange code (volatile 128-bit remote stores exist only in the PEER variants)
    peer = [n for n in names if n.startswith("_ZN4tsim15k_jacobi_tilesNILi512ELi2ELi2ELi4ELb1E")]
    assert len(peer) == 1
    psass = subprocess.run(["cuobjdump", "-sass", "-fun", peer[0], _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "STG.E.128.STRONG.SYS" in psass and "STRONG.SYS" not in sass


@pytest.mark.skipif(_capi.lib().tetsim_device_count() > 0, reason="a B200 is present")
def test_no_cpu_fallback(dragon):
    with pytest.raises(ts.TetSimError) as e:
        ts.SoftBody(dragon["tet_verts"], dragon["tet_ids"], dragon["tet_edge_ids"], None)
    assert e.value.code == _capi.E_CUDA
    assert "no CPU path" in str(e.value)


def test_malformed_meshes_rejected_before_touching_a_device(dragon):
    t = dragon["tet_ids"].copy()
    t[7] = -3
    with pytest.raises(ts.TetSimError) as e:
        ts.SoftBody(dragon["tet_verts"], t, None, None)
    assert e.value.code == _capi.E_INVALID
    with pytest.raises(ValueError):
        ts.SoftBody(dragon["tet_verts"][:-1], dragon["tet_ids"], None, None)
    with pytest.raises(ts.TetSimError):
        ts.SoftBody(dragon["tet_verts"], dragon["tet_ids"], None, None, solver="jacobi", cluster_size=100)


def test_level_schedule_and_colouring_host_tools(dragon):
    from oracle import oracle_np
    level, n = ts.level_schedule(dragon["tet_ids"], 1234)
    assert n == 703 and np.bincount(level).max() == 22
    ref_levels = oracle_np.level_schedule(1234, dragon["tet_ids"])
    order = np.argsort(level, kind="stable")
    assert np.array_equal(order, np.concatenate(ref_levels))
    ids = dragon["tet_ids"].reshape(-1, 4)
    for lv in ref_levels[:50]:                       # tets of one level share no vertex
        assert len(np.unique(ids[lv])) == 4 * len(lv)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "softbody_golden.npz"))["greedy_color"]
    color, ncol = ts.greedy_colors(dragon["tet_ids"], 1234)
    assert ncol == 32 and np.array_equal(color, gold)
    sizes = np.bincount(color)
    assert sizes[0] == 228 and sizes[-1] == 2
    for c in range(ncol):
        sel = ids[color == c]
        assert len(np.unique(sel)) == sel.size


def test_mesh_generators():
    from tetsim_b200 import mesh
    v, t = mesh.make_beam((6, 3, 2), h=0.5, jitter=0.2)
    x = v.reshape(-1, 3).astype(np.float64)
    tt = t.reshape(-1, 4)
    vol = np.linalg.det(x[tt[:, 1:]] - x[tt[:, :1]]) / 6
    assert len(tt) == 6 * 36 and np.all(vol > 0)
    assert abs(vol.sum() - 36 * 0.125) < 1e-5        # jitter moves interior vertices only
    d = mesh.load_dragon()
    v2, t2 = mesh.tile_bodies(d["tet_verts"], d["tet_ids"], 8, 8, y_shift=-0.4)
    assert t2.size // 4 == 245760 and v2.size // 3 == 78976          # BASELINE config 5 (64 copies)
    assert t2.max() == 78975


def test_js_host_shim_typechecks_and_matches_the_wrapper():
    """N4 (JS host integration) cannot be executed here -- no Node, no JS engine -- so it is checked statically:
    the N-API shim compiles (-fsyntax-only) against the documented Node-API signatures, every native.* call of the
    wrapper classes is exported by the shim, every C entry point the shim calls is declared by the header, and the
    wrapper exposes every member the reference's callers touch (src/main.js:53-68,82,88; src/Softbody.js:445-469;
    src/SoftbodyGPU.js:792-824)."""
    js = os.path.join(ROOT, "tetsim_b200", "js")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-DTETSIM_NAPI_TYPECHECK",
                        "-I" + os.path.join(ROOT, "include"), "-I" + js, os.path.join(js, "tetsim_napi.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    shim = open(os.path.join(js, "tetsim_napi.cc")).read()
    mjs = open(os.path.join(js, "softbody.mjs")).read()
    exported = set(re.findall(r'\{"(\w+)", nullptr,', shim))
    used = set(re.findall(r"native\.(\w+)", mjs))
    assert used and used <= exported, used - exported
    header = open(os.path.join(ROOT, "include", "tetsim_b200.h")).read()
    for sym in set(re.findall(r"\b(tetsim_\w+)\s*\(", shim)) - {"tetsim_napi"}:
        assert re.search(r"\b%s\s*\(" % sym, header), sym
    for member in ("edgeMesh", "visMesh", "userData = this", "updateEdgeMesh()", "updateVisMesh()", "endFrame()", "simulate(dt, physicsParams)",
                   "startGrab(pos)", "moveGrabbed(pos)", "endGrab()", "readToCPU(variable, buffer)", "get pos()", "get prevPos()", "get vel()",
                   "get invMass()", "get invRestPose()", "get invRestVolume()", "get volError()", "numParticles", "numElems", "grabId", "grabPos",
                   "export class SoftBody ", "export class SoftBodyGPU "):
        assert member in mjs, member
    # braces / parentheses balance (the cheapest syntax check available without an engine)
    code = re.sub(r"//[^\n]*", "", mjs)
    for a, b in ("{}", "()", "[]"):
        assert code.count(a) == code.count(b), (a, code.count(a), code.count(b))
