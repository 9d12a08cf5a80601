"""The oracle pinned to the REFERENCE'S OWN TEXT.

tools/transpile_reference.py mechanically re-emits /root/reference/src/Softbody.js (class SoftBody, every method) and
SoftBodyGPU.initPhysics as Python under JS number semantics (oracle/jsrt.py); tools/make_ref_golden.py ran the
scenarios of oracle/ref_scenarios.py on that code and committed tests/golden/ref_golden.npz.

  * CPU (here): the C restatement oracle/softbody_oracle.c must reproduce every stored array BIT FOR BIT
    (positions, prevPos, velocities, volError, grabId over free fall / contact + clamp / compliance / grab scenarios,
    the initPhysics arrays, skinning) and the polar oracle's reverse table must equal the reference's 9 x RGBA tables.
  * CPU, when /root/reference is present (this container): the transpile is re-run from scratch and executed live
    against the oracle, and its first checkpoints must equal the committed fixture (the fixture is not stale).
  * The WebGL solver likewise: JavaScript (initPhysics, simulate, the GPGPU runtime's compute) AND the seven GLSL passes are
    transpiled and executed texel by texel; the polar oracle reproduces those vectors bit for bit.
  * GPU: tests/test_parity_gpu.py::test_cuda_bitexact_reproduces_the_transpiled_reference (and ..._webgl_solver) run the same
    scenarios through the C ABI.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import ref_runner, ref_scenarios
from tetsim_b200 import mesh
from util import assert_bit_equal

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))


def check_against_golden(sc, step, d):
    n = sc["name"]
    assert_bit_equal(d["pos"], GOLD["%s_pos_%d" % (n, step)], "%s pos @%d" % (n, step))
    assert_bit_equal(d["prev"], GOLD["%s_prev_%d" % (n, step)], "%s prevPos @%d" % (n, step))
    assert_bit_equal(d["vel"], GOLD["%s_vel_%d" % (n, step)], "%s vel @%d" % (n, step))
    if "volError" in d:
        assert float(d["volError"]) == float(GOLD["%s_volError_%d" % (n, step)]), (n, step)
    assert int(d["grabId"]) == int(GOLD["%s_grabId_%d" % (n, step)]), (n, step)


@pytest.mark.parametrize("sc", ref_scenarios.SCENARIOS, ids=[s["name"] for s in ref_scenarios.SCENARIOS])
def test_c_oracle_reproduces_the_transpiled_reference(dragon, sc):
    body = oracle.SoftBodyOracle(ref_scenarios.shifted(dragon["tet_verts"], sc["shift"]), dragon["tet_ids"])
    seen = []
    ref_scenarios.run(sc, body, lambda b: dict(pos=b.pos, prev=b.prevPos, vel=b.vel, volError=b.volError, grabId=b.grabId),
                      lambda step, d: (seen.append(step), check_against_golden(sc, step, d)))
    assert tuple(seen) == tuple(sc["save"])


def test_c_oracle_init_physics_and_skinning_equal_the_transpiled_reference(dragon):
    o = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    assert_bit_equal(o.invRestPose, GOLD["invRestPose"], "invRestPose")
    assert_bit_equal(o.invRestVolume, GOLD["invRestVolume"], "invRestVolume")
    assert_bit_equal(o.invMass, GOLD["invMass"], "invMass")
    assert_bit_equal(oracle.skin(dragon["vis_verts"], dragon["tet_ids"], o.pos), GOLD["vis_pos_0"], "updateVisMesh @0")
    assert_bit_equal(oracle.skin(dragon["vis_verts"], dragon["tet_ids"], GOLD["free100_pos_100"]), GOLD["vis_pos_100"], "updateVisMesh @100")
    # updateEdgeMesh copies pos into the edge mesh buffer, which ALIASES the caller's `vertices` (src/Softbody.js:37,252-254)
    assert_bit_equal(GOLD["edge_pos_100"], GOLD["free100_pos_100"], "updateEdgeMesh")
    assert_bit_equal(GOLD["caller_vertices_100"], GOLD["free100_pos_100"], "aliasing quirk")


def test_polar_oracle_tables_equal_the_transpiled_initPhysics(dragon):
    """SoftBodyGPU.initPhysics (src/SoftbodyGPU.js:487-608): reverse tables (slot rule `<= 0.0`, :568), elems0, quats0,
    masses and volumes -- the polar oracle's CSR table must list, per particle, exactly the reference's slots in scan
    order (table 0 channel 0..3, table 1 ...) up to the first -1 (:306-318)."""
    N, M = dragon["tet_verts"].size // 3, dragon["tet_ids"].size // 4
    W = int(GOLD["gpu_texDim"])
    assert W == 62 and int(GOLD["gpu_biggestT"]) == 7   # console.log(biggestT), :607: max valence 32 -> tables 0..7
    tables = np.stack([GOLD["gpu_table_%d" % k].reshape(W * W, 4) for k in range(9)], axis=1).reshape(W * W, 36)
    po = oracle.PolarOracle(dragon["tet_verts"], dragon["tet_ids"], reference_table_bug=True)
    for v in range(N):
        row = tables[v]
        stop = np.flatnonzero(row == -1.0)
        n = int(stop[0]) if stop.size else 36
        mine = po.tblEntries[po.tblStart[v]:po.tblStart[v + 1]]
        assert n == mine.size and np.array_equal(row[:n].astype(np.int64), mine), v
    assert np.all(tables[N:] == -1.0)
    # the quirk itself: encoded id 0 (tet 0, corner 0 -> particle tetIds[0]) is overwritten by that particle's next corner
    p0 = int(dragon["tet_ids"][0])
    assert 0 not in set(po.tblEntries[po.tblStart[p0]:po.tblStart[p0 + 1]].tolist())
    # goal corners, quaternions, volumes, masses
    ids = dragon["tet_ids"].reshape(-1, 4)
    for k in range(4):
        e = GOLD["gpu_elems0_%d" % k].reshape(-1, 4)[:M, :3]
        assert_bit_equal(e, po.rest.reshape(M, 4, 3)[:, k, :], "elems0[%d]" % k)
        assert_bit_equal(e, dragon["tet_verts"].reshape(-1, 3)[ids[:, k]], "elems0[%d] = rest corner" % k)
    assert_bit_equal(GOLD["gpu_quats0"].reshape(-1, 4)[:M], po.quat.reshape(M, 4), "quats0")
    assert_bit_equal(GOLD["gpu_invRestVolumeAndColor"].reshape(-1, 4)[:M, 0], po.invRestVolume, "invRestVolume")
    assert np.all(GOLD["gpu_invRestVolumeAndColor"].reshape(-1, 4)[:M, 1] == -1.0)       # colour "undefined", :590
    assert_bit_equal(GOLD["gpu_invMass"].reshape(-1, 4)[:N, 0], GOLD["invMass"], "invMass (same as the CPU class)")
    assert_bit_equal(GOLD["gpu_pos0"].reshape(-1, 4)[:N, :3], dragon["tet_verts"].reshape(-1, 3), "pos0")
    assert np.array_equal(GOLD["gpu_elemToParticlesTable"].reshape(-1, 4)[:M].astype(np.int64), ids)


@pytest.mark.skipif(not ref_runner.reference_present(), reason="/root/reference is only present in the build container")
def test_live_transpile_matches_oracle_and_fixture(dragon):
    """Re-run the transpiler on the reference's text and execute it: equals the C oracle bit for bit, and the committed
    fixture's first checkpoints (so the fixture is what this code produces)."""
    assert ref_runner.ensure()
    for sc in ref_scenarios.SCENARIOS:
        short = dict(sc, steps=min(sc["steps"], 6 if sc["name"] != "grab30" else 12), save=tuple(s for s in sc["save"] if s <= 6) or (5,))
        if sc["name"] == "grab30":
            short["save"] = (5, 12)
        v = ref_scenarios.shifted(dragon["tet_verts"], sc["shift"])
        ref = ref_runner.RefSoftBody(v, dragon["tet_ids"], sc["params"])
        orc = oracle.SoftBodyOracle(v, dragon["tet_ids"])
        got_r, got_o = {}, {}
        ref_scenarios.run(short, ref, lambda b: dict(pos=b.pos, prev=b.prevPos, vel=b.vel, volError=b.volError, grabId=b.grabId),
                          lambda s, d: got_r.__setitem__(s, d))
        ref_scenarios.run(short, orc, lambda b: dict(pos=b.pos.copy(), prev=b.prevPos.copy(), vel=b.vel.copy(), volError=b.volError, grabId=b.grabId),
                          lambda s, d: got_o.__setitem__(s, d))
        assert got_r.keys() == got_o.keys() and got_r
        for s in got_r:
            for k in ("pos", "prev", "vel"):
                assert_bit_equal(got_o[s][k], got_r[s][k], "%s %s @%d" % (sc["name"], k, s))
            assert got_o[s]["volError"] == got_r[s]["volError"] and got_o[s]["grabId"] == got_r[s]["grabId"]
            if s in sc["save"]:
                check_against_golden(sc, s, got_r[s])
    g = ref_runner.RefSoftBodyGPUInit(dragon["tet_verts"], dragon["tet_ids"])
    for k in range(9):
        assert np.array_equal(g.tex("particleToElemVertsTable", k), GOLD["gpu_table_%d" % k])


def _polar_cases(dragon):
    return ref_scenarios.polar_scenarios(dragon["tet_verts"], dragon["tet_ids"], mesh)


@pytest.mark.parametrize("which", [0, 1], ids=["polar_dragon", "polar_beam"])
def test_polar_oracle_reproduces_the_transpiled_webgl_solver(dragon, which):
    """The WHOLE WebGL substep executed from the reference's text -- SoftBodyGPU.initPhysics + simulate and the GPGPU
    runtime's addVariable / addPass / compute transpiled from JavaScript, the seven passes transpiled from GLSL
    (tools/transpile_shaders.py), run texel by texel -- produced these vectors; the C restatement (oracle/polar_oracle.c)
    must reproduce positions, prevPos, velocities, quaternions and goal corners BIT FOR BIT, free fall and floor contact."""
    name, (v, t), params, steps, save = _polar_cases(dragon)[which]
    po = oracle.PolarOracle(v, t, reference_table_bug=True, **params)
    for s in range(1, steps + 1):
        po.simulate(ref_scenarios.FRAME_DT / 20, params)
        if s in save:
            for k, a in dict(pos=po.pos, prev=po.prevPos, vel=po.vel, quat=po.quat, rest=po.rest).items():
                assert_bit_equal(a, GOLD["%s_%s_%d" % (name, k, s)], "%s %s @%d" % (name, k, s))
    if which == 1:
        assert np.any(po.pos[1::3] == 0.0), "the beam is meant to reach the floor"


@pytest.mark.skipif(not ref_runner.reference_present(), reason="/root/reference is only present in the build container")
def test_live_transpiled_webgl_solver_matches_polar_oracle(dragon):
    """Re-run both transpilers (JavaScript and GLSL) on the reference's text and execute the result: a 12-tet body for 12
    substeps equals the polar oracle bit for bit, and the first checkpoint of the committed beam fixture is reproduced."""
    assert ref_runner.ensure()
    v, t = mesh.make_beam((2, 1, 1), h=0.25, y0=0.004)
    v = v.copy()
    v[1::3] += np.float32(0.01) * np.arange(v.size // 3, dtype=np.float32)
    params = dict(ref_scenarios.DEFAULTS)
    ref = ref_runner.RefSoftBodyGPU(v, t, params)
    po = oracle.PolarOracle(v, t, reference_table_bug=True)
    for _ in range(12):
        ref.simulate(ref_scenarios.FRAME_DT / 20, params)
        po.simulate(ref_scenarios.FRAME_DT / 20)
    for k, a, b in (("pos", ref.pos, po.pos), ("prev", ref.prevPos, po.prevPos), ("vel", ref.vel, po.vel), ("quat", ref.quat, po.quat),
                    ("rest", ref.rest, po.rest)):
        assert_bit_equal(a, b, k)
    name, (bv, bt), bp, _, _ = _polar_cases(dragon)[1]
    ref = ref_runner.RefSoftBodyGPU(bv, bt, bp)
    ref.simulate(ref_scenarios.FRAME_DT / 20, bp)
    assert_bit_equal(ref.pos, GOLD["polar_beam_pos_1"], "fixture pos @1")
    assert_bit_equal(ref.rest, GOLD["polar_beam_rest_1"], "fixture rest @1")


def test_transpiler_rejects_what_it_does_not_understand(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools"))
    import transpile_reference as tr
    for bad in ("f() { while (1) { } }", "f() { let a = 1; for (let i = 0; i < 2; i++) { let a = 2; } }", "f(a) { a[0] = b = 1; }",
                "f(a) { x = ++a; }", "f() { for (;;) { continue; } }"):
        p = tmp_path / "x.js"
        p.write_text("class K {\n %s \n}\n" % bad)
        with pytest.raises(tr.Unsupported):
            tr.transpile_class(str(p), "K")
    import transpile_shaders as tsh
    for bad in ("void main() { while (true) { } }", "void main() { float a = b ? 1.0 : ; }", "void main() { discard; }", "struct S { float a; };"):
        with pytest.raises(tsh.Unsupported):
            tsh.transpile_shader("x", bad, [])
    # and the evaluation-order rule that matters (src/Softbody.js:350-355: dst[dnr] = f(dst[dnr++]))
    p = tmp_path / "y.js"
    p.write_text("class K {\n f(d, n) { d[n] = 10 + d[n++]; return n; }\n}\n")
    text, _ = tr.transpile_class(str(p), "K")
    ns = {}
    exec("from oracle.jsrt import *\n" + text, ns)
    d = ns["JSArray"]([1.0, 2.0])
    assert ns["K"]().f(d, 0) == 1 and list(d) == [11.0, 2.0]
