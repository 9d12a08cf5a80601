"""CPU tests of the oracle itself (no GPU).  The reference has no golden vectors and cannot run here
(PARITY UNPINNED); these tests pin the C restatement the only ways available:
  1. against the anchors of a third, separately written emulation made at survey time,
  2. bit-for-bit against the independently structured numpy restatement (oracle/oracle_np.py),
  3. against the committed golden vectors (regression),
  4. against analytic properties of the constraint projection.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import oracle_np
from util import DT600, DT1200, vec_rel_err

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "softbody_golden.npz"))


def test_survey_anchors(dragon):
    sb = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    assert list(sb.tetIds[:4]) == [1, 0, 2, 521]
    np.testing.assert_array_equal(sb.invRestPose[:9], np.float32(
        [-10.637138, 0.8368553, -3.1084132, 5.9744034, -1.666426, -5.4478583, -0.014100595, -10.181579, 0.26633632]))
    assert sb.invRestVolume[0] == np.float32(4695.7603)
    np.testing.assert_array_equal(sb.invMass[:4], np.float32([0.93524957, 0.76354, 1.4777923, 0.8246292]))
    want_vol = {1: -0.16658050127720742, 10: -0.15752738889947235, 100: -0.15737990687763845}
    want_sum = {1: 1300.5853506604035, 10: 1298.7850487290088, 100: 1132.864895039551}
    for s in range(1, 101):
        sb.simulate(DT600)
        if s in want_vol:
            assert sb.volError == want_vol[s]
            assert float(np.sum(sb.pos.astype(np.float64))) == want_sum[s]
    np.testing.assert_array_equal(sb.pos[:3], np.float32([-0.06708563, 1.1470823, -0.081903815]))


def test_two_restatements_agree_bitwise(dragon):
    sb = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    nb = oracle_np.SoftBodyNP(dragon["tet_verts"], dragon["tet_ids"])
    assert np.array_equal(sb.invRestPose, nb.Q.reshape(-1))
    assert np.array_equal(sb.invRestVolume, nb.irv)
    assert np.array_equal(sb.invMass, nb.inv_mass)
    assert len(nb.levels) == 703 and max(len(l) for l in nb.levels) == 22
    for s in range(1, 11):
        sb.simulate(DT600)
        nb.simulate(DT600)
        assert np.array_equal(sb.pos, nb.pos.reshape(-1)), s
        assert np.array_equal(sb.vel, nb.vel.reshape(-1)), s
        assert sb.volError == nb.volError
    assert np.array_equal(nb.pos.reshape(-1), GOLD["gs_pos_10"])


def test_two_restatements_agree_with_contact_and_params(dragon):
    v = dragon["tet_verts"].reshape(-1, 3).copy()
    v[:, 1] -= np.float32(0.4545)
    kw = dict(gravity=-20.0, friction=250.0, devCompliance=2e-5, volCompliance=1e-6,
              worldBounds=(-0.9, -1.0, -2.5, 0.8, 1.45, 0.3))
    sb = oracle.SoftBodyOracle(v, dragon["tet_ids"], **kw)
    nb = oracle_np.SoftBodyNP(v, dragon["tet_ids"], **kw)
    for s in range(8):
        sb.simulate(DT600)
        nb.simulate(DT600)
    assert np.any(sb.pos.reshape(-1, 3)[:, 1] == 0.0)
    assert np.array_equal(sb.pos, nb.pos.reshape(-1))
    assert np.array_equal(sb.vel, nb.vel.reshape(-1))


def test_golden_regression(dragon):
    sb = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    assert np.array_equal(sb.invRestPose, GOLD["invRestPose"])
    assert np.array_equal(sb.invMass, GOLD["invMass"])
    for s in range(1, 101):
        sb.simulate(DT600)
        if s in (1, 10, 100):
            assert np.array_equal(sb.pos, GOLD["gs_pos_%d" % s])
            assert np.array_equal(sb.vel, GOLD["gs_vel_%d" % s])
            assert sb.volError == float(GOLD["gs_vol_error_%d" % s])
    jb = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    po = oracle.PolarOracle(dragon["tet_verts"], dragon["tet_ids"])
    for s in range(1, 101):
        jb.simulate_jacobi(DT1200, 1)
        po.simulate(DT1200)
        if s in (10, 100):
            assert np.array_equal(jb.pos, GOLD["jacobi1_pos_%d" % s])
            assert np.array_equal(po.pos, GOLD["polar_bug1_pos_%d" % s])


def test_level_order_equals_sequential_order(dragon):
    """F4 of the survey: the dependency-level order is bit-identical to the in-order sweep."""
    nb_levels = oracle_np.level_schedule(1234, dragon["tet_ids"])
    order = np.concatenate(nb_levels).astype(np.int32)
    a = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    b = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    for _ in range(20):
        a.simulate(DT600)
        b.simulate(DT600, order=order)
    assert np.array_equal(a.pos, b.pos)
    # ...whereas another order (reverse) is a different algorithm, far outside the 1e-4 tolerance
    c = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    a2 = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    rev = np.arange(3840, dtype=np.int32)[::-1].copy()
    for _ in range(100):
        a2.simulate(DT600)
        c.simulate(DT600, order=rev)
    assert vec_rel_err(c.pos, a2.pos) > 1e-3


def test_projection_properties():
    """One tet: the projection conserves linear momentum (sum m_i dx_i = 0, since the gradients sum
    to zero) and drives det F toward 1 with the default volCompliance = 0."""
    v = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32)
    t = np.array([0, 1, 2, 3], np.int32)
    sb = oracle.SoftBodyOracle(v, t, gravity=0.0)
    np.testing.assert_allclose(sb.invRestPose, np.eye(3, dtype=np.float32).reshape(-1), atol=0)
    assert sb.invRestVolume[0] == np.float32(6.0)
    sb.pos[:] = (sb.pos.reshape(4, 3) * np.float32([1.3, 0.9, 1.1]) + np.float32([0, 1, 0])).reshape(-1)  # stretch, off the floor
    sb.prevPos[:] = sb.pos
    before = sb.pos.copy()
    sb.simulate(DT600)
    m = 1.0 / sb.invMass.astype(np.float64)
    dp = (sb.pos.astype(np.float64) - before.astype(np.float64)).reshape(4, 3)
    assert np.max(np.abs((m[:, None] * dp).sum(0))) < 1e-6 * m.sum()
    x = sb.pos.reshape(4, 3).astype(np.float64)
    J = np.linalg.det((x[1:] - x[0]).T)
    assert abs(J - 1.0) < abs(1.3 * 0.9 * 1.1 - 1.0) * 0.05


def test_jacobi_is_order_independent_up_to_rounding(dragon):
    a = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    perm = np.random.default_rng(0).permutation(3840)
    ids = dragon["tet_ids"].reshape(-1, 4)[perm].reshape(-1)
    b = oracle.SoftBodyOracle(dragon["tet_verts"], ids)
    for _ in range(10):
        a.simulate_jacobi(DT1200, 2)
        b.simulate_jacobi(DT1200, 2)
    assert vec_rel_err(a.pos, b.pos) < 1e-5


def test_polar_table_bug_only_touches_one_vertex(dragon):
    a = oracle.PolarOracle(dragon["tet_verts"], dragon["tet_ids"], reference_table_bug=True)
    b = oracle.PolarOracle(dragon["tet_verts"], dragon["tet_ids"], reference_table_bug=False)
    assert a.tblEntries.size == b.tblEntries.size - 1
    v = int(dragon["tet_ids"][0])          # tetIds[0] = 1: the vertex whose corner 0 of tet 0 is lost
    na = np.diff(a.tblStart)
    nb_ = np.diff(b.tblStart)
    assert np.flatnonzero(na != nb_).tolist() == [v]
    assert 0 not in a.tblEntries[a.tblStart[v]:a.tblStart[v + 1]]
    assert nb_.max() == 32 <= 36


def test_polar_free_fall_and_volume(dragon):
    po = oracle.PolarOracle(dragon["tet_verts"], dragon["tet_ids"])
    y0 = po.pos.reshape(-1, 3)[:, 1].min()
    for _ in range(120):
        po.simulate(DT1200)
    y = po.pos.reshape(-1, 3)[:, 1].min()
    t = 119 * DT1200  # gravity enters the velocity one substep late (src/SoftbodyGPU.js:367-371)
    assert abs((y0 - y) - 0.5 * 9.81 * t * t) < 0.01
    assert np.all(np.isfinite(po.pos))
    q = po.quat.reshape(-1, 4)
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-6)


def test_skin_and_normals_oracle(dragon):
    pos = dragon["tet_verts"]
    s = oracle.skin(dragon["vis_verts"], dragon["tet_ids"], pos).reshape(-1, 3)
    vv = dragon["vis_verts"].reshape(-1, 4).astype(np.float64)
    ids = dragon["tet_ids"].reshape(-1, 4)[vv[:, 0].astype(int)]
    b = np.concatenate([vv[:, 1:], 1.0 - vv[:, 1:2] - vv[:, 2:3] - vv[:, 3:4]], axis=1)
    ref = (pos.reshape(-1, 3).astype(np.float64)[ids] * b[:, :, None]).sum(1)
    assert np.max(np.abs(s - ref)) < 1e-6
    n = oracle.vertex_normals(s.reshape(-1), dragon["vis_tri_ids"]).reshape(-1, 3)
    ln = np.linalg.norm(n, axis=1)
    assert np.all((np.abs(ln - 1.0) < 1e-6) | (ln == 0.0))


def test_projection_matches_the_published_constraints_by_finite_differences():
    """Pins the restated solveElem/applyToElem (src/Softbody.js:91-193) to the PUBLISHED algorithm rather than to its
    own code shape: Macklin & Mueller 2021 define the deviatoric constraint C_D = ||F||_F and the hydrostatic constraint
    C_H = det F - 1 - alpha_vol/alpha_dev with F = Ds Dm^-1, projected by XPBD, dlambda = -C / (sum_i w_i |grad_i C|^2 +
    alpha / (dt^2 V_rest)), dx_i = w_i grad_i C dlambda.  Here the gradients come from central finite differences of
    those two scalar functions in float64 -- no hand-derived gradient, no matrix layout shared with the oracle -- for
    random well-shaped tets with random deformations, and the per-tet displacement must agree with the oracle's to
    float32 rounding."""
    rng = np.random.default_rng(7)
    dt, dev_c, vol_c, density = 1.0 / 600.0, 1e-5, 2e-6, 1000.0
    for trial in range(20):
        rest = (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float) * rng.uniform(0.05, 0.3)
                + rng.normal(0, 0.01, (4, 3)) + rng.uniform(-1, 1, 3))
        if np.linalg.det((rest[1:] - rest[0]).T) < 0:
            rest[[1, 2]] = rest[[2, 1]]
        rest32 = rest.astype(np.float32)
        sb = oracle.SoftBodyOracle(rest32.reshape(-1), np.array([0, 1, 2, 3], np.int32), gravity=0.0, density=density,
                                   devCompliance=dev_c, volCompliance=vol_c)
        A = np.eye(3) + rng.normal(0, 0.15, (3, 3))                      # a random deformation of the rest shape
        cur32 = (rest32.astype(float) @ A.T + rng.normal(0, 0.002, (4, 3))).astype(np.float32)
        sb.pos[:] = cur32.reshape(-1)
        got = sb.jacobi_accumulate(np.array([0], np.int32), dt).reshape(4, 3).astype(float)

        r = rest32.astype(float)
        Dm_inv = np.linalg.inv((r[1:] - r[0]).T)
        V = np.linalg.det((r[1:] - r[0]).T) / 6.0
        w = np.full(4, 1.0 / (V / 4.0 * density))                        # lumped mass V/4*rho per vertex (:75-78)

        def F_of(x):
            return (x[1:] - x[0]).T @ Dm_inv

        def project(x, Cfun, alpha):
            C = Cfun(x)
            g = np.zeros((4, 3))
            h = 1e-6
            for i in range(4):
                for c in range(3):
                    xp, xm = x.copy(), x.copy()
                    xp[i, c] += h
                    xm[i, c] -= h
                    g[i, c] = (Cfun(xp) - Cfun(xm)) / (2 * h)
            dl = -C / ((w * (g * g).sum(1)).sum() + alpha / dt / dt / V)
            return x + g * (w * dl)[:, None]

        x0 = cur32.astype(float)
        x1 = project(x0, lambda x: np.sqrt((F_of(x) ** 2).sum()), dev_c)
        x2 = project(x1, lambda x: np.linalg.det(F_of(x)) - 1.0 - vol_c / dev_c, vol_c)
        want = x2 - x0
        scale = np.abs(want).max()
        assert scale > 1e-6
        assert np.abs(got - want).max() <= 2e-4 * scale + 2e-7, (trial, np.abs(got - want).max(), scale)


def test_polar_rotation_extraction_against_svd_polar_decomposition():
    """Pins the restated extractRotation (src/SoftbodyGPU.js:122-139, Mueller et al. 2016) to what it approximates: the
    rotation of the polar decomposition of A = sum_k cur_k rest_k^T, computed independently with numpy's SVD.  One tet,
    no gravity, positions held fixed.  The shader runs at most 9 iterations per substep from the identity and carries the
    result in the tet's quaternion and its incrementally rotated rest pose, so: (i) after ONE substep the rotation is
    within 1.2 % of its angle (the iteration converges linearly and is simply not finished -- measured up to 0.75 %),
    (ii) after four substeps it has converged: within 5e-4 rad (the float32 floor of an angle taken from a trace) of
    the applied rotation for a rigid motion -- whose goal positions are then the current positions, to 1e-6 -- and of
    the SVD polar rotation for rotation + 20 % stretch."""
    rng = np.random.default_rng(11)
    rest = np.array([[0, 0, 0], [0.2, 0, 0], [0, 0.25, 0], [0, 0, 0.15]], np.float32) + np.float32([0.3, 1.0, -0.2])
    ids = np.array([0, 1, 2, 3], np.int32)

    def rot(axis, ang):
        a = np.asarray(axis, float) / np.linalg.norm(axis)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K

    def quat_to_mat(q):
        x, y, z, w = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    def angle_between(Ra, Rb):
        return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1.0) / 2.0, -1.0, 1.0)))

    c = rest.astype(float).mean(0)
    for trial in range(12):
        ang = np.deg2rad(rng.uniform(5, 40))
        R0 = rot(rng.normal(size=3), ang)
        for stretch in (False, True):
            S = np.eye(3) + (np.diag(rng.uniform(-0.2, 0.2, 3)) if stretch else 0.0)
            cur = ((rest.astype(float) - c) @ (R0 @ S).T + c + [0.05, 0.3, -0.1]).astype(np.float32)
            # independent reference: polar rotation of A = sum (cur - c_cur)(rest - c_rest)^T
            x, r = cur.astype(float), rest.astype(float)
            U, _, Vt = np.linalg.svd((x - x.mean(0)).T @ (r - r.mean(0)))
            Rp = U @ np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))]) @ Vt
            if not stretch:
                assert angle_between(Rp, R0) < 1e-6
            po = oracle.PolarOracle(rest.reshape(-1), ids, gravity=0.0)
            for k in range(4):
                po.pos[:] = cur.reshape(-1)
                po.prevPos[:] = cur.reshape(-1)
                po.vel[:] = 0.0
                po.simulate(1.0 / 1200.0)
                err = angle_between(quat_to_mat(po.quat.astype(float)), Rp)
                if k == 0 and not stretch:
                    assert err < 1.2e-2 * ang, (trial, err, ang)
            assert err < 5e-4, (trial, stretch, err)
            if not stretch:
                assert np.abs(po.pos - cur.reshape(-1)).max() < 1e-6   # a rigidly moved tet is already at its goal


def test_init_physics_identities(dragon):
    """initPhysics (src/Softbody.js:60-87) against its defining identities, computed independently in float64:
    invRestPose_e * Dm_e = I, invRestVolume_e = 1 / (det Dm_e / 6), and the lumped masses add up to density * volume
    with every vertex receiving a quarter of each incident tet (the column-major layout of App. A is what makes the first
    identity come out)."""
    sb = oracle.SoftBodyOracle(dragon["tet_verts"], dragon["tet_ids"])
    x = dragon["tet_verts"].reshape(-1, 3).astype(np.float64)
    t = dragon["tet_ids"].reshape(-1, 4)
    Dm = np.stack([x[t[:, 1]] - x[t[:, 0]], x[t[:, 2]] - x[t[:, 0]], x[t[:, 3]] - x[t[:, 0]]], axis=2)  # columns = edges
    Q = sb.invRestPose.reshape(-1, 3, 3).transpose(0, 2, 1).astype(np.float64)   # stored column-major: A[9n + 3c + r]
    I = np.einsum("eij,ejk->eik", Q, Dm)
    cond = np.linalg.cond(Dm)
    assert np.max(np.abs(I - np.eye(3)) / cond[:, None, None]) < 2e-6            # f32 storage of an inverse: error ~ eps * cond
    V = np.linalg.det(Dm) / 6.0
    assert V.min() > 0
    np.testing.assert_allclose(sb.invRestVolume, 1.0 / V, rtol=2e-4)               # V itself comes from f32 edge vectors
    m = np.zeros(len(x))
    np.add.at(m, t.reshape(-1), np.repeat(V / 4.0 * 1000.0, 4))
    np.testing.assert_allclose(1.0 / sb.invMass.astype(np.float64), m, rtol=2e-4)
    assert abs((1.0 / sb.invMass.astype(np.float64)).sum() - 1000.0 * V.sum()) < 1e-4 * 1000.0 * V.sum()
