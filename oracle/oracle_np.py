"""Second, independently structured restatement of the reference CPU solver (TEST INFRASTRUCTURE).

Where ``softbody_oracle.c`` walks the tets one by one like src/Softbody.js does, this file
restates the same arithmetic in batched numpy: tets are grouped into the order-preserving
dependency levels of the sequential sweep (a tet's level is one more than the highest level of
any earlier tet sharing a vertex), and each level is processed as one vectorised batch.  Because
tets inside a level share no vertex and every earlier conflicting tet is in an earlier level, the
result must equal the sequential sweep bit for bit -- which is what tests/test_oracle.py asserts
against the C restatement.  Agreement of two differently written restatements is the only pin
available in round 1: the reference has no tests and cannot be executed in this image (since round 2 the oracle is also
pinned to the mechanically transpiled reference, tests/test_reference_pin.py).

JS arithmetic rule: f32 arrays, f64 expressions, one f32 rounding per store.
Reference lines: src/Softbody.js:60-87 (init), :91-166 (solveElem), :168-193 (applyToElem),
:195-240 (simulate).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64


def r32(a):
    """One typed-array store: f64 -> f32 -> (widened again for the next expression)."""
    return np.asarray(a, F64).astype(F32).astype(F64)


def level_schedule(num_verts: int, tet_ids: np.ndarray):
    """Order-preserving levels of the sequential sweep; returns a list of tet-index arrays."""
    ids = np.asarray(tet_ids).reshape(-1, 4)
    last = np.zeros(num_verts, np.int64)
    level = np.zeros(len(ids), np.int64)
    for j, t in enumerate(ids):
        lv = 1 + max(last[t[0]], last[t[1]], last[t[2]], last[t[3]])
        level[j] = lv
        last[t] = lv
    order = np.argsort(level, kind="stable")
    bounds = np.flatnonzero(np.diff(level[order])) + 1
    return np.split(order, bounds)


def _det_cols(c0, c1, c2):
    """matGetDeterminant on column vectors (src/Softbody.js:381-387), term order as written."""
    a11, a21, a31 = c0[:, 0], c0[:, 1], c0[:, 2]
    a12, a22, a32 = c1[:, 0], c1[:, 1], c1[:, 2]
    a13, a23, a33 = c2[:, 0], c2[:, 1], c2[:, 2]
    return a11 * a22 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - a13 * a22 * a31 - a12 * a21 * a33 - a11 * a23 * a32


def init_physics(verts, tet_ids, density):
    x = np.asarray(verts, F32).reshape(-1, 3).astype(F64)
    ids = np.asarray(tet_ids).reshape(-1, 4)
    c = [r32(x[ids[:, k + 1]] - x[ids[:, 0]]) for k in range(3)]  # Dm columns, f32
    det = _det_cols(*c)
    V = det / 6.0
    inv_det = 1.0 / det
    a11, a21, a31 = c[0][:, 0], c[0][:, 1], c[0][:, 2]
    a12, a22, a32 = c[1][:, 0], c[1][:, 1], c[1][:, 2]
    a13, a23, a33 = c[2][:, 0], c[2][:, 1], c[2][:, 2]
    Q = np.zeros((len(ids), 9), F64)  # column-major: Q[:, 3*col + row]
    Q[:, 0] = (a22 * a33 - a23 * a32) * inv_det
    Q[:, 3] = -(a12 * a33 - a13 * a32) * inv_det
    Q[:, 6] = (a12 * a23 - a13 * a22) * inv_det
    Q[:, 1] = -(a21 * a33 - a23 * a31) * inv_det
    Q[:, 4] = (a11 * a33 - a13 * a31) * inv_det
    Q[:, 7] = -(a11 * a23 - a13 * a21) * inv_det
    Q[:, 2] = (a21 * a32 - a22 * a31) * inv_det
    Q[:, 5] = -(a11 * a32 - a12 * a31) * inv_det
    Q[:, 8] = (a11 * a22 - a12 * a21) * inv_det
    Q = Q.astype(F32)
    pm = V / 4.0 * density
    mass = np.zeros(len(x), F32)
    for e in range(len(ids)):  # f32 accumulation in tet order
        for k in range(4):
            mass[ids[e, k]] = F32(F64(mass[ids[e, k]]) + pm[e])
    inv_mass = mass.copy()
    nz = mass != 0
    inv_mass[nz] = (1.0 / mass[nz].astype(F64)).astype(F32)
    return Q, (1.0 / V).astype(F32), inv_mass


def _F_from(x4, Qc):
    """F = Ds * Q with the reference's three rounded column accumulations (:363-379)."""
    P = [r32(x4[:, k + 1] - x4[:, 0]) for k in range(3)]
    F = []
    for k in range(3):
        col = r32(0.0 + P[0] * Qc[:, 3 * k + 0, None])
        col = r32(col + P[1] * Qc[:, 3 * k + 1, None])
        col = r32(col + P[2] * Qc[:, 3 * k + 2, None])
        F.append(col)
    return F


def _grads(cols, Qc, scale):
    """g_k = sum_j cols[j] * (scale * Q(row k-1, col j)), rounded after every term (:112-125,:144-157)."""
    g = []
    for row in range(3):
        a = r32(0.0 + cols[0] * (scale * Qc[:, 0 + row])[:, None])
        a = r32(a + cols[1] * (scale * Qc[:, 3 + row])[:, None])
        a = r32(a + cols[2] * (scale * Qc[:, 6 + row])[:, None])
        g.append(a)
    return g


def _apply(x4, w4, g123, Cval, compliance, dt, irv):
    """applyToElem (:168-193) on a batch; x4 (n,4,3) f64 holding f32 values, returns updated x4."""
    g0 = r32(0.0 + g123[0] * -1.0)
    g0 = r32(g0 + g123[1] * -1.0)
    g0 = r32(g0 + g123[2] * -1.0)
    g = [g0] + g123
    w = np.zeros(len(x4), F64)
    for i in range(4):
        l2 = g[i][:, 0] * g[i][:, 0] + g[i][:, 1] * g[i][:, 1] + g[i][:, 2] * g[i][:, 2]
        w = w + l2 * w4[:, i]
    alpha = compliance / dt / dt * irv
    act = (Cval != 0.0) & (w != 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        dl = -Cval / (w + alpha)
    out = x4.copy()
    for i in range(4):
        upd = r32(x4[:, i] + g[i] * (dl * w4[:, i])[:, None])
        out[:, i] = np.where(act[:, None], upd, x4[:, i])
    return out


def solve_batch(x4, w4, Qc, irv, dt, dev_c, vol_c):
    """solveElem (:91-166) for a batch of vertex-disjoint tets. Returns (new x4, vol - 1)."""
    F = _F_from(x4, Qc)
    def l2(c):
        return c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1] + c[:, 2] * c[:, 2]

    r_s = np.sqrt(l2(F[0]) + l2(F[1]) + l2(F[2]))
    with np.errstate(divide="ignore"):
        r_inv = 1.0 / r_s
    x4 = _apply(x4, w4, _grads(F, Qc, r_inv), r_s, dev_c, dt, irv)
    F = _F_from(x4, Qc)

    def cross(b, c):
        return r32(np.stack([b[:, 1] * c[:, 2] - b[:, 2] * c[:, 1],
                             b[:, 2] * c[:, 0] - b[:, 0] * c[:, 2],
                             b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]], axis=1))

    dF = [cross(F[1], F[2]), cross(F[2], F[0]), cross(F[0], F[1])]
    vol = _det_cols(*F)
    Cv = vol - 1.0 - vol_c / dev_c
    x4 = _apply(x4, w4, _grads(dF, Qc, np.ones(len(x4))), Cv, vol_c, dt, irv)
    return x4, vol - 1.0


class SoftBodyNP:
    def __init__(self, verts, tet_ids, gravity=-9.81, friction=1000.0, density=1000.0,
                 devCompliance=1.0 / 100000.0, volCompliance=0.0,
                 worldBounds=(-2.5, -1.0, -2.5, 2.5, 10.0, 2.5)):
        self.p = dict(gravity=gravity, friction=friction, devCompliance=devCompliance,
                      volCompliance=volCompliance, worldBounds=worldBounds)
        self.pos = np.asarray(verts, F32).reshape(-1, 3).copy()
        self.prev = self.pos.copy()
        self.vel = np.zeros_like(self.pos)
        self.ids = np.asarray(tet_ids).reshape(-1, 4).astype(np.int64)
        self.Q, self.irv, self.inv_mass = init_physics(self.pos, self.ids, density)
        self.levels = level_schedule(len(self.pos), self.ids)
        self.volError = 0.0

    def simulate(self, dt):
        p = self.p
        g = np.array([0.0, p["gravity"], 0.0])
        self.vel = (self.vel.astype(F64) + g * dt).astype(F32)
        self.prev = self.pos.copy()
        self.pos = (self.pos.astype(F64) + self.vel.astype(F64) * dt).astype(F32)
        x = self.pos.astype(F64)
        ve = np.zeros(len(self.ids), F64)
        Qd, irv, w = self.Q.astype(F64), self.irv.astype(F64), self.inv_mass.astype(F64)
        for lv in self.levels:
            t = self.ids[lv]
            x4, dv = solve_batch(x[t], w[t], Qd[lv], irv[lv], dt, p["devCompliance"], p["volCompliance"])
            x[t.reshape(-1)] = x4.reshape(-1, 3)
            ve[lv] = dv
        # volError is a sequential f64 sum in tet order (:163)
        acc = 0.0
        for v in ve:
            acc += v
        self.volError = acc / len(self.ids)
        lo, hi = np.array(p["worldBounds"][:3]), np.array(p["worldBounds"][3:])
        x = np.maximum(lo, np.minimum(hi, x))
        below = x[:, 1] < 0.0
        if below.any():
            x[below, 1] = 0.0
            prev = self.prev.astype(F64)
            k = min(1.0, dt * p["friction"])
            Fx = r32(prev[below, 0] - x[below, 0])
            Fz = r32(prev[below, 2] - x[below, 2])
            x[below, 0] = r32(x[below, 0] + Fx * k)
            x[below, 2] = r32(x[below, 2] + Fz * k)
        self.pos = x.astype(F32)
        self.vel = ((self.pos.astype(F64) - self.prev.astype(F64)) * (1.0 / dt)).astype(F32)
