/*
 * oracle/softbody_oracle.c -- CPU restatement of the reference CPU solver (TEST INFRASTRUCTURE).
 *
 * This file is the parity oracle for the XPBD Neo-Hookean substep path of zalo/TetSim,
 * class SoftBody in src/Softbody.js.  It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY PINNED TO THE REFERENCE'S OWN TEXT (round 2).  The reference ships no tests or golden vectors and no JavaScript
 * engine exists in the build image, so it cannot be executed as is; instead tools/transpile_reference.py mechanically
 * re-emits src/Softbody.js (class SoftBody, every method) as Python under JS number semantics (oracle/jsrt.py),
 * tools/make_ref_golden.py executes that on Dragon (free fall, contact + clamp, compliance, grab) and commits
 * tests/golden/ref_golden.npz, and tests/test_reference_pin.py requires this file to reproduce every vector BIT FOR BIT
 * (and re-runs the transpile live where /root/reference exists).  Older pins stay: (i) a second, independently structured
 * numpy restatement (oracle/oracle_np.py), (ii) analytic properties and (iii) a finite-difference XPBD projection of the
 * published constraint functions (tests/test_oracle.py).
 *
 * Arithmetic rule being restated (JavaScript typed-array semantics):
 *   - every Float32Array read widens f32 -> f64,
 *   - every expression is IEEE f64, evaluated left to right as written, never fused,
 *   - every Float32Array write rounds f64 -> f32 (round to nearest even).
 * Compile with -ffp-contract=off (see oracle/Makefile) so the compiler never forms an FMA.
 *
 * Each function cites the reference lines it follows (paths relative to the reference root).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

typedef struct OracleParams {
    double gravity;        /* src/main.js:23 */
    double friction;       /* src/main.js:28 */
    double density;        /* src/main.js:29 */
    double devCompliance;  /* src/main.js:30 */
    double volCompliance;  /* src/main.js:31 */
    double worldBounds[6]; /* src/main.js:32: lo.xyz, hi.xyz */
} OracleParams;

/* One f32 store: the only place a double becomes a float. */
static inline float st(double v) { return (float)v; }

/* Math.max / Math.min propagate NaN (C fmax/fmin do not). src/Softbody.js:352 */
static inline double js_max(double a, double b) { if (a != a || b != b) return NAN; return a > b ? a : b; }
static inline double js_min(double a, double b) { if (a != a || b != b) return NAN; return a < b ? a : b; }

/* a[3i..] += b[3j..] * s with an f32 store per component.  src/Softbody.js:316-321 */
static inline void axpy3(float *a, size_t i, const float *b, size_t j, double s) {
    a[3 * i + 0] = st((double)a[3 * i + 0] + (double)b[3 * j + 0] * s);
    a[3 * i + 1] = st((double)a[3 * i + 1] + (double)b[3 * j + 1] * s);
    a[3 * i + 2] = st((double)a[3 * i + 2] + (double)b[3 * j + 2] * s);
}

/* d[3k..] = (a[3i..] - b[3j..]) * s.  src/Softbody.js:323-328 */
static inline void diff3(float *d, size_t k, const float *a, size_t i, const float *b, size_t j, double s) {
    d[3 * k + 0] = st(((double)a[3 * i + 0] - (double)b[3 * j + 0]) * s);
    d[3 * k + 1] = st(((double)a[3 * i + 1] - (double)b[3 * j + 1]) * s);
    d[3 * k + 2] = st(((double)a[3 * i + 2] - (double)b[3 * j + 2]) * s);
}

static inline void zero3(float *a, size_t i) { a[3 * i] = 0.0f; a[3 * i + 1] = 0.0f; a[3 * i + 2] = 0.0f; }

/* src/Softbody.js:330-334 */
static inline double len2(const float *a, size_t i) {
    double a0 = a[3 * i], a1 = a[3 * i + 1], a2 = a[3 * i + 2];
    return a0 * a0 + a1 * a1 + a2 * a2;
}

/* a = b x c.  src/Softbody.js:343-348 */
static inline void cross3(float *a, size_t i, const float *b, size_t j, const float *c, size_t k) {
    double b0 = b[3 * j], b1 = b[3 * j + 1], b2 = b[3 * j + 2];
    double c0 = c[3 * k], c1 = c[3 * k + 1], c2 = c[3 * k + 2];
    a[3 * i + 0] = st(b1 * c2 - b2 * c1);
    a[3 * i + 1] = st(b2 * c0 - b0 * c2);
    a[3 * i + 2] = st(b0 * c1 - b1 * c0);
}

/* Column-major 3x3 determinant, term order as written.  src/Softbody.js:381-387 */
static inline double det3(const float *A, size_t n) {
    const float *m = A + 9 * n;
    double a11 = m[0], a12 = m[3], a13 = m[6];
    double a21 = m[1], a22 = m[4], a23 = m[7];
    double a31 = m[2], a32 = m[5], a33 = m[8];
    return a11 * a22 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - a13 * a22 * a31 - a12 * a21 * a33 - a11 * a23 * a32;
}

/* In-place adjugate inverse.  src/Softbody.js:389-410.
 * The zero-determinant branch of the reference indexes A[anr+i] BEFORE anr is scaled by 9
 * (src/Softbody.js:391-394), i.e. it zeroes floats n..n+8 of the whole array rather than
 * matrix n.  Restated literally; valid meshes never reach it. */
static inline void inv3(float *A, size_t n) {
    double det = det3(A, n);
    if (det == 0.0) {
        for (int i = 0; i < 9; i++) A[n + i] = 0.0f;
        return;
    }
    double invDet = 1.0 / det;
    float *m = A + 9 * n;
    double a11 = m[0], a12 = m[3], a13 = m[6];
    double a21 = m[1], a22 = m[4], a23 = m[7];
    double a31 = m[2], a32 = m[5], a33 = m[8];
    m[0] = st((a22 * a33 - a23 * a32) * invDet);
    m[3] = st(-(a12 * a33 - a13 * a32) * invDet);
    m[6] = st((a12 * a23 - a13 * a22) * invDet);
    m[1] = st(-(a21 * a33 - a23 * a31) * invDet);
    m[4] = st((a11 * a33 - a13 * a31) * invDet);
    m[7] = st(-(a11 * a23 - a13 * a21) * invDet);
    m[2] = st((a21 * a32 - a22 * a31) * invDet);
    m[5] = st(-(a11 * a32 - a12 * a31) * invDet);
    m[8] = st((a11 * a22 - a12 * a21) * invDet);
}

/* dst column k = A * (column k of B_n): three rounded accumulations per column.
 * src/Softbody.js:363-379 */
static inline void matmul3(float *dst, const float *A, const float *B, size_t n) {
    for (int k = 0; k < 3; k++) {
        double b0 = B[9 * n + 3 * k + 0], b1 = B[9 * n + 3 * k + 1], b2 = B[9 * n + 3 * k + 2];
        zero3(dst, k);
        axpy3(dst, k, A, 0, b0);
        axpy3(dst, k, A, 1, b1);
        axpy3(dst, k, A, 2, b2);
    }
}

/* SoftBody.initPhysics: rest-pose inverse, lumped masses, inverse rest volume.
 * src/Softbody.js:60-87.  invMass is accumulated in f32, in tet order. */
void oracle_init_physics(int numVerts, int numTets, const float *pos, const int *tetIds, double density,
                         float *invRestPose, float *invRestVolume, float *invMass) {
    for (int i = 0; i < numVerts; i++) invMass[i] = 0.0f;
    for (int i = 0; i < numTets; i++) {
        int id0 = tetIds[4 * i], id1 = tetIds[4 * i + 1], id2 = tetIds[4 * i + 2], id3 = tetIds[4 * i + 3];
        diff3(invRestPose, 3 * (size_t)i + 0, pos, id1, pos, id0, 1.0);
        diff3(invRestPose, 3 * (size_t)i + 1, pos, id2, pos, id0, 1.0);
        diff3(invRestPose, 3 * (size_t)i + 2, pos, id3, pos, id0, 1.0);
        double V = det3(invRestPose, i) / 6.0;
        inv3(invRestPose, i);
        double pm = V / 4.0 * density;
        invMass[id0] = st((double)invMass[id0] + pm);
        invMass[id1] = st((double)invMass[id1] + pm);
        invMass[id2] = st((double)invMass[id2] + pm);
        invMass[id3] = st((double)invMass[id3] + pm);
        invRestVolume[i] = st(1.0 / V);
    }
    for (int i = 0; i < numVerts; i++)
        if (invMass[i] != 0.0f) invMass[i] = st(1.0 / (double)invMass[i]);
}

/* Scratch the reference keeps on the object (src/Softbody.js:27-30). */
typedef struct Scratch { float P[9], F[9], dF[9], g[12]; } Scratch;

/* SoftBody.applyToElem.  src/Softbody.js:168-193.  `x` is the position array being corrected and
 * `ids` the 4 indices into it (global ids for Gauss-Seidel, 0..3 for the Jacobi local copy). */
static void apply_to_elem(Scratch *s, float *x, const int *ids, const float *w4, double C, double compliance,
                          double dt, double invRestVolume) {
    if (C == 0.0) return;
    float *g = s->g;
    zero3(g, 0);
    axpy3(g, 0, g, 1, -1.0);
    axpy3(g, 0, g, 2, -1.0);
    axpy3(g, 0, g, 3, -1.0);
    double w = 0.0;
    for (int i = 0; i < 4; i++) w += len2(g, i) * (double)w4[i];
    if (w == 0.0) return;
    double alpha = compliance / dt / dt * invRestVolume;
    double dlambda = -C / (w + alpha);
    for (int i = 0; i < 4; i++) axpy3(x, ids[i], g, i, dlambda * (double)w4[i]);
}

/* SoftBody.solveElem: deviatoric then hydrostatic constraint of one tet.  src/Softbody.js:91-166.
 * Returns vol - 1 (the term added to volError at :163). */
static double solve_elem(Scratch *s, float *x, const int *ids, const float *w4, const float *Q, size_t e,
                         float irv, double dt, double devCompliance, double volCompliance) {
    float *g = s->g;
    /* tr(F) = 3 :  C = ||F||_F */
    diff3(s->P, 0, x, ids[1], x, ids[0], 1.0);
    diff3(s->P, 1, x, ids[2], x, ids[0], 1.0);
    diff3(s->P, 2, x, ids[3], x, ids[0], 1.0);
    matmul3(s->F, s->P, Q, e);
    double r_s = sqrt(len2(s->F, 0) + len2(s->F, 1) + len2(s->F, 2));
    double r_s_inv = 1.0 / r_s;
    for (int k = 1; k <= 3; k++) { /* matIJ(ir,e,row,col) = ir[9e+3col+row]; row = k-1.  :112-125 */
        zero3(g, k);
        axpy3(g, k, s->F, 0, r_s_inv * (double)Q[9 * e + 0 + (k - 1)]);
        axpy3(g, k, s->F, 1, r_s_inv * (double)Q[9 * e + 3 + (k - 1)]);
        axpy3(g, k, s->F, 2, r_s_inv * (double)Q[9 * e + 6 + (k - 1)]);
    }
    apply_to_elem(s, x, ids, w4, r_s, devCompliance, dt, irv);

    /* det F = 1 : recomputed from the UPDATED positions.  :134-165 */
    diff3(s->P, 0, x, ids[1], x, ids[0], 1.0);
    diff3(s->P, 1, x, ids[2], x, ids[0], 1.0);
    diff3(s->P, 2, x, ids[3], x, ids[0], 1.0);
    matmul3(s->F, s->P, Q, e);
    cross3(s->dF, 0, s->F, 1, s->F, 2);
    cross3(s->dF, 1, s->F, 2, s->F, 0);
    cross3(s->dF, 2, s->F, 0, s->F, 1);
    for (int k = 1; k <= 3; k++) {
        zero3(g, k);
        axpy3(g, k, s->dF, 0, (double)Q[9 * e + 0 + (k - 1)]);
        axpy3(g, k, s->dF, 1, (double)Q[9 * e + 3 + (k - 1)]);
        axpy3(g, k, s->dF, 2, (double)Q[9 * e + 6 + (k - 1)]);
    }
    double vol = det3(s->F, 0);
    double C = vol - 1.0 - volCompliance / devCompliance;
    apply_to_elem(s, x, ids, w4, C, volCompliance, dt, irv);
    return vol - 1.0;
}

/* simulate() lines 198-202: semi-implicit predict over ALL vertices (no pinned test). */
static void predict(int N, float *pos, float *prev, float *vel, double gravity, double dt) {
    for (int i = 0; i < N; i++) {
        vel[3 * i + 0] = st((double)vel[3 * i + 0] + 0.0 * dt);
        vel[3 * i + 1] = st((double)vel[3 * i + 1] + gravity * dt);
        vel[3 * i + 2] = st((double)vel[3 * i + 2] + 0.0 * dt);
        prev[3 * i + 0] = pos[3 * i + 0];
        prev[3 * i + 1] = pos[3 * i + 1];
        prev[3 * i + 2] = pos[3 * i + 2];
        axpy3(pos, i, vel, i, dt);
    }
}

/* simulate() lines 213-239: bounds clamp, floor + friction, grab, velocity. */
static void post(int N, float *pos, const float *prev, float *vel, const OracleParams *p, double dt, int grabId,
                 const float *grabPos) {
    const double *lo = p->worldBounds, *hi = p->worldBounds + 3;
    for (int i = 0; i < N; i++) {
        for (int c = 0; c < 3; c++)
            pos[3 * i + c] = st(js_max(lo[c], js_min(hi[c], (double)pos[3 * i + c])));
        if (pos[3 * i + 1] < 0.0f) {
            pos[3 * i + 1] = 0.0f;
            float Fx = st((double)prev[3 * i + 0] - (double)pos[3 * i + 0]);
            float Fz = st((double)prev[3 * i + 2] - (double)pos[3 * i + 2]);
            double k = js_min(1.0, dt * p->friction);
            pos[3 * i + 0] = st((double)pos[3 * i + 0] + (double)Fx * k);
            pos[3 * i + 2] = st((double)pos[3 * i + 2] + (double)Fz * k);
        }
    }
    if (grabId >= 0) {
        pos[3 * grabId + 0] = grabPos[0];
        pos[3 * grabId + 1] = grabPos[1];
        pos[3 * grabId + 2] = grabPos[2];
    }
    double inv_dt = 1.0 / dt; /* multiply by the reciprocal, as :239 does */
    for (int i = 0; i < N; i++) diff3(vel, i, pos, i, prev, i, inv_dt);
}

/* SoftBody.simulate: ONE substep, Gauss-Seidel sweep in `order` (NULL = 0..M-1, the reference's
 * own order, src/Softbody.js:207-208).  A level- or colour-ordered sweep passes its permutation. */
void oracle_simulate(int numVerts, int numTets, float *pos, float *prev, float *vel, const float *invMass,
                     const float *invRestPose, const float *invRestVolume, const int *tetIds, const int *order,
                     double dt, const OracleParams *p, int grabId, const float *grabPos, double *volError) {
    Scratch s;
    predict(numVerts, pos, prev, vel, p->gravity, dt);
    double ve = 0.0;
    for (int k = 0; k < numTets; k++) {
        int e = order ? order[k] : k;
        const int *ids = tetIds + 4 * (size_t)e;
        float w4[4] = {invMass[ids[0]], invMass[ids[1]], invMass[ids[2]], invMass[ids[3]]};
        ve += solve_elem(&s, pos, ids, w4, invRestPose, e, invRestVolume[e], dt, p->devCompliance, p->volCompliance);
    }
    ve /= numTets;
    if (volError) *volError = ve;
    post(numVerts, pos, prev, vel, p, dt, grabId, grabPos);
}

/* Jacobi Neo-Hookean substep.  NOT in the reference (README.md:25 only names Jacobi for the
 * shape-matching variant); this is the semantics the B200 path defines and DESIGN.md states:
 *   per iteration, every tet runs the reference's solveElem arithmetic on a private copy of its
 *   four vertices taken from the iteration-start positions; dx = copy_after - copy_before (f32);
 *   per vertex, dx are summed in f32 in ascending (tet, slot) order and the vertex moves by
 *   sum / valence (valence = number of incident tet corners).
 * `acc` is caller scratch of 3*numVerts floats; `valence` has numVerts ints. */
void oracle_simulate_jacobi(int numVerts, int numTets, float *pos, float *prev, float *vel, const float *invMass,
                            const float *invRestPose, const float *invRestVolume, const int *tetIds,
                            const int *valence, float *acc, int iters, double dt, const OracleParams *p,
                            int grabId, const float *grabPos, double *volError) {
    Scratch s;
    static const int loc[4] = {0, 1, 2, 3};
    predict(numVerts, pos, prev, vel, p->gravity, dt);
    double ve = 0.0;
    for (int it = 0; it < iters; it++) {
        memset(acc, 0, sizeof(float) * 3 * (size_t)numVerts);
        ve = 0.0;
        for (int e = 0; e < numTets; e++) {
            const int *ids = tetIds + 4 * (size_t)e;
            float y[12], w4[4];
            for (int k = 0; k < 4; k++) {
                y[3 * k] = pos[3 * ids[k]]; y[3 * k + 1] = pos[3 * ids[k] + 1]; y[3 * k + 2] = pos[3 * ids[k] + 2];
                w4[k] = invMass[ids[k]];
            }
            ve += solve_elem(&s, y, loc, w4, invRestPose, e, invRestVolume[e], dt, p->devCompliance, p->volCompliance);
            for (int k = 0; k < 4; k++)
                for (int c = 0; c < 3; c++) {
                    float dx = st((double)y[3 * k + c] - (double)pos[3 * ids[k] + c]);
                    acc[3 * ids[k] + c] = acc[3 * ids[k] + c] + dx; /* f32 add */
                }
        }
        ve /= numTets;
        for (int i = 0; i < numVerts; i++) {
            if (valence[i] == 0) continue;
            float inv = 1.0f / (float)valence[i];
            for (int c = 0; c < 3; c++) pos[3 * i + c] = pos[3 * i + c] + acc[3 * i + c] * inv; /* f32 mul, f32 add */
        }
    }
    if (volError) *volError = ve;
    post(numVerts, pos, prev, vel, p, dt, grabId, grabPos);
}

/* The tet half of one Jacobi iteration on a SUBSET of the tets (tetList, numList entries): adds each
 * corner's dx into acc (3 floats per vertex, caller-zeroed).  Used by the multi-process tests to
 * restate what one rank of a tet-partitioned run contributes before the boundary all-reduce. */
void oracle_jacobi_accumulate(int numList, const int *tetList, const float *pos, const float *invMass,
                              const float *invRestPose, const float *invRestVolume, const int *tetIds, float *acc,
                              double dt, const OracleParams *p) {
    Scratch s;
    static const int loc[4] = {0, 1, 2, 3};
    for (int n = 0; n < numList; n++) {
        const int e = tetList[n];
        const int *ids = tetIds + 4 * (size_t)e;
        float y[12], w4[4];
        for (int k = 0; k < 4; k++) {
            y[3 * k] = pos[3 * ids[k]]; y[3 * k + 1] = pos[3 * ids[k] + 1]; y[3 * k + 2] = pos[3 * ids[k] + 2];
            w4[k] = invMass[ids[k]];
        }
        solve_elem(&s, y, loc, w4, invRestPose, e, invRestVolume[e], dt, p->devCompliance, p->volCompliance);
        for (int k = 0; k < 4; k++)
            for (int c = 0; c < 3; c++) {
                float dx = st((double)y[3 * k + c] - (double)pos[3 * ids[k] + c]);
                acc[3 * ids[k] + c] = acc[3 * ids[k] + c] + dx;
            }
    }
}

/* SoftBody.updateVisMesh without the normals: barycentric skinning.  src/Softbody.js:259-273.
 * visVerts = (tetNr, b0, b1, b2) per vertex, b3 = 1 - b0 - b1 - b2 evaluated in f64. */
void oracle_skin(int numVis, const float *visVerts, const int *tetIds, const float *pos, float *out) {
    for (int i = 0; i < numVis; i++) {
        size_t t = (size_t)(4.0 * (double)visVerts[4 * i]);
        double b0 = visVerts[4 * i + 1], b1 = visVerts[4 * i + 2], b2 = visVerts[4 * i + 3];
        double b3 = 1.0 - b0 - b1 - b2;
        zero3(out, i);
        axpy3(out, i, pos, tetIds[t + 0], b0);
        axpy3(out, i, pos, tetIds[t + 1], b1);
        axpy3(out, i, pos, tetIds[t + 2], b2);
        axpy3(out, i, pos, tetIds[t + 3], b3);
    }
}

/* three.js BufferGeometry.computeVertexNormals for an indexed geometry
 * (node_modules/three/build/three.module.js:11125-11215, three@0.160.0): per triangle
 * cb = (pC - pB) x (pA - pB) in f64 Vector3 math, accumulated into the f32 normal attribute
 * (each += is an f32 store), then every normal is normalised (x * (1 / (length || 1))). */
void oracle_vertex_normals(int numVerts, int numTris, const float *pos, const int *tri, float *nrm) {
    memset(nrm, 0, sizeof(float) * 3 * (size_t)numVerts);
    for (int t = 0; t < numTris; t++) {
        int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        double cbx = (double)pos[3 * c] - (double)pos[3 * b], cby = (double)pos[3 * c + 1] - (double)pos[3 * b + 1],
               cbz = (double)pos[3 * c + 2] - (double)pos[3 * b + 2];
        double abx = (double)pos[3 * a] - (double)pos[3 * b], aby = (double)pos[3 * a + 1] - (double)pos[3 * b + 1],
               abz = (double)pos[3 * a + 2] - (double)pos[3 * b + 2];
        double nx = cby * abz - cbz * aby, ny = cbz * abx - cbx * abz, nz = cbx * aby - cby * abx;
        /* nA, nB, nC are all read before any is written back (:11173-11183): with a repeated
         * index the later write wins rather than accumulating twice. */
        const int v[3] = {a, b, c};
        double n[3][3];
        for (int k = 0; k < 3; k++) {
            n[k][0] = (double)nrm[3 * v[k] + 0] + nx;
            n[k][1] = (double)nrm[3 * v[k] + 1] + ny;
            n[k][2] = (double)nrm[3 * v[k] + 2] + nz;
        }
        for (int k = 0; k < 3; k++) {
            nrm[3 * v[k] + 0] = st(n[k][0]);
            nrm[3 * v[k] + 1] = st(n[k][1]);
            nrm[3 * v[k] + 2] = st(n[k][2]);
        }
    }
    for (int i = 0; i < numVerts; i++) {
        double x = nrm[3 * i], y = nrm[3 * i + 1], z = nrm[3 * i + 2];
        double len = sqrt(x * x + y * y + z * z);
        double s = 1.0 / ((len != 0.0 && len == len) ? len : 1.0); /* length() || 1 */
        nrm[3 * i] = st(x * s); nrm[3 * i + 1] = st(y * s); nrm[3 * i + 2] = st(z * s);
    }
}

/* SoftBody.startGrab nearest-vertex search: first strict minimum of the f64 squared distance.
 * src/Softbody.js:279-291 (p is a plain JS array of doubles; pos is f32). */
int oracle_nearest_vertex(int numVerts, const float *pos, const double *p) {
    double minD2 = 1.7976931348623157e308; /* Number.MAX_VALUE */
    int id = -1;
    for (int i = 0; i < numVerts; i++) {
        double a0 = p[0] - (double)pos[3 * i], a1 = p[1] - (double)pos[3 * i + 1], a2 = p[2] - (double)pos[3 * i + 2];
        double d2 = a0 * a0 + a1 * a1 + a2 * a2;
        if (d2 < minD2) { minD2 = d2; id = i; }
    }
    return id;
}
