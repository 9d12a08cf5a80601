/*
 * oracle/polar_oracle.c -- CPU restatement of the reference WebGL solver (TEST INFRASTRUCTURE).
 *
 * Restates one substep of class SoftBodyGPU (src/SoftbodyGPU.js): the seven fragment-shader
 * passes K1..K7 (src/SoftbodyGPU.js:59-376) in the order MultiTargetGPUComputationRenderer.compute
 * runs them (src/MultiTargetGPUComputationRenderer.js:272-306: every pass reads the CURRENT target
 * of each dependency, `prev_<dep>` is the alternate target, and the written variable flips).
 * NOT part of the product: only tests/, smoke() and bench.py's CPU-baseline legs may load it.
 *
 * PARITY PINNED TO THE REFERENCE'S OWN TEXT, UP TO THE FLOAT MODEL (round 2).  The reference has no tests or golden vectors
 * and cannot be executed here (no JS engine / WebGL).  tools/transpile_reference.py (JavaScript: initPhysics, simulate, the
 * GPGPU runtime's addVariable / addPass / compute) and tools/transpile_shaders.py (GLSL: the seven passes) re-emit the
 * WHOLE solver mechanically as Python; oracle/ref_runner.py executes it texel by texel and tests/test_reference_pin.py
 * requires this file to reproduce positions, prevPos, velocities, quaternions and goal corners BIT FOR BIT (Dragon in
 * free fall, a beam with floor contact and friction).  What a transpile cannot settle: GLSL ES 3.00 `highp float` leaves
 * the precision of sin(), normalize(), length() and the freedom to fuse multiply-adds to the driver, so "the reference's
 * bits" are not defined even in principle; this file and the transpiled shaders' runtime (oracle/glslrt.py) fix ONE
 * IEEE-754 binary32 reading:
 *   - every operation is a separately rounded f32 operation, left to right (no FMA:
 *     compile with -ffp-contract=off),
 *   - length(v) = sqrtf(dot(v,v)), normalize(v) = v / length(v) (IEEE divide),
 *   - sin(x) = (float)sin((double)x),
 *   - clamp(x,lo,hi) = fminf(fmaxf(x,lo),hi).
 * Deliberate departures from the shader text (not exercised by the pinned scenarios): the grab uses the particle's linear
 * index instead of the broken indexFromUV decode (src/SoftbodyGPU.js:336-338), the bounds are a parameter (:347).
 * The CUDA path's BITEXACT mode performs the same operations and must match bit for bit.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct PolarParams {
    float gravity;        /* uniform, src/SoftbodyGPU.js:374 */
    float friction;       /* uniform, :357 */
    float worldBounds[6]; /* hard-coded in the shader (:347) to main.js:32's defaults; a parameter here */
} PolarParams;

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;

static inline v3 add3(v3 a, v3 b) { v3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static inline v3 sub3(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline v3 mul3(v3 a, float s) { v3 r = {a.x * s, a.y * s, a.z * s}; return r; }
static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 crs3(v3 a, v3 b) {
    v3 r = {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
    return r;
}
static inline v4 normalize4(v4 q) {
    float len = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    v4 r = {q.x / len, q.y / len, q.z / len, q.w / len};
    return r;
}

/* Rotate(), src/SoftbodyGPU.js:111-113 */
static inline v3 rotate(v3 p, v4 q) {
    v3 u = {q.x, q.y, q.z};
    v3 t = add3(crs3(u, p), mul3(p, q.w));
    return add3(p, mul3(crs3(u, t), 2.0f));
}

/* quat_mult(), Hamilton product in xyzw order, src/SoftbodyGPU.js:114-121 */
static inline v4 quat_mult(v4 a, v4 b) {
    v4 r;
    r.x = (a.w * b.x) + (a.x * b.w) + (a.y * b.z) - (a.z * b.y);
    r.y = (a.w * b.y) - (a.x * b.z) + (a.y * b.w) + (a.z * b.x);
    r.z = (a.w * b.z) + (a.x * b.y) - (a.y * b.x) + (a.z * b.w);
    r.w = (a.w * b.w) - (a.x * b.x) - (a.y * b.y) - (a.z * b.z);
    return r;
}

static inline float sin_f(float x) { return (float)sin((double)x); }

/* extractRotation(), Mueller et al. 2016, at most 9 iterations.  src/SoftbodyGPU.js:122-139.
 * A[c] is column c.  The half-angle cosine is sin(h + 1.57) as the shader writes it (:108). */
static v4 extract_rotation(const v3 A[3], v4 q) {
    const v3 ex = {1.0f, 0.0f, 0.0f}, ey = {0.0f, 1.0f, 0.0f}, ez = {0.0f, 0.0f, 1.0f};
    for (int iter = 0; iter < 9; iter++) {
        v3 X = rotate(ex, q), Y = rotate(ey, q), Z = rotate(ez, q);
        v3 num = add3(add3(crs3(X, A[0]), crs3(Y, A[1])), crs3(Z, A[2]));
        float den = dot3(X, A[0]) + dot3(Y, A[1]) + dot3(Z, A[2]) + 0.000000001f;
        v3 omega = mul3(num, 1.0f / fabsf(den));
        float w = sqrtf(dot3(omega, omega));
        if (w < 0.000000001f) break;
        float half = w * 0.5f;
        float s = sin_f(half), c = sin_f(half + 1.57f);
        v4 dq = {omega.x / w * s, omega.y / w * s, omega.z / w * s, c};
        q = quat_mult(dq, q);
    }
    return q;
}

static inline v3 ld3(const float *a, size_t i) { v3 r = {a[3 * i], a[3 * i + 1], a[3 * i + 2]}; return r; }
static inline void st3(float *a, size_t i, v3 v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }

/* SoftBodyGPU.initPhysics, the parts the passes read.  src/SoftbodyGPU.js:526-591.
 * rest[12e + 3k ..] = rest position of corner k of tet e (elems0[k], :535-546); quat = identity
 * (:548-551); invRestVolume = f32(1 / V), V = det/6 in f64 from f32 edge differences (:579-589). */
void oracle_polar_init(int numVerts, int numTets, const float *verts, const int *tetIds, float *rest, float *quat,
                       float *invRestVolume) {
    (void)numVerts;
    for (int e = 0; e < numTets; e++) {
        const int *id = tetIds + 4 * (size_t)e;
        for (int k = 0; k < 4; k++)
            for (int c = 0; c < 3; c++) rest[12 * (size_t)e + 3 * k + c] = verts[3 * id[k] + c];
        quat[4 * e] = 0.0f; quat[4 * e + 1] = 0.0f; quat[4 * e + 2] = 0.0f; quat[4 * e + 3] = 1.0f;
        float d[9];
        for (int k = 0; k < 3; k++)
            for (int c = 0; c < 3; c++)
                d[3 * k + c] = (float)((double)verts[3 * id[k + 1] + c] - (double)verts[3 * id[0] + c]);
        double a11 = d[0], a12 = d[3], a13 = d[6], a21 = d[1], a22 = d[4], a23 = d[7], a31 = d[2], a32 = d[5], a33 = d[8];
        double det = a11 * a22 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - a13 * a22 * a31 - a12 * a21 * a33 - a11 * a23 * a32;
        double V = det / 6.0;
        invRestVolume[e] = (float)(1.0 / V);
    }
}

/* Reverse (vertex -> tet corner) table, src/SoftbodyGPU.js:559-577, flattened to CSR.
 * The reference keeps 9 RGBA tables = 36 slots per vertex, initialised to -1, and puts the
 * encoded corner 4*tet+slot into the first slot whose value is <= 0.0 -- so the encoded value 0
 * (tet 0, corner 0) looks empty and is overwritten by that vertex's next corner (:568).
 * referenceTableBug != 0 reproduces that; 0 keeps every corner.  Corners beyond 36 per vertex are
 * silently dropped by the reference; here `cap` is that capacity (36), or <= 0 for unlimited.
 * Returns the number of entries written; start has numVerts + 1 ints. */
int oracle_polar_build_table(int numVerts, int numTets, const int *tetIds, int referenceTableBug, int cap,
                             int *start, int *entries) {
    int *count = (int *)calloc((size_t)numVerts, sizeof(int));
    /* pass 1: per-vertex list lengths under the reference's insertion rule */
    int **lists = (int **)calloc((size_t)numVerts, sizeof(int *));
    int *capv = (int *)calloc((size_t)numVerts, sizeof(int));
    for (int e = 0; e < numTets; e++)
        for (int k = 0; k < 4; k++) {
            int p = tetIds[4 * (size_t)e + k], v = 4 * e + k;
            if (count[p] == capv[p]) {
                capv[p] = capv[p] ? 2 * capv[p] : 8;
                lists[p] = (int *)realloc(lists[p], sizeof(int) * (size_t)capv[p]);
            }
            int n = count[p];
            /* the first slot holding a value <= 0 among the filled prefix, else the first free slot */
            int slot = n;
            if (referenceTableBug)
                for (int j = 0; j < n; j++)
                    if (lists[p][j] <= 0) { slot = j; break; }
            if (cap > 0 && slot >= cap) continue; /* table full: dropped */
            lists[p][slot] = v;
            if (slot == n) count[p] = n + 1;
        }
    int total = 0;
    for (int p = 0; p < numVerts; p++) {
        start[p] = total;
        for (int j = 0; j < count[p]; j++) entries[total++] = lists[p][j];
        free(lists[p]);
    }
    start[numVerts] = total;
    free(lists); free(capv); free(count);
    return total;
}

/* One substep = passes K1..K7.  dt, gravity and friction are f32 shader uniforms. */
void oracle_polar_simulate(int numVerts, int numTets, float *pos, float *prev, float *vel, float *rest, float *quat,
                           const float *invRestVolume, const int *tetIds, const int *tblStart, const int *tblEntries,
                           float dt, const PolarParams *p, int grabId, const float *grabPos) {
    /* K1 copyPrevPos (:59-64) and K2 xpbdIntegrate (:67-74): gravity is NOT applied here */
    for (int i = 0; i < numVerts; i++) {
        v3 x = ld3(pos, i);
        st3(prev, i, x);
        st3(pos, i, add3(x, mul3(ld3(vel, i), dt)));
    }
    /* K3 solveElem (:142-182) then K4 gatherElem (:217-262); both read the same old rest tet */
    for (int e = 0; e < numTets; e++) {
        const int *id = tetIds + 4 * (size_t)e;
        v3 cur[4], last[4];
        for (int k = 0; k < 4; k++) { cur[k] = ld3(pos, id[k]); last[k] = ld3(rest, 4 * (size_t)e + k); }
        v3 cc = mul3(add3(add3(add3(cur[0], cur[1]), cur[2]), cur[3]), 0.25f);
        v3 lc = mul3(add3(add3(add3(last[0], last[1]), last[2]), last[3]), 0.25f);
        v3 curC[4], lastC[4];
        for (int k = 0; k < 4; k++) { curC[k] = sub3(cur[k], cc); lastC[k] = sub3(last[k], lc); }
        /* TransposeMult(lastRest, current) (:90-105): column c of A = sum_k last_k[c] * cur_k */
        v3 A[3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int k = 0; k < 4; k++) {
            A[0] = add3(A[0], mul3(curC[k], lastC[k].x));
            A[1] = add3(A[1], mul3(curC[k], lastC[k].y));
            A[2] = add3(A[2], mul3(curC[k], lastC[k].z));
        }
        v4 ident = {0.0f, 0.0f, 0.0f, 1.0f};
        v4 rot = extract_rotation(A, ident);
        v4 qOld = {quat[4 * e], quat[4 * e + 1], quat[4 * e + 2], quat[4 * e + 3]};
        v4 qNew = normalize4(quat_mult(rot, qOld)); /* :181 */
        quat[4 * e] = qNew.x; quat[4 * e + 1] = qNew.y; quat[4 * e + 2] = qNew.z; quat[4 * e + 3] = qNew.w;
        /* K4: relative rotation between the new and the previous quaternion (:207,:237-239) */
        v4 conj = {-qOld.x, -qOld.y, -qOld.z, qOld.w};
        v4 rel = normalize4(quat_mult(qNew, normalize4(conj)));
        for (int k = 0; k < 4; k++) st3(rest, 4 * (size_t)e + k, add3(rotate(sub3(last[k], lc), rel), cc)); /* :253-256 */
    }
    /* K5 applyElem (:302-319): volume-weighted average of the goal corners, gather in table order */
    for (int i = 0; i < numVerts; i++) {
        v3 sumV = {0.0f, 0.0f, 0.0f};
        float sum = 0.0f;
        for (int j = tblStart[i]; j < tblStart[i + 1]; j++) {
            int e = tblEntries[j] / 4, k = tblEntries[j] % 4;
            float V = 1.0f / invRestVolume[e]; /* :220 */
            sumV = add3(sumV, mul3(ld3(rest, 4 * (size_t)e + k), V));
            sum = sum + V;
        }
        v3 r = {sumV.x / sum, sumV.y / sum, sumV.z / sum};
        st3(pos, i, r);
    }
    /* K6 collision (:340-354); the grab uses the linear vertex index (the shader's indexFromUV
     * decode at :336-338 is wrong for almost every texel and is deliberately not reproduced) */
    float fr = fminf(1.0f, dt * p->friction);
    for (int i = 0; i < numVerts; i++) {
        v3 x = ld3(pos, i);
        if (i == grabId) { x.x = grabPos[0]; x.y = grabPos[1]; x.z = grabPos[2]; }
        x.x = fminf(fmaxf(x.x, p->worldBounds[0]), p->worldBounds[3]);
        x.y = fminf(fmaxf(x.y, p->worldBounds[1]), p->worldBounds[4]);
        x.z = fminf(fmaxf(x.z, p->worldBounds[2]), p->worldBounds[5]);
        if (x.y < 0.0f) {
            x.y = 0.0f;
            v3 F = sub3(ld3(prev, i), x);
            x.x = x.x + F.x * fr;
            x.z = x.z + F.z * fr;
        }
        st3(pos, i, x);
    }
    /* K7 xpbdVelocity (:367-371): gravity enters the velocity AFTER the position update */
    for (int i = 0; i < numVerts; i++) {
        v3 d = sub3(ld3(pos, i), ld3(prev, i));
        v3 v = {d.x / dt + 0.0f * dt, d.y / dt + p->gravity * dt, d.z / dt + 0.0f * dt};
        st3(vel, i, v);
    }
}

/* The WebGL variant's render-time skinning of the embedded surface mesh (vertex shader patched into the vis material,
 * src/SoftbodyGPU.js:424-448): per surface vertex (tetNr, b0, b1, b2)
 *   lastTetWeight = 1.0 - (b0 + b1 + b2)                                   (:431; note the grouping, unlike the CPU class)
 *   position = ((p0 * b0 + p1 * b1) + p2 * b2) + p3 * lastTetWeight        (:432-435, vec4 arithmetic, f32)
 *   normal   = Rotate(objectNormal, tetQuaternion)                         (:438; objectNormal = the rest-pose vertex normal
 *              computeVertexNormals() left in the geometry at construction, :484-485 + :685)
 * Same f32 reading of GLSL as the passes above (every operation separately rounded). */
void oracle_polar_skin(int numVis, const float *visVerts, const int *tetIds, const float *pos, const float *quat,
                       const float *restNormals, float *outPos, float *outNrm) {
    for (int i = 0; i < numVis; i++) {
        const int e = (int)visVerts[4 * i];
        const float b0 = visVerts[4 * i + 1], b1 = visVerts[4 * i + 2], b2 = visVerts[4 * i + 3];
        const float b3 = 1.0f - ((b0 + b1) + b2);
        const v3 p0 = ld3(pos, (size_t)tetIds[4 * e]), p1 = ld3(pos, (size_t)tetIds[4 * e + 1]);
        const v3 p2 = ld3(pos, (size_t)tetIds[4 * e + 2]), p3 = ld3(pos, (size_t)tetIds[4 * e + 3]);
        const v3 r = add3(add3(add3(mul3(p0, b0), mul3(p1, b1)), mul3(p2, b2)), mul3(p3, b3));
        outPos[3 * i] = r.x; outPos[3 * i + 1] = r.y; outPos[3 * i + 2] = r.z;
        if (outNrm) {
            const v4 q = {quat[4 * e], quat[4 * e + 1], quat[4 * e + 2], quat[4 * e + 3]};
            const v3 n = rotate(ld3(restNormals, (size_t)i), q);
            outNrm[3 * i] = n.x; outNrm[3 * i + 1] = n.y; outNrm[3 * i + 2] = n.z;
        }
    }
}
