// kernels.cuh -- the substep kernels, templated on the arithmetic flavour (see device_math.cuh).
// Instantiated twice: kernels_exact.cu (-fmad=false) and kernels_fast.cu (FMA contraction on).
//
// HBM layout (all 16-byte records so every gather/scatter/stream access is one 128-bit transaction):
//   x4[v]   = (x, y, z, invMass)        vertex position + inverse mass   (src/Softbody.js:12,15)
//   prev4[v]= (x, y, z, -)              position at substep start         (:13)
//   vel4[v] = (vx, vy, vz, -)           velocity                          (:14)
//   tet stream, SoA of float4 planes in SOLVER order (level-sorted for GS, cluster-major for Jacobi):
//     A[e] = (Q0,Q1,Q2,Q3)  B[e] = (Q4,Q5,Q6,Q7)  C[e] = (Q8, invRestVolume, i0|i1<<16, i2|i3<<16)  [Jacobi: tile-local slots]
//     I[e] = (id0,id1,id2,id3) global vertex ids                                               [GS / gather solvers]
#pragma once
#include "device_math.cuh"
#include "launch.h"

namespace tsim {

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// ------------------------------------------------------------------------------------------------
// initPhysics, per-tet part (src/Softbody.js:64-80): Dm columns, V = det/6, adjugate inverse,
// per-corner mass V/4*rho, 1/V.  Always run in the EXACT flavour (one-off, must be bit-identical).
// ------------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void k_init_tets(int M, const float4 *__restrict__ x4, const int4 *__restrict__ ids, double density,
                            float *__restrict__ Q9, float *__restrict__ irv, double *__restrict__ pm) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M) return;
    int4 id = ids[e];
    float4 p0 = x4[id.x], p1 = x4[id.y], p2 = x4[id.z], p3 = x4[id.w];
    float a[3] = {p0.x, p0.y, p0.z}, b[3] = {p1.x, p1.y, p1.z}, c[3] = {p2.x, p2.y, p2.z}, d[3] = {p3.x, p3.y, p3.z};
    float m[9];
    ex_diff3(m + 0, b, a);
    ex_diff3(m + 3, c, a);
    ex_diff3(m + 6, d, a);
    double det = ex_det3(m);
    double V = det / 6.0;
    float q[9];
    if (det == 0.0) {
        // the reference's zero-determinant branch clobbers unrelated floats (src/Softbody.js:391-394);
        // here the tet simply gets a zero rest inverse and 1/V = inf, like the reference's :79
#pragma unroll
        for (int i = 0; i < 9; i++) q[i] = 0.0f;
    } else {
        double invDet = 1.0 / det;
        double a11 = m[0], a12 = m[3], a13 = m[6];
        double a21 = m[1], a22 = m[4], a23 = m[7];
        double a31 = m[2], a32 = m[5], a33 = m[8];
        q[0] = f32((a22 * a33 - a23 * a32) * invDet);
        q[3] = f32(-(a12 * a33 - a13 * a32) * invDet);
        q[6] = f32((a12 * a23 - a13 * a22) * invDet);
        q[1] = f32(-(a21 * a33 - a23 * a31) * invDet);
        q[4] = f32((a11 * a33 - a13 * a31) * invDet);
        q[7] = f32(-(a11 * a23 - a13 * a21) * invDet);
        q[2] = f32((a21 * a32 - a22 * a31) * invDet);
        q[5] = f32(-(a11 * a32 - a12 * a31) * invDet);
        q[8] = f32((a11 * a22 - a12 * a21) * invDet);
    }
#pragma unroll
    for (int i = 0; i < 9; i++) Q9[9 * (size_t)e + i] = q[i];
    irv[e] = f32(1.0 / V);
    pm[e] = V / 4.0 * density;
}

// initPhysics, per-vertex part (:75-78,:82-85): invMass[v] accumulates pm in f32 in tet order.  A
// vertex's corners are listed in ascending (tet, corner) order, which is exactly the order in which
// the reference's sequential loop touches that accumulator -> bit-identical.
template <bool EXACT>
__global__ void k_init_mass(int N, const int *__restrict__ cStart, const int *__restrict__ cEnt,
                            const double *__restrict__ pm, float4 *__restrict__ x4) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    float m = 0.0f;
    for (int j = cStart[v]; j < cStart[v + 1]; j++) m = f32((double)m + pm[cEnt[j] >> 2]);
    if (m != 0.0f) m = f32(1.0 / (double)m);
    x4[v].w = m;
}

// ------------------------------------------------------------------------------------------------
// predict: simulate() lines 198-202.  v += g*dt ; prev = x ; x += v*dt   (every vertex, no pin test)
// ------------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void k_predict(int N, float4 *__restrict__ x4, float4 *__restrict__ prev4, float4 *__restrict__ vel4,
                          const SubstepParams *__restrict__ sp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float4 x = x4[i], v = vel4[i];
    prev4[i] = x;
    if (EXACT) {
        double dt = sp->dt, g = sp->gravity;
        v.x = f32((double)v.x + 0.0 * dt);
        v.y = f32((double)v.y + g * dt);
        v.z = f32((double)v.z + 0.0 * dt);
        x.x = f32((double)x.x + (double)v.x * dt);
        x.y = f32((double)x.y + (double)v.y * dt);
        x.z = f32((double)x.z + (double)v.z * dt);
    } else {
        float dt = sp->dtF;
        v.y += sp->gDt;
        x.x = fmaf(v.x, dt, x.x);
        x.y = fmaf(v.y, dt, x.y);
        x.z = fmaf(v.z, dt, x.z);
    }
    vel4[i] = v;
    x4[i] = x;
}

// Math.max(lo, Math.min(hi, x)) with JS NaN propagation (src/Softbody.js:350-355)
__device__ __forceinline__ double js_clamp(double x, double lo, double hi) {
    double m = (x != x || hi != hi) ? __longlong_as_double(0x7ff8000000000000LL) : (hi < x ? hi : x);
    return (m != m || lo != lo) ? __longlong_as_double(0x7ff8000000000000LL) : (lo > m ? lo : m);
}

// post: simulate() lines 213-239 for one vertex: bounds clamp, floor + friction, grab, velocity.
template <bool EXACT>
__device__ __forceinline__ void post_vertex(int i, float4 &x, const float4 prev, float4 &v, const SubstepParams *sp) {
    if (EXACT) {
        double dt = sp->dt;
        x.x = f32(js_clamp((double)x.x, sp->lo[0], sp->hi[0]));
        x.y = f32(js_clamp((double)x.y, sp->lo[1], sp->hi[1]));
        x.z = f32(js_clamp((double)x.z, sp->lo[2], sp->hi[2]));
        if (x.y < 0.0f) {
            x.y = 0.0f;
            float Fx = f32((double)prev.x - (double)x.x), Fz = f32((double)prev.z - (double)x.z);
            double k = dt * sp->friction;
            k = (k != k) ? k : (k < 1.0 ? k : 1.0);
            x.x = f32((double)x.x + (double)Fx * k);
            x.z = f32((double)x.z + (double)Fz * k);
        }
        if (i == sp->grabId) { x.x = f32(sp->grab[0]); x.y = f32(sp->grab[1]); x.z = f32(sp->grab[2]); }
        double inv = sp->invDtD;
        v.x = f32(((double)x.x - (double)prev.x) * inv);
        v.y = f32(((double)x.y - (double)prev.y) * inv);
        v.z = f32(((double)x.z - (double)prev.z) * inv);
    } else {
        x.x = fmaxf(sp->loF[0], fminf(sp->hiF[0], x.x));
        x.y = fmaxf(sp->loF[1], fminf(sp->hiF[1], x.y));
        x.z = fmaxf(sp->loF[2], fminf(sp->hiF[2], x.z));
        if (x.y < 0.0f) {
            x.y = 0.0f;
            x.x = fmaf(prev.x - x.x, sp->fric, x.x);
            x.z = fmaf(prev.z - x.z, sp->fric, x.z);
        }
        if (i == sp->grabId) { x.x = sp->grabF[0]; x.y = sp->grabF[1]; x.z = sp->grabF[2]; }
        float inv = sp->invDt;
        v.x = (x.x - prev.x) * inv;
        v.y = (x.y - prev.y) * inv;
        v.z = (x.z - prev.z) * inv;
    }
}

template <bool EXACT>
__global__ void k_post(int N, float4 *__restrict__ x4, const float4 *__restrict__ prev4, float4 *__restrict__ vel4,
                       const SubstepParams *__restrict__ sp, const int *__restrict__ vertId) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float4 x = x4[i], p = prev4[i], v;
    v.w = 0.0f;
    post_vertex<EXACT>(vertId ? vertId[i] : i, x, p, v, sp);
    x4[i] = x;
    vel4[i] = v;
}

// ------------------------------------------------------------------------------------------------
// Gauss-Seidel Neo-Hookean, one dependency level (or colour) per launch, in place in global memory.
// Tets of one level share no vertex, so plain loads/stores are race-free.  Stream is level-sorted.
// ------------------------------------------------------------------------------------------------
template <bool EXACT>
__device__ __forceinline__ double gs_solve_one(float4 *x4, int4 id, float4 A, float4 B, float4 C,
                                               const SubstepParams *sp) {
    float4 p0 = x4[id.x], p1 = x4[id.y], p2 = x4[id.z], p3 = x4[id.w];
    const float Q[9] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x};
    double volm1;
    if (EXACT) {
        float y[12] = {p0.x, p0.y, p0.z, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z, p3.x, p3.y, p3.z};
        const float w4[4] = {p0.w, p1.w, p2.w, p3.w};
        volm1 = nh_solve_exact(y, w4, Q, C.y, sp->alphaDevD, sp->alphaVolD, sp->volOverDevD);
        p0.x = y[0]; p0.y = y[1]; p0.z = y[2];
        p1.x = y[3]; p1.y = y[4]; p1.z = y[5];
        p2.x = y[6]; p2.y = y[7]; p2.z = y[8];
        p3.x = y[9]; p3.y = y[10]; p3.z = y[11];
    } else {
        V3 p[4] = {{p0.x, p0.y, p0.z}, {p1.x, p1.y, p1.z}, {p2.x, p2.y, p2.z}, {p3.x, p3.y, p3.z}};
        const float w[4] = {p0.w, p1.w, p2.w, p3.w};
        volm1 = nh_solve_fast(p, w, Q, C.y, sp->alphaDev, sp->alphaVol, sp->gammaVol);
        p0.x = p[0].x; p0.y = p[0].y; p0.z = p[0].z;
        p1.x = p[1].x; p1.y = p[1].y; p1.z = p[1].z;
        p2.x = p[2].x; p2.y = p[2].y; p2.z = p[2].z;
        p3.x = p[3].x; p3.y = p[3].y; p3.z = p[3].z;
    }
    x4[id.x] = p0; x4[id.y] = p1; x4[id.z] = p2; x4[id.w] = p3;
    return volm1;
}

template <bool EXACT>
__global__ void k_gs_level(int begin, int end, float4 *__restrict__ x4, const int4 *__restrict__ I,
                           const float4 *__restrict__ A, const float4 *__restrict__ B, const float4 *__restrict__ C,
                           const int *__restrict__ order, double *__restrict__ volTerm,
                           const SubstepParams *__restrict__ sp) {
    int e = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= end) return;
    double vm1 = gs_solve_one<EXACT>(x4, I[e], ldg4(A + e), ldg4(B + e), ldg4(C + e), sp);
    if (volTerm) volTerm[order[e]] = vm1;
}

// ------------------------------------------------------------------------------------------------
// Gauss-Seidel Neo-Hookean, one CTA per connected component ("body"), the whole substep in ONE
// launch: predict -> level sweep -> post, with the body's vertices resident in shared memory
// (Dragon: 1,234 x 16 B = 19.7 KB) and a block barrier between dependency levels.  Tet records are
// streamed level by level from HBM/L2 with the next level's record prefetched into registers.
// Vertex ids in I are body-local.  This is the kernel behind BASELINE configs 1, 3 and 5.
// ------------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void k_gs_body(const BodyDesc *__restrict__ bodies, const int *__restrict__ levelStart,
                          float4 *__restrict__ x4, float4 *__restrict__ prev4, float4 *__restrict__ vel4,
                          const int4 *__restrict__ I, const float4 *__restrict__ A, const float4 *__restrict__ B,
                          const float4 *__restrict__ C, const int *__restrict__ order, double *__restrict__ volTerm,
                          const SubstepParams *__restrict__ sp, const int *__restrict__ vertId) {
    extern __shared__ float4 sx[];
    const BodyDesc bd = bodies[blockIdx.x];
    const int nv = bd.vertEnd - bd.vertBegin;
    const int tid = threadIdx.x, nt = blockDim.x;
    // predict (simulate() :198-202), positions land in shared memory
    for (int j = tid; j < nv; j += nt) {
        int i = bd.vertBegin + j;
        float4 x = x4[i], v = vel4[i];
        prev4[i] = x;
        if (EXACT) {
            double dt = sp->dt, g = sp->gravity;
            v.x = f32((double)v.x + 0.0 * dt);
            v.y = f32((double)v.y + g * dt);
            v.z = f32((double)v.z + 0.0 * dt);
            x.x = f32((double)x.x + (double)v.x * dt);
            x.y = f32((double)x.y + (double)v.y * dt);
            x.z = f32((double)x.z + (double)v.z * dt);
        } else {
            float dt = sp->dtF;
            v.y += sp->gDt;
            x.x = fmaf(v.x, dt, x.x);
            x.y = fmaf(v.y, dt, x.y);
            x.z = fmaf(v.z, dt, x.z);
        }
        sx[j] = x;
    }
    // this body's level offsets -> shared memory (one pointer chase fewer per level)
    int *sLevel = reinterpret_cast<int *>(sx + nv);
    const int nLev = bd.levelEnd - bd.levelBegin;
    for (int j = tid; j <= nLev; j += nt) sLevel[j] = levelStart[bd.levelBegin + j];
    __syncthreads();
    // level sweep.  A level is a handful of tets (Dragon: <= 22), so the sweep is a chain of ~700
    // dependent steps and each step's record fetch (HBM/L2, ~1 us) would sit on the critical path:
    // the first record of level l+1 is loaded into registers before level l is solved.
    int4 rI = make_int4(0, 0, 0, 0);
    float4 rA = make_float4(0.f, 0.f, 0.f, 0.f), rB = rA, rC = rA;
    int rO = 0;   // the tet's caller index (volError slot) travels with the record: a load of order[] inside the level would sit on its path
    if (nLev > 0 && sLevel[0] + tid < sLevel[1]) {
        const int t = sLevel[0] + tid;
        rI = I[t]; rA = ldg4(A + t); rB = ldg4(B + t); rC = ldg4(C + t);
        if (volTerm) rO = order[t];
    }
    for (int l = 0; l < nLev; l++) {
        const int b = sLevel[l], e = sLevel[l + 1];
        const int4 cI = rI;
        const float4 cA = rA, cB = rB, cC = rC;
        const int cO = rO;
        if (l + 1 < nLev && e + tid < sLevel[l + 2]) {  // prefetch for the next level
            const int t = e + tid;
            rI = I[t]; rA = ldg4(A + t); rB = ldg4(B + t); rC = ldg4(C + t);
            if (volTerm) rO = order[t];
        }
        if (b + tid < e) {
            double vm1 = gs_solve_one<EXACT>(sx, cI, cA, cB, cC, sp);
            if (volTerm) volTerm[cO] = vm1;
        }
        for (int t = b + tid + nt; t < e; t += nt) {  // levels wider than the CTA (rare)
            double vm1 = gs_solve_one<EXACT>(sx, I[t], ldg4(A + t), ldg4(B + t), ldg4(C + t), sp);
            if (volTerm) volTerm[order[t]] = vm1;
        }
        __syncthreads();
    }
    // post (simulate() :213-239)
    for (int j = tid; j < nv; j += nt) {
        int i = bd.vertBegin + j;
        float4 x = sx[j], p = prev4[i], v;
        v.w = 0.0f;
        post_vertex<EXACT>(vertId ? vertId[i] : i, x, p, v, sp);
        x4[i] = x;
        vel4[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Jacobi Neo-Hookean, gather formulation (the parity / reference-structure path; the throughput
// path is k_jacobi_tilesN in kernels_fast.cu).  Mirrors the WebGL solver's own split
// (src/SoftbodyGPU.js K4 writes 4 corner results per tet, K5 gathers them per vertex):
//   k_jacobi_tet    : every tet runs solveElem on a private copy, writes dx for its 4 corners
//   k_jacobi_gather : every vertex sums its corners' dx in ascending (tet, corner) order (f32) and
//                     moves by sum * (1/valence)
// ------------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void k_jacobi_tet(int M, const float4 *__restrict__ x4, const int4 *__restrict__ I,
                             const float4 *__restrict__ A, const float4 *__restrict__ B, const float4 *__restrict__ C,
                             float4 *__restrict__ dx /* [4M], corner-major per tet */, double *__restrict__ volTerm,
                             const SubstepParams *__restrict__ sp) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M) return;
    int4 id = I[e];
    float4 a = ldg4(A + e), b = ldg4(B + e), c = ldg4(C + e);
    float4 p0 = x4[id.x], p1 = x4[id.y], p2 = x4[id.z], p3 = x4[id.w];
    const float Q[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x};
    float4 d0, d1, d2, d3;
    double vm1;
    if (EXACT) {
        float y[12] = {p0.x, p0.y, p0.z, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z, p3.x, p3.y, p3.z};
        const float w4[4] = {p0.w, p1.w, p2.w, p3.w};
        vm1 = nh_solve_exact(y, w4, Q, c.y, sp->alphaDevD, sp->alphaVolD, sp->volOverDevD);
        d0 = make_float4(f32((double)y[0] - (double)p0.x), f32((double)y[1] - (double)p0.y), f32((double)y[2] - (double)p0.z), 0.f);
        d1 = make_float4(f32((double)y[3] - (double)p1.x), f32((double)y[4] - (double)p1.y), f32((double)y[5] - (double)p1.z), 0.f);
        d2 = make_float4(f32((double)y[6] - (double)p2.x), f32((double)y[7] - (double)p2.y), f32((double)y[8] - (double)p2.z), 0.f);
        d3 = make_float4(f32((double)y[9] - (double)p3.x), f32((double)y[10] - (double)p3.y), f32((double)y[11] - (double)p3.z), 0.f);
    } else {
        V3 p[4] = {{p0.x, p0.y, p0.z}, {p1.x, p1.y, p1.z}, {p2.x, p2.y, p2.z}, {p3.x, p3.y, p3.z}};
        const float w[4] = {p0.w, p1.w, p2.w, p3.w};
        vm1 = nh_solve_fast(p, w, Q, c.y, sp->alphaDev, sp->alphaVol, sp->gammaVol);
        d0 = make_float4(p[0].x - p0.x, p[0].y - p0.y, p[0].z - p0.z, 0.f);
        d1 = make_float4(p[1].x - p1.x, p[1].y - p1.y, p[1].z - p1.z, 0.f);
        d2 = make_float4(p[2].x - p2.x, p[2].y - p2.y, p[2].z - p2.z, 0.f);
        d3 = make_float4(p[3].x - p3.x, p[3].y - p3.y, p[3].z - p3.z, 0.f);
    }
    dx[4 * (size_t)e + 0] = d0; dx[4 * (size_t)e + 1] = d1; dx[4 * (size_t)e + 2] = d2; dx[4 * (size_t)e + 3] = d3;
    if (volTerm) volTerm[e] = vm1;
}

template <bool EXACT>
__global__ void k_jacobi_gather(int N, float4 *__restrict__ x4, const int *__restrict__ cStart,
                                const int *__restrict__ cEnt, const float4 *__restrict__ dx) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    int b = cStart[v], e = cStart[v + 1];
    if (b == e) return;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int j = b; j < e; j++) {
        float4 d = dx[cEnt[j]];
        sx = sx + d.x; sy = sy + d.y; sz = sz + d.z;
    }
    float inv = 1.0f / (float)(e - b);
    float4 x = x4[v];
    if (EXACT) {
        x.x = x.x + sx * inv; x.y = x.y + sy * inv; x.z = x.z + sz * inv;  // -fmad=false: mul then add
    } else {
        x.x = fmaf(sx, inv, x.x); x.y = fmaf(sy, inv, x.y); x.z = fmaf(sz, inv, x.z);
    }
    x4[v] = x;
}

// ------------------------------------------------------------------------------------------------
// Polar-decomposition shape matching (SoftBodyGPU).  Pass structure of the reference collapsed:
//   K1+K2 -> k_polar_integrate      (src/SoftbodyGPU.js:59-74)
//   K3+K4 -> k_polar_tet            (:80-262)  goal corners + quaternion per tet
//   K5+K6+K7 -> k_polar_vertex      (:272-376) volume-weighted gather, collision, velocity
// rest[4e+k] = (goal corner k of tet e, V_e)  -- the reference's `elems` MRT (:259-262).
// ------------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void k_polar_integrate(int N, float4 *__restrict__ x4, float4 *__restrict__ prev4,
                                  const float4 *__restrict__ vel4, const SubstepParams *__restrict__ sp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float4 x = x4[i], v = vel4[i];
    prev4[i] = x;
    float dt = sp->dtF;
    x.x = x.x + v.x * dt; x.y = x.y + v.y * dt; x.z = x.z + v.z * dt;
    x4[i] = x;
}

template <bool EXACT>
__global__ void k_polar_tet(int M, const float4 *__restrict__ x4, const int4 *__restrict__ I,
                            float4 *__restrict__ rest, float4 *__restrict__ quat) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M) return;
    int4 id = I[e];
    float4 p0 = x4[id.x], p1 = x4[id.y], p2 = x4[id.z], p3 = x4[id.w];
    float4 r0 = rest[4 * (size_t)e], r1 = rest[4 * (size_t)e + 1], r2 = rest[4 * (size_t)e + 2], r3 = rest[4 * (size_t)e + 3];
    float4 q4 = quat[e];
    V3 cur[4] = {{p0.x, p0.y, p0.z}, {p1.x, p1.y, p1.z}, {p2.x, p2.y, p2.z}, {p3.x, p3.y, p3.z}};
    V3 last[4] = {{r0.x, r0.y, r0.z}, {r1.x, r1.y, r1.z}, {r2.x, r2.y, r2.z}, {r3.x, r3.y, r3.z}};
    Q4 q = {q4.x, q4.y, q4.z, q4.w};
    polar_solve<EXACT>(cur, last, q);
    quat[e] = make_float4(q.x, q.y, q.z, q.w);
    rest[4 * (size_t)e + 0] = make_float4(last[0].x, last[0].y, last[0].z, r0.w);
    rest[4 * (size_t)e + 1] = make_float4(last[1].x, last[1].y, last[1].z, r1.w);
    rest[4 * (size_t)e + 2] = make_float4(last[2].x, last[2].y, last[2].z, r2.w);
    rest[4 * (size_t)e + 3] = make_float4(last[3].x, last[3].y, last[3].z, r3.w);
}

template <bool EXACT>
__global__ void k_polar_vertex(int N, float4 *__restrict__ x4, const float4 *__restrict__ prev4,
                               float4 *__restrict__ vel4, const int *__restrict__ tStart,
                               const int *__restrict__ tEnt, const float4 *__restrict__ rest,
                               const SubstepParams *__restrict__ sp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f, sum = 0.0f;
    for (int j = tStart[i]; j < tStart[i + 1]; j++) {  // K5, :305-319
        float4 g = rest[tEnt[j]];
        sx = sx + g.x * g.w; sy = sy + g.y * g.w; sz = sz + g.z * g.w;
        sum = sum + g.w;
    }
    float4 x = x4[i], p = prev4[i];
    x.x = sx / sum; x.y = sy / sum; x.z = sz / sum;
    // K6, :340-354 (grab by linear index; see DESIGN.md on the reference's indexFromUV)
    if (i == sp->grabId) { x.x = sp->grabF[0]; x.y = sp->grabF[1]; x.z = sp->grabF[2]; }
    x.x = fminf(fmaxf(x.x, sp->loF[0]), sp->hiF[0]);
    x.y = fminf(fmaxf(x.y, sp->loF[1]), sp->hiF[1]);
    x.z = fminf(fmaxf(x.z, sp->loF[2]), sp->hiF[2]);
    if (x.y < 0.0f) {
        x.y = 0.0f;
        float fr = fminf(1.0f, sp->dtF * sp->frictionF);
        x.x = x.x + (p.x - x.x) * fr;
        x.z = x.z + (p.z - x.z) * fr;
    }
    x4[i] = x;
    // K7, :367-371: gravity enters the velocity after the position update
    float dt = sp->dtF;
    float4 v;
    v.x = (x.x - p.x) / dt + 0.0f * dt;
    v.y = (x.y - p.y) / dt + sp->gravityF * dt;
    v.z = (x.z - p.z) / dt + 0.0f * dt;
    v.w = 0.0f;
    vel4[i] = v;
}

// ------------------------------------------------------------------------------------------------
// updateVisMesh (src/Softbody.js:259-273): barycentric skinning, and three.js computeVertexNormals
// as a per-vertex gather over incident triangles in ascending triangle order (each (vertex,
// triangle) pair once: the reference reads nA,nB,nC before writing any, three.module.js:11173-11183).
// ------------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void k_skin(int nVis, const float4 *__restrict__ vis, const int4 *__restrict__ ids,
                       const float4 *__restrict__ x4, float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nVis) return;
    float4 vv = vis[i];
    int4 id = ids[(int)vv.x];
    float4 p0 = x4[id.x], p1 = x4[id.y], p2 = x4[id.z], p3 = x4[id.w];
    if (EXACT) {
        double b0 = vv.y, b1 = vv.z, b2 = vv.w, b3 = 1.0 - b0 - b1 - b2;
        float o[3] = {0.0f, 0.0f, 0.0f};
        const float a0[3] = {p0.x, p0.y, p0.z}, a1[3] = {p1.x, p1.y, p1.z}, a2[3] = {p2.x, p2.y, p2.z}, a3[3] = {p3.x, p3.y, p3.z};
        ex_axpy3(o, a0, b0); ex_axpy3(o, a1, b1); ex_axpy3(o, a2, b2); ex_axpy3(o, a3, b3);
        out[3 * (size_t)i] = o[0]; out[3 * (size_t)i + 1] = o[1]; out[3 * (size_t)i + 2] = o[2];
    } else {
        float b0 = vv.y, b1 = vv.z, b2 = vv.w, b3 = 1.0f - b0 - b1 - b2;
        out[3 * (size_t)i] = fmaf(p3.x, b3, fmaf(p2.x, b2, fmaf(p1.x, b1, p0.x * b0)));
        out[3 * (size_t)i + 1] = fmaf(p3.y, b3, fmaf(p2.y, b2, fmaf(p1.y, b1, p0.y * b0)));
        out[3 * (size_t)i + 2] = fmaf(p3.z, b3, fmaf(p2.z, b2, fmaf(p1.z, b1, p0.z * b0)));
    }
}

// The WebGL variant's vertex-shader skinning (src/SoftbodyGPU.js:424-448): position = ((p0 b0 + p1 b1) + p2 b2) + p3 (1 - (b0 + b1 + b2))
// in f32, normal = Rotate(rest normal, quaternion of the surface vertex's tet).  quatOf(e) abstracts where the tet's
// quaternion lives (a plain array in the reference-structure solver, the tile blocks in the tiled one).
template <bool EXACT>
__global__ void k_skin_polar(int nVis, const float4 *__restrict__ vis, const int4 *__restrict__ ids, const float4 *__restrict__ x4,
                             const float4 *__restrict__ quat, const unsigned char *__restrict__ tileTets, const int *__restrict__ tetRecord,
                             int T, const float *__restrict__ restNrm, float *__restrict__ outPos, float *__restrict__ outNrm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nVis) return;
    const float4 vv = vis[i];
    const int e = (int)vv.x;
    const int4 id = ids[e];
    const float4 p0 = x4[id.x], p1 = x4[id.y], p2 = x4[id.z], p3 = x4[id.w];
    const float b0 = vv.y, b1 = vv.z, b2 = vv.w, b3 = 1.0f - ((b0 + b1) + b2);
    outPos[3 * (size_t)i] = ((p0.x * b0 + p1.x * b1) + p2.x * b2) + p3.x * b3;   // EXACT: this unit is compiled with -fmad=false
    outPos[3 * (size_t)i + 1] = ((p0.y * b0 + p1.y * b1) + p2.y * b2) + p3.y * b3;
    outPos[3 * (size_t)i + 2] = ((p0.z * b0 + p1.z * b1) + p2.z * b2) + p3.z * b3;
    if (outNrm) {
        float4 q4;
        if (tetRecord) { const int r = tetRecord[e]; q4 = *reinterpret_cast<const float4 *>(tileTets + (size_t)(r / T) * T * 80 + (size_t)T * 48 + (size_t)(r % T) * 16); }
        else q4 = quat[e];
        const V3 n = pl_rotate({restNrm[3 * (size_t)i], restNrm[3 * (size_t)i + 1], restNrm[3 * (size_t)i + 2]}, {q4.x, q4.y, q4.z, q4.w});
        outNrm[3 * (size_t)i] = n.x; outNrm[3 * (size_t)i + 1] = n.y; outNrm[3 * (size_t)i + 2] = n.z;
    }
}

template <bool EXACT>
__global__ void k_normals(int nVis, const float *__restrict__ pos, const int *__restrict__ tri,
                          const int *__restrict__ vtStart, const int *__restrict__ vtEnt, float *__restrict__ nrm) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nVis) return;
    float nx = 0.0f, ny = 0.0f, nz = 0.0f;
    for (int j = vtStart[v]; j < vtStart[v + 1]; j++) {
        int t = vtEnt[j];
        int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        double cbx = (double)pos[3 * c] - (double)pos[3 * b], cby = (double)pos[3 * c + 1] - (double)pos[3 * b + 1],
               cbz = (double)pos[3 * c + 2] - (double)pos[3 * b + 2];
        double abx = (double)pos[3 * a] - (double)pos[3 * b], aby = (double)pos[3 * a + 1] - (double)pos[3 * b + 1],
               abz = (double)pos[3 * a + 2] - (double)pos[3 * b + 2];
        double x = cby * abz - cbz * aby, y = cbz * abx - cbx * abz, z = cbx * aby - cby * abx;
        nx = f32((double)nx + x); ny = f32((double)ny + y); nz = f32((double)nz + z);
    }
    double x = nx, y = ny, z = nz;
    double len = sqrt(x * x + y * y + z * z);
    double s = 1.0 / ((len != 0.0 && len == len) ? len : 1.0);
    nrm[3 * (size_t)v] = f32(x * s); nrm[3 * (size_t)v + 1] = f32(y * s); nrm[3 * (size_t)v + 2] = f32(z * s);
}

}  // namespace tsim
