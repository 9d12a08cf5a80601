// device_math.cuh -- per-tet and per-vertex arithmetic of the substep, in two flavours.
//
//   EXACT = true : the reference's arithmetic operation for operation:
//                  Neo-Hookean = JS typed-array semantics (f64 expressions, one f32 rounding per
//                  store); polar = separately rounded f32.  The translation unit that instantiates
//                  it is compiled with -fmad=false so nothing is contracted.
//   EXACT = false: f32 with FMA contraction, rsqrt / fast reciprocal, algebraically regrouped to
//                  cut the instruction count -- the throughput flavour (no tensor cores: there is
//                  no dense contraction in this path, only 3x3 products per tet).
//
// Layout conventions follow the reference: invRestPose is column-major, Q[3*col + row]
// (src/Softbody.js:359-361).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsim {

// Per-launch parameters.  Lives in device memory, refreshed by a cudaMemcpyAsync ahead of every
// simulate/step so a captured CUDA graph never has to be re-instantiated when the GUI-mutable
// physicsParams (src/main.js:37-42) change.
struct SubstepParams {
    double dt, gravity, friction, devCompliance, volCompliance;
    double lo[3], hi[3];
    double grab[3];
    int grabId;
    int pad_;
    // derived, reference (f64) flavour: the per-launch constants of applyToElem / solveElem / the velocity update, formed ONCE
    // with the same IEEE operations in the same order as the reference forms them per tet / per call (so the result is the
    // same double): compliance / dt / dt (src/Softbody.js:187), volCompliance / devCompliance (:161), 1.0 / dt (:239).
    // Each f64 division on the device is a ~25-instruction dependent sequence on every tet's critical path otherwise.
    double alphaDevD, alphaVolD, volOverDevD, invDtD;
    // derived, f32 flavour
    float dtF, gDt, invDt, fric, alphaDev, alphaVol, gammaVol, gravityF, frictionF;
    float loF[3], hiF[3], grabF[3];
};

__device__ __forceinline__ float f32(double v) { return __double2float_rn(v); }

// ---------------------------------------------------------------------------------------------
// EXACT Neo-Hookean: restates solveElem/applyToElem (src/Softbody.js:91-193).
// y = 4 private vertex copies (12 floats), w4 = their inverse masses.
// ---------------------------------------------------------------------------------------------
struct ExactScratch {
    float P[9], F[9], dF[9], g[12];
};

__device__ __forceinline__ void ex_axpy3(float *a, const float *b, double s) {  // vecAdd :316-321
    a[0] = f32((double)a[0] + (double)b[0] * s);
    a[1] = f32((double)a[1] + (double)b[1] * s);
    a[2] = f32((double)a[2] + (double)b[2] * s);
}
__device__ __forceinline__ void ex_diff3(float *d, const float *a, const float *b) {  // vecSetDiff :323-328
    d[0] = f32(((double)a[0] - (double)b[0]) * 1.0);
    d[1] = f32(((double)a[1] - (double)b[1]) * 1.0);
    d[2] = f32(((double)a[2] - (double)b[2]) * 1.0);
}
__device__ __forceinline__ double ex_len2(const float *a) {  // vecLengthSquared :330-334
    double a0 = a[0], a1 = a[1], a2 = a[2];
    return a0 * a0 + a1 * a1 + a2 * a2;
}
__device__ __forceinline__ void ex_cross3(float *a, const float *b, const float *c) {  // :343-348
    double b0 = b[0], b1 = b[1], b2 = b[2], c0 = c[0], c1 = c[1], c2 = c[2];
    a[0] = f32(b1 * c2 - b2 * c1);
    a[1] = f32(b2 * c0 - b0 * c2);
    a[2] = f32(b0 * c1 - b1 * c0);
}
__device__ __forceinline__ double ex_det3(const float *m) {  // matGetDeterminant :381-387
    double a11 = m[0], a12 = m[3], a13 = m[6];
    double a21 = m[1], a22 = m[4], a23 = m[7];
    double a31 = m[2], a32 = m[5], a33 = m[8];
    return a11 * a22 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - a13 * a22 * a31 - a12 * a21 * a33 - a11 * a23 * a32;
}
__device__ __forceinline__ void ex_matmul3(float *dst, const float *A, const float *B) {  // :363-379
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double b0 = B[3 * k], b1 = B[3 * k + 1], b2 = B[3 * k + 2];
        dst[3 * k] = 0.0f; dst[3 * k + 1] = 0.0f; dst[3 * k + 2] = 0.0f;
        ex_axpy3(dst + 3 * k, A + 0, b0);
        ex_axpy3(dst + 3 * k, A + 3, b1);
        ex_axpy3(dst + 3 * k, A + 6, b2);
    }
}
__device__ __forceinline__ void ex_apply(ExactScratch &s, float *y, const float *w4, double C, double complianceOverDt2,
                                         double irv) {  // applyToElem :168-193; complianceOverDt2 = compliance / dt / dt
    if (C == 0.0) return;
    float *g = s.g;
    g[0] = 0.0f; g[1] = 0.0f; g[2] = 0.0f;
    ex_axpy3(g, g + 3, -1.0);
    ex_axpy3(g, g + 6, -1.0);
    ex_axpy3(g, g + 9, -1.0);
    double w = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) w += ex_len2(g + 3 * i) * (double)w4[i];
    if (w == 0.0) return;
    double alpha = complianceOverDt2 * irv;
    double dlambda = -C / (w + alpha);
#pragma unroll
    for (int i = 0; i < 4; i++) ex_axpy3(y + 3 * i, g + 3 * i, dlambda * (double)w4[i]);
}
// Returns vol - 1 (the volError term, :163).
__device__ __forceinline__ double nh_solve_exact(float *y, const float *w4, const float *Q, float irv, double alphaDevD,
                                                 double alphaVolD, double volOverDevD) {
    ExactScratch s;
    ex_diff3(s.P + 0, y + 3, y);
    ex_diff3(s.P + 3, y + 6, y);
    ex_diff3(s.P + 6, y + 9, y);
    ex_matmul3(s.F, s.P, Q);
    double r_s = sqrt(ex_len2(s.F) + ex_len2(s.F + 3) + ex_len2(s.F + 6));
    double r_s_inv = 1.0 / r_s;
#pragma unroll
    for (int k = 1; k <= 3; k++) {
        float *g = s.g + 3 * k;
        g[0] = 0.0f; g[1] = 0.0f; g[2] = 0.0f;
        ex_axpy3(g, s.F + 0, r_s_inv * (double)Q[0 + (k - 1)]);
        ex_axpy3(g, s.F + 3, r_s_inv * (double)Q[3 + (k - 1)]);
        ex_axpy3(g, s.F + 6, r_s_inv * (double)Q[6 + (k - 1)]);
    }
    ex_apply(s, y, w4, r_s, alphaDevD, (double)irv);

    ex_diff3(s.P + 0, y + 3, y);
    ex_diff3(s.P + 3, y + 6, y);
    ex_diff3(s.P + 6, y + 9, y);
    ex_matmul3(s.F, s.P, Q);
    ex_cross3(s.dF + 0, s.F + 3, s.F + 6);
    ex_cross3(s.dF + 3, s.F + 6, s.F + 0);
    ex_cross3(s.dF + 6, s.F + 0, s.F + 3);
#pragma unroll
    for (int k = 1; k <= 3; k++) {
        float *g = s.g + 3 * k;
        g[0] = 0.0f; g[1] = 0.0f; g[2] = 0.0f;
        ex_axpy3(g, s.dF + 0, (double)Q[0 + (k - 1)]);
        ex_axpy3(g, s.dF + 3, (double)Q[3 + (k - 1)]);
        ex_axpy3(g, s.dF + 6, (double)Q[6 + (k - 1)]);
    }
    double vol = ex_det3(s.F);
    double C = vol - 1.0 - volOverDevD;
    ex_apply(s, y, w4, C, alphaVolD, (double)irv);
    return vol - 1.0;
}

// ---------------------------------------------------------------------------------------------
// FAST Neo-Hookean: same two constraints, f32/FMA, regrouped:
//   unscaled gradients G = F Q^T are formed once and the 1/||F|| factor is folded into the step
//   scale; g0 = -(g1+g2+g3); the hydrostatic gradient uses det F = F0 . (F1 x F2).
// p[4] in/out (positions), w[4] inverse masses.  Returns det F - 1 (sampled between the two
// projections like the reference, :159-163).
// ---------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 fma3(V3 a, float s, V3 c) { return {fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z)}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return {fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x)};
}

__device__ __forceinline__ float nh_solve_fast(V3 p[4], const float w[4], const float Q[9], float irv, float alphaDev,
                                               float alphaVol, float gammaVol) {
    // ---- deviatoric: C = ||F||_F ----
    V3 P0 = p[1] - p[0], P1 = p[2] - p[0], P2 = p[3] - p[0];
    V3 F0 = fma3(P2, Q[2], fma3(P1, Q[1], P0 * Q[0]));
    V3 F1 = fma3(P2, Q[5], fma3(P1, Q[4], P0 * Q[3]));
    V3 F2 = fma3(P2, Q[8], fma3(P1, Q[7], P0 * Q[6]));
    float rs2 = dot(F0, F0) + dot(F1, F1) + dot(F2, F2);
    {
        V3 G1 = fma3(F2, Q[6], fma3(F1, Q[3], F0 * Q[0]));
        V3 G2 = fma3(F2, Q[7], fma3(F1, Q[4], F0 * Q[1]));
        V3 G3 = fma3(F2, Q[8], fma3(F1, Q[5], F0 * Q[2]));
        V3 G0 = {-(G1.x + G2.x + G3.x), -(G1.y + G2.y + G3.y), -(G1.z + G2.z + G3.z)};
        // w_true = (1/rs2) * sum w_i |G_i|^2 ; dlambda = -rs / (w_true + alpha)
        // step_i = G_i * (1/rs) * dlambda * w_i = -G_i * w_i / (wG/rs2 + alpha) = -G_i * w_i * rs2 / (wG + alpha*rs2)
        float wG = fmaf(w[3], dot(G3, G3), fmaf(w[2], dot(G2, G2), fmaf(w[1], dot(G1, G1), w[0] * dot(G0, G0))));
        float den = fmaf(alphaDev * irv, rs2, wG);
        float s = (rs2 > 0.0f && wG > 0.0f) ? -__fdividef(rs2, den) : 0.0f;
        p[0] = fma3(G0, s * w[0], p[0]);
        p[1] = fma3(G1, s * w[1], p[1]);
        p[2] = fma3(G2, s * w[2], p[2]);
        p[3] = fma3(G3, s * w[3], p[3]);
    }
    // ---- hydrostatic: C = det F - 1 - volC/devC, from the UPDATED positions ----
    P0 = p[1] - p[0]; P1 = p[2] - p[0]; P2 = p[3] - p[0];
    F0 = fma3(P2, Q[2], fma3(P1, Q[1], P0 * Q[0]));
    F1 = fma3(P2, Q[5], fma3(P1, Q[4], P0 * Q[3]));
    F2 = fma3(P2, Q[8], fma3(P1, Q[7], P0 * Q[6]));
    V3 D0 = cross(F1, F2), D1 = cross(F2, F0), D2 = cross(F0, F1);
    float vol = dot(F0, D0);
    V3 G1 = fma3(D2, Q[6], fma3(D1, Q[3], D0 * Q[0]));
    V3 G2 = fma3(D2, Q[7], fma3(D1, Q[4], D0 * Q[1]));
    V3 G3 = fma3(D2, Q[8], fma3(D1, Q[5], D0 * Q[2]));
    V3 G0 = {-(G1.x + G2.x + G3.x), -(G1.y + G2.y + G3.y), -(G1.z + G2.z + G3.z)};
    float wG = fmaf(w[3], dot(G3, G3), fmaf(w[2], dot(G2, G2), fmaf(w[1], dot(G1, G1), w[0] * dot(G0, G0))));
    float C = vol - gammaVol;
    float s = (C != 0.0f && wG > 0.0f) ? -__fdividef(C, fmaf(alphaVol, irv, wG)) : 0.0f;
    p[0] = fma3(G0, s * w[0], p[0]);
    p[1] = fma3(G1, s * w[1], p[1]);
    p[2] = fma3(G2, s * w[2], p[2]);
    p[3] = fma3(G3, s * w[3], p[3]);
    return vol - 1.0f;
}

// FAST Neo-Hookean, rest-metric form -- what the throughput (tile) kernel runs.  Same two projections,
// with the rest inverse Q eliminated algebraically (F = P Q, P = current edge matrix):
//   deviatoric : C = ||F||,  dC/dP = F Q^T / C = P B / C  with B = Q Q^T (symmetric, 6 floats per tet)
//                and ||F||^2 = tr(P^T P B) = sum_k P_k . (P B)_k         -> F is never formed
//   hydrostatic: det F = det P det Q,  d(det F)/dP = det Q cof(P)         -> no Q at all
// 170 FP instructions per tet instead of 290 (round-1 ncu + ablation: the tile kernel is bound by the
// FP32 pipe, not by HBM, once the stream is prefetched), and the tet stream shrinks to 7 floats.
// Algebraically identical to nh_solve_fast; roundings differ at the 1e-7 level.
__device__ __forceinline__ float nh_solve_fast_metric(V3 p[4], const float w[4], const float Bm[6] /* B00 B01 B02 B11 B12 B22 */,
                                                      float irv, float detQ, float alphaDev, float alphaVol,
                                                      float gammaVol) {
    V3 P0 = p[1] - p[0], P1 = p[2] - p[0], P2 = p[3] - p[0];
    {
        V3 G1 = fma3(P2, Bm[2], fma3(P1, Bm[1], P0 * Bm[0]));
        V3 G2 = fma3(P2, Bm[4], fma3(P1, Bm[3], P0 * Bm[1]));
        V3 G3 = fma3(P2, Bm[5], fma3(P1, Bm[4], P0 * Bm[2]));
        float rs2 = dot(P0, G1) + dot(P1, G2) + dot(P2, G3);
        V3 G0 = {-(G1.x + G2.x + G3.x), -(G1.y + G2.y + G3.y), -(G1.z + G2.z + G3.z)};
        float wG = fmaf(w[3], dot(G3, G3), fmaf(w[2], dot(G2, G2), fmaf(w[1], dot(G1, G1), w[0] * dot(G0, G0))));
        float den = fmaf(alphaDev * irv, rs2, wG);
        float s = (rs2 > 0.0f && wG > 0.0f) ? -__fdividef(rs2, den) : 0.0f;
        p[0] = fma3(G0, s * w[0], p[0]);
        p[1] = fma3(G1, s * w[1], p[1]);
        p[2] = fma3(G2, s * w[2], p[2]);
        p[3] = fma3(G3, s * w[3], p[3]);
    }
    P0 = p[1] - p[0]; P1 = p[2] - p[0]; P2 = p[3] - p[0];
    V3 c1 = cross(P1, P2), c2 = cross(P2, P0), c3 = cross(P0, P1);
    float vol = dot(P0, c1) * detQ;
    V3 c0 = {-(c1.x + c2.x + c3.x), -(c1.y + c2.y + c3.y), -(c1.z + c2.z + c3.z)};
    float wC = fmaf(w[3], dot(c3, c3), fmaf(w[2], dot(c2, c2), fmaf(w[1], dot(c1, c1), w[0] * dot(c0, c0))));
    float C = vol - gammaVol;
    // true gradients are detQ * c_i:  dlambda = -C / (detQ^2 wC + alpha);  step_i = c_i * (detQ * dlambda * w_i)
    float den = fmaf(detQ * detQ, wC, alphaVol * irv);
    float s = (C != 0.0f && wC > 0.0f) ? -__fdividef(C * detQ, den) : 0.0f;
    p[0] = fma3(c0, s * w[0], p[0]);
    p[1] = fma3(c1, s * w[1], p[1]);
    p[2] = fma3(c2, s * w[2], p[2]);
    p[3] = fma3(c3, s * w[3], p[3]);
    return vol - 1.0f;
}

// Same two projections as nh_solve_fast_metric, restated for the tile kernel's inner loop: returns the four corner
// displacements d[i] = x_i(after both projections) - q[i] directly (the kernel scatters displacements, so the
// positions themselves are never rebuilt), the second stage works on the edge matrix only
// (P'_k = P_k + d_k - d_0), the gradient of corner 0 is kept with flipped sign (nG0 = G1 + G2 + G3, the sign goes
// into its scale) and both step scales are branch-free (reciprocal always formed, result selected):
// ~178 FP instructions, no divergent code.  VOL = false drops the det F - 1 sample.
__device__ __forceinline__ float rcp_approx(float x) {  // MUFU.RCP, no range fix-up (the result is selected away when unusable)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
template <bool VOL>
__device__ __forceinline__ float nh_solve_tile(const V3 q[4], const float w[4], const float Bm[6], float detQ2, float detQ,
                                               float alphaDev6, float alphaVol6, float gammaVol, V3 d[4]) {
    // detQ2 = det Q squared (stored: the invRestVolume slot is redundant, 1 / V = 6 det Q); alphaDev6 / alphaVol6 =
    // 6 compliance / dt^2, so compliance / dt^2 * invRestVolume = alpha6 * detQ.
    V3 P0 = q[1] - q[0], P1 = q[2] - q[0], P2 = q[3] - q[0];
    V3 G1 = fma3(P2, Bm[2], fma3(P1, Bm[1], P0 * Bm[0]));
    V3 G2 = fma3(P2, Bm[4], fma3(P1, Bm[3], P0 * Bm[1]));
    V3 G3 = fma3(P2, Bm[5], fma3(P1, Bm[4], P0 * Bm[2]));
    float rs2 = P0.x * G1.x;
    rs2 = fmaf(P0.y, G1.y, rs2); rs2 = fmaf(P0.z, G1.z, rs2);
    rs2 = fmaf(P1.x, G2.x, rs2); rs2 = fmaf(P1.y, G2.y, rs2); rs2 = fmaf(P1.z, G2.z, rs2);
    rs2 = fmaf(P2.x, G3.x, rs2); rs2 = fmaf(P2.y, G3.y, rs2); rs2 = fmaf(P2.z, G3.z, rs2);
    const V3 nG0 = {G1.x + G2.x + G3.x, G1.y + G2.y + G3.y, G1.z + G2.z + G3.z};
    const float wG = fmaf(w[3], dot(G3, G3), fmaf(w[2], dot(G2, G2), fmaf(w[1], dot(G1, G1), w[0] * dot(nG0, nG0))));
    const float r1 = rs2 * rcp_approx(fmaf(alphaDev6 * detQ, rs2, wG));
    const float s = (rs2 > 0.0f && wG > 0.0f) ? -r1 : 0.0f;
    const float s1 = s * w[1], s2 = s * w[2], s3 = s * w[3];
    const V3 d0 = nG0 * (-(s * w[0]));
    // edges after the deviatoric projection
    P0 = fma3(G1, s1, P0) - d0; P1 = fma3(G2, s2, P1) - d0; P2 = fma3(G3, s3, P2) - d0;
    const V3 c1 = cross(P1, P2), c2 = cross(P2, P0), c3 = cross(P0, P1);
    const float vol = dot(P0, c1) * detQ;
    const V3 nc0 = {c1.x + c2.x + c3.x, c1.y + c2.y + c3.y, c1.z + c2.z + c3.z};
    const float wC = fmaf(w[3], dot(c3, c3), fmaf(w[2], dot(c2, c2), fmaf(w[1], dot(c1, c1), w[0] * dot(nc0, nc0))));
    const float C = vol - gammaVol;
    // true gradients are detQ * c_i: w = detQ^2 wC (the reference's `w == 0 -> return` tests exactly this, so an all-zero
    // record -- an unused slot or a zero-volume rest tet -- is a no-op);  dlambda = -C / (w + alpha);  step_i = c_i * (detQ * dlambda * w_i)
    const float wt = detQ2 * wC;
    const float r2 = (C * detQ) * rcp_approx(fmaf(alphaVol6, detQ, wt));
    const float t = (C != 0.0f && wt > 0.0f) ? -r2 : 0.0f;
    d[0] = fma3(nc0, -(t * w[0]), d0);
    d[1] = fma3(c1, t * w[1], G1 * s1);
    d[2] = fma3(c2, t * w[2], G2 * s2);
    d[3] = fma3(c3, t * w[3], G3 * s3);
    return VOL ? vol - 1.0f : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// Polar-decomposition shape matching (src/SoftbodyGPU.js:80-262), f32.
// EXACT: every op separately rounded (TU compiled -fmad=false), IEEE div/sqrt, sin via double.
// FAST : FMA contraction allowed, sincosf, rsqrtf.
// ---------------------------------------------------------------------------------------------
struct Q4 { float x, y, z, w; };

__device__ __forceinline__ V3 pl_add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 pl_sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 pl_mul(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float pl_dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 pl_cross(V3 a, V3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
template <bool EXACT>
__device__ __forceinline__ Q4 pl_normalize(Q4 q) {
    float d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    if (EXACT) {
        float len = sqrtf(d);
        return {q.x / len, q.y / len, q.z / len, q.w / len};
    } else {
        float r = rsqrtf(d);
        return {q.x * r, q.y * r, q.z * r, q.w * r};
    }
}
__device__ __forceinline__ V3 pl_rotate(V3 p, Q4 q) {  // Rotate(), :111-113
    V3 u = {q.x, q.y, q.z};
    V3 t = pl_add(pl_cross(u, p), pl_mul(p, q.w));
    return pl_add(p, pl_mul(pl_cross(u, t), 2.0f));
}
__device__ __forceinline__ Q4 pl_qmul(Q4 a, Q4 b) {  // quat_mult(), :114-121
    Q4 r;
    r.x = (a.w * b.x) + (a.x * b.w) + (a.y * b.z) - (a.z * b.y);
    r.y = (a.w * b.y) - (a.x * b.z) + (a.y * b.w) + (a.z * b.x);
    r.z = (a.w * b.z) + (a.x * b.y) - (a.y * b.x) + (a.z * b.w);
    r.w = (a.w * b.w) - (a.x * b.x) - (a.y * b.y) - (a.z * b.z);
    return r;
}
template <bool EXACT>
__device__ __forceinline__ Q4 pl_extract_rotation(const V3 A[3], Q4 q) {  // extractRotation(), :122-139
    const V3 ex = {1.0f, 0.0f, 0.0f}, ey = {0.0f, 1.0f, 0.0f}, ez = {0.0f, 0.0f, 1.0f};
    for (int iter = 0; iter < 9; iter++) {
        V3 X = pl_rotate(ex, q), Y = pl_rotate(ey, q), Z = pl_rotate(ez, q);
        V3 num = pl_add(pl_add(pl_cross(X, A[0]), pl_cross(Y, A[1])), pl_cross(Z, A[2]));
        float den = pl_dot(X, A[0]) + pl_dot(Y, A[1]) + pl_dot(Z, A[2]) + 0.000000001f;
        V3 omega = pl_mul(num, 1.0f / fabsf(den));
        float w = sqrtf(pl_dot(omega, omega));
        if (w < 0.000000001f) break;
        float half = w * 0.5f;
        float s, c;
        if (EXACT) {
            s = (float)sin((double)half);
            c = (float)sin((double)(half + 1.57f));  // the shader's cosine, :108
        } else {
            s = __sinf(half);
            c = __sinf(half + 1.57f);
        }
        Q4 dq = {omega.x / w * s, omega.y / w * s, omega.z / w * s, c};
        q = pl_qmul(dq, q);
    }
    return q;
}
// FAST flavour of extractRotation for the tiled kernel: the three rotated basis vectors are the columns of R(q), formed
// directly from the quaternion's products (same polynomial as three Rotate() calls, a quarter of the instructions), one
// reciprocal per iteration, and the loop leaves as soon as the increment is at the float32 noise floor: the reference's own
// exit `w < 1e-9` (:131) is unreachable in float32 once |A| ~ edge^2 (the residual rotation of a converged iterate is
// ~1e-7), so its last five or six of nine iterations only stir rounding noise.  kPolarEps is the measured trade-off:
// tests/test_parity_gpu.py keeps the result within the polar tolerance of the oracle (which always runs the shader's loop).
// The floor scales with the mesh: A = sum (cur - cc)(last - lc)^T is built from positions that carry half an ulp of |x|
// each, so omega = num / den cannot be known better than ~ 2^-24 |x| / L (L^2 = sum |last - lc|^2 ~ den): on the Dragon
// (|x| ~ 1, edges ~ 0.1) that IS ~1e-6, on the 1 cm beam it is ~1e-5, where a fixed 1e-6 never triggers and all nine
// iterations run.  noise2 = (k 2^-24)^2 |cc|^2 (k = 4; TETSIM_POLAR_NOISE_K overrides, 0 = fixed floor only); the loop leaves
// when |omega|^2 den < noise2 or |omega| < kPolarEps.  Measured (profiles/r2_polar_exit_rule.txt): no visible change of the
// error against the nine-iteration BITEXACT path for k = 0..8; kernel 0.416 -> 0.300 ms in free fall, 0.50 -> 0.42 ms on the
// crumpling beam (the reference's iteration converges only linearly there, so most warps still need all nine).
constexpr float kPolarEps = 1.0e-6f;
__device__ __forceinline__ Q4 pl_extract_rotation_fast(const V3 A[3], Q4 q, float noise2) {
    for (int iter = 0; iter < 9; iter++) {
        const float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z, xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
        const float wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
        const V3 X = {1.0f - 2.0f * (yy + zz), 2.0f * (xy + wz), 2.0f * (xz - wy)};
        const V3 Y = {2.0f * (xy - wz), 1.0f - 2.0f * (xx + zz), 2.0f * (yz + wx)};
        const V3 Z = {2.0f * (xz + wy), 2.0f * (yz - wx), 1.0f - 2.0f * (xx + yy)};
        const V3 num = pl_add(pl_add(pl_cross(X, A[0]), pl_cross(Y, A[1])), pl_cross(Z, A[2]));
        const float den = pl_dot(X, A[0]) + pl_dot(Y, A[1]) + pl_dot(Z, A[2]) + 0.000000001f;
        const V3 omega = pl_mul(num, __fdividef(1.0f, fabsf(den)));
        const float w2 = pl_dot(omega, omega);
        if (w2 < kPolarEps * kPolarEps || w2 * fabsf(den) < noise2) break;
        const float rw = rsqrtf(w2), w = w2 * rw, half = w * 0.5f;
        const float s = __sinf(half) * rw, c = __sinf(half + 1.57f);   // the shader's cosine, :108
        const Q4 dq = {omega.x * s, omega.y * s, omega.z * s, c};
        q = pl_qmul(dq, q);
    }
    return q;
}
// K3 + K4 for one tet: cur[4] current corner positions, last[4] in/out goal corners, quat in/out.
template <bool EXACT, bool TILED = false>
__device__ __forceinline__ void polar_solve(const V3 cur[4], V3 last[4], Q4 &quat, float noiseK2 = 0.0f) {
    V3 cc = pl_mul(pl_add(pl_add(pl_add(cur[0], cur[1]), cur[2]), cur[3]), 0.25f);
    V3 lc = pl_mul(pl_add(pl_add(pl_add(last[0], last[1]), last[2]), last[3]), 0.25f);
    V3 A[3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
    for (int k = 0; k < 4; k++) {  // TransposeMult(lastRest, current), :90-105
        V3 c = pl_sub(cur[k], cc), l = pl_sub(last[k], lc);
        A[0] = pl_add(A[0], pl_mul(c, l.x));
        A[1] = pl_add(A[1], pl_mul(c, l.y));
        A[2] = pl_add(A[2], pl_mul(c, l.z));
    }
    Q4 ident = {0.0f, 0.0f, 0.0f, 1.0f};
    Q4 rot = (!EXACT && TILED) ? pl_extract_rotation_fast(A, ident, noiseK2 * pl_dot(cc, cc)) : pl_extract_rotation<EXACT>(A, ident);
    Q4 qOld = quat;
    Q4 qNew = pl_normalize<EXACT>(pl_qmul(rot, qOld));  // :181
    quat = qNew;
    if (!EXACT && TILED) {
        // throughput form: rel = normalize(qNew (x) conj(qOld)) IS the extracted rotation when qOld is a unit quaternion
        // ((rot qOld) conj(qOld) = rot), so the second product and two of the three normalisations drop out, and the four
        // goal corners are turned by the rotation MATRIX of rel (25 + 4 x 9 instructions instead of 4 x 30).  Differs from
        // the shader's sequence at the 1e-7 level; tests/test_parity_gpu.py holds it to the polar tolerance.
        const Q4 r = pl_normalize<false>(rot);
        const float xx = r.x * r.x, yy = r.y * r.y, zz = r.z * r.z, xy = r.x * r.y, xz = r.x * r.z, yz = r.y * r.z;
        const float wx = r.w * r.x, wy = r.w * r.y, wz = r.w * r.z;
        const V3 X = {1.0f - 2.0f * (yy + zz), 2.0f * (xy + wz), 2.0f * (xz - wy)};
        const V3 Y = {2.0f * (xy - wz), 1.0f - 2.0f * (xx + zz), 2.0f * (yz + wx)};
        const V3 Z = {2.0f * (xz + wy), 2.0f * (yz - wx), 1.0f - 2.0f * (xx + yy)};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const V3 l = pl_sub(last[k], lc);
            last[k] = {fmaf(Z.x, l.z, fmaf(Y.x, l.y, fmaf(X.x, l.x, cc.x))), fmaf(Z.y, l.z, fmaf(Y.y, l.y, fmaf(X.y, l.x, cc.y))),
                       fmaf(Z.z, l.z, fmaf(Y.z, l.y, fmaf(X.z, l.x, cc.z)))};
        }
        return;
    }
    Q4 conj = {-qOld.x, -qOld.y, -qOld.z, qOld.w};
    Q4 rel = pl_normalize<EXACT>(pl_qmul(qNew, pl_normalize<EXACT>(conj)));  // :207,:239
#pragma unroll
    for (int k = 0; k < 4; k++) last[k] = pl_add(pl_rotate(pl_sub(last[k], lc), rel), cc);  // :253-256
}

}  // namespace tsim
