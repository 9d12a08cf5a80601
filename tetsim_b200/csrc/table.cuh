// table.cuh -- builds the KernelTable of one arithmetic flavour from the templates in kernels.cuh.
#pragma once
#include "kernels.cuh"
#include "launch.h"

namespace tsim {

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

template <bool E>
struct Launchers {
    static constexpr int TB = 256;
    static void init_tets(cudaStream_t s, int M, const float4 *x4, const int4 *ids, double density, float *Q9,
                          float *irv, double *pm) {
        if (M > 0) k_init_tets<E><<<cdiv(M, TB), TB, 0, s>>>(M, x4, ids, density, Q9, irv, pm);
    }
    static void init_mass(cudaStream_t s, int N, const int *cStart, const int *cEnt, const double *pm, float4 *x4) {
        if (N > 0) k_init_mass<E><<<cdiv(N, TB), TB, 0, s>>>(N, cStart, cEnt, pm, x4);
    }
    static void predict(cudaStream_t s, int N, float4 *x4, float4 *prev4, float4 *vel4, const SubstepParams *sp) {
        if (N > 0) k_predict<E><<<cdiv(N, TB), TB, 0, s>>>(N, x4, prev4, vel4, sp);
    }
    static void post(cudaStream_t s, int N, float4 *x4, const float4 *prev4, float4 *vel4, const SubstepParams *sp,
                     const int *vertId) {
        if (N > 0) k_post<E><<<cdiv(N, TB), TB, 0, s>>>(N, x4, prev4, vel4, sp, vertId);
    }
    static void gs_level(cudaStream_t s, int begin, int end, float4 *x4, const int4 *I, const float4 *A,
                         const float4 *B, const float4 *C, const int *order, double *volTerm,
                         const SubstepParams *sp) {
        int n = end - begin;
        if (n <= 0) return;
        int tb = n < 128 ? 32 * cdiv(n, 32) : 128;
        k_gs_level<E><<<cdiv(n, tb), tb, 0, s>>>(begin, end, x4, I, A, B, C, order, volTerm, sp);
    }
    static void gs_body(cudaStream_t s, int numBodies, int threads, size_t smemBytes, const BodyDesc *bodies,
                        const int *levelStart, float4 *x4, float4 *prev4, float4 *vel4, const int4 *I,
                        const float4 *A, const float4 *B, const float4 *C, const int *order, double *volTerm,
                        const SubstepParams *sp, const int *vertId) {
        if (numBodies <= 0) return;
        static size_t configured[64] = {0};  // per device: the opt-in is a per-device function attribute
        int dev = 0;
        cudaGetDevice(&dev);
        size_t &c = configured[dev < 0 || dev >= 64 ? 0 : dev];
        if (smemBytes > 48 * 1024 && smemBytes > c) {
            cudaFuncSetAttribute(k_gs_body<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
            c = smemBytes;
        }
        k_gs_body<E><<<numBodies, threads, smemBytes, s>>>(bodies, levelStart, x4, prev4, vel4, I, A, B, C, order,
                                                          volTerm, sp, vertId);
    }
    static int gs_body_max_smem() { return 200 * 1024; }
    static void jacobi_tet(cudaStream_t s, int M, const float4 *x4, const int4 *I, const float4 *A, const float4 *B,
                           const float4 *C, float4 *dx, double *volTerm, const SubstepParams *sp) {
        if (M > 0) k_jacobi_tet<E><<<cdiv(M, 128), 128, 0, s>>>(M, x4, I, A, B, C, dx, volTerm, sp);
    }
    static void jacobi_gather(cudaStream_t s, int N, float4 *x4, const int *cStart, const int *cEnt,
                              const float4 *dx) {
        if (N > 0) k_jacobi_gather<E><<<cdiv(N, TB), TB, 0, s>>>(N, x4, cStart, cEnt, dx);
    }
    static void polar_integrate(cudaStream_t s, int N, float4 *x4, float4 *prev4, const float4 *vel4,
                                const SubstepParams *sp) {
        if (N > 0) k_polar_integrate<E><<<cdiv(N, TB), TB, 0, s>>>(N, x4, prev4, vel4, sp);
    }
    static void polar_tet(cudaStream_t s, int M, const float4 *x4, const int4 *I, float4 *rest, float4 *quat) {
        if (M > 0) k_polar_tet<E><<<cdiv(M, 128), 128, 0, s>>>(M, x4, I, rest, quat);
    }
    static void polar_vertex(cudaStream_t s, int N, float4 *x4, const float4 *prev4, float4 *vel4, const int *tStart,
                             const int *tEnt, const float4 *rest, const SubstepParams *sp) {
        if (N > 0) k_polar_vertex<E><<<cdiv(N, TB), TB, 0, s>>>(N, x4, prev4, vel4, tStart, tEnt, rest, sp);
    }
    static void skin(cudaStream_t s, int nVis, const float4 *vis, const int4 *ids, const float4 *x4, float *out) {
        if (nVis > 0) k_skin<E><<<cdiv(nVis, TB), TB, 0, s>>>(nVis, vis, ids, x4, out);
    }
    static void skin_polar(cudaStream_t s, int nVis, const float4 *vis, const int4 *ids, const float4 *x4, const float4 *quat,
                           const unsigned char *tileTets, const int *tetRecord, int T, const float *restNrm, float *outPos, float *outNrm) {
        if (nVis > 0) k_skin_polar<E><<<cdiv(nVis, TB), TB, 0, s>>>(nVis, vis, ids, x4, quat, tileTets, tetRecord, T, restNrm, outPos, outNrm);
    }
    static void normals(cudaStream_t s, int nVis, const float *pos, const int *tri, const int *vtStart,
                        const int *vtEnt, float *nrm) {
        if (nVis > 0) k_normals<E><<<cdiv(nVis, TB), TB, 0, s>>>(nVis, pos, tri, vtStart, vtEnt, nrm);
    }
    static const KernelTable *table() {
        static const KernelTable t = {init_tets,  init_mass,     predict,         post,      gs_level,
                                      gs_body,    gs_body_max_smem, jacobi_tet,   jacobi_gather,
                                      polar_integrate, polar_tet, polar_vertex,   skin,      normals, skin_polar};
        return &t;
    }
};

}  // namespace tsim
