// kernels_fast.cu -- FAST_F32 flavour (FMA contraction on) + the clustered Jacobi throughput path
// + small utility kernels.  sm_100a only.
#include "table.cuh"

namespace tsim {

const KernelTable *fast_kernels() { return Launchers<false>::table(); }

// =================================================================================================
// Clustered Jacobi Neo-Hookean -- the throughput kernel (BASELINE config 4).
//
// One CTA = one tile of T consecutive tets of the (Morton-sorted) tet stream, one tet per thread.
//   1. the tile's vertex list (unique vertices touched by its T tets, ~0.35 T of them) is gathered
//      from HBM/L2 into shared memory as float4 (x,y,z,invMass): one LDG.128 per tile vertex;
//   2. each thread streams its tet record with three coalesced LDG.128 (A,B,C planes = 48 B/tet:
//      Q 36 B, invRestVolume 4 B, four 16-bit TILE-LOCAL vertex slots 8 B), gathers its 4 corners
//      from shared memory (LDS.128), runs both Neo-Hookean projections in registers and parks the
//      12 floats of corner dx in shared memory (3 x STS.128, 48-byte records -> conflict-free);
//   3. the tile's vertices then sum their corners' dx straight out of shared memory.  Corner lists
//      are stored as jagged diagonals (tile vertices sorted by descending tile valence, i-th corner
//      of all vertices contiguous) so the 16-bit index loads are coalesced and a warp's lanes run
//      the same trip count.  No shared-memory float atomics (sm_100a has no native one: ATOMS.CAST
//      spin loops), no global atomics in the default flush: each tile writes its partial sums to
//      its own slice of `part` with plain coalesced STG.128 and the vertex kernel adds the few
//      (~1.6) slices of a vertex in a fixed order -> bit-reproducible run to run.
//      (deterministic = 0 flushes with one REDG.E.ADD.F32x4 per tile vertex instead.)
// Algorithmic traffic per launch: 56 B/tet + 32 B/vertex; actual DRAM traffic is
// lower on the tet stream (48 B) and higher on the vertex side (tile overlap).
// =================================================================================================
template <int T>
__global__ void __launch_bounds__(T) k_jacobi_cluster(int firstCluster, ClusterArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sdx = reinterpret_cast<float *>(smem_raw);                          // [T * 12]
    float4 *sx = reinterpret_cast<float4 *>(smem_raw + (size_t)T * 48);        // [maxTileVerts]
    uint16_t *scol = reinterpret_cast<uint16_t *>(sx + a.maxTileVerts);        // [colStride]

    const int c = firstCluster + blockIdx.x;
    const int t = threadIdx.x;
    const int v0 = a.clVertStart[c];
    const int nl = a.clVertStart[c + 1] - v0;
    const size_t rec = (size_t)c * T + t;

    // tet record: issue the streaming loads first so they overlap the tile gather
    const float4 A = ldg4(a.A + rec), B = ldg4(a.B + rec), C = ldg4(a.C + rec);

    for (int j = t; j < nl; j += T) sx[j] = a.x4[a.clVerts[v0 + j]];
    if (t < a.colStride) scol[t] = a.colOff[(size_t)c * a.colStride + t];
    __syncthreads();

    const unsigned s01 = __float_as_uint(C.z), s23 = __float_as_uint(C.w);
    const float4 q0 = sx[s01 & 0xffffu], q1 = sx[s01 >> 16], q2 = sx[s23 & 0xffffu], q3 = sx[s23 >> 16];
    V3 p[4] = {{q0.x, q0.y, q0.z}, {q1.x, q1.y, q1.z}, {q2.x, q2.y, q2.z}, {q3.x, q3.y, q3.z}};
    const float w[4] = {q0.w, q1.w, q2.w, q3.w};
    const float Q[9] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x};
    const SubstepParams *sp = a.sp;
    float vm1 = nh_solve_fast(p, w, Q, C.y, sp->alphaDev, sp->alphaVol, sp->gammaVol);

    float4 *d4 = reinterpret_cast<float4 *>(sdx + t * 12);
    d4[0] = make_float4(p[0].x - q0.x, p[0].y - q0.y, p[0].z - q0.z, p[1].x - q1.x);
    d4[1] = make_float4(p[1].y - q1.y, p[1].z - q1.z, p[2].x - q2.x, p[2].y - q2.y);
    d4[2] = make_float4(p[2].z - q2.z, p[3].x - q3.x, p[3].y - q3.y, p[3].z - q3.z);

    if (a.volAcc) {  // volError (src/Softbody.js:163): warp-reduce, one double atomic per warp
        float s = (C.y != 0.0f) ? vm1 : 0.0f;  // padding records carry invRestVolume = 0
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((t & 31) == 0) atomicAdd(a.volAcc, (double)s);
    }
    __syncthreads();

    const uint16_t *jd = a.jds + (size_t)c * 4 * T;
    for (int j = t; j < nl; j += T) {
        const int val = a.clVal[v0 + j];
        float ax = 0.0f, ay = 0.0f, az = 0.0f;
#pragma unroll 4
        for (int i = 0; i < val; i++) {
            const unsigned e = jd[scol[i] + j];
            const float *d = sdx + 3 * e;
            ax += d[0]; ay += d[1]; az += d[2];
        }
        if (a.acc) atomicAdd(a.acc + a.clVerts[v0 + j], make_float4(ax, ay, az, 0.0f));
        else a.part[v0 + j] = make_float4(ax, ay, az, 0.0f);
    }
}

size_t jacobi_cluster_smem(int clusterSize, int maxTileVerts, int colStride) {
    return (size_t)clusterSize * 48 + (size_t)maxTileVerts * 16 + (size_t)((colStride * 2 + 15) / 16) * 16;
}

template <int T>
static void launch_cluster_T(cudaStream_t s, int first, int n, const ClusterArgs &a) {
    size_t smem = jacobi_cluster_smem(T, a.maxTileVerts, a.colStride);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaFuncSetAttribute(k_jacobi_cluster<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    k_jacobi_cluster<T><<<n, T, smem, s>>>(first, a);
}

void launch_jacobi_cluster(cudaStream_t s, int clusterSize, int firstCluster, int numClusters, const ClusterArgs &a) {
    if (numClusters <= 0) return;
    switch (clusterSize) {
        case 128: launch_cluster_T<128>(s, firstCluster, numClusters, a); break;
        case 256: launch_cluster_T<256>(s, firstCluster, numClusters, a); break;
        case 512: launch_cluster_T<512>(s, firstCluster, numClusters, a); break;
        default: break;
    }
}

// Vertex side of the clustered Jacobi: x += (sum of the vertex's tile partials) / valence, optionally
// fused with post (simulate() :213-239) and with the NEXT substep's predict (:198-202) so a substep
// inside tetsim_step costs exactly two launches.
template <int MODE>
__global__ void k_jacobi_apply(int begin, int end, ApplyArgs a) {
    int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    if (a.bsum && i >= a.boundaryBegin) {
        float4 s = a.bsum[i - a.boundaryBegin];
        sx = s.x; sy = s.y; sz = s.z;
    } else if (a.acc) {
        float4 s = a.acc[i];
        a.acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        sx = s.x; sy = s.y; sz = s.z;
    } else {
        for (int j = a.vpStart[i]; j < a.vpStart[i + 1]; j++) {
            float4 s = ldg4(a.part + a.vpSlot[j]);
            sx += s.x; sy += s.y; sz += s.z;
        }
    }
    float4 x = a.x4[i];
    const float inv = a.invVal[i];
    x.x = fmaf(sx, inv, x.x); x.y = fmaf(sy, inv, x.y); x.z = fmaf(sz, inv, x.z);
    if (MODE >= 1) {
        const SubstepParams *sp = a.sp;
        float4 p = a.prev4[i], v;
        v.w = 0.0f;
        post_vertex<false>(a.vertId ? a.vertId[i] : i, x, p, v, sp);
        if (MODE == 2) {
            a.prev4[i] = x;
            v.y += sp->gDt;
            const float dt = sp->dtF;
            x.x = fmaf(v.x, dt, x.x); x.y = fmaf(v.y, dt, x.y); x.z = fmaf(v.z, dt, x.z);
        }
        a.vel4[i] = v;
    }
    a.x4[i] = x;
}

void launch_jacobi_apply(cudaStream_t s, int begin, int end, int mode, const ApplyArgs &a) {
    int n = end - begin;
    if (n <= 0) return;
    const int TB = 256;
    if (mode == 0) k_jacobi_apply<0><<<cdiv(n, TB), TB, 0, s>>>(begin, end, a);
    else if (mode == 1) k_jacobi_apply<1><<<cdiv(n, TB), TB, 0, s>>>(begin, end, a);
    else k_jacobi_apply<2><<<cdiv(n, TB), TB, 0, s>>>(begin, end, a);
}

__global__ void k_boundary_pack(int boundaryBegin, int nB, const int *__restrict__ vpStart,
                                const int *__restrict__ vpSlot, const float4 *__restrict__ part,
                                float4 *__restrict__ bsum) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nB) return;
    int i = boundaryBegin + b;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int j = vpStart[i]; j < vpStart[i + 1]; j++) {
        float4 s = ldg4(part + vpSlot[j]);
        sx += s.x; sy += s.y; sz += s.z;
    }
    bsum[b] = make_float4(sx, sy, sz, 0.0f);
}
void launch_boundary_pack(cudaStream_t s, int boundaryBegin, int nB, const int *vpStart, const int *vpSlot,
                          const float4 *part, float4 *bsum) {
    if (nB > 0) k_boundary_pack<<<cdiv(nB, 256), 256, 0, s>>>(boundaryBegin, nB, vpStart, vpSlot, part, bsum);
}

// =================================================================================================
// Utility kernels
// =================================================================================================
__global__ void k_pack3(int N, const float4 *__restrict__ src, const int *__restrict__ perm, float *__restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float4 v = src[i];
    size_t o = 3 * (size_t)(perm ? perm[i] : i);
    dst[o] = v.x; dst[o + 1] = v.y; dst[o + 2] = v.z;
}
void launch_pack3(cudaStream_t s, int N, const float4 *src, const int *perm, float *dst3) {
    if (N > 0) k_pack3<<<cdiv(N, 256), 256, 0, s>>>(N, src, perm, dst3);
}
__global__ void k_unpack3(int N, const float *__restrict__ src, const int *__restrict__ perm, float4 *__restrict__ dst,
                          int keepW) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    size_t o = 3 * (size_t)(perm ? perm[i] : i);
    float w = keepW ? dst[i].w : 0.0f;
    dst[i] = make_float4(src[o], src[o + 1], src[o + 2], w);
}
void launch_unpack3(cudaStream_t s, int N, const float *src3, const int *perm, float4 *dst, int keepW) {
    if (N > 0) k_unpack3<<<cdiv(N, 256), 256, 0, s>>>(N, src3, perm, dst, keepW);
}

// volError = (sequential f64 sum of the per-tet terms in tet order) / M  (src/Softbody.js:206-209)
__global__ void k_sum_sequential(int M, const double *__restrict__ terms, double *__restrict__ out) {
    double acc = 0.0;
    for (int i = 0; i < M; i++) acc = __dadd_rn(acc, terms[i]);
    *out = __ddiv_rn(acc, (double)M);
}
void launch_sum_sequential(cudaStream_t s, int M, const double *terms, double *out) {
    k_sum_sequential<<<1, 1, 0, s>>>(M, terms, out);
}

// startGrab (src/Softbody.js:279-291): first strict minimum of the f64 squared distance.
// Pass 1: atomicMin over the (non-negative) double's bit pattern; pass 2: smallest caller-side
// vertex id among the vertices that attain it.
__device__ __forceinline__ double grab_d2(float4 x, const double *p) {
    double a0 = __dsub_rn(p[0], (double)x.x), a1 = __dsub_rn(p[1], (double)x.y), a2 = __dsub_rn(p[2], (double)x.z);
    return __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(a2, a2));
}
__global__ void k_nearest_pass1(int N, const float4 *__restrict__ x4, const double *__restrict__ p,
                                unsigned long long *__restrict__ best) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double d2 = grab_d2(x4[i], p);
    if (d2 < 1.7976931348623157e308) atomicMin(best, (unsigned long long)__double_as_longlong(d2));
}
__global__ void k_nearest_pass2(int N, const float4 *__restrict__ x4, const int *__restrict__ vertId,
                                const double *__restrict__ p, const unsigned long long *__restrict__ best,
                                int *__restrict__ outId) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double d2 = grab_d2(x4[i], p);
    if (d2 < 1.7976931348623157e308 && (unsigned long long)__double_as_longlong(d2) == *best)
        atomicMin(outId, vertId ? vertId[i] : i);
}
void launch_nearest_vertex(cudaStream_t s, int N, const float4 *x4, const int *vertId, const double *p3, int *outId,
                           unsigned long long *scratch) {
    cudaMemsetAsync(scratch, 0xff, sizeof(unsigned long long), s);
    cudaMemsetAsync(outId, 0x7f, sizeof(int), s);  // 0x7f7f7f7f: larger than any vertex id
    if (N <= 0) return;
    k_nearest_pass1<<<cdiv(N, 256), 256, 0, s>>>(N, x4, p3, scratch);
    k_nearest_pass2<<<cdiv(N, 256), 256, 0, s>>>(N, x4, vertId, p3, scratch, outId);
}

// SoftBodyGPU.initPhysics (src/SoftbodyGPU.js:535-551): goal corners start at the rest positions,
// quaternions at identity; .w carries the tet volume V = 1 / invRestVolume the shader forms at :220.
__global__ void k_fill_rest(int M, const float4 *__restrict__ x4, const int4 *__restrict__ I,
                            const float *__restrict__ irv, float4 *__restrict__ rest, float4 *__restrict__ quat) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M) return;
    int4 id = I[e];
    float V = __fdiv_rn(1.0f, irv[e]);
    float4 p;
    p = x4[id.x]; rest[4 * (size_t)e + 0] = make_float4(p.x, p.y, p.z, V);
    p = x4[id.y]; rest[4 * (size_t)e + 1] = make_float4(p.x, p.y, p.z, V);
    p = x4[id.z]; rest[4 * (size_t)e + 2] = make_float4(p.x, p.y, p.z, V);
    p = x4[id.w]; rest[4 * (size_t)e + 3] = make_float4(p.x, p.y, p.z, V);
    quat[e] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
}
void launch_fill_rest(cudaStream_t s, int M, const float4 *x4, const int4 *I, const float *irv, float4 *rest,
                      float4 *quat) {
    if (M > 0) k_fill_rest<<<cdiv(M, 256), 256, 0, s>>>(M, x4, I, irv, rest, quat);
}

// Tet stream in solver order: A/B/C planes from the tet-order Q9/invRestVolume arrays.
// order[i] < 0 marks a padding record (all zero: F = 0 -> no correction).
__global__ void k_build_stream(int count, const int *__restrict__ order, const float *__restrict__ Q9,
                               const float *__restrict__ irv, float4 *__restrict__ A, float4 *__restrict__ B,
                               float4 *__restrict__ C) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int e = order ? order[i] : i;
    float4 c = C[i];  // .z/.w may already hold packed tile slots
    if (e < 0) {
        A[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        B[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        C[i] = make_float4(0.f, 0.f, c.z, c.w);
        return;
    }
    const float *q = Q9 + 9 * (size_t)e;
    A[i] = make_float4(q[0], q[1], q[2], q[3]);
    B[i] = make_float4(q[4], q[5], q[6], q[7]);
    C[i] = make_float4(q[8], irv[e], c.z, c.w);
}
void launch_build_stream(cudaStream_t s, int count, const int *order, const float *Q9, const float *irv, float4 *A,
                         float4 *B, float4 *C) {
    if (count > 0) k_build_stream<<<cdiv(count, 256), 256, 0, s>>>(count, order, Q9, irv, A, B, C);
}

}  // namespace tsim
