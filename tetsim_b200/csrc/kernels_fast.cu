// kernels_fast.cu -- FAST_F32 flavour (FMA contraction on) + the clustered Jacobi throughput path
// + small utility kernels.  sm_100a only.
#include <mutex>

#include "table.cuh"

namespace tsim {

const KernelTable *fast_kernels() { return Launchers<false>::table(); }

// =================================================================================================
// Clustered Jacobi Neo-Hookean -- the throughput kernel (BASELINE config 4).
//
// Persistent CTAs (grid = SMs x resident CTAs), each walking tiles c = blockIdx.x, + gridDim.x, ...
// One tile = up to T consecutive tets of the Hilbert-sorted tet stream (closed early if it would touch
// more than a capped number of vertices).  Default shape (k_jacobi_tilesN): TWO tets per thread -- T = 512 with
// 256 threads (64 registers, 4 CTAs per SM) or T = 256 with 128 threads -- so every thread carries two independent
// dependency chains, the per-tile overhead (two barriers, prefetch issue, loop control) is paid once per two tets
// and the corner sums below keep all warps busy.  HBM layout per tile:
//   tet block  T*48 B: planes A[T] float4 (B00,B01,B02,B11), B[T] float4 (B12,B22,invRestVolume,detQ),
//              C[T] uint4 (4 vertex slots, 4 scatter destinations as 16-bit byte offsets); B = Q Q^T is
//              the rest metric (see nh_solve_tile) -- 28 B of physics + 16 B of indices per tet;
//   meta block (variable size): tile vertex ids, tile valences, jagged-diagonal offsets.
// Data movement, all asynchronous and issued ahead of use:
//   * tet block: one cp.async.bulk.prefetch.L2 per tile two tiles ahead, then three coalesced
//     LDG.128 per tet straight into registers (staging the stream in shared memory would spend
//     two passes of the 128 B/clk shared-memory pipe, which gather + scatter already load to 60 %);
//   * meta block: ONE cp.async.bulk (TMA, UBLKCP) on an mbarrier, S tiles ahead;
//   * the tile's vertex records float4(x,y,z,invMass): indexed gather with cp.async 16 B (LDGSTS),
//     L2 -> shared memory without register staging, S-1 tiles ahead, issued by the upper half of the CTA.
// Per tile:  gather 4 corners (LDS.128) -> both Neo-Hookean projections in registers, corner displacements formed
//   directly (nh_solve_tile) -> each corner's dx is stored (STS.128) at its precomputed slot of a jagged-diagonal
//   buffer (entry (i, j) = i-th corner of tile vertex j; vertices sorted by descending tile valence)
//   -> barrier -> thread j sums entries (0..val_j, j): consecutive lanes read consecutive 16-B
//   entries (conflict-free LDS.128, no index loads), in ascending (tet, corner) order
//   -> one coalesced STG.128 of the tile's partial sum per tile vertex
//   (-> on a multi-GPU handle with the fused peer exchange: the partial of a rank-shared vertex is also stored into
//   the sharers' receive buffers over NVLink, see PeerArgs in launch.h).
// No shared-memory float atomics (sm_100a has none natively: ATOMS.CAST spin loops) and, in the
// default flush, no global atomics either: the vertex kernel adds the ~2.4 tile partials of a vertex in a
// fixed order -> results are bit-reproducible run to run.  (deterministic = 0 flushes with one
// REDG.E.ADD.F32x4 per tile vertex instead.)
// Algorithmic traffic per launch: 56 B/tet + 32 B/vertex; the kernel's own
// DRAM traffic is lower on the tet stream (48 B) and adds ~3 B/tet of metadata and the partial sums.
// Round-1 measurements: 0.141 ms per 10M-tet launch = 0.68 of the measured HBM peak; the
// binding limits are FP32 issue + shared-memory wavefronts (64 % issue-active), not DRAM (3.65 TB/s).
// =================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_u32(b)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Peer-memory exchange, sender side for ONE active boundary vertex b (layout and protocol: launch.h, PeerArgs): add this
// rank's tile partials in their fixed order, keep the sum, store it into every sharer's receive buffer; the caller
// that completes the rank's last push publishes the epoch.
__device__ __forceinline__ void peer_push_vertex(const PeerArgs &a, int b, unsigned e, bool coherentPartials) {
    const int i = a.boundaryBegin + b;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = a.vpStart[i]; j < a.vpStart[i + 1]; j++) {
        const float4 s = coherentPartials ? __ldcg(a.part + a.vpSlot[j]) : __ldg(a.part + a.vpSlot[j]);
        sum.x += s.x; sum.y += s.y; sum.z += s.z;
    }
    a.bsum[b] = sum;
    for (int j = a.pxStart[b]; j < a.pxStart[b + 1]; j++) {
        const int q = a.pxPeer[j];
        float4 *dst = reinterpret_cast<float4 *>(a.peerBase[q] + kPeerRecvOff) + (size_t)(e & 1u) * a.remoteTotal[q] + a.pxEntry[j];
        *dst = sum;  // NVLink store into the sharer's receive buffer
    }
}
__device__ __forceinline__ void peer_publish(const PeerArgs &a, unsigned *ctl, unsigned e) {
    __threadfence_system();  // every push of this rank is ordered before the flags
    for (int q = 0; q < a.numPeers; q++) st_release_sys(reinterpret_cast<unsigned *>(a.peerBase[q]) + a.remoteSlot[q], e);
    ctl[1] = 0u;
    *reinterpret_cast<volatile unsigned *>(ctl) = e;
}
// Receiver side: wait until every sharer has published epoch e (a block calls this with all its threads).
__device__ __forceinline__ void peer_wait_block(const PeerArgs &a, unsigned *ctl, unsigned e) {
    // a wait that timed out once is never repeated (the error is sticky and reported by the host): no pile-up of timeouts
    if ((int)threadIdx.x < a.numPeers && *reinterpret_cast<volatile unsigned *>(ctl + 2) == 0u) {
        const unsigned *flag = reinterpret_cast<const unsigned *>(a.self) + threadIdx.x;
        const unsigned long long t0 = global_timer_ns();
        while ((int)(ld_acquire_sys(flag) - e) < 0) {
            if (global_timer_ns() - t0 > a.timeoutNs) { atomicExch(ctl + 2, 1u); break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
}
__device__ __forceinline__ float4 peer_reduce_vertex(const PeerArgs &a, int b, unsigned e) {
    const float4 *recv = reinterpret_cast<const float4 *>(a.self + kPeerRecvOff) + (size_t)(e & 1u) * a.selfTotal;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int j = a.srcStart[b]; j < a.srcStart[b + 1]; j++) {
        const int k = a.src[j];
        const float4 v = k < a.numBoundary ? a.bsum[k] : __ldcg(recv + (k - a.numBoundary));  // written by a peer: not through L1
        sx += v.x; sy += v.y; sz += v.z;
    }
    return make_float4(sx, sy, sz, 0.0f);
}

// ---- fused form: self-validating 32-byte entries (launch.h) ----
__device__ __forceinline__ void st_volatile_u4(void *p, unsigned x, unsigned y, unsigned z, unsigned w) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_u4(const void *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// receiver: the partials one sharer delivered for receive entry E, added in the sharer's order.  Spins until the
// entries carry this epoch's tag; after a timeout the sticky error flag is set and whatever is there is used.
__device__ __forceinline__ float4 peer_poll_entry(const PeerArgs &a, unsigned *ctl, int E, unsigned e) {
    const unsigned char *base = a.self + kPeerRecvOff + (((size_t)(e & 1u) * a.selfTotal + E) * kPeerK) * 32;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    int n = 1;
    unsigned long long t0 = 0ull;
    for (int i = 0; i < n; i++) {
        uint4 u, v;
        for (unsigned spin = 0;; spin++) {
            u = ld_volatile_u4(base + i * 32);
            v = ld_volatile_u4(base + i * 32 + 16);
            const unsigned tag = u.y;
            if (tag / (unsigned)kPeerK == (e & 0x0fffffffu) && u.w == tag && v.y == tag && v.w == tag) { n = (int)(tag % (unsigned)kPeerK) + 1; break; }
            if ((spin & 63u) == 63u) {
                if (*reinterpret_cast<volatile unsigned *>(ctl + 2)) break;  // already failed once: do not pile up timeouts
                const unsigned long long now = global_timer_ns();
                if (t0 == 0ull) t0 = now;
                else if (now - t0 > a.timeoutNs) { atomicExch(ctl + 2, 1u); break; }
                __nanosleep(100);
            }
        }
        sx += __uint_as_float(u.x); sy += __uint_as_float(u.z); sz += __uint_as_float(v.x);
    }
    return make_float4(sx, sy, sz, 0.0f);
}

// the same on 32-bit shared-window addresses (converted once per worker, not per call)
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void *src, uint32_t bytes, uint32_t b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t b, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(b), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void cp_async16_a(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int T, int S>
struct TileSmem {
    static constexpr int TET_BYTES = T * 48;
    // byte offsets inside the worker's shared-memory region (computed, never indexed: stays in registers)
    int sxBytes, metaStride, sdx, sx0, meta0, bars, total;
    __host__ __device__ TileSmem(int metaStride_, int maxTileVertsPad, int maxTileEntries) {
        sxBytes = maxTileVertsPad * 16;
        metaStride = metaStride_;
        sdx = 0;
        sx0 = sdx + ((maxTileEntries + 1) * 16 + 127) / 128 * 128;  // + one spare entry for padding records
        meta0 = sx0 + S * sxBytes;
        bars = meta0 + (S + 1) * metaStride;
        total = (bars + (S + 1) * 8 + 127) & ~127;
    }
    __host__ __device__ int sx(int buf) const { return sx0 + buf * sxBytes; }
    __host__ __device__ int meta(int slot) const { return meta0 + slot * metaStride; }
};

__device__ __forceinline__ float4 ldg_stream4(const void *p) {  // read-once stream: do not pollute L1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int N>
__device__ __forceinline__ void cp_async_wait_pending() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One persistent worker (a CTA when WARP_SCOPE = false, a single warp when true) walking tiles
// first, first + stride, ...  NT = threads of the worker, TPT = tets per thread, T = NT * TPT.
// S-stage ring: at tile k the tet block and vertex gather of tile k+S-1 and the meta block of tile
// k+S are put in flight, so S-1 tiles of HBM/L2 latency are covered by math.
template <int NT, int TPT, int S, bool WARP_SCOPE, bool DBG = false, bool PEER = false>
__device__ __forceinline__ void tile_worker(const TileArgs &a, unsigned char *ws, const int tid, const int first,
                                            const int stride) {
    constexpr int T = NT * TPT;
    const int dbg = DBG ? a.debugSkip : 0;  // ablation switches exist only in the DBG instantiation (tetsim_time_kernel)
    const bool trackVol = a.volAcc != nullptr;
    const TileSmem<T, S> L(a.metaStride, a.maxTileVertsPad, a.maxTileEntries);
    uint64_t *metaFull = reinterpret_cast<uint64_t *>(ws + L.bars);  // [S + 1]
    unsigned char *const sdx = ws + L.sdx;
    const uint32_t wsa = smem_u32(ws);           // shared-window address of the worker's region
    const uint32_t barA = wsa + (uint32_t)L.bars;  // mbarrier i at barA + 8 i
    auto sync = [&]() { if (WARP_SCOPE) __syncwarp(); else __syncthreads(); };

    if (tid == 0) {
        for (int i = 0; i < S + 1; i++) mbar_init(metaFull + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    sync();

    const SubstepParams *sp = a.sp;
    const float alphaDev = sp->alphaDev * 6.0f, alphaVol = sp->alphaVol * 6.0f, gammaVol = sp->gammaVol;  // invRestVolume = 6 det Q
    // fused peer push: the epoch does not change while this kernel runs (its last CTA advances it) -- read it once
    const uint4 *pushRec = PEER ? a.px->pushRec : nullptr;
    const unsigned pushEpoch = PEER ? *reinterpret_cast<volatile unsigned *>(a.px->self + kPeerCtlOff) + 1u : 0u;

    auto issue_meta = [&](uint32_t o, uint32_t end, int slot) {  // one thread; block = [16*o, 16*end)
        const uint32_t bytes = (end - o) * 16u;
        mbar_expect_tx_a(barA + 8u * (uint32_t)slot, bytes);
        bulk_g2s_a(wsa + (uint32_t)L.meta(slot), a.meta + (size_t)o * 16, bytes, barA + 8u * (uint32_t)slot);
    };
    // The tet stream is read once, straight into registers (3.5 coalesced LDG.128 per tet: staging it in
    // shared memory would cost a write and a read of the SM's 128 B/clk shared-memory pipe, which the
    // gather/scatter traffic of this kernel already loads heavily).  Its HBM latency is hidden by a
    // bulk L2 prefetch of the whole 56*T-byte block issued PF tiles ahead.
    constexpr int PF = 2;
    auto prefetch_tets = [&](int tile) {  // one thread
        bulk_prefetch_l2(a.tets + (size_t)tile * TileSmem<T, S>::TET_BYTES, (uint32_t)TileSmem<T, S>::TET_BYTES);
    };
    auto issue_gather = [&](int slot, int buf) {  // all threads; always commits exactly one group
        const unsigned char *m = ws + L.meta(slot);
        const int nl = reinterpret_cast<const int *>(m)[1];
        const int *ids = reinterpret_cast<const int *>(m + reinterpret_cast<const int *>(m)[3]);
        const uint32_t sxa = wsa + (uint32_t)L.sx(buf);
        if (!(dbg & 4))
            for (int j = tid; j < nl; j += NT) cp_async16_a(sxa + 16u * (uint32_t)j, a.x4 + ids[j]);
        cp_async_commit();
    };

    // prologue: meta of tiles 0..S-1, tet blocks + gathers of tiles 0..S-2.  Thread 0 keeps the block
    // range of the NEXT meta it will issue in registers (loaded an iteration early, never stalls).
    const bool issuer = WARP_SCOPE ? tid == 0 : tid == NT / 2;  // the thread that issues bulk copies in the loop
    uint32_t nOff = 0, nEnd = 0;
    if (issuer) {
#pragma unroll
        for (int i = 0; i < S; i++)
            if (first + i * stride < a.numTiles) issue_meta(a.metaOff[first + i * stride], a.metaOff[first + i * stride + 1], i);
#pragma unroll
        for (int i = 1; i <= PF; i++)
            if (first + i * stride < a.numTiles) prefetch_tets(first + i * stride);
        if (first + S * stride < a.numTiles) { nOff = a.metaOff[first + S * stride]; nEnd = a.metaOff[first + S * stride + 1]; }
    }
#pragma unroll
    for (int i = 0; i < S - 1; i++) {
        if (first + i * stride < a.numTiles) { mbar_wait_a(barA + 8u * (uint32_t)i, 0); issue_gather(i, i); }
        else cp_async_commit();
    }

    // The tile's records (3 x LDG.128 per tet), issued at the top of the tile's iteration: they land while the CTA waits for
    // its gathers and the barrier.  (Issuing them one phase earlier -- for the NEXT tile, right before this tile's corner
    // sums, into the same registers -- was measured three times, rounds 1 and 2, and is 3 % slower each time.)
    // Issuing the next tile's GATHER in their shadow (right after the first barrier, by all threads) is worse still: +8 %.
    // Staggering the start of co-resident CTAs changes nothing.  (profiles/r2_tile_experiments.txt)
    float4 rA[TPT], rB[TPT], rC[TPT];
    auto load_records = [&](int tile) {
        const unsigned char *tb = a.tets + (size_t)tile * TileSmem<T, S>::TET_BYTES;
#pragma unroll
        for (int u = 0; u < TPT; u++) {
            const int t = tid + NT * u;
            if (DBG && (dbg & 8)) {  // measurement only: synthetic record, no HBM stream
                const float f = 1.0f + 1e-3f * (float)(t & 7);
                rA[u] = make_float4(f, 0.01f, 0.02f, f); rB[u] = make_float4(0.03f, f, 1.0f, 1.0f);
                rC[u] = make_float4(__uint_as_float(((t * 16) & 0x3ff) | (((t * 16 + 16) & 0x3ff) << 16)),
                                    __uint_as_float(((t * 16 + 32) & 0x3ff) | (((t * 16 + 48) & 0x3ff) << 16)),
                                    __uint_as_float((unsigned)(t * 64) | (unsigned)(t * 64 + 16) << 16),
                                    __uint_as_float((unsigned)(t * 64 + 32) | (unsigned)(t * 64 + 48) << 16));
                continue;
            }
            rA[u] = ldg_stream4(tb + t * 16);
            rB[u] = ldg_stream4(tb + T * 16 + t * 16);
            rC[u] = ldg_stream4(tb + T * 32 + t * 16);
        }
    };
    int k = 0;
    for (int c = first; c < a.numTiles; c += stride, k++) {
        const int cur = k % S, mcur = k % (S + 1);
        load_records(c);
        cp_async_wait_pending<S - 2>();
        sync();  // this tile's gathers (all threads') landed; previous tile's corner sums are finished

        // ---- put tile k+S-1 (vertex gather) and tile k+S (meta) in flight ----
        // The corner sums further down keep only the first ceil(nl/32) warps busy; the issue work is
        // therefore handed to the LAST warps of the CTA, which would otherwise idle at the next barrier.
        auto prefetch_next = [&]() {
            const int kn = k + S - 1, cn = c + (S - 1) * stride;
            constexpr int NI = WARP_SCOPE ? NT : NT / 2;      // threads that issue
            const int it = WARP_SCOPE ? tid : tid - (NT - NI);  // their index, < 0 for the others
            if (cn < a.numTiles) {
                const int mslot = kn % (S + 1), buf = kn % S;
                if (it >= 0) {
                    mbar_wait_a(barA + 8u * (uint32_t)mslot, (kn / (S + 1)) & 1);
                    const unsigned char *m = ws + L.meta(mslot);
                    const int nl = reinterpret_cast<const int *>(m)[1];
                    const int *ids = reinterpret_cast<const int *>(m + reinterpret_cast<const int *>(m)[3]);
                    const uint32_t sxa = wsa + (uint32_t)L.sx(buf);
                    if (!(dbg & 4))
                        for (int j = it; j < nl; j += NI) cp_async16_a(sxa + 16u * (uint32_t)j, a.x4 + ids[j]);
                }
            }
            cp_async_commit();
            if (it == 0) {
                if (c + (PF + 1) * stride < a.numTiles) prefetch_tets(c + (PF + 1) * stride);
                if (c + S * stride < a.numTiles) {
                    issue_meta(nOff, nEnd, (k + S) % (S + 1));
                    if (c + (S + 1) * stride < a.numTiles) { nOff = a.metaOff[c + (S + 1) * stride]; nEnd = a.metaOff[c + (S + 1) * stride + 1]; }
                }
            }
        };

        // ---- per-tet solve ----
        const unsigned char *sxb = ws + L.sx(cur);
        float vsum = 0.0f;
#pragma unroll
        for (int u = 0; u < TPT; u++) {
            const float4 A = rA[u], B = rB[u], C = rC[u];
            const unsigned s01 = __float_as_uint(C.x), s23 = __float_as_uint(C.y);
            const uint2 D = make_uint2(__float_as_uint(C.z), __float_as_uint(C.w));
            const float4 q0 = *reinterpret_cast<const float4 *>(sxb + (s01 & 0xffffu));
            const float4 q1 = *reinterpret_cast<const float4 *>(sxb + (s01 >> 16));
            const float4 q2 = *reinterpret_cast<const float4 *>(sxb + (s23 & 0xffffu));
            const float4 q3 = *reinterpret_cast<const float4 *>(sxb + (s23 >> 16));
            const V3 q[4] = {{q0.x, q0.y, q0.z}, {q1.x, q1.y, q1.z}, {q2.x, q2.y, q2.z}, {q3.x, q3.y, q3.z}};
            const float w[4] = {q0.w, q1.w, q2.w, q3.w};
            const float Bm[6] = {A.x, A.y, A.z, A.w, B.x, B.y};
            V3 d[4] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
            float vm1 = 0.0f;
            if (!(dbg & 2)) vm1 = nh_solve_tile<true>(q, w, Bm, B.z, B.w, alphaDev, alphaVol, gammaVol, d);
            if ((dbg & 16) && d[0].x + d[1].y + d[2].z + d[3].x != 123.456f) continue;  // measurement only: no scatter
            *reinterpret_cast<float4 *>(sdx + (D.x & 0xffffu)) = make_float4(d[0].x, d[0].y, d[0].z, 0.f);
            *reinterpret_cast<float4 *>(sdx + (D.x >> 16)) = make_float4(d[1].x, d[1].y, d[1].z, 0.f);
            *reinterpret_cast<float4 *>(sdx + (D.y & 0xffffu)) = make_float4(d[2].x, d[2].y, d[2].z, 0.f);
            *reinterpret_cast<float4 *>(sdx + (D.y >> 16)) = make_float4(d[3].x, d[3].y, d[3].z, 0.f);
            if (trackVol) vsum += (B.z != 0.0f) ? vm1 : 0.0f;  // padding records carry invRestVolume = 0
        }
        if (trackVol) {  // volError (src/Softbody.js:163): warp-reduce, one double atomic per warp
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
            if ((tid & 31) == 0) atomicAdd(a.volAcc, (double)vsum);
        }
        sync();
        prefetch_next();

        // ---- per-tile-vertex sum of corner dx (grouped rows, mesh_prep.cpp): thread j owns tile vertex j, its entries sit
        // 144 bytes apart from a per-group base -> LDS.128 with immediate offsets, four in flight at a time ----
        const unsigned char *m = ws + L.meta(mcur);
        const int v0 = reinterpret_cast<const int *>(m)[0];
        const int nl = reinterpret_cast<const int *>(m)[1];
        const uint16_t *gbase = reinterpret_cast<const uint16_t *>(m + 16);
        if (!(dbg & 1))
            for (int j = tid; j < nl; j += NT) {
                const int val = m[a.metaValOff + j];
                const unsigned char *p = sdx + gbase[j >> 3] + ((j & 7) << 4);
                float ax = 0.0f, ay = 0.0f, az = 0.0f;
                int i = 0;
#pragma unroll 1
                for (; i + 2 <= val; i += 2, p += 2 * 144) {  // (four per trip spills at the 64-register cap)
                    const float4 d0 = *reinterpret_cast<const float4 *>(p), d1 = *reinterpret_cast<const float4 *>(p + 144);
                    ax += d0.x; ay += d0.y; az += d0.z;
                    ax += d1.x; ay += d1.y; az += d1.z;
                }
                if (i < val) {
                    const float4 d0 = *reinterpret_cast<const float4 *>(p);
                    ax += d0.x; ay += d0.y; az += d0.z;
                }
                if (a.acc) atomicAdd(a.acc + reinterpret_cast<const int *>(m + reinterpret_cast<const int *>(m)[3])[j],
                                     make_float4(ax, ay, az, 0.0f));
                else a.part[v0 + j] = make_float4(ax, ay, az, 0.0f);
                if (PEER) {  // multi-GPU, fused exchange: a rank-shared vertex's tile partial goes straight to the sharers
                    const int slot = v0 + j;
                    if (slot < a.pxSlots) {  // boundary tiles come first, so do their partial slots
                        // ONE 32-byte record per pushing slot (two parallel 16-byte loads, addresses precomputed at
                        // tetsim_set_peers): a dependent chain slot -> CSR -> peer table -> base cost ~3 us in the first
                        // tile of every CTA that owns boundary tiles (measured, profiles/r2_peer_experiments.txt)
                        const uint4 r0 = __ldg(pushRec + 2 * slot), r1 = __ldg(pushRec + 2 * slot + 1);
                        if (r0.x != 0xffffffffu) {
                            const unsigned i = r0.x & 0xffu, n1 = (r0.x >> 8) & 0xffu, sharers = r0.x >> 16;
                            const unsigned tag = (pushEpoch & 0x0fffffffu) * (unsigned)kPeerK + n1;
                            const unsigned long long p0 = ((unsigned long long)r0.w << 32) | r0.z, p1 = ((unsigned long long)r1.w << 32) | r1.z;
                            unsigned char *d0 = reinterpret_cast<unsigned char *>(p0) + (size_t)(pushEpoch & 1u) * r1.x * 32;
                            st_volatile_u4(d0, __float_as_uint(ax), tag, __float_as_uint(ay), tag);
                            st_volatile_u4(d0 + 16, __float_as_uint(az), tag, 0u, tag);
                            if (sharers > 1) {
                                unsigned char *d1 = reinterpret_cast<unsigned char *>(p1) + (size_t)(pushEpoch & 1u) * r1.y * 32;
                                st_volatile_u4(d1, __float_as_uint(ax), tag, __float_as_uint(ay), tag);
                                st_volatile_u4(d1 + 16, __float_as_uint(az), tag, 0u, tag);
                            }
                            if (sharers > 2) {  // a vertex on a corner of the partition: the remaining sharers from the CSR
                                const PeerArgs &px = *a.px;
                                const int b = (int)r0.y;
                                for (int t = px.pxStart[b] + 2; t < px.pxStart[b + 1]; t++) {
                                    const int q = px.pxPeer[t];
                                    unsigned char *dst = px.peerBase[q] + kPeerRecvOff +
                                                         (((size_t)(pushEpoch & 1u) * px.remoteTotal[q] + px.pxEntry[t]) * kPeerK + i) * 32;
                                    st_volatile_u4(dst, __float_as_uint(ax), tag, __float_as_uint(ay), tag);
                                    st_volatile_u4(dst + 16, __float_as_uint(az), tag, 0u, tag);
                                }
                            }
                        }
                    }
                }
            }
    }
    if (PEER && !WARP_SCOPE) {
        // every push of this CTA is issued (and used the epoch read above): take a ticket, the last CTA advances the
        // epoch.  (The vertex kernel used to do this with one same-address atomic per block -- 3,400 per launch at two
        // ranks; moving it here was worth 9 % of the substep, profiles/r2_peer_experiments.txt.)
        __syncthreads();
        if (tid == 0) {
            unsigned *ctl = reinterpret_cast<unsigned *>(a.px->self + kPeerCtlOff);
            const unsigned e = *reinterpret_cast<volatile unsigned *>(ctl) + 1u;
            if (atomicAdd(ctl + 1, 1u) == gridDim.x - 1) {
                ctl[1] = 0u;
                *reinterpret_cast<volatile unsigned *>(ctl) = e;
            }
        }
    }
}

// CTA tiles: T tets per tile, one tet per thread, two __syncthreads per tile.  DBG = true compiles the ablation
// switches of TileArgs::debugSkip in (tetsim_time_kernel with TETSIM_TILE_DEBUG only).
template <int T, int S, int MINB, bool DBG = false>
__global__ void __launch_bounds__(T, MINB) k_jacobi_tiles(TileArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (a.tileBegin + (int)blockIdx.x >= a.numTiles) return;
    tile_worker<T, 1, S, false, DBG>(a, smem, threadIdx.x, a.tileBegin + blockIdx.x, gridDim.x);
}

// Same with TPT tets per thread (T/TPT threads per tile): per-tile overhead (barriers, prefetch issue,
// loop control) is shared by TPT times the tets and every thread carries TPT independent chains.
constexpr int tilesN_minb(int T, int TPT, int MINB) { return MINB > 0 ? MINB : 1024 / T; }
// PEER = true adds the fused peer-memory push (TileArgs::px must be set); the single-GPU kernels carry none of it.
template <int T, int TPT, int S, int MINB, bool PEER = false>
__global__ void __launch_bounds__(T / TPT, tilesN_minb(T, TPT, MINB)) k_jacobi_tilesN(TileArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (a.tileBegin + (int)blockIdx.x >= a.numTiles) return;
    tile_worker<T / TPT, TPT, S, false, false, PEER>(a, smem, threadIdx.x, a.tileBegin + blockIdx.x, gridDim.x);
}

// Warp tiles: every warp is its own worker with private staging and mbarriers (tile = 32 * TPL tets);
// no block-level barrier anywhere, so warps drift apart and cover each other's load and sum phases.
template <int TPL, int S>
__global__ void __launch_bounds__(TPL == 1 ? 832 : 448, 1) k_jacobi_warptiles(TileArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int first = a.tileBegin + blockIdx.x * wpb + wid;
    if (first >= a.numTiles) return;
    const TileSmem<32 * TPL, S> L(a.metaStride, a.maxTileVertsPad, a.maxTileEntries);
    tile_worker<32, TPL, S, true>(a, smem + (size_t)wid * L.total, lane, first, gridDim.x * wpb);
}

template <int T, int S>
static size_t tile_smem_bytes(const TileArgs &a) { return (size_t)TileSmem<T, S>(a.metaStride, a.maxTileVertsPad, a.maxTileEntries).total; }

// Kernel shape per tile size.  Defaults are the round-1 sweep winners on the 10M-tet beam (profiles/r1_tile_sweep.txt):
// two tets per thread (independent dependency chains, half the per-tile barrier/prefetch overhead per tet, corner sums
// spread over all warps), two pipeline stages, registers capped so 4 x 256 threads (T = 512) stay resident per SM.
// TETSIM_TILE_TPT / TETSIM_TILE_STAGES / TETSIM_TILE_MINB override them for experiments (tools/tile_sweep.py).
struct TileShape { int tpt, stages, minb; };
static TileShape tile_shape(int clusterSize) {
    TileShape sh{clusterSize >= 128 ? 2 : 1, 0, 0};
    if (const char *e = getenv("TETSIM_TILE_TPT")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4) sh.tpt = v; }
    if (clusterSize < 128) sh.tpt = 1;  // warp tiles
    sh.stages = sh.tpt == 1 ? 3 : 2;
    if (const char *e = getenv("TETSIM_TILE_STAGES")) { int v = atoi(e); if (v >= 2 && v <= 4) sh.stages = v; }
    sh.minb = (clusterSize == 512 && sh.tpt == 2) ? 4 : 0;
    if (const char *e = getenv("TETSIM_TILE_MINB")) sh.minb = atoi(e);
    return sh;
}
static int tile_stages(int clusterSize) { return tile_shape(clusterSize).stages; }

size_t jacobi_tiles_smem(int clusterSize, const TileArgs &a) {
    const int S = tile_stages(clusterSize);
#define TS_CASE(T_) case T_: return S == 2 ? tile_smem_bytes<T_, 2>(a) : (S == 3 ? tile_smem_bytes<T_, 3>(a) : tile_smem_bytes<T_, 4>(a));
    switch (clusterSize) { TS_CASE(32) TS_CASE(64) TS_CASE(128) TS_CASE(256) TS_CASE(512) default: return 0; }
#undef TS_CASE
}

// Launch configuration is cached per (kernel instantiation, device): attributes such as the dynamic
// shared-memory opt-in are per device, and a process may hold handles on several GPUs.
struct LaunchCache { size_t smem = 0; int n = 0, sms = 0; };
static std::mutex g_launchMu;   // handles on different host threads share the per-device launch caches below
static int current_device() { int d = 0; cudaGetDevice(&d); return d < 0 || d >= 64 ? 0 : d; }

template <int T, int S, int MINB, bool DBG = false>
static void launch_tiles_T(cudaStream_t s, const TileArgs &a) {
    const size_t smem = tile_smem_bytes<T, S>(a);
    std::lock_guard<std::mutex> lk(g_launchMu);
    static LaunchCache cache[64];
    LaunchCache &lc = cache[current_device()];
    if (smem != lc.smem) {
        cudaFuncSetAttribute(k_jacobi_tiles<T, S, MINB, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lc.n, k_jacobi_tiles<T, S, MINB, DBG>, T, smem);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&lc.sms, cudaDevAttrMultiProcessorCount, dev);
        if (lc.n < 1) lc.n = 1;
        lc.smem = smem;
    }
    int grid = lc.sms * lc.n;  // persistent: every CTA resident, tiles strided over the grid
    if (grid > a.numTiles - a.tileBegin) grid = a.numTiles - a.tileBegin;
    k_jacobi_tiles<T, S, MINB, DBG><<<grid, T, smem, s>>>(a);
}

template <int T, int TPT, int S, int MINB, bool PEER = false>
static void launch_tilesN(cudaStream_t s, const TileArgs &a) {
    const size_t smem = tile_smem_bytes<T, S>(a);
    std::lock_guard<std::mutex> lk(g_launchMu);
    static LaunchCache cache[64];
    LaunchCache &lc = cache[current_device()];
    if (smem != lc.smem) {
        cudaFuncSetAttribute(k_jacobi_tilesN<T, TPT, S, MINB, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lc.n, k_jacobi_tilesN<T, TPT, S, MINB, PEER>, T / TPT, smem);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&lc.sms, cudaDevAttrMultiProcessorCount, dev);
        if (lc.n < 1) lc.n = 1;
        lc.smem = smem;
    }
    int grid = lc.sms * lc.n;
    if (grid > a.numTiles - a.tileBegin) grid = a.numTiles - a.tileBegin;
    k_jacobi_tilesN<T, TPT, S, MINB, PEER><<<grid, T / TPT, smem, s>>>(a);
}

// Warp tiles: one CTA per SM holding as many warps as shared memory and registers allow.
template <int TPL, int S>
static void launch_warptiles(cudaStream_t s, const TileArgs &a) {
    const size_t perWarp = tile_smem_bytes<32 * TPL, S>(a);
    std::lock_guard<std::mutex> lk(g_launchMu);
    static LaunchCache cache[64];
    LaunchCache &lc = cache[current_device()];
    if (perWarp != lc.smem) {
        int dev = 0, maxSmem = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&lc.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, k_jacobi_warptiles<TPL, S>);
        const int byRegs = 65536 / (32 * (fa.numRegs > 0 ? fa.numRegs : 64));
        int warps = (int)((size_t)maxSmem / perWarp);
        if (warps > byRegs) warps = byRegs;
        if (warps > (TPL == 1 ? 832 : 448) / 32) warps = (TPL == 1 ? 832 : 448) / 32;
        if (const char *w = getenv("TETSIM_WARPS_PER_SM")) { int v = atoi(w); if (v >= 1 && v < warps) warps = v; }
        if (warps < 1) warps = 1;
        cudaFuncSetAttribute(k_jacobi_warptiles<TPL, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(perWarp * warps));
        lc.n = warps;
        lc.smem = perWarp;
    }
    int grid = lc.sms;
    if ((long long)grid * lc.n > a.numTiles - a.tileBegin) grid = (a.numTiles - a.tileBegin + lc.n - 1) / lc.n;
    k_jacobi_warptiles<TPL, S><<<grid, 32 * lc.n, perWarp * lc.n, s>>>(a);
}

// The k_jacobi_tilesN instantiations that exist: (tile size, tets per thread, stages, CTAs per SM the registers are capped
// for; 0 = 1024 / T).  The first of each tile size is the default shape, the rest are reachable through the
// TETSIM_TILE_* overrides (tools/tile_sweep.py).
#define TILESN_SHAPES(X)                                                              \
    X(128, 2, 2, 0) X(128, 2, 3, 0)                                                   \
    X(256, 2, 2, 0) X(256, 2, 3, 0) X(256, 2, 2, 6) X(256, 2, 2, 8) X(256, 4, 2, 0)   \
    X(512, 2, 2, 4) X(512, 2, 2, 0) X(512, 2, 2, 3) X(512, 2, 3, 4) X(512, 4, 2, 0) X(512, 4, 2, 3)

// The fused peer push exists only in the k_jacobi_tilesN instantiations (two or more tets per thread).
bool jacobi_tiles_has_peer_push(int clusterSize) {
    const TileShape sh = tile_shape(clusterSize);
#define X(T_, TPT_, S_, MINB_) if (clusterSize == T_ && sh.tpt == TPT_ && sh.stages == S_ && sh.minb == MINB_) return true;
    TILESN_SHAPES(X)
#undef X
    return false;
}

void launch_jacobi_tiles(cudaStream_t s, int clusterSize, const TileArgs &a) {
    if (a.numTiles - a.tileBegin <= 0) return;
    const TileShape sh = tile_shape(clusterSize);
    const int S = sh.stages;
    if (a.debugSkip && clusterSize == 256) { launch_tiles_T<256, 3, 4, true>(s, a); return; }  // ablations: one instantiation
#define X(T_, TPT_, S_, MINB_)                                                                       \
    if (clusterSize == T_ && sh.tpt == TPT_ && S == S_ && sh.minb == MINB_) {                        \
        if (a.px) launch_tilesN<T_, TPT_, S_, MINB_, true>(s, a); else launch_tilesN<T_, TPT_, S_, MINB_>(s, a); \
        return;                                                                                      \
    }
    TILESN_SHAPES(X)
#undef X
    // no such instantiation (one tet per thread, warp tiles): the kernels below carry no peer push
    switch (clusterSize) {
        case 32: S == 2 ? launch_warptiles<1, 2>(s, a) : (S == 3 ? launch_warptiles<1, 3>(s, a) : launch_warptiles<1, 4>(s, a)); break;
        case 64: S == 2 ? launch_warptiles<2, 2>(s, a) : (S == 3 ? launch_warptiles<2, 3>(s, a) : launch_warptiles<2, 4>(s, a)); break;
        case 128: S == 2 ? launch_tiles_T<128, 2, 6>(s, a) : (S == 3 ? launch_tiles_T<128, 3, 5>(s, a) : launch_tiles_T<128, 4, 4>(s, a)); break;
        case 256: S == 2 ? launch_tiles_T<256, 2, 4>(s, a) : (S == 3 ? launch_tiles_T<256, 3, 4>(s, a) : launch_tiles_T<256, 4, 2>(s, a)); break;
        case 512: launch_tiles_T<512, 2, 2>(s, a); break;
        default: break;
    }
}

// Tile-major tet blocks from the tet-order rest data (device) and the host-built slot/destination words.
template <int T>
__global__ void k_build_tiles(int numRecords, const int *__restrict__ order, const float *__restrict__ Q9,
                              const float *__restrict__ irv, const uint4 *__restrict__ aux,
                              unsigned char *__restrict__ tets) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= numRecords) return;
    const int tile = r / T, t = r % T;
    unsigned char *tb = tets + (size_t)tile * T * 48;
    const int e = order[r];
    const uint4 x = aux[r];
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;  // padding record: B = 0, invRestVolume = 0 -> no correction
    if (e >= 0) {
        // rest metric B = Q Q^T (Q column-major: Q(i,k) = q[3k+i]), formed in f64 and rounded once
        const float *q = Q9 + 9 * (size_t)e;
        double b[3][3];
        for (int i = 0; i < 3; i++)
            for (int j = i; j < 3; j++)
                b[i][j] = (double)q[i] * (double)q[j] + (double)q[3 + i] * (double)q[3 + j] + (double)q[6 + i] * (double)q[6 + j];
        const float rv = irv[e];
        A = make_float4((float)b[0][0], (float)b[0][1], (float)b[0][2], (float)b[1][1]);
        const float dq = rv / 6.0f;  // det Q = 1 / det Dm = invRestVolume / 6
        // a zero-volume rest tet (1 / V = inf) stays an all-zero record: no correction, like the reference's early returns
        if (fabsf(rv) < INFINITY && rv == rv) B = make_float4((float)b[1][2], (float)b[2][2], dq * dq, dq);
        else A = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    reinterpret_cast<float4 *>(tb)[t] = A;
    reinterpret_cast<float4 *>(tb + T * 16)[t] = B;
    reinterpret_cast<uint4 *>(tb + T * 32)[t] = x;
}
void launch_build_tiles(cudaStream_t s, int clusterSize, int numRecords, const int *order, const float *Q9,
                        const float *irv, const uint4 *aux, unsigned char *tets) {
    if (numRecords <= 0) return;
    const int g = cdiv(numRecords, 256);
    if (clusterSize == 32) k_build_tiles<32><<<g, 256, 0, s>>>(numRecords, order, Q9, irv, aux, tets);
    else if (clusterSize == 64) k_build_tiles<64><<<g, 256, 0, s>>>(numRecords, order, Q9, irv, aux, tets);
    else if (clusterSize == 128) k_build_tiles<128><<<g, 256, 0, s>>>(numRecords, order, Q9, irv, aux, tets);
    else if (clusterSize == 256) k_build_tiles<256><<<g, 256, 0, s>>>(numRecords, order, Q9, irv, aux, tets);
    else k_build_tiles<512><<<g, 256, 0, s>>>(numRecords, order, Q9, irv, aux, tets);
}

// Vertex side of the clustered Jacobi: x += (sum of the vertex's tile partials) / valence, optionally
// fused with post (simulate() :213-239) and with the NEXT substep's predict (:198-202) so a substep
// inside tetsim_step costs exactly two launches.
template <int MODE, bool PEER>
__global__ void __launch_bounds__(256) k_jacobi_apply(int begin, int end, ApplyArgs a) {
    // fused peer exchange: the rank-shared vertices are the LAST of the handle's numbering, and they may have to wait for
    // a sharer's entries -- their blocks are scheduled first so that the wait overlaps the interior vertices' work
    const unsigned blk = PEER ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
    int i = begin + blk * blockDim.x + threadIdx.x;
    unsigned epoch = 0u;
    if (PEER) epoch = *reinterpret_cast<volatile unsigned *>(a.px->self + kPeerCtlOff);  // advanced by the tile kernel's last CTA
    if (i >= end) return;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    if (PEER && i >= a.boundaryBegin) {
        // each sharer's partials in that sharer's order, the sharers in ascending rank order: identical on every sharer
        const PeerArgs &px = *a.px;
        unsigned *ctl = reinterpret_cast<unsigned *>(px.self + kPeerCtlOff);
        const int b = i - a.boundaryBegin;
        for (int j = px.srcStart[b]; j < px.srcStart[b + 1]; j++) {
            const int k = px.src[j];
            float4 s;
            if (k < px.numBoundary) {  // this rank's own partials
                s = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int t = a.vpStart[i]; t < a.vpStart[i + 1]; t++) {
                    const float4 v = ldg4(a.part + a.vpSlot[t]);
                    s.x += v.x; s.y += v.y; s.z += v.z;
                }
            } else {
                s = peer_poll_entry(px, ctl, k - px.numBoundary, epoch);
            }
            sx += s.x; sy += s.y; sz += s.z;
        }
    } else if (a.bsum && i >= a.boundaryBegin) {
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
            // post of this substep + predict of the next: the velocity lives in registers only.  Storing it would be
            // a dead 16 B/vertex write -- the next substep's post recomputes v from x and prev, and the LAST substep
            // of a tetsim_step call always runs MODE 1, which stores the velocity the caller (and the next call's
            // predict) sees.
            a.prev4[i] = x;
            v.y += sp->gDt;
            const float dt = sp->dtF;
            x.x = fmaf(v.x, dt, x.x); x.y = fmaf(v.y, dt, x.y); x.z = fmaf(v.z, dt, x.z);
        } else {
            a.vel4[i] = v;
        }
    }
    a.x4[i] = x;
}

void launch_jacobi_apply(cudaStream_t s, int begin, int end, int mode, const ApplyArgs &a) {
    int n = end - begin;
    if (n <= 0) return;
    const int TB = 256;
    if (a.px) {  // multi-GPU, fused peer exchange: poll + rank-ordered reduce for the rank-shared vertices
        if (mode == 0) k_jacobi_apply<0, true><<<cdiv(n, TB), TB, 0, s>>>(begin, end, a);
        else if (mode == 1) k_jacobi_apply<1, true><<<cdiv(n, TB), TB, 0, s>>>(begin, end, a);
        else k_jacobi_apply<2, true><<<cdiv(n, TB), TB, 0, s>>>(begin, end, a);
        return;
    }
    if (mode == 0) k_jacobi_apply<0, false><<<cdiv(n, TB), TB, 0, s>>>(begin, end, a);
    else if (mode == 1) k_jacobi_apply<1, false><<<cdiv(n, TB), TB, 0, s>>>(begin, end, a);
    else k_jacobi_apply<2, false><<<cdiv(n, TB), TB, 0, s>>>(begin, end, a);
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

__global__ void k_halo_pack(int n, const int *__restrict__ sendIdx, const float4 *__restrict__ bsum,
                            float4 *__restrict__ send) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) send[i] = bsum[sendIdx[i]];
}
void launch_halo_pack(cudaStream_t s, int n, const int *sendIdx, const float4 *bsum, float4 *send) {
    if (n > 0) k_halo_pack<<<cdiv(n, 256), 256, 0, s>>>(n, sendIdx, bsum, send);
}
// every sharer adds the sharers' partial sums in ascending rank order -> bit-identical replicas
__global__ void k_halo_reduce(int nB, const int *__restrict__ srcStart, const int *__restrict__ src,
                              const float4 *__restrict__ recv, float4 *__restrict__ bsum) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nB) return;
    const int s0 = srcStart[b], s1 = srcStart[b + 1];
    if (s0 == s1) return;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int j = s0; j < s1; j++) {
        const int k = src[j];
        const float4 v = k < nB ? bsum[k] : recv[k - nB];  // k == b for the own term: read before the write below
        sx += v.x; sy += v.y; sz += v.z;
    }
    bsum[b] = make_float4(sx, sy, sz, 0.0f);
}
void launch_halo_reduce(cudaStream_t s, int nB, const int *srcStart, const int *src, const float4 *recv, float4 *bsum) {
    if (nB > 0) k_halo_reduce<<<cdiv(nB, 256), 256, 0, s>>>(nB, srcStart, src, recv, bsum);
}

// ---- peer-memory exchange (layout and protocol: launch.h, PeerArgs) ----

__global__ void k_peer_push(PeerArgs a) {
    unsigned *ctl = reinterpret_cast<unsigned *>(a.self + kPeerCtlOff);
    // every block reads the epoch before it takes its ticket, the last ticket holder advances it: no block sees the new value
    const unsigned e = *reinterpret_cast<volatile unsigned *>(ctl) + 1u;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < a.numBoundary && a.pxStart[b + 1] > a.pxStart[b]) {  // active: touched by this rank's tets
        if (a.acc) {  // atomic flush: the accumulator holds the sum
            const int i = a.boundaryBegin + b;
            float4 sum = a.acc[i];
            sum.w = 0.f;
            a.acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            a.bsum[b] = sum;
            for (int j = a.pxStart[b]; j < a.pxStart[b + 1]; j++) {
                const int q = a.pxPeer[j];
                float4 *dst = reinterpret_cast<float4 *>(a.peerBase[q] + kPeerRecvOff) + (size_t)(e & 1u) * a.remoteTotal[q] + a.pxEntry[j];
                *dst = sum;
            }
        } else {
            peer_push_vertex(a, b, e, false);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();  // this block's remote stores are ordered before the ticket
        const unsigned ticket = atomicAdd(ctl + 1, 1u);
        if (ticket == gridDim.x - 1) peer_publish(a, ctl, e);
    }
}
void launch_peer_push(cudaStream_t s, const PeerArgs &a) {
    if (a.numBoundary > 0) k_peer_push<<<cdiv(a.numBoundary, 256), 256, 0, s>>>(a);
}

__global__ void k_peer_reduce(PeerArgs a) {
    unsigned *ctl = reinterpret_cast<unsigned *>(a.self + kPeerCtlOff);
    const unsigned e = *reinterpret_cast<volatile unsigned *>(ctl);  // advanced by this iteration's push
    peer_wait_block(a, ctl, e);
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.numBoundary || a.srcStart[b] == a.srcStart[b + 1]) return;
    a.bsum[b] = peer_reduce_vertex(a, b, e);  // the own term is bsum[b] itself: read before this write
}
void launch_peer_reduce(cudaStream_t s, const PeerArgs &a) {
    if (a.numBoundary > 0) k_peer_reduce<<<cdiv(a.numBoundary, 256), 256, 0, s>>>(a);
}

// =================================================================================================
// Gauss-Seidel in the reference order, FAST arithmetic: one CTA per body, FOUR LANES PER TET.
//
// The exact-order sweep is a chain of dependency levels (Dragon: 703 levels of <= 22 tets, 5.5 on average), so a substep
// costs (levels) x (latency of one tet's solve), whatever the width of the machine.  k_gs_body (kernels.cuh) runs one tet
// per thread: ~350 dependent-ish instructions, ~1,070 cycles per level.  Here a tet is solved by a quad of lanes -- lane c
// of the quad owns component c of every vector (x, y, z; the fourth lane idles with zeros) -- in the rest-metric form of
// the tile kernel (nh_solve_tile): per lane ~95 instructions; the 3x3 algebra that couples components (||F||^2, the
// weighted gradient norms, the cofactors, det F) crosses lanes with warp shuffles: two butterfly reductions of two values
// each and one rotation of the three updated edge vectors.  Corner displacements are added in place in shared memory
// (tets of one level share no vertex).  Records (metric, det Q, 4 body-local vertex ids) come from L2 with the next
// level's record prefetched into registers, as in k_gs_body.
// =================================================================================================
struct QuadRec { float4 A, B; int4 I; };   // A = (B00,B01,B02,B11)  B = (B12,B22,detQ, caller tet index as bits)  I = body-local vertex ids

__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// All 32 lanes call this together (full-mask shuffles); `valid` quads store.  Returns det F - 1 (lane-uniform in the quad).
__device__ __forceinline__ float gs_solve_quad(float4 *sx, const QuadRec &r, bool valid, int c, int lane, float alphaDev,
                                               float alphaVol, float gammaVol) {
    const float *f = reinterpret_cast<const float *>(sx);
    const int cc = c < 3 ? c : 0;
    const float m = c < 3 ? 1.0f : 0.0f;                        // the idle fourth lane carries zeros through the algebra
    const float q0 = f[4 * r.I.x + cc] * m, q1 = f[4 * r.I.y + cc] * m, q2 = f[4 * r.I.z + cc] * m, q3 = f[4 * r.I.w + cc] * m;
    const float w0 = f[4 * r.I.x + 3], w1 = f[4 * r.I.y + 3], w2 = f[4 * r.I.z + 3], w3 = f[4 * r.I.w + 3];
    float P0 = q1 - q0, P1 = q2 - q0, P2 = q3 - q0;
    const float G1 = fmaf(P2, r.A.z, fmaf(P1, r.A.y, P0 * r.A.x));
    const float G2 = fmaf(P2, r.B.x, fmaf(P1, r.A.w, P0 * r.A.y));
    const float G3 = fmaf(P2, r.B.y, fmaf(P1, r.B.x, P0 * r.A.z));
    const float nG0 = G1 + G2 + G3;
    const float rs2 = quad_sum(fmaf(P2, G3, fmaf(P1, G2, P0 * G1)));
    const float wG = quad_sum(fmaf(w3, G3 * G3, fmaf(w2, G2 * G2, fmaf(w1, G1 * G1, w0 * (nG0 * nG0)))));
    const float detQ = r.B.z, irv = 6.0f * detQ;   // 1 / V = 6 det Q
    const float r1 = rs2 * rcp_approx(fmaf(alphaDev * irv, rs2, wG));
    const float s = (rs2 > 0.0f && wG > 0.0f) ? -r1 : 0.0f;
    const float e1 = G1 * (s * w1), e2 = G2 * (s * w2), e3 = G3 * (s * w3), d0 = nG0 * (-(s * w0));
    P0 = P0 + e1 - d0; P1 = P1 + e2 - d0; P2 = P2 + e3 - d0;
    // cofactors: component c of a x b needs components c+1, c+2 of a and b -- held by the two other lanes of the quad
    const int base = lane & ~3, l1 = base | (c < 3 ? (c + 1) % 3 : 0), l2 = base | (c < 3 ? (c + 2) % 3 : 0);
    const float P0a = __shfl_sync(0xffffffffu, P0, l1), P0b = __shfl_sync(0xffffffffu, P0, l2);
    const float P1a = __shfl_sync(0xffffffffu, P1, l1), P1b = __shfl_sync(0xffffffffu, P1, l2);
    const float P2a = __shfl_sync(0xffffffffu, P2, l1), P2b = __shfl_sync(0xffffffffu, P2, l2);
    const float c1 = (P1a * P2b - P1b * P2a) * m, c2 = (P2a * P0b - P2b * P0a) * m, c3 = (P0a * P1b - P0b * P1a) * m;
    const float nc0 = c1 + c2 + c3;
    const float vol = quad_sum(P0 * c1) * detQ;
    const float wC = quad_sum(fmaf(w3, c3 * c3, fmaf(w2, c2 * c2, fmaf(w1, c1 * c1, w0 * (nc0 * nc0)))));
    const float C = vol - gammaVol;
    const float wt = (detQ * detQ) * wC;   // the true weighted gradient norm: the reference's `w == 0 -> return` (:184) tests this
    const float r2 = (C * detQ) * rcp_approx(fmaf(alphaVol, irv, wt));
    const float t = (C != 0.0f && wt > 0.0f) ? -r2 : 0.0f;
    if (valid && c < 3) {
        float *g = reinterpret_cast<float *>(sx);
        g[4 * r.I.x + c] = q0 + fmaf(nc0, -(t * w0), d0);
        g[4 * r.I.y + c] = q1 + fmaf(c1, t * w1, e1);
        g[4 * r.I.z + c] = q2 + fmaf(c2, t * w2, e2);
        g[4 * r.I.w + c] = q3 + fmaf(c3, t * w3, e3);
    }
    return vol - 1.0f;
}

__global__ void k_gs_body_quads(const BodyDesc *__restrict__ bodies, const int *__restrict__ levelStart,
                                float4 *__restrict__ x4, float4 *__restrict__ prev4, float4 *__restrict__ vel4,
                                const int4 *__restrict__ I, const float4 *__restrict__ A, const float4 *__restrict__ B,
                                const int *__restrict__ order, double *__restrict__ volTerm,
                                const SubstepParams *__restrict__ sp, const int *__restrict__ vertId) {
    extern __shared__ float4 sxq[];
    const BodyDesc bd = bodies[blockIdx.x];
    const int nv = bd.vertEnd - bd.vertBegin;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, c = tid & 3, quad = tid >> 2, nq = nt >> 2;
    const float dt = sp->dtF, gDt = sp->gDt;
    for (int j = tid; j < nv; j += nt) {   // predict (simulate() :198-202)
        const int i = bd.vertBegin + j;
        float4 x = x4[i], v = vel4[i];
        prev4[i] = x;
        v.y += gDt;
        x.x = fmaf(v.x, dt, x.x); x.y = fmaf(v.y, dt, x.y); x.z = fmaf(v.z, dt, x.z);
        sxq[j] = x;
    }
    int *sLevel = reinterpret_cast<int *>(sxq + nv);
    const int nLev = bd.levelEnd - bd.levelBegin;
    for (int j = tid; j <= nLev; j += nt) sLevel[j] = levelStart[bd.levelBegin + j];
    __syncthreads();
    const float alphaDev = sp->alphaDev, alphaVol = sp->alphaVol, gammaVol = sp->gammaVol;
    const QuadRec zero = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_int4(0, 0, 0, 0)};
    auto fetch = [&](int t) { QuadRec r; r.A = ldg4(A + t); r.B = ldg4(B + t); r.I = __ldg(I + t); return r; };
    // A level lasts ~300 cycles, an L2 hit ~700-1000: the record a quad needs at level l is loaded FOUR levels earlier into
    // a ring of four register sets (the level loop is unrolled by four so the ring is addressed statically).  With the
    // one-level look-ahead of k_gs_body the sweep ran at one L2 latency per level.
    auto fetch_level = [&](int l) {
        if (l < nLev) { const int t = sLevel[l] + quad; if (t < sLevel[l + 1]) return fetch(t); }
        return zero;
    };
    auto do_level = [&](int l, QuadRec &ring) {
        const int b = sLevel[l], e = sLevel[l + 1];
        const QuadRec cur = ring;
        ring = fetch_level(l + 4);
        if (b + (tid & ~31) / 4 < e) {   // warp-uniform: this warp has a tet in the first pass
            const bool valid = b + quad < e;
            const float vm1 = gs_solve_quad(sxq, cur, valid, c, lane, alphaDev, alphaVol, gammaVol);
            if (volTerm && valid && c == 0) volTerm[__float_as_int(cur.B.w)] = (double)vm1;   // tet index rides in the record: no load on the level's path
        }
        for (int t0 = b + nq; t0 < e; t0 += nq) {   // levels wider than the CTA's quads (block-uniform trip count)
            if (t0 + (tid & ~31) / 4 < e) {
                const bool valid = t0 + quad < e;
                const QuadRec r = valid ? fetch(t0 + quad) : zero;
                const float vm1 = gs_solve_quad(sxq, r, valid, c, lane, alphaDev, alphaVol, gammaVol);
                if (volTerm && valid && c == 0) volTerm[__float_as_int(r.B.w)] = (double)vm1;
            }
        }
        __syncthreads();
    };
    QuadRec R0 = fetch_level(0), R1 = fetch_level(1), R2 = fetch_level(2), R3 = fetch_level(3);
    for (int l = 0; l < nLev; l += 4) {
        do_level(l, R0);
        if (l + 1 < nLev) do_level(l + 1, R1);
        if (l + 2 < nLev) do_level(l + 2, R2);
        if (l + 3 < nLev) do_level(l + 3, R3);
    }
    for (int j = tid; j < nv; j += nt) {   // post (simulate() :213-239)
        const int i = bd.vertBegin + j;
        float4 x = sxq[j], p = prev4[i], v;
        v.w = 0.0f;
        post_vertex<false>(vertId ? vertId[i] : i, x, p, v, sp);
        x4[i] = x;
        vel4[i] = v;
    }
}

void launch_gs_body_quads(cudaStream_t s, int numBodies, int threads, size_t smemBytes, const BodyDesc *bodies,
                          const int *levelStart, float4 *x4, float4 *prev4, float4 *vel4, const int4 *I, const float4 *A,
                          const float4 *B, const int *order, double *volTerm, const SubstepParams *sp, const int *vertId) {
    if (numBodies <= 0) return;
    std::lock_guard<std::mutex> lk(g_launchMu);
    static size_t configured[64] = {0};
    size_t &cfg = configured[current_device()];
    if (smemBytes > 48 * 1024 && smemBytes > cfg) {
        cudaFuncSetAttribute(k_gs_body_quads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
        cfg = smemBytes;
    }
    k_gs_body_quads<<<numBodies, threads, smemBytes, s>>>(bodies, levelStart, x4, prev4, vel4, I, A, B, order, volTerm, sp, vertId);
}

// Level-sorted GS stream in the rest-metric form: A = (B00,B01,B02,B11), B = (B12,B22,detQ, caller tet index), B = Q Q^T.
__global__ void k_build_stream_metric(int count, const int *__restrict__ order, const float *__restrict__ Q9,
                                      const float *__restrict__ irv, float4 *__restrict__ A, float4 *__restrict__ B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int e = order ? order[i] : i;
    const float *q = Q9 + 9 * (size_t)e;
    double b[3][3];
    for (int r = 0; r < 3; r++)
        for (int k = r; k < 3; k++)
            b[r][k] = (double)q[r] * (double)q[k] + (double)q[3 + r] * (double)q[3 + k] + (double)q[6 + r] * (double)q[6 + k];
    const float rv = irv[e];
    const bool ok = fabsf(rv) < INFINITY && rv == rv;   // a zero-volume rest tet is a no-op, like the reference's C == 0 / w == 0 exits
    A[i] = ok ? make_float4((float)b[0][0], (float)b[0][1], (float)b[0][2], (float)b[1][1]) : make_float4(0.f, 0.f, 0.f, 0.f);
    B[i] = make_float4(ok ? (float)b[1][2] : 0.f, ok ? (float)b[2][2] : 0.f, ok ? rv / 6.0f : 0.f, __int_as_float(e));
}
void launch_build_stream_metric(cudaStream_t s, int count, const int *order, const float *Q9, const float *irv, float4 *A, float4 *B) {
    if (count > 0) k_build_stream_metric<<<cdiv(count, 256), 256, 0, s>>>(count, order, Q9, irv, A, B);
}

// =================================================================================================
// Tiled polar-decomposition shape matching (SoftBodyGPU, src/SoftbodyGPU.js:59-376) -- the throughput form of BASELINE
// config 2.  Same tiling as the Neo-Hookean tile kernel (ClusterPlan: Hilbert-ordered tiles, vertex tile staged in shared
// memory, grouped-row corner buffer); per tet it streams what the reference keeps in its `elems` / `quats` render targets:
//   tile block, T * 80 B, plane-major:  R0 R1 R2 (goal corners, 12 floats, read AND written back -- the reference rotates
//   its rest tet incrementally, :253-262), Qt (quaternion, read + written), C (4 vertex slots + 4 corner destinations);
//   plus one float per record, vol[] = rest volume V (negative = "drop corner 0", the reference's table quirk :568;
//   0 = unused record slot) -- 84 B read + 64 B written per tet, the 148 B of the algorithmic figure.
// K3 + K4 run per tet in registers (polar_solve), every corner's goal * V and V go to the corner buffer (STS.128, .w = V --
// the reference's vec4(goal, V), :259-262), per-tile-vertex sums leave as one float4 partial (sum goal V, sum V).  K5's
// per-particle gather over <= 36 scattered texels becomes a sum of ~2.4 tile partials in k_polar_vertex_tiles, so the
// 64 B / tet `elems` state is never re-read through a pointer chase.  Algorithmic bytes:
// 148 B / tet + 32 B / vertex per substep.
// =================================================================================================
template <int T>
__global__ void __launch_bounds__(T) k_polar_tiles(PolarTileArgs a) {
    extern __shared__ __align__(128) unsigned char psm[];
    const int tile = blockIdx.x, tid = threadIdx.x;
    float4 *sx = reinterpret_cast<float4 *>(psm);                                   // [maxTileVertsPad]
    unsigned char *sdx = psm + (size_t)a.maxTileVertsPad * 16;                      // [(maxTileEntries + 1) * 16]
    unsigned char *sm = sdx + (((size_t)a.maxTileEntries + 1) * 16 + 127) / 128 * 128;  // metadata block
    const uint4 *mg = reinterpret_cast<const uint4 *>(a.meta + (size_t)a.metaOff[tile] * 16);
    const int mlen = (int)(a.metaOff[tile + 1] - a.metaOff[tile]);
    for (int i = tid; i < mlen; i += T) reinterpret_cast<uint4 *>(sm)[i] = __ldg(mg + i);
    // this tet's record (coalesced LDG.128 per plane), in flight while the vertex tile is gathered
    unsigned char *tb = a.tets + (size_t)tile * T * 80;
    const float4 r0 = ldg_stream4(tb + tid * 16), r1 = ldg_stream4(tb + T * 16 + tid * 16), r2 = ldg_stream4(tb + T * 32 + tid * 16);
    const float4 qt = ldg_stream4(tb + T * 48 + tid * 16), cc = ldg_stream4(tb + T * 64 + tid * 16);
    const float vol = __ldg(a.vol + (size_t)tile * T + tid);
    __syncthreads();
    const int v0 = reinterpret_cast<const int *>(sm)[0], nl = reinterpret_cast<const int *>(sm)[1];
    const int *ids = reinterpret_cast<const int *>(sm + reinterpret_cast<const int *>(sm)[3]);
    for (int j = tid; j < nl; j += T) sx[j] = a.x4[ids[j]];
    __syncthreads();
    if (vol != 0.0f) {
        const unsigned s01 = __float_as_uint(cc.x), s23 = __float_as_uint(cc.y), d01 = __float_as_uint(cc.z), d23 = __float_as_uint(cc.w);
        const unsigned char *sxb = reinterpret_cast<const unsigned char *>(sx);
        const float4 p0 = *reinterpret_cast<const float4 *>(sxb + (s01 & 0xffffu)), p1 = *reinterpret_cast<const float4 *>(sxb + (s01 >> 16));
        const float4 p2 = *reinterpret_cast<const float4 *>(sxb + (s23 & 0xffffu)), p3 = *reinterpret_cast<const float4 *>(sxb + (s23 >> 16));
        const V3 cur[4] = {{p0.x, p0.y, p0.z}, {p1.x, p1.y, p1.z}, {p2.x, p2.y, p2.z}, {p3.x, p3.y, p3.z}};
        V3 last[4] = {{r0.x, r0.y, r0.z}, {r0.w, r1.x, r1.y}, {r1.z, r1.w, r2.x}, {r2.y, r2.z, r2.w}};
        Q4 q = {qt.x, qt.y, qt.z, qt.w};
        polar_solve<false, true>(cur, last, q, a.noiseK2);
        reinterpret_cast<float4 *>(tb)[tid] = make_float4(last[0].x, last[0].y, last[0].z, last[1].x);
        reinterpret_cast<float4 *>(tb + T * 16)[tid] = make_float4(last[1].y, last[1].z, last[2].x, last[2].y);
        reinterpret_cast<float4 *>(tb + T * 32)[tid] = make_float4(last[2].z, last[3].x, last[3].y, last[3].z);
        reinterpret_cast<float4 *>(tb + T * 48)[tid] = make_float4(q.x, q.y, q.z, q.w);
        const float V = fabsf(vol), V0 = vol < 0.0f ? 0.0f : V;   // corner 0 of tet 0 is dropped from its particle's average (:568)
        *reinterpret_cast<float4 *>(sdx + (d01 & 0xffffu)) = make_float4(last[0].x * V0, last[0].y * V0, last[0].z * V0, V0);
        *reinterpret_cast<float4 *>(sdx + (d01 >> 16)) = make_float4(last[1].x * V, last[1].y * V, last[1].z * V, V);
        *reinterpret_cast<float4 *>(sdx + (d23 & 0xffffu)) = make_float4(last[2].x * V, last[2].y * V, last[2].z * V, V);
        *reinterpret_cast<float4 *>(sdx + (d23 >> 16)) = make_float4(last[3].x * V, last[3].y * V, last[3].z * V, V);
    }
    __syncthreads();
    const uint16_t *gbase = reinterpret_cast<const uint16_t *>(sm + 16);
    for (int j = tid; j < nl; j += T) {
        const int val = sm[a.metaValOff + j];
        const unsigned char *p = sdx + gbase[j >> 3] + ((j & 7) << 4);
        float ax = 0.0f, ay = 0.0f, az = 0.0f, aw = 0.0f;
        for (int i = 0; i < val; i++, p += 144) {
            const float4 d = *reinterpret_cast<const float4 *>(p);
            ax += d.x; ay += d.y; az += d.z; aw += d.w;
        }
        a.part[v0 + j] = make_float4(ax, ay, az, aw);
    }
}

size_t polar_tiles_smem(const PolarTileArgs &a) {
    return (size_t)a.maxTileVertsPad * 16 + (((size_t)a.maxTileEntries + 1) * 16 + 127) / 128 * 128 + (size_t)a.metaStride;
}
void launch_polar_tiles(cudaStream_t s, int clusterSize, const PolarTileArgs &a) {
    if (a.numTiles <= 0) return;
    const size_t smem = polar_tiles_smem(a);
    std::lock_guard<std::mutex> lk(g_launchMu);
#define PT_CASE(T_)                                                                                                   \
    case T_: {                                                                                                        \
        static size_t set[64];                                                                                        \
        const int dev = current_device();                                                                             \
        if (set[dev] != smem) { cudaFuncSetAttribute(k_polar_tiles<T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set[dev] = smem; } \
        k_polar_tiles<T_><<<a.numTiles, T_, smem, s>>>(a);                                                            \
        break;                                                                                                        \
    }
    switch (clusterSize) { PT_CASE(128) PT_CASE(256) PT_CASE(512) default: break; }
#undef PT_CASE
}

// K5 (average of the tile partials) + K6 (grab, bounds, floor + friction) + K7 (velocity, late gravity) per particle,
// src/SoftbodyGPU.js:272-376; mode 2 also runs K1 + K2 of the NEXT substep (prev = pos; pos += vel * dt), the velocity
// then never leaves the registers.
template <int MODE>
__global__ void k_polar_vertex_tiles(int N, float4 *__restrict__ x4, float4 *__restrict__ prev4, float4 *__restrict__ vel4,
                                     const int *__restrict__ vpStart, const int *__restrict__ vpSlot,
                                     const float4 *__restrict__ part, const int *__restrict__ vertId,
                                     const SubstepParams *__restrict__ sp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f, sw = 0.0f;
    for (int j = vpStart[i]; j < vpStart[i + 1]; j++) {
        const float4 s = ldg4(part + vpSlot[j]);
        sx += s.x; sy += s.y; sz += s.z; sw += s.w;
    }
    float4 x = x4[i];
    const float4 p = prev4[i];
    const float inv = 1.0f / sw;   // a particle no tet references gets 0 / 0 = NaN, like the shader's unused texels
    x.x = sx * inv; x.y = sy * inv; x.z = sz * inv;
    if ((vertId ? vertId[i] : i) == sp->grabId) { x.x = sp->grabF[0]; x.y = sp->grabF[1]; x.z = sp->grabF[2]; }
    x.x = fminf(fmaxf(x.x, sp->loF[0]), sp->hiF[0]);
    x.y = fminf(fmaxf(x.y, sp->loF[1]), sp->hiF[1]);
    x.z = fminf(fmaxf(x.z, sp->loF[2]), sp->hiF[2]);
    if (x.y < 0.0f) {
        x.y = 0.0f;
        const float fr = fminf(1.0f, sp->dtF * sp->frictionF);
        x.x = fmaf(p.x - x.x, fr, x.x);
        x.z = fmaf(p.z - x.z, fr, x.z);
    }
    const float dt = sp->dtF, idt = sp->invDt;
    float4 v = make_float4((x.x - p.x) * idt, fmaf(sp->gravityF, dt, (x.y - p.y) * idt), (x.z - p.z) * idt, 0.0f);
    if (MODE == 2) {
        prev4[i] = x;
        x.x = fmaf(v.x, dt, x.x); x.y = fmaf(v.y, dt, x.y); x.z = fmaf(v.z, dt, x.z);
    } else {
        vel4[i] = v;
    }
    x4[i] = x;
}
void launch_polar_vertex_tiles(cudaStream_t s, int N, int mode, float4 *x4, float4 *prev4, float4 *vel4, const int *vpStart,
                               const int *vpSlot, const float4 *part, const int *vertId, const SubstepParams *sp) {
    if (N <= 0) return;
    if (mode == 2) k_polar_vertex_tiles<2><<<cdiv(N, 256), 256, 0, s>>>(N, x4, prev4, vel4, vpStart, vpSlot, part, vertId, sp);
    else k_polar_vertex_tiles<1><<<cdiv(N, 256), 256, 0, s>>>(N, x4, prev4, vel4, vpStart, vpSlot, part, vertId, sp);
}

// Tile blocks of the polar solver from the caller-order rest data: record r = tet order[r] (or an unused slot).
template <int T>
__global__ void k_build_polar_tiles(int numRecords, const int *__restrict__ order, const float4 *__restrict__ x4, const int4 *__restrict__ ids,
                                    const float *__restrict__ irv, const uint4 *__restrict__ aux, int dropTet0Corner0,
                                    unsigned char *__restrict__ tets, float *__restrict__ vol) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= numRecords) return;
    const int tile = r / T, t = r % T;
    unsigned char *tb = tets + (size_t)tile * T * 80;
    const int e = order[r];
    float4 R0 = make_float4(0.f, 0.f, 0.f, 0.f), R1 = R0, R2 = R0, Q = make_float4(0.f, 0.f, 0.f, 1.f);
    float Vs = 0.0f;
    if (e >= 0) {
        const int4 id = ids[e];
        const float4 a0 = x4[id.x], a1 = x4[id.y], a2 = x4[id.z], a3 = x4[id.w];
        R0 = make_float4(a0.x, a0.y, a0.z, a1.x); R1 = make_float4(a1.y, a1.z, a2.x, a2.y); R2 = make_float4(a2.z, a3.x, a3.y, a3.z);
        const float V = __fdiv_rn(1.0f, irv[e]);   // the shader's 1.0 / invRestVolume, :220
        Vs = (dropTet0Corner0 && e == 0) ? -V : V;
    }
    reinterpret_cast<float4 *>(tb)[t] = R0;
    reinterpret_cast<float4 *>(tb + T * 16)[t] = R1;
    reinterpret_cast<float4 *>(tb + T * 32)[t] = R2;
    reinterpret_cast<float4 *>(tb + T * 48)[t] = Q;
    reinterpret_cast<uint4 *>(tb + T * 64)[t] = aux[r];
    vol[r] = Vs;
}
void launch_build_polar_tiles(cudaStream_t s, int clusterSize, int numRecords, const int *order, const float4 *x4, const int4 *ids,
                              const float *irv, const uint4 *aux, int dropTet0Corner0, unsigned char *tets, float *vol) {
    if (numRecords <= 0) return;
    const int g = cdiv(numRecords, 256);
    if (clusterSize == 128) k_build_polar_tiles<128><<<g, 256, 0, s>>>(numRecords, order, x4, ids, irv, aux, dropTet0Corner0, tets, vol);
    else if (clusterSize == 256) k_build_polar_tiles<256><<<g, 256, 0, s>>>(numRecords, order, x4, ids, irv, aux, dropTet0Corner0, tets, vol);
    else k_build_polar_tiles<512><<<g, 256, 0, s>>>(numRecords, order, x4, ids, irv, aux, dropTet0Corner0, tets, vol);
}

// =================================================================================================
// Utility kernels
// =================================================================================================
__global__ void k_pack3(int N, const float4 *__restrict__ src, const int *__restrict__ perm, float *__restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (perm && perm[i] < 0) return;  // replica this rank does not maintain
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
    if (perm && perm[i] < 0) return;
    size_t o = 3 * (size_t)(perm ? perm[i] : i);
    float w = keepW ? dst[i].w : 0.0f;
    dst[i] = make_float4(src[o], src[o + 1], src[o + 2], w);
}
void launch_unpack3(cudaStream_t s, int N, const float *src3, const int *perm, float4 *dst, int keepW) {
    if (N > 0) k_unpack3<<<cdiv(N, 256), 256, 0, s>>>(N, src3, perm, dst, keepW);
}
// tetsim_set_state: the arrays the caller passed (mask bit 0 pos, 1 prevPos, 2 vel) sit in ONE staging buffer, array k at
// stage + k * stride; one launch unpacks them all into the 16-byte records (pos keeps its invMass lane).
__global__ void k_unpack_state(int N, size_t stride, const float *__restrict__ stage, const int *__restrict__ perm,
                               float4 *__restrict__ x4, float4 *__restrict__ prev4, float4 *__restrict__ vel4, int mask) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (perm && perm[i] < 0) return;  // replica this rank does not maintain
    const size_t o = 3 * (size_t)(perm ? perm[i] : i);
    if (mask & 1) { const float *p = stage + o; x4[i] = make_float4(p[0], p[1], p[2], x4[i].w); }
    if (mask & 2) { const float *p = stage + stride + o; prev4[i] = make_float4(p[0], p[1], p[2], 0.0f); }
    if (mask & 4) { const float *p = stage + 2 * stride + o; vel4[i] = make_float4(p[0], p[1], p[2], 0.0f); }
}
void launch_unpack_state(cudaStream_t s, int N, size_t stride, const float *stage, const int *perm, float4 *x4,
                         float4 *prev4, float4 *vel4, int mask) {
    if (N > 0 && mask) k_unpack_state<<<cdiv(N, 256), 256, 0, s>>>(N, stride, stage, perm, x4, prev4, vel4, mask);
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
__global__ void k_nearest_pass1(int N, const float4 *__restrict__ x4, const int *__restrict__ vertId,
                                const double *__restrict__ p, unsigned long long *__restrict__ best) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (vertId && vertId[i] < 0) return;  // replica this rank does not maintain
    double d2 = grab_d2(x4[i], p);
    if (d2 < 1.7976931348623157e308) atomicMin(best, (unsigned long long)__double_as_longlong(d2));
}
__global__ void k_nearest_pass2(int N, const float4 *__restrict__ x4, const int *__restrict__ vertId,
                                const double *__restrict__ p, const unsigned long long *__restrict__ best,
                                int *__restrict__ outId) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (vertId && vertId[i] < 0) return;
    double d2 = grab_d2(x4[i], p);
    if (d2 < 1.7976931348623157e308 && (unsigned long long)__double_as_longlong(d2) == *best)
        atomicMin(outId, vertId ? vertId[i] : i);
}
void launch_nearest_vertex(cudaStream_t s, int N, const float4 *x4, const int *vertId, const double *p3, int *outId,
                           unsigned long long *scratch) {
    cudaMemsetAsync(scratch, 0xff, sizeof(unsigned long long), s);
    cudaMemsetAsync(outId, 0x7f, sizeof(int), s);  // 0x7f7f7f7f: larger than any vertex id
    if (N <= 0) return;
    k_nearest_pass1<<<cdiv(N, 256), 256, 0, s>>>(N, x4, vertId, p3, scratch);
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
