// launch.h -- host-callable launchers for the kernels in kernels.cuh.  Each arithmetic flavour
// lives in its own translation unit (kernels_exact.cu is built with -fmad=false) and exports one
// table of launchers; the host code picks a table once at create time.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsim {

struct SubstepParams;
// One connected component ("body") of the mesh for the one-CTA-per-body Gauss-Seidel kernel.
struct BodyDesc {
    int vertBegin, vertEnd;    // body's vertex range (internal numbering)
    int tetBegin;              // first record of the body in the level-sorted stream
    int levelBegin, levelEnd;  // range into levelStart[] (record offsets into the stream)
};

struct KernelTable {
    void (*init_tets)(cudaStream_t, int M, const float4 *x4, const int4 *ids, double density, float *Q9, float *irv,
                      double *pm);
    void (*init_mass)(cudaStream_t, int N, const int *cStart, const int *cEnt, const double *pm, float4 *x4);
    void (*predict)(cudaStream_t, int N, float4 *x4, float4 *prev4, float4 *vel4, const SubstepParams *sp);
    void (*post)(cudaStream_t, int N, float4 *x4, const float4 *prev4, float4 *vel4, const SubstepParams *sp,
                 const int *vertId);
    void (*gs_level)(cudaStream_t, int begin, int end, float4 *x4, const int4 *I, const float4 *A, const float4 *B,
                     const float4 *C, const int *order, double *volTerm, const SubstepParams *sp);
    void (*gs_body)(cudaStream_t, int numBodies, int threads, size_t smemBytes, const BodyDesc *bodies,
                    const int *levelStart, float4 *x4, float4 *prev4, float4 *vel4, const int4 *I, const float4 *A,
                    const float4 *B, const float4 *C, const int *order, double *volTerm, const SubstepParams *sp,
                    const int *vertId);
    int (*gs_body_max_smem)();
    void (*jacobi_tet)(cudaStream_t, int M, const float4 *x4, const int4 *I, const float4 *A, const float4 *B,
                       const float4 *C, float4 *dx, double *volTerm, const SubstepParams *sp);
    void (*jacobi_gather)(cudaStream_t, int N, float4 *x4, const int *cStart, const int *cEnt, const float4 *dx);
    void (*polar_integrate)(cudaStream_t, int N, float4 *x4, float4 *prev4, const float4 *vel4,
                            const SubstepParams *sp);
    void (*polar_tet)(cudaStream_t, int M, const float4 *x4, const int4 *I, float4 *rest, float4 *quat);
    void (*polar_vertex)(cudaStream_t, int N, float4 *x4, const float4 *prev4, float4 *vel4, const int *tStart,
                         const int *tEnt, const float4 *rest, const SubstepParams *sp);
    void (*skin)(cudaStream_t, int nVis, const float4 *vis, const int4 *ids, const float4 *x4, float *out);
    void (*normals)(cudaStream_t, int nVis, const float *pos, const int *tri, const int *vtStart, const int *vtEnt,
                    float *nrm);
    void (*skin_polar)(cudaStream_t, int nVis, const float4 *vis, const int4 *ids, const float4 *x4, const float4 *quat,
                       const unsigned char *tileTets, const int *tetRecord, int T, const float *restNrm, float *outPos, float *outNrm);
};

const KernelTable *exact_kernels();
const KernelTable *fast_kernels();

// ---- FAST-only throughput path: clustered Jacobi (kernels_fast.cu) ----
// Tile-major HBM layout consumed by the persistent tile kernel.  Per tile of T tets:
//   tet block  (T * 56 B, one bulk async copy): planes A[T] float4 (Q0..Q3), B[T] float4 (Q4..Q7),
//              C[T] float4 (Q8, invRestVolume, slot01, slot23), D[T] uint2 (dest01, dest23)
//   meta block (variable size, one bulk async copy): see ClusterPlan::tileMeta in mesh_prep.h
struct PeerArgs;
struct TileArgs {
    const float4 *x4;               // handle-local vertex records (x, y, z, invMass)
    const unsigned char *tets;      // [numTiles * T * 56]
    const unsigned char *meta;      // variable-size blocks, block c at 16 * metaOff[c]
    const uint32_t *metaOff;        // [numTiles + 1]
    int tileBegin, numTiles;        // this launch walks tiles [tileBegin, numTiles)
    int metaStride, metaValOff, colStride, maxTileVertsPad;  // metaStride = largest block
    int maxTileEntries;             // sixteen-byte entries of the largest tile's corner buffer (ClusterPlan::maxTileEntries)
    float4 *part;                   // per-tile partial dx sums (deterministic flush)
    float4 *acc;                    // global accumulator (atomic flush), or NULL
    double *volAcc;                 // sum over tets of det F - 1, or NULL
    const SubstepParams *sp;
    int debugSkip;                  // measurement only (tetsim_time_kernel + TETSIM_TILE_DEBUG): 1 no vertex phase, 2 no math, 4 no gather
    const PeerArgs *px;             // DEVICE copy of the peer-exchange arguments: fused push of the boundary sums, or NULL
    int pxSlots;                    // fused push: partial slots [0, pxSlots) belong to boundary tiles (by value: no load on the way)
};
void launch_jacobi_tiles(cudaStream_t, int clusterSize, const TileArgs &a);
bool jacobi_tiles_has_peer_push(int clusterSize);  // can this tile size (and the TETSIM_TILE_* overrides in force) push?
size_t jacobi_tiles_smem(int clusterSize, const TileArgs &a);
// record r of the tile stream: Q/irv of tet order[r] (or zeros when order[r] < 0) + aux[r] -> tet blocks
void launch_build_tiles(cudaStream_t, int clusterSize, int numRecords, const int *order, const float *Q9,
                        const float *irv, const uint4 *aux, unsigned char *tets);

struct ApplyArgs {
    float4 *x4, *prev4, *vel4;
    const int *vpStart, *vpSlot;    // vertex -> partial-sum slots (CSR into part[])
    const float4 *part;
    float4 *acc;                    // atomic-flush accumulator (read and re-zeroed) or NULL
    const float *invVal;            // 1 / valence
    const float4 *bsum;             // all-reduced boundary sums for vertices >= boundaryBegin, or NULL
    const PeerArgs *px;             // DEVICE copy: wait for the sharers' flags and reduce in rank order here (fused), or NULL
    int boundaryBegin;
    const SubstepParams *sp;
    const int *vertId;              // handle-local -> caller's vertex id (for the grab test)
};
// mode: 0 = apply only, 1 = apply + post, 2 = apply + post + predict of the next substep
void launch_jacobi_apply(cudaStream_t, int begin, int end, int mode, const ApplyArgs &a);
// boundary vertices: bsum[b] = sum of this rank's partials (input of the all-reduce)
void launch_boundary_pack(cudaStream_t, int boundaryBegin, int numBoundary, const int *vpStart, const int *vpSlot,
                          const float4 *part, float4 *bsum);

// neighbour exchange: send[i] = bsum[sendIdx[i]];  bsum[b] = sum over srcs (own bsum / recv entries) in rank order
void launch_halo_pack(cudaStream_t, int n, const int *sendIdx, const float4 *bsum, float4 *send);
void launch_halo_reduce(cudaStream_t, int numBoundary, const int *srcStart, const int *src, const float4 *recv,
                        float4 *bsum);

// Peer-memory exchange (TetSimOptions.exchange = 2): no NCCL, no side stream.  Every rank owns one exchange
// allocation, mapped into its sharers with cudaIpc:
//   [0, 256)  uint32 flag[64]   flag[s] = last epoch whose entries the s-th peer (ascending rank) has delivered
//   [256, ..) uint32 ctl[]      ctl[0] epoch of this rank, ctl[1] block ticket, ctl[2] error (wait timed out)
//   [1024, ..) float4 recv[2][total]  receive buffer, double-buffered by epoch parity
// k_peer_push sums this rank's partials of every active boundary vertex, keeps the sum in bsum[] and STORES it over
// NVLink into the receive buffer of every sharer; the last block to finish publishes the epoch into the sharers'
// flags (release at system scope).  k_peer_reduce (after the interior tiles) waits for the sharers' flags (acquire)
// and adds the contributions of all sharers in ascending rank order -- identical on every sharer.
constexpr int kPeerFlagBytes = 256, kPeerCtlOff = 256, kPeerRecvOff = 1024, kPeerMaxPeers = 64;
struct PeerArgs {
    int numBoundary, boundaryBegin, numPeers;
    const int *vpStart, *vpSlot;          // boundary vertex -> partial-sum slots (deterministic flush)
    const float4 *part;
    float4 *acc;                          // atomic-flush accumulator (read and re-zeroed) or NULL
    float4 *bsum;                         // [numBoundary] own sums in, reduced sums out
    const int *pxStart, *pxPeer, *pxEntry;  // push destinations (ClusterPlan)
    unsigned char *const *peerBase;       // [numPeers] mapped exchange allocation of each peer
    const int *remoteTotal, *remoteSlot;  // [numPeers] parity stride of the peer's buffer; my flag index there
    unsigned char *self;                  // this rank's exchange allocation
    int selfTotal;                        // parity stride of my receive buffer
    const int *srcStart, *src;            // reduce sources (ClusterPlan hxSrcStart / hxSrc)
    unsigned long long timeoutNs;         // give up waiting after this long (sets ctl[2])
    // fused push: everything a push needs in ONE 32-byte record per boundary-tile partial slot (two parallel 16-byte loads):
    //   rec[2 s]     = {i | (count - 1) << 8 | sharers << 16 (or ~0: not shared), boundary id, address of sharer 0's entry (lo, hi)}
    //   rec[2 s + 1] = {parity stride of sharer 0, of sharer 1 (units of 32 B), address of sharer 1's entry (lo, hi)}
    // entry addresses are those of the parity-0 half; a third and further sharer (partition corners) come from the CSR.
    const uint4 *pushRec;
};
// Fused form (deterministic flush; the default of exchange = 2): the tile kernel itself pushes, with NO synchronisation
// at all on the sending side.  A thread that has just formed a tile's partial sum of a rank-shared vertex stores it, next
// to its normal place in part[], straight into the receive buffer of every sharer as one self-validating 32-byte entry
//   uint4 {x, tag, y, tag}, uint4 {z, tag, 0, tag}     tag = epoch * 16 + (number of partials of this vertex - 1)
// (every 8-byte half carries the tag, the granularity NVLink stores are not torn at -- the scheme of NCCL's LL
// protocol), slot (entry * kPeerK + partial index) of the parity half of the buffer.  Boundary tiles run first, so the
// partials cross NVLink while the interior tiles are still being solved, inside ONE launch and without fences, flags
// or tickets.
// The last CTA of the tile kernel advances the epoch (ticket in ctl[1]); the vertex kernel reads it, schedules the blocks
// holding the rank-shared vertices FIRST, polls their entries, adds each sharer's partials in that sharer's own order and
// the sharers' sums in ascending rank order (so every sharer computes the identical value): 2 launches per iteration, as
// on a single GPU.  Measured at 2 ranks on the 10M-tet beam (profiles/r2_peer_experiments.txt): epoch advanced by the
// vertex kernel's 3,400 blocks 85.9 G tet/s -> by the tile kernel's 592 CTAs 94.0; boundary blocks first 90.1; both +
// one-record pushes 99.1 (NCCL all-reduce 95.1, NCCL neighbour exchange 87.3).
constexpr int kPeerK = 16;                // most tile partials a rank may hold for one shared vertex (checked at create)
void launch_peer_push(cudaStream_t, const PeerArgs &a);
void launch_peer_reduce(cudaStream_t, const PeerArgs &a);

// ---- FAST-only: Gauss-Seidel in the reference order with four lanes per tet (kernels_fast.cu) ----
void launch_gs_body_quads(cudaStream_t, int numBodies, int threads, size_t smemBytes, const BodyDesc *bodies, const int *levelStart,
                          float4 *x4, float4 *prev4, float4 *vel4, const int4 *I, const float4 *A, const float4 *B,
                          const int *order, double *volTerm, const SubstepParams *sp, const int *vertId);
void launch_build_stream_metric(cudaStream_t, int count, const int *order, const float *Q9, const float *irv, float4 *A, float4 *B);

// ---- FAST-only throughput path of the polar solver: tiled shape matching (kernels_fast.cu) ----
struct PolarTileArgs {
    const float4 *x4;
    unsigned char *tets;            // [numTiles * T * 80] R0 R1 R2 Qt C planes per tile (read and written)
    const float *vol;               // [numTiles * T] rest volume per record (negative: drop corner 0; 0: unused slot)
    const unsigned char *meta;      // the ClusterPlan's per-tile metadata blocks
    const uint32_t *metaOff;
    int numTiles, metaStride, metaValOff, maxTileVertsPad, maxTileEntries;
    float4 *part;                   // per tile vertex: (sum goal * V, sum V)
    float noiseK2;                  // (k 2^-24)^2: the rotation extraction's mesh-scaled exit (device_math.cuh); 0 = fixed floor only
};
void launch_polar_tiles(cudaStream_t, int clusterSize, const PolarTileArgs &a);
size_t polar_tiles_smem(const PolarTileArgs &a);
// mode 1: K5 + K6 + K7; mode 2: + K1 + K2 of the next substep
void launch_polar_vertex_tiles(cudaStream_t, int N, int mode, float4 *x4, float4 *prev4, float4 *vel4, const int *vpStart,
                               const int *vpSlot, const float4 *part, const int *vertId, const SubstepParams *sp);
void launch_build_polar_tiles(cudaStream_t, int clusterSize, int numRecords, const int *order, const float4 *x4, const int4 *ids,
                              const float *irv, const uint4 *aux, int dropTet0Corner0, unsigned char *tets, float *vol);

// ---- utility kernels (kernels_fast.cu) ----
void launch_pack3(cudaStream_t, int N, const float4 *src, const int *perm, float *dst3);       // dst[perm[i]] = src[i].xyz
void launch_unpack3(cudaStream_t, int N, const float *src3, const int *perm, float4 *dst, int keepW);
// array k (0 pos, 1 prevPos, 2 vel; mask bit k = present) at stage + k * stride -> the 16-byte state records, one launch
void launch_unpack_state(cudaStream_t, int N, size_t stride, const float *stage, const int *perm, float4 *x4,
                         float4 *prev4, float4 *vel4, int mask);
void launch_sum_sequential(cudaStream_t, int M, const double *terms, double *out);             // out = (sum in index order) / M
void launch_nearest_vertex(cudaStream_t, int N, const float4 *x4, const int *vertId, const double *p3, int *outId,
                           unsigned long long *scratch);
void launch_fill_rest(cudaStream_t, int M, const float4 *x4, const int4 *I, const float *irv, float4 *rest,
                      float4 *quat);
void launch_build_stream(cudaStream_t, int count, const int *order, const float *Q9, const float *irv, float4 *A,
                         float4 *B, float4 *C);

}  // namespace tsim
