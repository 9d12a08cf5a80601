/*
 * tetsim_b200.h -- C ABI of libtetsim_b200.so: a B200 (sm_100a) XPBD tetrahedral-FEM substep solver
 * that stands in for the solver classes of zalo/TetSim.
 *
 * The reference has no FFI layer; its boundary is two duck-typed JavaScript classes,
 *   SoftBody     src/Softbody.js:3-298      (CPU, Neo-Hookean XPBD, Gauss-Seidel)
 *   SoftBodyGPU  src/SoftbodyGPU.js:4-712   (WebGL, polar-decomposition shape matching, Jacobi)
 * driven by Main.update (src/main.js:74-96).  Each entry point below names the reference member
 * it replaces.  An N-API shim (tetsim_b200/js/) maps the JS classes onto these
 * calls one to one; in this repository the same calls are bound with ctypes (tetsim_b200/_capi.py).
 *
 * Conventions
 *   - every function returns 0 on success or a negative TETSIM_E_* code; tetsim_last_error()
 *     then returns a thread-local message (the reference's init() "error string or null"
 *     convention, src/MultiTargetGPUComputationRenderer.js:178-190, surfaced as a code + text);
 *   - all input arrays are COPIED at create (the reference keeps references to tetIds/visVerts,
 *     src/Softbody.js:19,46 -- copying is stricter, never looser);
 *   - numbers that are JS `number`s in the reference (dt, physicsParams fields, grab position)
 *     are `double` here, typed-array contents are `float`/`int32_t`;
 *   - a handle is used from one host thread at a time (the reference is single-threaded,
 *     src/World.js:73); simulate/step enqueue on the handle's CUDA stream and return, the get_*
 *     calls synchronise that stream;
 *   - there is NO CPU fallback: every call fails with TETSIM_E_CUDA when no sm_100 device is usable.
 */
#ifndef TETSIM_B200_H
#define TETSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TETSIM_VERSION 100

enum {
    TETSIM_OK = 0,
    TETSIM_E_INVALID = -1, /* bad argument / malformed mesh            */
    TETSIM_E_CUDA = -2,    /* CUDA runtime error or no usable device   */
    TETSIM_E_NCCL = -3,    /* NCCL could not be loaded or failed       */
    TETSIM_E_STATE = -4,   /* call not valid for this handle's solver  */
    TETSIM_E_NOMEM = -5
};

/* physicsParams of the reference, src/main.js:22-36.  Read on EVERY simulate/step call because the
 * GUI mutates it between frames (src/main.js:37-42).  density is consumed at create only
 * (src/Softbody.js:32).  devCompliance/volCompliance are read per call here; the reference reads
 * them through the object captured at construction (src/Softbody.js:130,161,165), which is the
 * same object main.js passes to simulate(), so the two coincide. */
typedef struct TetSimParams {
    double gravity;        /* -9.81   */
    double friction;       /* 1000.0  */
    double density;        /* 1000.0  */
    double devCompliance;  /* 1/100000 */
    double volCompliance;  /* 0.0     */
    double worldBounds[6]; /* lo.xyz, hi.xyz = -2.5,-1,-2.5, 2.5,10,2.5 */
} TetSimParams;

/* Which substep algorithm the handle runs. */
enum TetSimSolver {
    /* SoftBody.simulate exactly: Gauss-Seidel in tet-index order, executed as the order-preserving
     * dependency-level schedule (bit-identical to the sequential sweep, src/Softbody.js:206-209). */
    TETSIM_NH_GS_EXACT = 0,
    /* Gauss-Seidel in greedy graph-colour order ("Graph Coloring", README.md:25 TODO). */
    TETSIM_NH_GS_COLOR = 1,
    /* Jacobi Neo-Hookean: per-tet solveElem on a private copy, dx averaged by valence. */
    TETSIM_NH_JACOBI = 2,
    /* SoftBodyGPU.simulate: polar-decomposition shape matching, volume-weighted Jacobi average
     * (src/SoftbodyGPU.js:59-376). */
    TETSIM_POLAR_JACOBI = 3
};

enum TetSimArithmetic {
    /* f32 with FMA contraction, rsqrt/rcp approximations: the throughput mode. */
    TETSIM_ARITH_FAST_F32 = 0,
    /* The reference's arithmetic, operation for operation: f64 expressions with f32 stores for the
     * Neo-Hookean solvers (JS typed arrays), separately rounded f32 for the polar solver (GLSL
     * highp).  Compiled with -fmad=false; bit-identical to oracle/ for the GS and gather paths. */
    TETSIM_ARITH_BITEXACT = 1
};

typedef struct TetSimOptions {
    int32_t solver;            /* enum TetSimSolver, default TETSIM_NH_GS_EXACT                       */
    int32_t arithmetic;        /* enum TetSimArithmetic, default TETSIM_ARITH_FAST_F32                */
    int32_t iters;             /* Jacobi iterations per substep (>= 1), default 1                     */
    int32_t deterministic;     /* 1 (default): sums in a fixed order, run-to-run bit-reproducible;
                                  0: Jacobi dx flushed with float atomics (order not reproducible)   */
    int32_t referenceTableBug; /* polar: 1 (default) reproduces src/SoftbodyGPU.js:568 (corner 0 of
                                  tet 0 dropped from its vertex's average); 0 averages every corner  */
    int32_t reorder;           /* Jacobi: 1 (default) = sort tets along a Morton curve of their
                                  centroids before clustering; 0 = keep the caller's order           */
    int32_t clusterSize;       /* Jacobi: tets per tile: 32, 64 (one warp per tile) or 128, 256
                                  (default), 512 (one CTA per tile)                                    */
    int32_t trackVolError;     /* -1 auto (on for GS, off for Jacobi), 0 off, 1 on                    */
    int32_t device;            /* CUDA device ordinal, -1 = the calling thread's current device       */
    int32_t rank;              /* multi-GPU Jacobi: this process's rank, 0 when worldSize == 1        */
    int32_t worldSize;         /* multi-GPU Jacobi: number of tet partitions / processes, default 1   */
    int32_t exchange;          /* multi-GPU: 0 = ncclAllReduce of the boundary dx over all ranks (default),
                                  1 = neighbour exchange: grouped ncclSend/ncclRecv with the ranks that share
                                  vertices with this one, sharers' sums added in ascending rank order,
                                  2 = peer-memory exchange: the tile kernel stores its partial sums of rank-shared
                                  vertices over NVLink straight into the sharers' receive buffers (cudaIpc mappings,
                                  tetsim_get_ipc_handle / tetsim_set_peers) as self-validating tagged entries, the
                                  vertex kernel polls them; no NCCL, no ncclUniqueId, same rank-ordered sums as 1 */
    void *stream;              /* cudaStream_t to enqueue on; NULL = a stream owned by the handle     */
    const void *ncclUniqueId;  /* 128-byte ncclUniqueId from tetsim_nccl_unique_id (rank 0's), or NULL */
} TetSimOptions;

typedef struct TetSimInfo {
    int32_t numVerts, numTets;
    int32_t solver, arithmetic, iters;
    int32_t numLevels;        /* GS: dependency levels (EXACT) or colours (COLOR)                     */
    int32_t maxLevelSize;
    int32_t numComponents;    /* connected components (independent bodies) found in the mesh          */
    int32_t bodyKernel;       /* GS: 1 = one CTA per component with positions in shared memory        */
    int32_t numClusters;      /* Jacobi: CTA tiles on this rank                                       */
    int32_t clusterSize;
    int32_t localTets;        /* tets solved by this rank                                             */
    int32_t localVerts;       /* vertices resident on this rank                                       */
    int32_t boundaryVerts;    /* vertices shared between ranks (all-reduced each iteration)            */
    int32_t maxValence;
    int32_t launchesPerSubstep; /* kernels of this library launched per substep on this rank          */
    int64_t deviceBytes;      /* device memory held by the handle                                     */
    int64_t sumLocalVerts;    /* Jacobi: sum over clusters of tile vertex counts                      */
    int64_t kernelLaunches;   /* kernels launched by simulate/step since create (graph replays count) */
    int64_t tileMetaBytes;    /* Jacobi: bytes of per-tile metadata streamed per launch               */
    int32_t maxTileVerts;     /* Jacobi: largest tile vertex count                                    */
    int32_t boundaryTiles;    /* Jacobi multi-GPU: tiles touching rank-shared vertices (run first)     */
} TetSimInfo;

typedef struct tetsim tetsim_t;

const char *tetsim_last_error(void);
int tetsim_version(void);
/* Number of CUDA devices this process can use with compute capability 10.x; < 0 on error. */
int tetsim_device_count(void);
void tetsim_default_params(TetSimParams *p);   /* src/main.js:22-36 */
void tetsim_default_options(TetSimOptions *o);

/* new SoftBody(vertices, tetIds, ..., physicsParams, ...) / new SoftBodyGPU(...):
 * src/Softbody.js:4-32 + initPhysics :60-87;  src/SoftbodyGPU.js:5-55 + initPhysics :487-608.
 * verts = 3*numVerts floats, tetIds = 4*numTets ints.  The mesh may hold any number of disconnected
 * bodies (the reference's physicsScene.softBodies[], src/main.js:51,80-84, concatenated). */
int tetsim_create(const float *verts, int32_t numVerts, const int32_t *tetIds, int32_t numTets,
                  const TetSimParams *params, const TetSimOptions *options, tetsim_t **out);
void tetsim_destroy(tetsim_t *h);

/* softBody.simulate(dt, physicsParams): ONE substep.  src/Softbody.js:195-240, src/SoftbodyGPU.js:610-641 */
int tetsim_simulate(tetsim_t *h, double dt, const TetSimParams *params);
/* The substep loop of Main.update, src/main.js:79-84: dt = frameDt / numSubsteps (frameDt =
 * timeScale * timeStep, in double), then numSubsteps x simulate.  One CUDA-graph launch. */
int tetsim_step(tetsim_t *h, double frameDt, int32_t numSubsteps, const TetSimParams *params);
int tetsim_synchronize(tetsim_t *h);

/* Readable state of the reference objects: .pos .prevPos .vel (src/Softbody.js:12-14);
 * SoftBodyGPU.readToCPU(pos) (src/SoftbodyGPU.js:649-653).  out = 3*numVerts floats.
 * On a multi-GPU handle only vertices resident on this rank are written (others are left
 * untouched); tetsim_get_resident marks them. */
int tetsim_get_positions(tetsim_t *h, float *out);
int tetsim_get_prev_positions(tetsim_t *h, float *out);
int tetsim_get_velocities(tetsim_t *h, float *out);
int tetsim_get_resident(tetsim_t *h, uint8_t *outNumVerts);
/* Overwrite state (checkpoint / resume; also the per-frame upload of the end-to-end benchmark).
 * Any pointer may be NULL = leave unchanged.  The copies are ENQUEUED (a dedicated copy stream into a double-buffered
 * staging area, then one unpack kernel on the handle's stream): with pinned host memory the call returns at once and
 * the arrays must stay unchanged until the next tetsim_synchronize / tetsim_get_* on the handle returns. */
int tetsim_set_state(tetsim_t *h, const float *pos, const float *prevPos, const float *vel);
/* Asynchronous form of tetsim_get_positions (SoftBodyGPU.readToCPU without the stall, src/SoftbodyGPU.js:649-653):
 * the positions as of everything enqueued so far are packed and copied to `out` on a second copy stream;
 * tetsim_wait_positions (or tetsim_synchronize) returns when `out` is complete.  What the caller enqueues in between --
 * the next frame's tetsim_set_state / tetsim_step -- overlaps the download (PCIe is full duplex). */
int tetsim_get_positions_async(tetsim_t *h, float *out);
int tetsim_wait_positions(tetsim_t *h);
/* Rank-local state access for multi-GPU hosts: the arrays hold ONLY the vertices resident on this rank, in the
 * handle's own order -- tetsim_get_resident_ids lists the caller's vertex id of each (TetSimInfo.localVerts entries,
 * -1 = a replica this rank does not maintain: written as NaN, ignored on upload).  Each rank moves 1/worldSize of the
 * state instead of all of it.  Valid on single-GPU handles too (localVerts = numVerts, the solver's internal order). */
int tetsim_get_resident_ids(tetsim_t *h, int32_t *outLocalVerts);
int tetsim_set_state_resident(tetsim_t *h, const float *pos, const float *prevPos, const float *vel);
int tetsim_get_positions_resident(tetsim_t *h, float *out);
int tetsim_get_positions_resident_async(tetsim_t *h, float *out);
/* .invRestPose (9 per tet, column-major) .invRestVolume .invMass of initPhysics, src/Softbody.js:60-87.
 * Any pointer may be NULL. */
int tetsim_get_rest(tetsim_t *h, float *invRestPose, float *invRestVolume, float *invMass);
/* .volError after the last substep (src/Softbody.js:163,206-209). */
int tetsim_get_vol_error(tetsim_t *h, double *out);
/* Polar solver state: the per-tet goal corners `elems` (12 floats per tet) and `quats` (xyzw),
 * src/SoftbodyGPU.js:54-55.  Any pointer may be NULL. */
int tetsim_get_polar_state(tetsim_t *h, float *rest12, float *quat4);

/* startGrab / moveGrabbed / endGrab, src/Softbody.js:279-298, src/SoftbodyGPU.js:692-712.
 * startGrab picks the nearest vertex on the device (first strict minimum of the f64 squared
 * distance, like the reference loop) and returns its index through outGrabId (may be NULL). */
int tetsim_start_grab(tetsim_t *h, const double p[3], int32_t *outGrabId);
/* The two halves of startGrab for multi-GPU hosts (tetsim_start_grab fails with TETSIM_E_STATE there: a rank sees only
 * its own vertices).  tetsim_nearest_vertex searches the vertices this rank maintains and returns the caller's id and
 * the f64 squared distance (id -1 / +inf when none); the host takes the minimum over ranks (ties: smallest id) and
 * passes the winner to tetsim_set_grab on EVERY rank, so replicas of a shared vertex are pinned consistently. */
int tetsim_nearest_vertex(tetsim_t *h, const double p[3], int32_t *outId, double *outD2);
int tetsim_set_grab(tetsim_t *h, int32_t grabId, const double p[3]);
int tetsim_move_grabbed(tetsim_t *h, const double p[3]);
int tetsim_end_grab(tetsim_t *h);

/* updateVisMesh, src/Softbody.js:259-277: barycentric skinning of the embedded surface mesh.
 * visVerts = (tetNr, b0, b1, b2) per surface vertex; triIds may be NULL/0 to skip the normals
 * (three@0.160.0 BufferGeometry.computeVertexNormals, three.module.js:11125-11215).
 * The surface mesh is uploaded on the first call and cached; the cache is keyed on the arrays' CONTENT (a 64-bit hash
 * per call), so mutating visVerts / triIds in place, or a new array at the same address, is picked up. */
int tetsim_skin(tetsim_t *h, const float *visVerts, int32_t numVis, const int32_t *triIds, int32_t numTris,
                float *outPos, float *outNormals);

/* The WebGL variant's render-time skinning, i.e. the vertex shader SoftBodyGPU patches into its vis material
 * (src/SoftbodyGPU.js:424-448): position = ((p0 b0 + p1 b1) + p2 b2) + p3 (1 - (b0 + b1 + b2)) in f32, and -- instead of
 * computeVertexNormals every frame -- normal = Rotate(rest normal, quaternion of the surface vertex's tet).
 * restNormals = 3*numVis floats: the normals computeVertexNormals left in the geometry at construction (:484-485 via
 * updateVisMesh :685; tetsim_skin on the initial state returns exactly those); NULL / outNormals NULL skips the normals.
 * POLAR_JACOBI handles only.  The surface mesh is cached by content, like tetsim_skin. */
int tetsim_skin_gpu(tetsim_t *h, const float *visVerts, int32_t numVis, const float *restNormals, float *outPos,
                    float *outNormals);

int tetsim_get_info(tetsim_t *h, TetSimInfo *info);
/* Measurement aid for bench.py's roofline line: launches the dominant kernel of the handle (the
 * clustered Jacobi tile kernel) `reps` times back to back on the handle's stream between two CUDA
 * events and returns the mean duration.  The kernel is idempotent (reads positions, writes the
 * per-tile partial sums), so simulation state is untouched.  algorithmicBytes (may be NULL) receives
 * 56 * localTets + 32 * localVerts, the per-launch figure of BASELINE.md section 2. */
int tetsim_time_kernel(tetsim_t *h, int32_t reps, double *msPerLaunch, int64_t *algorithmicBytes);

/* Multi-GPU plumbing (one process per GPU).  Rank 0 calls tetsim_nccl_unique_id and broadcasts the
 * 128 bytes out of band (torch.distributed in this repo); every rank passes them in TetSimOptions. */
int tetsim_nccl_unique_id(void *out128);
/* Peer-memory exchange (exchange = 2).  After tetsim_create every rank calls tetsim_get_ipc_handle, the ranks
 * all-gather the TETSIM_PEER_BLOB_BYTES-byte blobs by any means (torch.distributed in bench.py), and every rank
 * passes the worldSize blobs, in rank order, to tetsim_set_peers, which maps the exchange buffers of the ranks it
 * shares vertices with.  simulate/step fail with TETSIM_E_STATE until then.  Ranks must issue the same sequence of
 * simulate/step calls (as with NCCL); a rank that waits longer than TETSIM_PEER_TIMEOUT_MS (default 10000) for a
 * sharer gives up and the next tetsim_synchronize / tetsim_get_* reports TETSIM_E_STATE.  All ranks must be idle
 * (synchronized) before any of them destroys its handle.  Handles of one process may also be peers of each other
 * (the blob carries the owner's pointer), which is how the single-process test drives the protocol. */
#define TETSIM_PEER_BLOB_BYTES 128
int tetsim_get_ipc_handle(tetsim_t *h, void *outBlob);
int tetsim_set_peers(tetsim_t *h, const void *blobs /* worldSize * TETSIM_PEER_BLOB_BYTES */);

/* Host-side mesh tools used by tests and the benchmark (the reference has none; README.md:25). */
/* Order-preserving dependency levels of the sequential sweep: level[numTets] out; returns #levels. */
int tetsim_level_schedule(const int32_t *tetIds, int32_t numTets, int32_t numVerts, int32_t *level);
/* Greedy colouring in tet order (smallest colour unused by any tet sharing a vertex). */
int tetsim_greedy_colors(const int32_t *tetIds, int32_t numTets, int32_t numVerts, int32_t *color);

/* Connected components of the mesh over shared vertices = the independent bodies of the scene (the reference's
 * physicsScene.softBodies[], src/main.js:51,80-84, concatenated into one mesh): vertComp[numVerts] receives the body of
 * every vertex, bodies numbered by their first vertex; returns the number of bodies.  Bodies never interact (the reference
 * has no body-body collision), so a scene shards across GPUs by bodies with no exchange at all: tetsim_b200.mesh.shard_bodies. */
int tetsim_connected_components(const int32_t *tetIds, int32_t numTets, int32_t numVerts, int32_t *vertComp);

/* The tet partition a multi-GPU Jacobi handle of (rank, worldSize) would use, without touching a
 * GPU: tets are (optionally Morton-) ordered, cut into tiles of clusterSize and each rank takes a
 * contiguous run of tiles.  counts[0..3] = {localTets, interiorVerts, boundaryVerts, tiles};
 * localToCaller (capacity numVerts, may be NULL) lists resident vertices, interior first then the
 * boundary set, which is identical and identically ordered on every rank; localTets (capacity
 * numTets, may be NULL) lists this rank's caller tet indices in solver order. */
int tetsim_plan_partition(const float *verts, int32_t numVerts, const int32_t *tetIds, int32_t numTets,
                          int32_t clusterSize, int32_t reorder, int32_t rank, int32_t worldSize, int32_t counts[4],
                          int32_t *localToCaller, int32_t *localTets);

/* The neighbour lists the halo / peer-memory exchanges of (rank, worldSize) would use, without touching a GPU:
 * returns the number of peers P (<= capacity, else TETSIM_E_INVALID); peers[P] ascending ranks; segStart[P + 1]
 * offsets of the per-peer segments of this rank's send/receive buffers (entries); remoteOff[P], remoteTotal[P],
 * remoteSlot[P]: where this rank's segment starts inside peer q's receive buffer, that buffer's size, and this
 * rank's index in q's peer list -- what the peer-memory exchange derives locally instead of communicating. */
int tetsim_plan_halo(const float *verts, int32_t numVerts, const int32_t *tetIds, int32_t numTets, int32_t clusterSize,
                     int32_t reorder, int32_t rank, int32_t worldSize, int32_t capacity, int32_t *peers,
                     int32_t *segStart, int32_t *remoteOff, int32_t *remoteTotal, int32_t *remoteSlot);

#ifdef __cplusplus
}
#endif
#endif /* TETSIM_B200_H */
