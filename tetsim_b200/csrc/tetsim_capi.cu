// tetsim_capi.cu -- the C ABI of libtetsim_b200.so (include/tetsim_b200.h): handle, device state,
// solver scheduling, CUDA-graph capture of the substep loop, multi-GPU exchange.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/tetsim_b200.h"
#include "device_math.cuh"
#include "launch.h"
#include "mesh_prep.h"

using namespace tsim;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(TETSIM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                           std::to_string(__LINE__) + ")");                              \
    } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL, loaded lazily (the library must load on hosts without it; a torch process already has
// libnccl.so.2 mapped and dlopen returns that copy).
// ------------------------------------------------------------------------------------------------
namespace {
struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string why;
    bool load() {
        if (lib) return true;
        const char *env = getenv("TETSIM_NCCL_LIB");
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { why = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        Send = (decltype(Send))dlsym(lib, "ncclSend");
        Recv = (decltype(Recv))dlsym(lib, "ncclRecv");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !Send || !Recv || !GroupStart || !GroupEnd || !CommDestroy || !GetErrorString) {
            why = "libnccl is missing a required symbol";
            lib = nullptr;
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;
constexpr int kNcclFloat = 7, kNcclSum = 0;

template <class T>
struct DevBuf {  // owning device array; frees on scope exit so early error returns do not leak
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void **)&p, count * sizeof(T));
    }
    cudaError_t upload(const std::vector<T> &h, cudaStream_t s) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    size_t bytes() const { return n * sizeof(T); }
};
}  // namespace

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct tetsim {
    TetSimOptions opt{};
    TetSimParams params{};
    int N = 0, M = 0;          // caller's mesh
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    const KernelTable *K = nullptr;   // flavour of the substep kernels
    const KernelTable *KX = nullptr;  // exact flavour (init, skinning)

    // vertex state in the handle's internal numbering
    int nInt = 0;
    DevBuf<float4> x4, prev4, vel4;
    DevBuf<int> vertId;               // internal -> caller id; empty when identity
    std::vector<int> h_vertId;

    // caller-order rest data (initPhysics)
    DevBuf<float> Q9, irv, invMass;
    DevBuf<int4> ids;                 // caller tet order; vertex ids in INTERNAL numbering
    DevBuf<int> cStart, cEnt;         // vertex -> corners (internal numbering), when the solver needs it

    // solver streams
    DevBuf<float4> A, B, C;
    DevBuf<int4> I;
    DevBuf<int> order, levelStart;
    std::vector<int> h_levelStart;
    DevBuf<BodyDesc> bodies;
    int numLevels = 0, maxLevelSize = 0, numComponents = 1, numBodies = 0, bodyThreads = 0;
    size_t bodySmem = 0;
    bool bodyKernel = false;
    bool gsQuads = false;                 // FAST GS: k_gs_body_quads (four lanes per tet, rest-metric stream)
    DevBuf<double> volTerm, volOut;
    bool trackVol = false;
    // Jacobi gather path
    DevBuf<float4> dx;
    // Jacobi cluster path
    ClusterPlan plan;
    DevBuf<int> vpStart, vpSlot;
    DevBuf<unsigned char> tileTets, tileMeta;
    DevBuf<float> tileVol;                // tiled polar solver: rest volume per record
    DevBuf<uint32_t> metaOff;
    DevBuf<float4> part, acc, bsum;
    DevBuf<float> invVal;
    bool clustered = false;
    // polar
    bool polarTiled = false;               // FAST arithmetic: the tiled kernels (k_polar_tiles); else the reference's CSR gather order
    DevBuf<float4> rest, quat;
    DevBuf<int> tStart, tEnt;
    // per-launch parameters, staging, grab
    DevBuf<SubstepParams> sp;
    // host <-> device state transfer.  Two copy streams (one per DMA direction) beside the compute stream and two staging
    // buffers per direction: an upload never waits for the previous download (PCIe is full duplex), the three arrays of
    // one tetsim_set_state go out back to back and ONE kernel unpacks them, and a download started with
    // tetsim_get_positions_async overlaps whatever the caller enqueues next.
    cudaStream_t h2dStream = nullptr, d2hStream = nullptr;
    DevBuf<float> stageIn[2], stageOut[2];
    cudaEvent_t evIn[2] = {nullptr, nullptr}, evInFree[2] = {nullptr, nullptr}, evPacked[2] = {nullptr, nullptr}, evOut[2] = {nullptr, nullptr};
    unsigned inSeq = 0, outSeq = 0;
    int pendingOut = -1;                    // staging slot of the download started last (tetsim_wait_positions)
    DevBuf<double> grabP;
    DevBuf<int> grabOut;
    DevBuf<unsigned long long> grabScratch;
    int grabId = -1;
    double grabPos[3] = {0, 0, 0};
    // skinning cache
    DevBuf<float4> visV;
    DevBuf<int> visTri, vtStart, vtEnt;
    DevBuf<float> visPos, visNrm;
    uint64_t visHashV = 0, visHashT = 0;   // content hashes of the cached surface mesh (a pointer is no identity: arrays get mutated in place)
    int visN = -1, visT = -1;
    DevBuf<float> visRestNrm;              // rest-pose normals of the WebGL variant's skinning (tetsim_skin_gpu)
    uint64_t visHashN = 0;
    DevBuf<int> tetRecord;                 // tiled polar solver: caller tet -> record slot (where its quaternion lives)
    // graphs
    std::map<int, cudaGraphExec_t> graphs;
    std::map<int, int64_t> graphLaunches;   // kernels inside each captured graph
    int64_t enq = 0;                        // kernels enqueued by the enqueue_substeps call in progress
    int64_t totalLaunches = 0;              // kernels launched by simulate/step since create
    int launchesPerSubstep = 0;
    // multi-GPU
    NcclComm comm = nullptr;
    cudaStream_t commStream = nullptr;      // the all-reduce runs here, beside the interior tiles
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    bool halo = false;                      // neighbour exchange instead of the all-reduce
    DevBuf<int> hxSendIdx, hxSrcStart, hxSrc;
    DevBuf<float4> hxSend, hxRecv;
    // peer-memory exchange (exchange = 2): this rank's exchange allocation, the sharers' mapped ones
    bool peer = false, peersSet = false;
    DevBuf<unsigned char> peerBuf;
    std::vector<void *> peerOpened;         // bases returned by cudaIpcOpenMemHandle (closed at destroy)
    DevBuf<unsigned char *> peerBase;       // [peers] mapped exchange allocation of each sharer
    DevBuf<int> pxStart, pxPeer, pxEntry, pxRemoteTotal, pxRemoteSlot;
    DevBuf<uint4> pushRec;                  // fused push: two uint4 per boundary-tile partial slot (PeerArgs::pushRec)
    DevBuf<PeerArgs> pxArgs;                // device copy of peer_args(h), read by the tile and vertex kernels
    bool peerFused = false;                 // the tile kernel pushes, the vertex kernel polls + reduces (2 launches/iteration)
    // TETSIM_TRACE=1: in-situ per-launch timing of the clustered Jacobi substep (events between launches, no graph);
    // the way to see where a multi-GPU substep spends its time, where ncu cannot be used
    bool trace = false;
    std::vector<std::pair<const char *, cudaEvent_t>> marks;
    int maxValence = 0;

    int64_t deviceBytes() const {
        return (int64_t)(x4.bytes() + prev4.bytes() + vel4.bytes() + vertId.bytes() + Q9.bytes() + irv.bytes() +
                         invMass.bytes() + ids.bytes() + cStart.bytes() + cEnt.bytes() + A.bytes() + B.bytes() +
                         C.bytes() + I.bytes() + order.bytes() + levelStart.bytes() + bodies.bytes() +
                         volTerm.bytes() + dx.bytes() + vpStart.bytes() + vpSlot.bytes() + tileTets.bytes() +
                         tileMeta.bytes() + metaOff.bytes() + part.bytes() + acc.bytes() +
                         bsum.bytes() + invVal.bytes() + rest.bytes() + quat.bytes() + tStart.bytes() + tEnt.bytes() +
                         stageIn[0].bytes() + stageIn[1].bytes() + stageOut[0].bytes() + stageOut[1].bytes() + peerBuf.bytes() + pushRec.bytes() + visV.bytes() + visTri.bytes() + vtStart.bytes() + vtEnt.bytes() +
                         visPos.bytes() + visNrm.bytes());
    }
};

namespace {
// What ranks hand each other for the peer-memory exchange (tetsim_get_ipc_handle / tetsim_set_peers).
struct PeerBlob {
    unsigned char ipc[64];      // cudaIpcMemHandle_t of the allocation holding the exchange buffer
    uint64_t offset;            // of the exchange buffer inside that allocation
    uint64_t rawPtr;            // the owner's own pointer: used instead of the IPC mapping inside the owner's process
    int32_t pid, device, rank, total, numPeers;
    uint32_t magic;
    unsigned char pad[TETSIM_PEER_BLOB_BYTES - 64 - 16 - 20 - 4];
};
static_assert(sizeof(PeerBlob) == TETSIM_PEER_BLOB_BYTES, "PeerBlob layout");
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
constexpr uint32_t kPeerMagic = 0x54455450u;  // "PTET"

// Offset of p inside its cudaMalloc allocation: an IPC handle always names the whole allocation and small buffers
// may be carved out of a shared block.  cuMemGetAddressRange lives in the driver library the runtime already loaded.
size_t allocation_offset(const void *p) {
    typedef int (*Fn)(unsigned long long *, size_t *, unsigned long long);
    static Fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        if (void *lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL)) fn = (Fn)dlsym(lib, "cuMemGetAddressRange_v2");
    }
    unsigned long long base = 0;
    size_t size = 0;
    if (fn && fn(&base, &size, (unsigned long long)(uintptr_t)p) == 0 && base) return (size_t)((uintptr_t)p - (uintptr_t)base);
    return 0;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

void fill_substep_params(SubstepParams &s, double dt, const TetSimParams &p, int grabId, const double grab[3]) {
    memset(&s, 0, sizeof(s));
    s.dt = dt; s.gravity = p.gravity; s.friction = p.friction;
    s.devCompliance = p.devCompliance; s.volCompliance = p.volCompliance;
    for (int c = 0; c < 3; c++) { s.lo[c] = p.worldBounds[c]; s.hi[c] = p.worldBounds[3 + c]; s.grab[c] = grab[c]; }
    s.grabId = grabId;
    s.alphaDevD = p.devCompliance / dt / dt;   // ((compliance / dt) / dt), the reference's grouping
    s.alphaVolD = p.volCompliance / dt / dt;
    s.volOverDevD = p.volCompliance / p.devCompliance;
    s.invDtD = 1.0 / dt;
    s.dtF = (float)dt;
    s.gDt = (float)(p.gravity * dt);
    s.invDt = (float)(1.0 / dt);
    double k = dt * p.friction;
    s.fric = (float)(k < 1.0 ? k : 1.0);
    s.alphaDev = (float)(p.devCompliance / dt / dt);
    s.alphaVol = (float)(p.volCompliance / dt / dt);
    s.gammaVol = (float)(1.0 + p.volCompliance / p.devCompliance);
    s.gravityF = (float)p.gravity;
    s.frictionF = (float)p.friction;
    for (int c = 0; c < 3; c++) { s.loF[c] = (float)p.worldBounds[c]; s.hiF[c] = (float)p.worldBounds[3 + c]; s.grabF[c] = (float)grab[c]; }
}

// ---- solver builders -----------------------------------------------------------------------------

// Gauss-Seidel schedules (EXACT levels or greedy colours).
int build_gs(tetsim *h, const std::vector<float> &verts, const std::vector<int> &tetIds) {
    const int N = h->N, M = h->M;
    std::vector<int> level((size_t)M);
    int nl;
    if (h->opt.solver == TETSIM_NH_GS_EXACT) nl = level_schedule(N, M, tetIds.data(), level.data());
    else {
        nl = greedy_colors(N, M, tetIds.data(), level.data());
        if (nl < 0) return fail(TETSIM_E_INVALID, "greedy colouring needs more than 256 colours");
    }
    std::vector<int> comp;
    h->numComponents = connected_components(N, M, tetIds.data(), comp);
    // component sizes -> can every body live in one CTA's shared memory?
    std::vector<int> compVerts((size_t)h->numComponents, 0);
    for (int v = 0; v < N; v++) compVerts[comp[v]]++;
    int maxCompVerts = 0;
    for (int c : compVerts) maxCompVerts = std::max(maxCompVerts, c);
    const char *forceLevel = getenv("TETSIM_FORCE_LEVEL_KERNEL");
    // vertices (16 B each) + the body's level table (at most numLevels + 1 ints) must fit in one CTA's shared memory
    h->bodyKernel = (size_t)maxCompVerts * sizeof(float4) + ((size_t)nl + 1) * sizeof(int) <= (size_t)h->K->gs_body_max_smem() && !(forceLevel && forceLevel[0] == '1');

    // internal vertex numbering: bodies contiguous
    std::vector<int> c2i((size_t)N);
    if (h->bodyKernel && h->numComponents > 1) {
        h->h_vertId.resize((size_t)N);
        std::iota(h->h_vertId.begin(), h->h_vertId.end(), 0);
        std::stable_sort(h->h_vertId.begin(), h->h_vertId.end(), [&](int a, int b) { return comp[a] < comp[b]; });
        bool identity = true;
        for (int i = 0; i < N; i++) { c2i[h->h_vertId[i]] = i; identity = identity && h->h_vertId[i] == i; }
        if (identity) h->h_vertId.clear();
    } else {
        std::iota(c2i.begin(), c2i.end(), 0);
    }
    // solver order: (body, level, tet) when bodies are separate CTAs, else (level, tet)
    std::vector<int> order((size_t)M);
    std::iota(order.begin(), order.end(), 0);
    auto tetComp = [&](int e) { return comp[tetIds[4 * (size_t)e]]; };
    if (h->bodyKernel)
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            int ca = tetComp(a), cb = tetComp(b);
            return ca != cb ? ca < cb : level[a] < level[b];
        });
    else
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return level[a] < level[b]; });

    std::vector<int4> I((size_t)M);
    h->h_levelStart.clear();
    std::vector<BodyDesc> bodies;
    h->maxLevelSize = 0;
    if (h->bodyKernel) {
        std::vector<int> vertBegin((size_t)h->numComponents + 1, 0);
        for (int c = 0; c < h->numComponents; c++) vertBegin[c + 1] = vertBegin[c] + compVerts[c];
        int pos = 0;
        for (int c = 0; c < h->numComponents; c++) {
            BodyDesc bd;
            bd.vertBegin = vertBegin[c]; bd.vertEnd = vertBegin[c + 1];
            bd.tetBegin = pos;
            bd.levelBegin = (int)h->h_levelStart.size();
            int lastLevel = -1, runStart = pos;
            while (pos < M && tetComp(order[pos]) == c) {
                int e = order[pos];
                if (level[e] != lastLevel) {
                    if (lastLevel >= 0) h->maxLevelSize = std::max(h->maxLevelSize, pos - runStart);
                    h->h_levelStart.push_back(pos);
                    lastLevel = level[e];
                    runStart = pos;
                }
                const int *t = tetIds.data() + 4 * (size_t)e;
                I[pos] = make_int4(c2i[t[0]] - bd.vertBegin, c2i[t[1]] - bd.vertBegin, c2i[t[2]] - bd.vertBegin, c2i[t[3]] - bd.vertBegin);
                pos++;
            }
            if (lastLevel >= 0) h->maxLevelSize = std::max(h->maxLevelSize, pos - runStart);
            bd.levelEnd = (int)h->h_levelStart.size();
            h->h_levelStart.push_back(pos);  // sentinel: end of this body's last level
            if (bd.tetBegin != pos || bd.vertEnd > bd.vertBegin) bodies.push_back(bd);
        }
        h->numBodies = (int)bodies.size();
        h->numLevels = nl;
        h->bodyThreads = maxCompVerts > 512 ? 128 : 64;
        h->bodyThreads = std::max(h->bodyThreads, std::min(256, 32 * ((h->maxLevelSize + 31) / 32)));
        int maxBodyLevels = 0;
        for (const BodyDesc &bd : bodies) maxBodyLevels = std::max(maxBodyLevels, bd.levelEnd - bd.levelBegin);
        h->bodySmem = (size_t)maxCompVerts * sizeof(float4) + ((size_t)maxBodyLevels + 1) * sizeof(int);
        h->launchesPerSubstep = 1;
    } else {
        int lastLevel = -1, runStart = 0;
        for (int pos = 0; pos < M; pos++) {
            int e = order[pos];
            if (level[e] != lastLevel) {
                if (lastLevel >= 0) h->maxLevelSize = std::max(h->maxLevelSize, pos - runStart);
                h->h_levelStart.push_back(pos);
                lastLevel = level[e];
                runStart = pos;
            }
            const int *t = tetIds.data() + 4 * (size_t)e;
            I[pos] = make_int4(t[0], t[1], t[2], t[3]);
        }
        if (M > 0) h->maxLevelSize = std::max(h->maxLevelSize, M - runStart);
        h->h_levelStart.push_back(M);
        h->numLevels = nl;
        h->launchesPerSubstep = 2 + nl;
    }
    cudaStream_t s = h->stream;
    CK(h->order.upload(order, s));
    CK(h->I.upload(I, s));
    CK(h->levelStart.upload(h->h_levelStart, s));
    if (h->bodyKernel) CK(h->bodies.upload(bodies, s));
    CK(h->A.alloc((size_t)M)); CK(h->B.alloc((size_t)M));
    // FAST arithmetic, one CTA per body: the four-lanes-per-tet kernel on the rest-metric stream (k_gs_body_quads)
    // (narrow levels only: the exact-order schedule has <= 22 tets per level on Dragon; a colour class of 228 tets keeps one thread per tet)
    h->gsQuads = h->bodyKernel && h->opt.arithmetic == TETSIM_ARITH_FAST_F32 && h->maxLevelSize <= 64;
    if (const char *e = getenv("TETSIM_GS_NO_QUADS")) if (e[0] == '1') h->gsQuads = false;
    if (h->gsQuads) {
        launch_build_stream_metric(s, M, h->order.p, h->Q9.p, h->irv.p, h->A.p, h->B.p);
        h->bodyThreads = std::max(64, std::min(512, 32 * ((4 * h->maxLevelSize + 31) / 32)));
        if (const char *e = getenv("TETSIM_GS_THREADS")) { const int v = atoi(e); if (v >= 32 && v <= 1024 && v % 32 == 0) h->bodyThreads = v; }
    } else {
        CK(h->C.alloc((size_t)M));
        CK(cudaMemsetAsync(h->C.p, 0, h->C.bytes(), s));
        launch_build_stream(s, M, h->order.p, h->Q9.p, h->irv.p, h->A.p, h->B.p, h->C.p);
    }
    if (h->trackVol) { CK(h->volTerm.alloc((size_t)M)); CK(cudaMemsetAsync(h->volTerm.p, 0, h->volTerm.bytes(), s)); }
    (void)verts;
    return TETSIM_OK;
}

int build_jacobi_gather(tetsim *h, const std::vector<int> &tetIds) {
    const int M = h->M;
    cudaStream_t s = h->stream;
    CK(h->A.alloc((size_t)M)); CK(h->B.alloc((size_t)M)); CK(h->C.alloc((size_t)M));
    CK(cudaMemsetAsync(h->C.p, 0, h->C.bytes(), s));
    launch_build_stream(s, M, nullptr, h->Q9.p, h->irv.p, h->A.p, h->B.p, h->C.p);
    CK(h->dx.alloc(4 * (size_t)M));
    if (h->trackVol) { CK(h->volTerm.alloc((size_t)M)); CK(cudaMemsetAsync(h->volTerm.p, 0, h->volTerm.bytes(), s)); }
    h->launchesPerSubstep = 2 + 2 * h->opt.iters;
    (void)tetIds;
    return TETSIM_OK;
}

int build_jacobi_cluster(tetsim *h, const std::vector<float> &verts, const std::vector<int> &tetIds) {
    const int N = h->N, M = h->M;
    std::vector<int> rankStart;
    std::vector<int> order = solver_order(N, M, verts.data(), tetIds.data(), h->opt.reorder != 0, h->opt.worldSize, rankStart);
    std::string err;
    if (!build_cluster_plan(N, M, tetIds.data(), order, rankStart, h->opt.clusterSize, h->opt.rank, h->opt.worldSize, h->plan, err))
        return fail(TETSIM_E_INVALID, err);
    ClusterPlan &P = h->plan;
    h->clustered = true;
    cudaStream_t s = h->stream;
    const size_t nRec = (size_t)P.numClusters * P.T;
    CK(h->order.upload(P.recordTet, s));
    {
        DevBuf<uint4> aux;
        CK(aux.alloc(nRec));
        if (nRec) CK(cudaMemcpyAsync(aux.p, P.recordAux.data(), nRec * sizeof(uint4), cudaMemcpyHostToDevice, s));
        CK(h->tileTets.alloc(nRec * 48));
        launch_build_tiles(s, P.T, (int)nRec, h->order.p, h->Q9.p, h->irv.p, aux.p, h->tileTets.p);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(s));
        aux.release();
    }
    CK(h->tileMeta.upload(P.tileMeta, s));
    CK(h->metaOff.upload(P.metaOff, s));
    CK(h->vpStart.upload(P.vpStart, s));
    CK(h->vpSlot.upload(P.vpSlot, s));
    CK(h->invVal.upload(P.invValence, s));
    if (h->opt.deterministic) CK(h->part.alloc(P.clVerts.size()));
    else { CK(h->acc.alloc((size_t)P.numLocalVerts)); CK(cudaMemsetAsync(h->acc.p, 0, h->acc.bytes(), s)); }
    if (P.numBoundary > 0) { CK(h->bsum.alloc((size_t)P.numBoundary)); CK(cudaMemsetAsync(h->bsum.p, 0, h->bsum.bytes(), s)); }
    if (h->trackVol) { CK(h->volTerm.alloc(1)); CK(cudaMemsetAsync(h->volTerm.p, 0, sizeof(double), s)); }
    h->h_vertId = P.localToCaller;
    h->halo = h->opt.worldSize > 1 && h->opt.exchange == 1 && P.haloOk;
    if (h->halo) {
        CK(h->hxSendIdx.upload(P.hxSendIdx, s));
        CK(h->hxSrcStart.upload(P.hxSrcStart, s));
        CK(h->hxSrc.upload(P.hxSrc, s));
        CK(h->hxSend.alloc(P.hxSendIdx.size()));
        CK(h->hxRecv.alloc(P.hxSendIdx.size()));
        // replicas of boundary vertices this rank's tets never touch receive no sums: not maintained here
        for (int b = 0; b < P.numBoundary; b++)
            if (!P.boundaryActive[b]) h->h_vertId[(size_t)P.numInterior + b] = -1;
    }
    h->peer = h->opt.worldSize > 1 && h->opt.exchange == 2;
    if (h->peer) {
        if (!P.haloOk) return fail(TETSIM_E_STATE, "the peer-memory exchange needs the neighbour lists (worldSize <= 64)");
        const size_t total = P.hxSendIdx.size();  // entries I receive == entries I send
        const char *unfused = getenv("TETSIM_PEER_UNFUSED");
        // the fused push sends tile partials, so it needs the deterministic flush (per-tile partial sums).  The choice
        // must be the same on every rank (it fixes the layout of the receive buffers): options + environment only.
        h->peerFused = h->opt.deterministic != 0 && !(unfused && unfused[0] == '1') && jacobi_tiles_has_peer_push(P.T);
        if (h->peerFused && P.pxMaxPartials > kPeerK)
            return fail(TETSIM_E_STATE, "a rank-shared vertex has " + std::to_string(P.pxMaxPartials) + " tile partials on this rank (limit " + std::to_string(kPeerK) + "): use a larger clusterSize or exchange = 1");
        const size_t entryBytes = h->peerFused ? (size_t)kPeerK * 32 : sizeof(float4);
        CK(h->peerBuf.alloc(kPeerRecvOff + 2 * std::max<size_t>(total, 1) * entryBytes));
        CK(cudaMemsetAsync(h->peerBuf.p, 0, h->peerBuf.bytes(), s));
        CK(h->hxSrcStart.upload(P.hxSrcStart, s));
        CK(h->hxSrc.upload(P.hxSrc, s));
        CK(h->pxStart.upload(P.pxStart, s));
        CK(h->pxPeer.upload(P.pxPeer, s));
        CK(h->pxEntry.upload(P.pxEntry, s));
        CK(h->pxRemoteTotal.upload(P.pxRemoteTotal, s));
        CK(h->pxRemoteSlot.upload(P.pxRemoteSlot, s));
        CK(h->peerBase.alloc(std::max<size_t>(P.hxPeers.size(), 1)));
        CK(h->pxArgs.alloc(1));
        for (int b = 0; b < P.numBoundary; b++)
            if (!P.boundaryActive[b]) h->h_vertId[(size_t)P.numInterior + b] = -1;
    }
    bool identity = (int)h->h_vertId.size() == N;
    for (int i = 0; identity && i < N; i++) identity = h->h_vertId[i] == i;
    if (identity) h->h_vertId.clear();
    h->maxValence = P.maxValence;
    h->launchesPerSubstep = 2 * h->opt.iters + (h->opt.worldSize > 1 && !h->peerFused ? 2 * h->opt.iters : 0);
    CK(cudaStreamSynchronize(s));
    return TETSIM_OK;
}

// Tiled polar solver: the tiling (and with it the internal vertex numbering) is decided before the state is uploaded.
int plan_polar_tiles(tetsim *h, const std::vector<float> &verts, const std::vector<int> &tetIds) {
    std::vector<int> rankStart;
    std::vector<int> order = solver_order(h->N, h->M, verts.data(), tetIds.data(), h->opt.reorder != 0, 1, rankStart);
    std::string err;
    const int T = h->opt.clusterSize < 128 ? 128 : h->opt.clusterSize;
    if (!build_cluster_plan(h->N, h->M, tetIds.data(), order, rankStart, T, 0, 1, h->plan, err)) return fail(TETSIM_E_INVALID, err);
    h->h_vertId = h->plan.localToCaller;
    bool identity = (int)h->h_vertId.size() == h->N;
    for (int i = 0; identity && i < h->N; i++) identity = h->h_vertId[i] == i;
    if (identity) h->h_vertId.clear();
    return TETSIM_OK;
}

int build_polar(tetsim *h, const std::vector<int> &tetIds) {
    cudaStream_t s = h->stream;
    if (h->polarTiled) {
        const ClusterPlan &P = h->plan;
        const size_t nRec = (size_t)P.numClusters * P.T;
        CK(h->order.upload(P.recordTet, s));
        DevBuf<uint4> aux;
        CK(aux.alloc(nRec));
        if (nRec) CK(cudaMemcpyAsync(aux.p, P.recordAux.data(), nRec * sizeof(uint4), cudaMemcpyHostToDevice, s));
        CK(h->tileTets.alloc(nRec * 80));
        CK(h->tileVol.alloc(std::max<size_t>(nRec, 1)));
        launch_build_polar_tiles(s, P.T, (int)nRec, h->order.p, h->x4.p, h->ids.p, h->irv.p, aux.p, h->opt.referenceTableBug != 0, h->tileTets.p, h->tileVol.p);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(s));
        aux.release();
        CK(h->tileMeta.upload(P.tileMeta, s));
        CK(h->metaOff.upload(P.metaOff, s));
        CK(h->vpStart.upload(P.vpStart, s));
        CK(h->vpSlot.upload(P.vpSlot, s));
        CK(h->part.alloc(std::max<size_t>(P.clVerts.size(), 1)));
        h->launchesPerSubstep = 2;
        CK(cudaStreamSynchronize(s));
        return TETSIM_OK;
    }
    CornerTable t = build_reference_table(h->N, h->M, tetIds.data(), h->opt.referenceTableBug != 0, 36);
    CK(h->tStart.upload(t.start, s));
    CK(h->tEnt.upload(t.ent, s));
    CK(h->rest.alloc(4 * (size_t)h->M));
    CK(h->quat.alloc((size_t)h->M));
    launch_fill_rest(s, h->M, h->x4.p, h->ids.p, h->irv.p, h->rest.p, h->quat.p);
    h->launchesPerSubstep = 3;
    CK(cudaStreamSynchronize(s));
    return TETSIM_OK;
}

PolarTileArgs polar_tile_args(const tetsim *h) {
    const ClusterPlan &P = h->plan;
    PolarTileArgs a{};
    a.x4 = h->x4.p; a.tets = h->tileTets.p; a.vol = h->tileVol.p; a.meta = h->tileMeta.p; a.metaOff = h->metaOff.p;
    a.numTiles = P.numClusters; a.metaStride = P.metaStride; a.metaValOff = P.metaValOff;
    a.maxTileVertsPad = P.maxTileVertsPad; a.maxTileEntries = P.maxTileEntries; a.part = h->part.p;
    float k = 4.0f;
    if (const char *e = getenv("TETSIM_POLAR_NOISE_K")) k = (float)atof(e);
    a.noiseK2 = (k * 5.9604645e-8f) * (k * 5.9604645e-8f);
    return a;
}

// ---- substep scheduling --------------------------------------------------------------------------

TileArgs tile_args(const tetsim *h) {
    const ClusterPlan &P = h->plan;
    TileArgs a{};
    a.x4 = h->x4.p; a.tets = h->tileTets.p; a.meta = h->tileMeta.p; a.tileBegin = 0; a.numTiles = P.numClusters;
    a.metaOff = h->metaOff.p; a.metaStride = P.metaStride; a.metaValOff = P.metaValOff;
    a.colStride = P.colStride; a.maxTileVertsPad = P.maxTileVertsPad; a.maxTileEntries = P.maxTileEntries;
    a.part = h->part.p; a.acc = nullptr; a.volAcc = nullptr; a.sp = h->sp.p;
    return a;
}

void trace_mark(tetsim *h, const char *label) {  // the interval ending here is charged to `label`
    cudaEvent_t ev = nullptr;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, h->stream);
    h->marks.emplace_back(label, ev);
}
#define TR(label) do { if (h->trace) trace_mark(h, label); } while (0)

// Print and reset the trace (stream must be idle).
void trace_report(tetsim *h) {
    if (h->marks.empty()) return;
    std::map<std::string, std::pair<int, double>> agg;
    std::vector<std::string> orderSeen;
    for (size_t i = 1; i < h->marks.size(); i++) {
        if (std::string(h->marks[i].first) == "begin") continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->marks[i - 1].second, h->marks[i].second) != cudaSuccess) continue;
        auto it = agg.find(h->marks[i].first);
        if (it == agg.end()) { orderSeen.push_back(h->marks[i].first); it = agg.emplace(h->marks[i].first, std::make_pair(0, 0.0)).first; }
        it->second.first++; it->second.second += ms;
    }
    double total = 0.0;
    for (auto &kv : agg) total += kv.second.second;
    fprintf(stderr, "tetsim trace, rank %d of %d (event-to-event intervals, launches not graphed):\n", h->opt.rank, h->opt.worldSize);
    for (const std::string &k : orderSeen)
        fprintf(stderr, "  %-28s n=%6d  avg %8.2f us  share %5.1f %%\n", k.c_str(), agg[k].first,
                1e3 * agg[k].second / std::max(agg[k].first, 1), 100.0 * agg[k].second / std::max(total, 1e-12));
    for (auto &m : h->marks) cudaEventDestroy(m.second);
    h->marks.clear();
    cudaGetLastError();
}

PeerArgs peer_args(const tetsim *h) {
    const ClusterPlan &P = h->plan;
    PeerArgs a{};
    a.numBoundary = P.numBoundary; a.boundaryBegin = P.numInterior; a.numPeers = (int)P.hxPeers.size();
    a.vpStart = h->vpStart.p; a.vpSlot = h->vpSlot.p; a.part = h->part.p; a.acc = h->acc.p; a.bsum = h->bsum.p;
    a.pxStart = h->pxStart.p; a.pxPeer = h->pxPeer.p; a.pxEntry = h->pxEntry.p;
    a.peerBase = h->peerBase.p; a.remoteTotal = h->pxRemoteTotal.p; a.remoteSlot = h->pxRemoteSlot.p;
    a.self = h->peerBuf.p; a.selfTotal = (int)P.hxSendIdx.size();
    a.srcStart = h->hxSrcStart.p; a.src = h->hxSrc.p;
    a.pushRec = h->pushRec.p;
    unsigned long long ms = 10000ull;
    if (const char *e = getenv("TETSIM_PEER_TIMEOUT_MS")) { long v = atol(e); if (v > 0) ms = (unsigned long long)v; }
    a.timeoutNs = ms * 1000000ull;
    return a;
}

// Enqueue `count` substeps.  Inside one call dt and the parameters are constant, which is what lets
// the clustered Jacobi path fuse a substep's post with the next substep's predict.
int enqueue_substeps(tetsim *h, int count) {
    cudaStream_t s = h->stream;
    h->enq = 0;
    const KernelTable *K = h->K;
    const SubstepParams *sp = h->sp.p;
    const int *vid = h->vertId.p;
    for (int step = 0; step < count; step++) {
        switch (h->opt.solver) {
            case TETSIM_NH_GS_EXACT:
            case TETSIM_NH_GS_COLOR:
                if (h->bodyKernel) {
                    h->enq += h->numBodies > 0;
                    if (h->gsQuads)
                        launch_gs_body_quads(s, h->numBodies, h->bodyThreads, h->bodySmem, h->bodies.p, h->levelStart.p, h->x4.p,
                                             h->prev4.p, h->vel4.p, h->I.p, h->A.p, h->B.p, h->order.p, h->volTerm.p, sp, vid);
                    else
                    K->gs_body(s, h->numBodies, h->bodyThreads, h->bodySmem, h->bodies.p, h->levelStart.p, h->x4.p,
                               h->prev4.p, h->vel4.p, h->I.p, h->A.p, h->B.p, h->C.p, h->order.p, h->volTerm.p, sp, vid);
                } else {
                    K->predict(s, h->nInt, h->x4.p, h->prev4.p, h->vel4.p, sp);
                    h->enq += 2 + (int64_t)h->h_levelStart.size() - 1;
                    for (size_t l = 0; l + 1 < h->h_levelStart.size(); l++)
                        K->gs_level(s, h->h_levelStart[l], h->h_levelStart[l + 1], h->x4.p, h->I.p, h->A.p, h->B.p,
                                    h->C.p, h->order.p, h->volTerm.p, sp);
                    K->post(s, h->nInt, h->x4.p, h->prev4.p, h->vel4.p, sp, vid);
                }
                break;
            case TETSIM_NH_JACOBI:
                if (!h->clustered) {
                    K->predict(s, h->nInt, h->x4.p, h->prev4.p, h->vel4.p, sp);
                    h->enq += 2 + 2 * h->opt.iters;
                    for (int it = 0; it < h->opt.iters; it++) {
                        K->jacobi_tet(s, h->M, h->x4.p, h->ids.p, h->A.p, h->B.p, h->C.p, h->dx.p, h->volTerm.p, sp);
                        K->jacobi_gather(s, h->nInt, h->x4.p, h->cStart.p, h->cEnt.p, h->dx.p);
                    }
                    K->post(s, h->nInt, h->x4.p, h->prev4.p, h->vel4.p, sp, vid);
                } else {
                    const ClusterPlan &P = h->plan;
                    if (step == 0) { TR("begin"); K->predict(s, h->nInt, h->x4.p, h->prev4.p, h->vel4.p, sp); h->enq++; TR("predict"); }
                    TileArgs ca = tile_args(h);
                    ca.acc = h->acc.p; ca.volAcc = h->volTerm.p;
                    ApplyArgs aa{};
                    aa.x4 = h->x4.p; aa.prev4 = h->prev4.p; aa.vel4 = h->vel4.p;
                    aa.vpStart = h->vpStart.p; aa.vpSlot = h->vpSlot.p; aa.part = h->part.p; aa.acc = h->acc.p;
                    aa.invVal = h->invVal.p; aa.sp = sp; aa.vertId = vid;
                    aa.boundaryBegin = P.numInterior;
                    const bool multi = h->opt.worldSize > 1 && P.numBoundary > 0;
                    for (int it = 0; it < h->opt.iters; it++) {
                        if (h->trackVol) CK(cudaMemsetAsync(h->volTerm.p, 0, sizeof(double), s));
                        const bool last = it == h->opt.iters - 1;
                        const int mode = !last ? 0 : (step + 1 < count ? 2 : 1);
                        if (!multi) {
                            launch_jacobi_tiles(s, P.T, ca);
                            TR("tiles");
                            h->enq += 2;  // tile kernel + vertex kernel
                        } else if (h->peer && h->peerFused) {
                            // ONE tile launch: boundary tiles first, every tile stores its partials of rank-shared vertices
                            // straight into the sharers' buffers; the vertex kernel below polls theirs and reduces
                            TileArgs cf = ca;
                            cf.px = h->pxArgs.p;
                            cf.pxSlots = (int)P.pxSlotIdx.size();
                            launch_jacobi_tiles(s, P.T, cf);
                            TR("tiles + peer push");
                            aa.px = h->pxArgs.p;
                            h->enq += 2;
                        } else if (h->peer) {
                            // boundary tiles -> push this rank's boundary sums into the sharers' buffers (+ flag)
                            // -> interior tiles (the sharers' stores arrive meanwhile) -> wait + rank-ordered reduce
                            TileArgs cb = ca;
                            cb.numTiles = P.numBoundaryTiles;
                            launch_jacobi_tiles(s, P.T, cb);
                            TR("boundary tiles");
                            PeerArgs pa = peer_args(h);
                            launch_peer_push(s, pa);
                            TR("peer push");
                            TileArgs ci = ca;
                            ci.tileBegin = P.numBoundaryTiles;
                            launch_jacobi_tiles(s, P.T, ci);
                            TR("interior tiles");
                            launch_peer_reduce(s, pa);
                            TR("peer wait + reduce");
                            aa.bsum = h->bsum.p;
                            h->enq += 4 + (P.numBoundaryTiles > 0 && P.numBoundaryTiles < P.numClusters ? 1 : 0);
                        } else {
                            // 1. the tiles that touch rank-shared vertices, then this rank's boundary sums
                            TileArgs cb = ca;
                            cb.numTiles = P.numBoundaryTiles;
                            launch_jacobi_tiles(s, P.T, cb);
                            TR("boundary tiles");
                            if (h->acc.p) {
                                CK(cudaMemcpyAsync(h->bsum.p, h->acc.p + P.numInterior, h->bsum.bytes(), cudaMemcpyDeviceToDevice, s));
                                CK(cudaMemsetAsync(h->acc.p + P.numInterior, 0, h->bsum.bytes(), s));
                            } else {
                                launch_boundary_pack(s, P.numInterior, P.numBoundary, h->vpStart.p, h->vpSlot.p, h->part.p, h->bsum.p);
                                h->enq++;
                            }
                            if (h->halo) { launch_halo_pack(s, (int)P.hxSendIdx.size(), h->hxSendIdx.p, h->bsum.p, h->hxSend.p); h->enq++; }
                            TR("boundary pack");
                            // 2. all-reduce (or neighbour exchange) over NVLink on the side stream ...
                            CK(cudaEventRecord(h->evFork, s));
                            CK(cudaStreamWaitEvent(h->commStream, h->evFork, 0));
                            int rc = 0;
                            if (!h->halo) {
                                rc = g_nccl.AllReduce(h->bsum.p, h->bsum.p, (size_t)P.numBoundary * 4, kNcclFloat, kNcclSum, h->comm, h->commStream);
                            } else {  // pairwise exchange with the ranks that share vertices with this one
                                rc = g_nccl.GroupStart();
                                for (size_t q = 0; q < P.hxPeers.size() && rc == 0; q++) {
                                    const size_t off = (size_t)P.hxSegStart[q], n = (size_t)(P.hxSegStart[q + 1] - P.hxSegStart[q]) * 4;
                                    rc = g_nccl.Send(h->hxSend.p + off, n, kNcclFloat, P.hxPeers[q], h->comm, h->commStream);
                                    if (rc == 0) rc = g_nccl.Recv(h->hxRecv.p + off, n, kNcclFloat, P.hxPeers[q], h->comm, h->commStream);
                                }
                                int rc2 = g_nccl.GroupEnd();
                                if (rc == 0) rc = rc2;
                            }
                            if (rc != 0) return fail(TETSIM_E_NCCL, std::string("NCCL exchange: ") + g_nccl.GetErrorString(rc));
                            CK(cudaEventRecord(h->evJoin, h->commStream));
                            // 3. ... while the interior tiles run on the main stream
                            TileArgs ci = ca;
                            ci.tileBegin = P.numBoundaryTiles;
                            launch_jacobi_tiles(s, P.T, ci);
                            TR("interior tiles");
                            CK(cudaStreamWaitEvent(s, h->evJoin, 0));
                            if (h->halo) { launch_halo_reduce(s, P.numBoundary, h->hxSrcStart.p, h->hxSrc.p, h->hxRecv.p, h->bsum.p); h->enq++; }
                            TR("exchange wait + reduce");
                            aa.bsum = h->bsum.p;
                            h->enq += 2 + (P.numBoundaryTiles > 0 && P.numBoundaryTiles < P.numClusters ? 1 : 0);
                        }
                        launch_jacobi_apply(s, 0, h->nInt, mode, aa);
                        TR("vertex kernel");
                    }
                }
                break;
            case TETSIM_POLAR_JACOBI:
                if (h->polarTiled) {
                    if (step == 0) { K->polar_integrate(s, h->nInt, h->x4.p, h->prev4.p, h->vel4.p, sp); h->enq++; }
                    launch_polar_tiles(s, h->plan.T, polar_tile_args(h));
                    launch_polar_vertex_tiles(s, h->nInt, step + 1 < count ? 2 : 1, h->x4.p, h->prev4.p, h->vel4.p, h->vpStart.p,
                                              h->vpSlot.p, h->part.p, vid, sp);
                    h->enq += 2;
                    break;
                }
                h->enq += 3;
                K->polar_integrate(s, h->nInt, h->x4.p, h->prev4.p, h->vel4.p, sp);
                K->polar_tet(s, h->M, h->x4.p, h->ids.p, h->rest.p, h->quat.p);
                K->polar_vertex(s, h->nInt, h->x4.p, h->prev4.p, h->vel4.p, h->tStart.p, h->tEnt.p, h->rest.p, sp);
                break;
            default:
                return fail(TETSIM_E_STATE, "unknown solver");
        }
    }
    CK(cudaGetLastError());
    return TETSIM_OK;
}

int run_substeps(tetsim *h, double dt, int count, const TetSimParams *params) {
    if (!h) return fail(TETSIM_E_INVALID, "null handle");
    if (count < 1) return fail(TETSIM_E_INVALID, "numSubsteps must be >= 1");
    if (!(dt == dt)) return fail(TETSIM_E_INVALID, "dt is NaN");
    if (h->peer && !h->peersSet) return fail(TETSIM_E_STATE, "exchange = 2: call tetsim_set_peers before simulate/step");
    DeviceGuard g(h->device);
    if (params) h->params = *params;
    SubstepParams hs;
    fill_substep_params(hs, dt, h->params, h->grabId, h->grabPos);
    // pageable source: the runtime stages the 300 bytes before returning, so `hs` may die here
    CK(cudaMemcpyAsync(h->sp.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, h->stream));
    const char *noGraph = getenv("TETSIM_NO_GRAPH");
    if (h->trace || (noGraph && noGraph[0] == '1')) {
        int rc = enqueue_substeps(h, count);
        h->totalLaunches += h->enq;
        return rc;
    }
    auto it = h->graphs.find(count);
    if (it == h->graphs.end()) {
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_substeps(h, count);
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (rc != TETSIM_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return fail(TETSIM_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
        cudaGraphExec_t exec = nullptr;
        ce = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return fail(TETSIM_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
        it = h->graphs.emplace(count, exec).first;
        h->graphLaunches[count] = h->enq;
    }
    CK(cudaGraphLaunch(it->second, h->stream));
    h->totalLaunches += h->graphLaunches[count];
    return TETSIM_OK;
}

// Peer-memory exchange: did a wait for a sharer time out?  (stream already synchronized by the caller)
int peer_check(tetsim *h) {
    if (!h->peer || !h->peerBuf.p) return TETSIM_OK;
    unsigned err = 0;
    CK(cudaMemcpy(&err, h->peerBuf.p + kPeerCtlOff + 8, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) return fail(TETSIM_E_STATE, "peer-memory exchange: a sharer's boundary sums did not arrive within the timeout (ranks out of step, or a peer died); state is invalid");
    return TETSIM_OK;
}

// ---- state transfer (see the comment on tetsim::h2dStream) ----
int xfer_init(tetsim *h) {
    if (h->h2dStream) return TETSIM_OK;
    CK(cudaStreamCreateWithFlags(&h->h2dStream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->d2hStream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        CK(cudaEventCreateWithFlags(&h->evIn[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->evInFree[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->evPacked[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->evOut[b], cudaEventDisableTiming));
    }
    return TETSIM_OK;
}

// resident = the arrays are in the handle's own vertex order (tetsim_get_resident_ids), nInt entries; else caller order, N.
int upload_state(tetsim *h, const float *pos, const float *prevPos, const float *vel, bool resident) {
    if (!h) return fail(TETSIM_E_INVALID, "null handle");
    DeviceGuard g(h->device);
    if (int rc = xfer_init(h)) return rc;
    const float *src[3] = {pos, prevPos, vel};
    const size_t n3 = 3 * (size_t)(resident ? h->nInt : h->N);
    int mask = 0;
    for (int k = 0; k < 3; k++) mask |= src[k] ? 1 << k : 0;
    if (!mask || n3 == 0) return TETSIM_OK;
    const int b = (int)(h->inSeq++ & 1u);
    if (h->stageIn[b].n < 3 * n3) CK(h->stageIn[b].alloc(3 * n3));
    CK(cudaStreamWaitEvent(h->h2dStream, h->evInFree[b], 0));   // the unpack that last read this staging buffer is done
    for (int k = 0; k < 3; k++)
        if (src[k]) CK(cudaMemcpyAsync(h->stageIn[b].p + k * n3, src[k], n3 * sizeof(float), cudaMemcpyHostToDevice, h->h2dStream));
    CK(cudaEventRecord(h->evIn[b], h->h2dStream));
    CK(cudaStreamWaitEvent(h->stream, h->evIn[b], 0));
    launch_unpack_state(h->stream, h->nInt, n3, h->stageIn[b].p, resident ? nullptr : h->vertId.p, h->x4.p, h->prev4.p, h->vel4.p, mask);
    CK(cudaEventRecord(h->evInFree[b], h->stream));
    CK(cudaGetLastError());
    return TETSIM_OK;
}

int download_begin(tetsim *h, const float4 *src, float *out, bool resident) {
    if (!h || !out) return fail(TETSIM_E_INVALID, "null argument");
    DeviceGuard g(h->device);
    if (int rc = xfer_init(h)) return rc;
    const size_t n3 = 3 * (size_t)(resident ? h->nInt : h->N);
    const int b = (int)(h->outSeq++ & 1u);
    if (h->stageOut[b].n < std::max<size_t>(n3, 1)) CK(h->stageOut[b].alloc(std::max<size_t>(n3, 1)));
    CK(cudaStreamWaitEvent(h->stream, h->evOut[b], 0));         // the copy that last read this staging buffer is done
    if (!resident && h->nInt != h->N) CK(cudaMemsetAsync(h->stageOut[b].p, 0xff, n3 * sizeof(float), h->stream));  // non-resident -> NaN
    launch_pack3(h->stream, h->nInt, src, resident ? nullptr : h->vertId.p, h->stageOut[b].p);
    CK(cudaEventRecord(h->evPacked[b], h->stream));
    CK(cudaStreamWaitEvent(h->d2hStream, h->evPacked[b], 0));
    if (n3) CK(cudaMemcpyAsync(out, h->stageOut[b].p, n3 * sizeof(float), cudaMemcpyDeviceToHost, h->d2hStream));
    CK(cudaEventRecord(h->evOut[b], h->d2hStream));
    h->pendingOut = b;
    CK(cudaGetLastError());
    return TETSIM_OK;
}

int download_wait(tetsim *h) {
    if (!h) return fail(TETSIM_E_INVALID, "null handle");
    DeviceGuard g(h->device);
    if (h->pendingOut >= 0) { CK(cudaEventSynchronize(h->evOut[h->pendingOut])); h->pendingOut = -1; }
    return peer_check(h);
}

int fetch3(tetsim *h, const float4 *src, float *out, bool resident = false) {
    if (int rc = download_begin(h, src, out, resident)) return rc;
    return download_wait(h);
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char *tetsim_last_error(void) { return g_err.c_str(); }
int tetsim_version(void) { return TETSIM_VERSION; }

int tetsim_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { fail(TETSIM_E_CUDA, cudaGetErrorString(e)); return TETSIM_E_CUDA; }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

void tetsim_default_params(TetSimParams *p) {
    if (!p) return;
    p->gravity = -9.81;
    p->friction = 1000.0;
    p->density = 1000.0;
    p->devCompliance = 1.0 / 100000.0;
    p->volCompliance = 0.0;
    const double wb[6] = {-2.5, -1.0, -2.5, 2.5, 10.0, 2.5};
    memcpy(p->worldBounds, wb, sizeof(wb));
}

void tetsim_default_options(TetSimOptions *o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->solver = TETSIM_NH_GS_EXACT;
    o->arithmetic = TETSIM_ARITH_FAST_F32;
    o->iters = 1;
    o->deterministic = 1;
    o->referenceTableBug = 1;
    o->reorder = 1;
    o->clusterSize = 256;
    o->trackVolError = -1;
    o->device = -1;
    o->rank = 0;
    o->worldSize = 1;
    o->exchange = 0;
    o->stream = nullptr;
    o->ncclUniqueId = nullptr;
}

int tetsim_create(const float *verts, int32_t numVerts, const int32_t *tetIds, int32_t numTets,
                  const TetSimParams *params, const TetSimOptions *options, tetsim_t **out) {
    if (!out) return fail(TETSIM_E_INVALID, "out is null");
    *out = nullptr;
    if (numVerts < 0 || numTets < 0 || (numVerts > 0 && !verts) || (numTets > 0 && !tetIds))
        return fail(TETSIM_E_INVALID, "null or negative-sized mesh arrays");
    if ((int64_t)numTets * 4 > INT32_MAX) return fail(TETSIM_E_INVALID, "more than 2^29 tets");
    TetSimOptions opt;
    tetsim_default_options(&opt);
    if (options) opt = *options;
    TetSimParams prm;
    tetsim_default_params(&prm);
    if (params) prm = *params;
    if (opt.solver < TETSIM_NH_GS_EXACT || opt.solver > TETSIM_POLAR_JACOBI) return fail(TETSIM_E_INVALID, "unknown solver");
    if (opt.arithmetic != TETSIM_ARITH_FAST_F32 && opt.arithmetic != TETSIM_ARITH_BITEXACT)
        return fail(TETSIM_E_INVALID, "unknown arithmetic");
    if (opt.iters < 1) return fail(TETSIM_E_INVALID, "iters must be >= 1");
    if (opt.clusterSize != 32 && opt.clusterSize != 64 && opt.clusterSize != 128 && opt.clusterSize != 256 && opt.clusterSize != 512)
        return fail(TETSIM_E_INVALID, "clusterSize must be 32, 64 (warp tiles) or 128, 256, 512 (CTA tiles)");
    if (opt.worldSize < 1 || opt.rank < 0 || opt.rank >= opt.worldSize) return fail(TETSIM_E_INVALID, "bad rank/worldSize");
    const bool clustered = opt.solver == TETSIM_NH_JACOBI && opt.arithmetic == TETSIM_ARITH_FAST_F32;
    if (opt.worldSize > 1 && !clustered)
        return fail(TETSIM_E_STATE, "worldSize > 1 is only supported by the FAST_F32 Jacobi solver; shard independent bodies across processes instead");
    if (opt.exchange < 0 || opt.exchange > 2) return fail(TETSIM_E_INVALID, "exchange must be 0 (all-reduce), 1 (neighbour exchange) or 2 (peer memory)");
    if (opt.exchange == 2 && opt.worldSize > kPeerMaxPeers) return fail(TETSIM_E_INVALID, "the peer-memory exchange supports at most 64 ranks");
    if (opt.worldSize > 1 && opt.exchange != 2 && !opt.ncclUniqueId) return fail(TETSIM_E_INVALID, "worldSize > 1 needs ncclUniqueId");
    for (int64_t c = 0; c < 4 * (int64_t)numTets; c++)
        if (tetIds[c] < 0 || tetIds[c] >= numVerts) return fail(TETSIM_E_INVALID, "tet " + std::to_string(c / 4) + " references vertex " + std::to_string(tetIds[c]) + " outside [0, numVerts)");
    for (int e = 0; e < numTets; e++) {
        const int32_t *t = tetIds + 4 * (size_t)e;
        if (t[0] == t[1] || t[0] == t[2] || t[0] == t[3] || t[1] == t[2] || t[1] == t[3] || t[2] == t[3])
            return fail(TETSIM_E_INVALID, "tet " + std::to_string(e) + " repeats a vertex");
    }

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(TETSIM_E_CUDA, std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0") + " (this library has no CPU path)");
    int dev = opt.device;
    if (dev < 0) CK(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(TETSIM_E_INVALID, "device ordinal out of range");
    int ccMajor = 0;
    CK(cudaDeviceGetAttribute(&ccMajor, cudaDevAttrComputeCapabilityMajor, dev));
    if (ccMajor != 10) return fail(TETSIM_E_CUDA, "device " + std::to_string(dev) + " is not compute capability 10.x (kernels are built for sm_100a only)");

    DeviceGuard guard(dev);
    tetsim *h = new (std::nothrow) tetsim();
    if (!h) return fail(TETSIM_E_NOMEM, "out of host memory");
    struct Cleanup { tetsim *h; bool armed = true; ~Cleanup() { if (armed) tetsim_destroy(h); } } cleanup{h};
    h->opt = opt; h->params = prm; h->N = numVerts; h->M = numTets; h->device = dev;
    { const char *tr = getenv("TETSIM_TRACE"); h->trace = tr && tr[0] == '1'; }
    h->KX = exact_kernels();
    h->K = opt.arithmetic == TETSIM_ARITH_BITEXACT ? exact_kernels() : fast_kernels();
    if (opt.stream) h->stream = (cudaStream_t)opt.stream;
    else { CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->ownStream = true; }
    h->trackVol = opt.trackVolError < 0 ? (opt.solver == TETSIM_NH_GS_EXACT || opt.solver == TETSIM_NH_GS_COLOR || (opt.solver == TETSIM_NH_JACOBI && !clustered))
                                        : opt.trackVolError != 0;
    if (opt.solver == TETSIM_POLAR_JACOBI) h->trackVol = false;
    cudaStream_t s = h->stream;

    std::vector<float> hv(verts, verts + 3 * (size_t)numVerts);
    std::vector<int> ht(tetIds, tetIds + 4 * (size_t)numTets);

    // ---- initPhysics on the device, caller numbering (src/Softbody.js:60-87) ----
    {
        std::vector<float4> hx((size_t)numVerts);
        for (int v = 0; v < numVerts; v++) hx[v] = make_float4(hv[3 * (size_t)v], hv[3 * (size_t)v + 1], hv[3 * (size_t)v + 2], 0.0f);
        DevBuf<float4> gx;
        DevBuf<int4> gids;
        DevBuf<double> pm;
        DevBuf<int> gcs, gce;
        CK(gx.upload(hx, s));
        CK(gids.alloc((size_t)numTets));
        if (numTets) CK(cudaMemcpyAsync(gids.p, ht.data(), (size_t)numTets * sizeof(int4), cudaMemcpyHostToDevice, s));
        CornerTable ct = build_corner_table(numVerts, numTets, ht.data());
        h->maxValence = ct.maxValence;
        CK(gcs.upload(ct.start, s));
        CK(gce.upload(ct.ent, s));
        CK(h->Q9.alloc(9 * (size_t)numTets)); CK(h->irv.alloc((size_t)numTets)); CK(pm.alloc((size_t)numTets));
        h->KX->init_tets(s, numTets, gx.p, gids.p, prm.density, h->Q9.p, h->irv.p, pm.p);
        h->KX->init_mass(s, numVerts, gcs.p, gce.p, pm.p, gx.p);
        CK(cudaGetLastError());
        // pull invMass back once: it is part of every solver's vertex record and of get_rest
        CK(cudaMemcpyAsync(hx.data(), gx.p, (size_t)numVerts * sizeof(float4), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        std::vector<float> im((size_t)numVerts);
        for (int v = 0; v < numVerts; v++) im[v] = hx[v].w;
        CK(h->invMass.upload(im, s));

        // ---- solver-specific build (decides the internal vertex numbering) ----
        int rc = TETSIM_OK;
        {   // polar: the tiled kernels in FAST arithmetic unless a particle has more corners than the reference's 36 table slots
            const char *csr = getenv("TETSIM_POLAR_CSR");
            h->polarTiled = opt.solver == TETSIM_POLAR_JACOBI && opt.arithmetic == TETSIM_ARITH_FAST_F32 && h->maxValence <= 36 &&
                            numTets > 0 && !(csr && csr[0] == '1');
        }
        if (opt.solver == TETSIM_NH_GS_EXACT || opt.solver == TETSIM_NH_GS_COLOR) rc = build_gs(h, hv, ht);
        else if (h->polarTiled) rc = plan_polar_tiles(h, hv, ht);
        else if (clustered) {
            if (opt.worldSize > 1 && opt.exchange != 2) {
                if (!g_nccl.load()) return fail(TETSIM_E_NCCL, g_nccl.why);
                NcclUniqueId id;
                memcpy(&id, opt.ncclUniqueId, sizeof(id));
                int nrc = g_nccl.CommInitRank(&h->comm, opt.worldSize, id, opt.rank);
                if (nrc != 0) return fail(TETSIM_E_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(nrc));
                CK(cudaStreamCreateWithFlags(&h->commStream, cudaStreamNonBlocking));
                CK(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming));
            }
            rc = build_jacobi_cluster(h, hv, ht);
        }
        if (rc != TETSIM_OK) return rc;

        // vertex state in internal numbering
        const bool identity = h->h_vertId.empty();
        h->nInt = identity ? numVerts : (int)h->h_vertId.size();
        std::vector<float4> hxi((size_t)h->nInt);
        // replicas of rank-shared vertices this rank's tets never touch (h_vertId < 0, worldSize >= 3 with the neighbour
        // and peer exchanges) are not maintained: NaN records that no tile reads; pack/unpack/nearest skip them through vertId
        const float qnan = std::nanf("");
        for (int i = 0; i < h->nInt; i++) {
            const int c = identity ? i : h->h_vertId[i];
            hxi[i] = c >= 0 ? hx[c] : make_float4(qnan, qnan, qnan, 0.0f);
        }
        CK(h->x4.upload(hxi, s));
        CK(h->prev4.upload(hxi, s));
        CK(h->vel4.alloc((size_t)h->nInt));
        if (h->nInt) CK(cudaMemsetAsync(h->vel4.p, 0, h->vel4.bytes(), s));
        if (!identity) CK(h->vertId.upload(h->h_vertId, s));
        // tet ids in internal numbering (caller tet order) for the gather solvers / skinning
        if (!clustered || opt.worldSize == 1) {
            std::vector<int> c2i;
            const int *map = nullptr;
            if (!identity) {
                c2i.assign((size_t)numVerts, -1);
                for (int i = 0; i < h->nInt; i++) if (h->h_vertId[i] >= 0) c2i[h->h_vertId[i]] = i;
                map = c2i.data();
            }
            std::vector<int4> hi((size_t)numTets);
            for (int e = 0; e < numTets; e++) {
                const int *t = ht.data() + 4 * (size_t)e;
                hi[e] = map ? make_int4(map[t[0]], map[t[1]], map[t[2]], map[t[3]]) : make_int4(t[0], t[1], t[2], t[3]);
            }
            CK(h->ids.upload(hi, s));
            CK(cudaStreamSynchronize(s));
        }
        if (opt.solver == TETSIM_NH_JACOBI && !clustered) {
            h->cStart.p = gcs.p; h->cStart.n = gcs.n; gcs.p = nullptr;  // adopt the corner table
            h->cEnt.p = gce.p; h->cEnt.n = gce.n; gce.p = nullptr;
            rc = build_jacobi_gather(h, ht);
        } else if (opt.solver == TETSIM_POLAR_JACOBI) {
            rc = build_polar(h, ht);
        }
        CK(cudaStreamSynchronize(s));
        gx.release(); gids.release(); pm.release(); gcs.release(); gce.release();
        if (rc != TETSIM_OK) return rc;
    }
    CK(h->sp.alloc(1));
    CK(h->volOut.alloc(1));
    CK(h->grabP.alloc(3)); CK(h->grabOut.alloc(1)); CK(h->grabScratch.alloc(1));
    CK(cudaStreamSynchronize(s));
    cleanup.armed = false;
    *out = h;
    return TETSIM_OK;
}

void tetsim_destroy(tetsim_t *h) {
    if (!h) return;
    DeviceGuard g(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->h2dStream) { cudaStreamSynchronize(h->h2dStream); cudaStreamDestroy(h->h2dStream); }
    if (h->d2hStream) { cudaStreamSynchronize(h->d2hStream); cudaStreamDestroy(h->d2hStream); }
    for (int b = 0; b < 2; b++)
        for (cudaEvent_t e : {h->evIn[b], h->evInFree[b], h->evPacked[b], h->evOut[b]}) if (e) cudaEventDestroy(e);
    for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);
    if (h->trace) trace_report(h);
    for (void *m : h->peerOpened) cudaIpcCloseMemHandle(m);
    if (h->comm) g_nccl.CommDestroy(h->comm);
    if (h->evFork) cudaEventDestroy(h->evFork);
    if (h->evJoin) cudaEventDestroy(h->evJoin);
    if (h->commStream) cudaStreamDestroy(h->commStream);
    DevBuf<float4> *f4[] = {&h->x4, &h->prev4, &h->vel4, &h->A, &h->B, &h->C, &h->dx, &h->part, &h->acc, &h->bsum, &h->rest, &h->quat, &h->visV};
    for (auto *b : f4) b->release();
    DevBuf<int> *i1[] = {&h->vertId, &h->cStart, &h->cEnt, &h->order, &h->levelStart, &h->vpStart, &h->vpSlot, &h->tStart, &h->tEnt, &h->grabOut, &h->visTri, &h->vtStart, &h->vtEnt};
    for (auto *b : i1) b->release();
    h->visRestNrm.release(); h->tetRecord.release(); h->tileVol.release();
    DevBuf<float> *f1[] = {&h->Q9, &h->irv, &h->invMass, &h->invVal, &h->stageIn[0], &h->stageIn[1], &h->stageOut[0], &h->stageOut[1], &h->visPos, &h->visNrm};
    for (auto *b : f1) b->release();
    h->ids.release(); h->I.release(); h->bodies.release(); h->volTerm.release(); h->volOut.release();
    h->tileTets.release(); h->tileMeta.release(); h->metaOff.release();
    h->hxSendIdx.release(); h->hxSrcStart.release(); h->hxSrc.release(); h->hxSend.release(); h->hxRecv.release(); h->sp.release(); h->grabP.release(); h->grabScratch.release();
    if (h->ownStream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int tetsim_simulate(tetsim_t *h, double dt, const TetSimParams *params) { return run_substeps(h, dt, 1, params); }

int tetsim_step(tetsim_t *h, double frameDt, int32_t numSubsteps, const TetSimParams *params) {
    if (numSubsteps < 1) return fail(TETSIM_E_INVALID, "numSubsteps must be >= 1");
    return run_substeps(h, frameDt / (double)numSubsteps, numSubsteps, params);  // src/main.js:79
}

int tetsim_synchronize(tetsim_t *h) {
    if (!h) return fail(TETSIM_E_INVALID, "null handle");
    DeviceGuard g(h->device);
    CK(cudaStreamSynchronize(h->stream));
    if (h->d2hStream) CK(cudaStreamSynchronize(h->d2hStream));
    h->pendingOut = -1;
    if (h->trace) trace_report(h);
    return peer_check(h);
}

int tetsim_get_positions(tetsim_t *h, float *out) { return h ? fetch3(h, h->x4.p, out) : fail(TETSIM_E_INVALID, "null handle"); }
int tetsim_get_prev_positions(tetsim_t *h, float *out) { return h ? fetch3(h, h->prev4.p, out) : fail(TETSIM_E_INVALID, "null handle"); }
int tetsim_get_velocities(tetsim_t *h, float *out) { return h ? fetch3(h, h->vel4.p, out) : fail(TETSIM_E_INVALID, "null handle"); }

int tetsim_get_resident(tetsim_t *h, uint8_t *out) {
    if (!h || !out) return fail(TETSIM_E_INVALID, "null argument");
    if (h->h_vertId.empty()) { memset(out, 1, (size_t)h->N); return TETSIM_OK; }
    memset(out, 0, (size_t)h->N);
    for (int v : h->h_vertId) if (v >= 0) out[v] = 1;
    return TETSIM_OK;
}

int tetsim_set_state(tetsim_t *h, const float *pos, const float *prevPos, const float *vel) { return upload_state(h, pos, prevPos, vel, false); }
int tetsim_set_state_resident(tetsim_t *h, const float *pos, const float *prevPos, const float *vel) { return upload_state(h, pos, prevPos, vel, true); }

int tetsim_get_positions_async(tetsim_t *h, float *out) { return h ? download_begin(h, h->x4.p, out, false) : fail(TETSIM_E_INVALID, "null handle"); }
int tetsim_get_positions_resident(tetsim_t *h, float *out) { return h ? fetch3(h, h->x4.p, out, true) : fail(TETSIM_E_INVALID, "null handle"); }
int tetsim_get_positions_resident_async(tetsim_t *h, float *out) { return h ? download_begin(h, h->x4.p, out, true) : fail(TETSIM_E_INVALID, "null handle"); }
int tetsim_wait_positions(tetsim_t *h) { return download_wait(h); }

int tetsim_get_resident_ids(tetsim_t *h, int32_t *out) {
    if (!h || !out) return fail(TETSIM_E_INVALID, "null argument");
    if (h->h_vertId.empty()) { for (int i = 0; i < h->nInt; i++) out[i] = i; }
    else std::copy(h->h_vertId.begin(), h->h_vertId.end(), out);
    return TETSIM_OK;
}

int tetsim_get_rest(tetsim_t *h, float *invRestPose, float *invRestVolume, float *invMass) {
    if (!h) return fail(TETSIM_E_INVALID, "null handle");
    DeviceGuard g(h->device);
    CK(cudaStreamSynchronize(h->stream));
    if (invRestPose && h->M) CK(cudaMemcpy(invRestPose, h->Q9.p, h->Q9.bytes(), cudaMemcpyDeviceToHost));
    if (invRestVolume && h->M) CK(cudaMemcpy(invRestVolume, h->irv.p, h->irv.bytes(), cudaMemcpyDeviceToHost));
    if (invMass && h->N) CK(cudaMemcpy(invMass, h->invMass.p, h->invMass.bytes(), cudaMemcpyDeviceToHost));
    return TETSIM_OK;
}

int tetsim_get_vol_error(tetsim_t *h, double *out) {
    if (!h || !out) return fail(TETSIM_E_INVALID, "null argument");
    if (!h->trackVol) return fail(TETSIM_E_STATE, "volError tracking is off for this handle (TetSimOptions.trackVolError)");
    DeviceGuard g(h->device);
    if (h->clustered) {
        double sum = 0.0;
        CK(cudaMemcpyAsync(&sum, h->volTerm.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        *out = sum / (double)h->M;  // this rank's tets over the GLOBAL tet count
        return TETSIM_OK;
    }
    launch_sum_sequential(h->stream, h->M, h->volTerm.p, h->volOut.p);
    CK(cudaMemcpyAsync(out, h->volOut.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return TETSIM_OK;
}

int tetsim_get_polar_state(tetsim_t *h, float *rest12, float *quat4) {
    if (!h) return fail(TETSIM_E_INVALID, "null handle");
    if (h->opt.solver != TETSIM_POLAR_JACOBI) return fail(TETSIM_E_STATE, "not a POLAR_JACOBI handle");
    DeviceGuard g(h->device);
    CK(cudaStreamSynchronize(h->stream));
    if (h->polarTiled) {  // tile-major planes -> the caller's tet order
        const ClusterPlan &P = h->plan;
        const size_t T = (size_t)P.T;
        std::vector<unsigned char> blocks(h->tileTets.bytes());
        if (!blocks.empty()) CK(cudaMemcpy(blocks.data(), h->tileTets.p, blocks.size(), cudaMemcpyDeviceToHost));
        for (size_t r = 0; r < P.recordTet.size(); r++) {
            const int e = P.recordTet[r];
            if (e < 0) continue;
            const unsigned char *tb = blocks.data() + (r / T) * T * 80;
            const size_t t = r % T;
            if (rest12) {
                const float *R0 = reinterpret_cast<const float *>(tb) + 4 * t, *R1 = reinterpret_cast<const float *>(tb + T * 16) + 4 * t,
                            *R2 = reinterpret_cast<const float *>(tb + T * 32) + 4 * t;
                float *o = rest12 + 12 * (size_t)e;
                for (int k = 0; k < 4; k++) { o[k] = R0[k]; o[4 + k] = R1[k]; o[8 + k] = R2[k]; }
            }
            if (quat4) memcpy(quat4 + 4 * (size_t)e, tb + T * 48 + 16 * t, 16);
        }
        return TETSIM_OK;
    }
    if (rest12) {
        std::vector<float4> r(4 * (size_t)h->M);
        if (h->M) CK(cudaMemcpy(r.data(), h->rest.p, h->rest.bytes(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < r.size(); i++) { rest12[3 * i] = r[i].x; rest12[3 * i + 1] = r[i].y; rest12[3 * i + 2] = r[i].z; }
    }
    if (quat4 && h->M) CK(cudaMemcpy(quat4, h->quat.p, h->quat.bytes(), cudaMemcpyDeviceToHost));
    return TETSIM_OK;
}

// Nearest resident vertex (first strict minimum of the f64 squared distance, src/Softbody.js:279-291) without touching
// the grab state.  On a multi-GPU handle every rank searches the vertices it maintains; the caller reduces (d2, id) over
// the ranks (smallest d2, ties to the smallest id) and hands the winner to every rank with tetsim_set_grab.
int tetsim_nearest_vertex(tetsim_t *h, const double p[3], int32_t *outId, double *outD2) {
    if (!h || !p || !outId) return fail(TETSIM_E_INVALID, "null argument");
    DeviceGuard g(h->device);
    CK(cudaMemcpyAsync(h->grabP.p, p, 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_nearest_vertex(h->stream, h->nInt, h->x4.p, h->vertId.p, h->grabP.p, h->grabOut.p, h->grabScratch.p);
    int id = -1;
    unsigned long long bits = 0;
    CK(cudaMemcpyAsync(&id, h->grabOut.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&bits, h->grabScratch.p, sizeof(bits), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (id < 0 || id >= h->N) id = -1;  // no finite distance: grabId stays -1 like the reference loop
    *outId = id;
    if (outD2) { double d2; memcpy(&d2, &bits, sizeof(d2)); *outD2 = id >= 0 ? d2 : INFINITY; }
    return TETSIM_OK;
}

int tetsim_set_grab(tetsim_t *h, int32_t grabId, const double p[3]) {
    if (!h || !p) return fail(TETSIM_E_INVALID, "null argument");
    if (grabId < -1 || grabId >= h->N) return fail(TETSIM_E_INVALID, "grabId outside [-1, numVerts)");
    h->grabId = grabId;
    memcpy(h->grabPos, p, sizeof(h->grabPos));
    return TETSIM_OK;
}

int tetsim_start_grab(tetsim_t *h, const double p[3], int32_t *outGrabId) {
    if (!h || !p) return fail(TETSIM_E_INVALID, "null argument");
    if (h->opt.worldSize > 1)
        return fail(TETSIM_E_STATE, "startGrab on a multi-GPU handle: every rank sees only its own vertices -- call tetsim_nearest_vertex on every rank, "
                                    "reduce (d2, id) over the ranks and pass the winner to tetsim_set_grab on every rank");
    int32_t id = -1;
    if (int rc = tetsim_nearest_vertex(h, p, &id, nullptr)) return rc;
    h->grabId = id;
    memcpy(h->grabPos, p, sizeof(h->grabPos));
    if (outGrabId) *outGrabId = id;
    return TETSIM_OK;
}

int tetsim_move_grabbed(tetsim_t *h, const double p[3]) {
    if (!h || !p) return fail(TETSIM_E_INVALID, "null argument");
    memcpy(h->grabPos, p, sizeof(h->grabPos));
    return TETSIM_OK;
}

int tetsim_end_grab(tetsim_t *h) {
    if (!h) return fail(TETSIM_E_INVALID, "null handle");
    h->grabId = -1;
    return TETSIM_OK;
}

// 64-bit content hash (word-wise multiply-xorshift; ~1 GB/s per core is plenty for a 0.5 MB surface mesh per frame)
static uint64_t content_hash(const void *data, size_t bytes) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    uint64_t hsh = 0x9e3779b97f4a7c15ull ^ bytes;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) { uint64_t w; memcpy(&w, p + i, 8); hsh = (hsh ^ w) * 0xff51afd7ed558ccdull; hsh ^= hsh >> 32; }
    for (; i < bytes; i++) { hsh = (hsh ^ p[i]) * 0x100000001b3ull; }
    return hsh;
}

// The embedded surface mesh on the device, re-uploaded whenever its CONTENT changes.
static int ensure_vis(tetsim *h, const float *visVerts, int32_t numVis) {
    const uint64_t hv = content_hash(visVerts, (size_t)numVis * 16);
    if (numVis == h->visN && hv == h->visHashV) return TETSIM_OK;
    for (int i = 0; i < numVis; i++) {
        float t = visVerts[4 * (size_t)i];
        if (!(t >= 0.0f && t < (float)h->M)) return fail(TETSIM_E_INVALID, "visVerts tet index out of range");
    }
    CK(h->visV.alloc((size_t)numVis));
    if (numVis) CK(cudaMemcpyAsync(h->visV.p, visVerts, (size_t)numVis * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    CK(h->visPos.alloc(3 * (size_t)std::max(numVis, 1)));
    CK(h->visNrm.alloc(3 * (size_t)std::max(numVis, 1)));
    h->visHashV = hv; h->visN = numVis; h->visT = -1; h->visHashN = 0;
    return TETSIM_OK;
}

int tetsim_skin_gpu(tetsim_t *h, const float *visVerts, int32_t numVis, const float *restNormals, float *outPos, float *outNormals) {
    if (!h || !visVerts || !outPos || numVis < 0) return fail(TETSIM_E_INVALID, "null argument");
    if (h->opt.solver != TETSIM_POLAR_JACOBI) return fail(TETSIM_E_STATE, "tetsim_skin_gpu rotates normals by the tets' quaternions: POLAR_JACOBI handles only");
    DeviceGuard g(h->device);
    cudaStream_t s = h->stream;
    if (int rc = ensure_vis(h, visVerts, numVis)) return rc;
    const bool wantN = outNormals && restNormals;
    if (wantN) {
        const uint64_t hn = content_hash(restNormals, (size_t)numVis * 12);
        if (hn != h->visHashN || h->visRestNrm.n != 3 * (size_t)std::max(numVis, 1)) {
            CK(h->visRestNrm.alloc(3 * (size_t)std::max(numVis, 1)));
            if (numVis) CK(cudaMemcpyAsync(h->visRestNrm.p, restNormals, (size_t)numVis * 12, cudaMemcpyHostToDevice, s));
            h->visHashN = hn;
        }
        if (h->polarTiled && h->tetRecord.n == 0) {   // caller tet -> record slot of the tile blocks
            std::vector<int> rec((size_t)std::max(h->M, 1), 0);
            for (size_t r = 0; r < h->plan.recordTet.size(); r++) if (h->plan.recordTet[r] >= 0) rec[h->plan.recordTet[r]] = (int)r;
            CK(h->tetRecord.upload(rec, s));
        }
    }
    h->K->skin_polar(s, numVis, h->visV.p, h->ids.p, h->x4.p, h->quat.p, h->tileTets.p, h->polarTiled ? h->tetRecord.p : nullptr,
                     h->plan.T, h->visRestNrm.p, h->visPos.p, wantN ? h->visNrm.p : nullptr);
    if (numVis) CK(cudaMemcpyAsync(outPos, h->visPos.p, 3 * (size_t)numVis * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (wantN && numVis) CK(cudaMemcpyAsync(outNormals, h->visNrm.p, 3 * (size_t)numVis * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    return TETSIM_OK;
}

int tetsim_skin(tetsim_t *h, const float *visVerts, int32_t numVis, const int32_t *triIds, int32_t numTris,
                float *outPos, float *outNormals) {
    if (!h || !visVerts || !outPos || numVis < 0) return fail(TETSIM_E_INVALID, "null argument");
    if (h->ids.n == 0 && h->M > 0) return fail(TETSIM_E_STATE, "skinning is not available on a multi-GPU handle");
    DeviceGuard g(h->device);
    cudaStream_t s = h->stream;
    const bool wantN = outNormals && triIds && numTris > 0;
    if (int rc = ensure_vis(h, visVerts, numVis)) return rc;
    const uint64_t ht = wantN ? content_hash(triIds, (size_t)numTris * 12) : 0;
    if (wantN && (ht != h->visHashT || numTris != h->visT)) {
        // vertex -> triangles, each (vertex, triangle) pair once, ascending triangle order
        std::vector<int> start((size_t)numVis + 1, 0), ent;
        for (int t = 0; t < numTris; t++) {
            const int *v = triIds + 3 * (size_t)t;
            for (int k = 0; k < 3; k++) {
                if (v[k] < 0 || v[k] >= numVis) return fail(TETSIM_E_INVALID, "triangle index out of range");
                bool dup = false;
                for (int j = 0; j < k; j++) dup = dup || v[j] == v[k];
                if (!dup) start[(size_t)v[k] + 1]++;
            }
        }
        for (int v = 0; v < numVis; v++) start[v + 1] += start[v];
        ent.resize((size_t)start[numVis]);
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int t = 0; t < numTris; t++) {
            const int *v = triIds + 3 * (size_t)t;
            for (int k = 0; k < 3; k++) {
                bool dup = false;
                for (int j = 0; j < k; j++) dup = dup || v[j] == v[k];
                if (!dup) ent[fill[v[k]]++] = t;
            }
        }
        std::vector<int> tri(triIds, triIds + 3 * (size_t)numTris);
        CK(h->visTri.upload(tri, s));
        CK(h->vtStart.upload(start, s));
        CK(h->vtEnt.upload(ent, s));
        CK(cudaStreamSynchronize(s));
        h->visHashT = ht; h->visT = numTris;
    }
    h->K->skin(s, numVis, h->visV.p, h->ids.p, h->x4.p, h->visPos.p);
    if (numVis) CK(cudaMemcpyAsync(outPos, h->visPos.p, 3 * (size_t)numVis * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (wantN) {
        h->KX->normals(s, numVis, h->visPos.p, h->visTri.p, h->vtStart.p, h->vtEnt.p, h->visNrm.p);
        if (numVis) CK(cudaMemcpyAsync(outNormals, h->visNrm.p, 3 * (size_t)numVis * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    return TETSIM_OK;
}

int tetsim_get_info(tetsim_t *h, TetSimInfo *info) {
    if (!h || !info) return fail(TETSIM_E_INVALID, "null argument");
    memset(info, 0, sizeof(*info));
    info->numVerts = h->N; info->numTets = h->M;
    info->solver = h->opt.solver; info->arithmetic = h->opt.arithmetic; info->iters = h->opt.iters;
    info->numLevels = h->numLevels; info->maxLevelSize = h->maxLevelSize;
    info->numComponents = h->numComponents; info->bodyKernel = h->bodyKernel ? 1 : 0;
    info->numClusters = h->plan.numClusters; info->clusterSize = (h->clustered || h->polarTiled) ? h->plan.T : 0;
    info->localTets = h->clustered ? h->plan.localTets : h->M;
    info->localVerts = h->nInt;
    info->boundaryVerts = h->plan.numBoundary;
    info->maxValence = h->maxValence;
    info->launchesPerSubstep = h->launchesPerSubstep;
    info->deviceBytes = h->deviceBytes();
    info->sumLocalVerts = (int64_t)h->plan.clVerts.size();
    info->kernelLaunches = h->totalLaunches;
    info->tileMetaBytes = (int64_t)h->plan.tileMeta.size();
    info->maxTileVerts = h->clustered ? h->plan.maxTileVerts : 0;
    info->boundaryTiles = h->plan.numBoundaryTiles;
    return TETSIM_OK;
}

int tetsim_time_kernel(tetsim_t *h, int32_t reps, double *msPerLaunch, int64_t *algorithmicBytes) {
    if (!h || !msPerLaunch || reps < 1) return fail(TETSIM_E_INVALID, "bad argument");
    if (h->polarTiled) {
        // k_polar_tiles advances the per-tet goal corners and quaternions: snapshot them, time, put them back
        DeviceGuard g(h->device);
        cudaStream_t s = h->stream;
        DevBuf<unsigned char> keep;
        CK(keep.alloc(h->tileTets.n));
        CK(cudaMemcpyAsync(keep.p, h->tileTets.p, h->tileTets.bytes(), cudaMemcpyDeviceToDevice, s));
        const PolarTileArgs pa = polar_tile_args(h);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        // every timed launch starts from the same state (a second launch on unchanged positions would find its rotation
        // already extracted and leave the iteration early): restore the snapshot, untimed, before each one
        launch_polar_tiles(s, h->plan.T, pa);
        float ms = 0.f;
        for (int r = 0; r < reps; r++) {
            CK(cudaMemcpyAsync(h->tileTets.p, keep.p, h->tileTets.bytes(), cudaMemcpyDeviceToDevice, s));
            CK(cudaEventRecord(e0, s));
            launch_polar_tiles(s, h->plan.T, pa);
            CK(cudaEventRecord(e1, s));
            CK(cudaEventSynchronize(e1));
            float one = 0.f;
            CK(cudaEventElapsedTime(&one, e0, e1));
            ms += one;
        }
        CK(cudaMemcpyAsync(h->tileTets.p, keep.p, h->tileTets.bytes(), cudaMemcpyDeviceToDevice, s));
        CK(cudaStreamSynchronize(s));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        CK(cudaGetLastError());
        *msPerLaunch = (double)ms / reps;
        if (algorithmicBytes) *algorithmicBytes = 148ll * h->M + 32ll * h->nInt;   // SURVEY.md section 8(d), polar variant
        return TETSIM_OK;
    }
    if (!h->clustered) return fail(TETSIM_E_STATE, "tetsim_time_kernel times the tile kernels only (FAST Jacobi or FAST polar handles)");
    DeviceGuard g(h->device);
    cudaStream_t s = h->stream;
    const ClusterPlan &P = h->plan;
    TileArgs ca = tile_args(h);
    if (const char *dbg = getenv("TETSIM_TILE_DEBUG")) ca.debugSkip = atoi(dbg);
    DevBuf<float4> scratch;  // atomic-flush handles have no partial-sum array: give the kernel one
    if (!ca.part) { CK(scratch.alloc(P.clVerts.size())); ca.part = scratch.p; }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch_jacobi_tiles(s, P.T, ca);  // warm-up
    CK(cudaEventRecord(e0, s));
    for (int r = 0; r < reps; r++) launch_jacobi_tiles(s, P.T, ca);
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    scratch.release();
    CK(cudaGetLastError());
    *msPerLaunch = (double)ms / reps;
    // 56 B per tet + 32 B per vertex per launch
    if (algorithmicBytes) *algorithmicBytes = 56ll * P.localTets + 32ll * h->nInt;
    return TETSIM_OK;
}

int tetsim_nccl_unique_id(void *out128) {
    if (!out128) return fail(TETSIM_E_INVALID, "null argument");
    if (!g_nccl.load()) return fail(TETSIM_E_NCCL, g_nccl.why);
    NcclUniqueId id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return fail(TETSIM_E_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
    memcpy(out128, &id, sizeof(id));
    return TETSIM_OK;
}

int tetsim_get_ipc_handle(tetsim_t *h, void *outBlob) {
    if (!h || !outBlob) return fail(TETSIM_E_INVALID, "null argument");
    if (!h->peer) return fail(TETSIM_E_STATE, "tetsim_get_ipc_handle needs a multi-GPU handle created with exchange = 2");
    DeviceGuard g(h->device);
    PeerBlob b;
    memset(&b, 0, sizeof(b));
    cudaIpcMemHandle_t ipc;
    CK(cudaIpcGetMemHandle(&ipc, h->peerBuf.p));
    memcpy(b.ipc, &ipc, sizeof(ipc));
    b.offset = allocation_offset(h->peerBuf.p);
    b.rawPtr = (uint64_t)(uintptr_t)h->peerBuf.p;
    b.pid = (int32_t)getpid(); b.device = h->device; b.rank = h->opt.rank;
    b.total = (int32_t)h->plan.hxSendIdx.size(); b.numPeers = (int32_t)h->plan.hxPeers.size();
    b.magic = kPeerMagic;
    memcpy(outBlob, &b, sizeof(b));
    return TETSIM_OK;
}

int tetsim_set_peers(tetsim_t *h, const void *blobs) {
    if (!h || !blobs) return fail(TETSIM_E_INVALID, "null argument");
    if (!h->peer) return fail(TETSIM_E_STATE, "tetsim_set_peers needs a multi-GPU handle created with exchange = 2");
    if (h->peersSet) return fail(TETSIM_E_STATE, "peers are already set");
    DeviceGuard g(h->device);
    const ClusterPlan &P = h->plan;
    std::vector<unsigned char *> base(std::max<size_t>(P.hxPeers.size(), 1), nullptr);
    for (size_t qi = 0; qi < P.hxPeers.size(); qi++) {
        const int q = P.hxPeers[qi];
        PeerBlob b;
        memcpy(&b, (const unsigned char *)blobs + (size_t)q * sizeof(PeerBlob), sizeof(b));
        if (b.magic != kPeerMagic || b.rank != q)
            return fail(TETSIM_E_INVALID, "blob " + std::to_string(q) + " is not rank " + std::to_string(q) + "'s tetsim_get_ipc_handle output");
        // both sides derive each other's layout from the same global partition: cross-check it
        if (b.total != P.pxRemoteTotal[qi])
            return fail(TETSIM_E_STATE, "rank " + std::to_string(q) + " reports " + std::to_string(b.total) + " exchange entries, this rank's plan expects " + std::to_string(P.pxRemoteTotal[qi]) + " (different mesh or options?)");
        if (b.pid == (int32_t)getpid()) {  // a handle of this process: its pointer is valid here
            if (b.device != h->device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, h->device, b.device));
                if (!can) return fail(TETSIM_E_CUDA, "device " + std::to_string(h->device) + " cannot access device " + std::to_string(b.device));
                cudaError_t pe = cudaDeviceEnablePeerAccess(b.device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CK(pe);
                cudaGetLastError();
            }
            base[qi] = (unsigned char *)(uintptr_t)b.rawPtr;
        } else {
            cudaIpcMemHandle_t ipc;
            memcpy(&ipc, b.ipc, sizeof(ipc));
            void *mapped = nullptr;
            CK(cudaIpcOpenMemHandle(&mapped, ipc, cudaIpcMemLazyEnablePeerAccess));
            h->peerOpened.push_back(mapped);
            base[qi] = (unsigned char *)mapped + b.offset;
        }
    }
    CK(cudaMemcpyAsync(h->peerBase.p, base.data(), base.size() * sizeof(unsigned char *), cudaMemcpyHostToDevice, h->stream));
    if (h->peerFused) {  // one self-contained 32-byte record per pushing slot (PeerArgs::pushRec)
        std::vector<uint4> rec(2 * std::max<size_t>(P.pxSlotIdx.size(), 1), make_uint4(0xffffffffu, 0u, 0u, 0u));
        for (int b = 0; b < P.numBoundary; b++) {
            const int id = P.numInterior + b, n = P.vpStart[id + 1] - P.vpStart[id];
            const int t0 = P.pxStart[b], sharers = P.pxStart[b + 1] - t0;
            if (n == 0 || sharers == 0) continue;
            for (int i = 0; i < n; i++) {
                const size_t slot = (size_t)P.vpSlot[P.vpStart[id] + i];
                uint4 r0 = make_uint4((unsigned)i | (unsigned)(n - 1) << 8 | (unsigned)sharers << 16, (unsigned)b, 0u, 0u), r1 = make_uint4(0u, 0u, 0u, 0u);
                for (int k = 0; k < 2 && k < sharers; k++) {
                    const int q = P.pxPeer[t0 + k];
                    const unsigned long long addr = (unsigned long long)(uintptr_t)(base[q] + kPeerRecvOff) +
                                                    ((unsigned long long)P.pxEntry[t0 + k] * kPeerK + (unsigned)i) * 32ull;
                    const unsigned stride = (unsigned)P.pxRemoteTotal[q] * (unsigned)kPeerK;  // parity stride in 32-byte units
                    if (k == 0) { r0.z = (unsigned)addr; r0.w = (unsigned)(addr >> 32); r1.x = stride; }
                    else { r1.z = (unsigned)addr; r1.w = (unsigned)(addr >> 32); r1.y = stride; }
                }
                rec[2 * slot] = r0;
                rec[2 * slot + 1] = r1;
            }
        }
        CK(h->pushRec.upload(rec, h->stream));
    }
    const PeerArgs pa = peer_args(h);
    CK(cudaMemcpyAsync(h->pxArgs.p, &pa, sizeof(pa), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->peersSet = true;
    return TETSIM_OK;
}

int tetsim_level_schedule(const int32_t *tetIds, int32_t numTets, int32_t numVerts, int32_t *level) {
    if (!tetIds || !level) return fail(TETSIM_E_INVALID, "null argument");
    return level_schedule(numVerts, numTets, tetIds, level);
}
int tetsim_greedy_colors(const int32_t *tetIds, int32_t numTets, int32_t numVerts, int32_t *color) {
    if (!tetIds || !color) return fail(TETSIM_E_INVALID, "null argument");
    int n = greedy_colors(numVerts, numTets, tetIds, color);
    return n < 0 ? fail(TETSIM_E_INVALID, "more than 256 colours") : n;
}

int tetsim_connected_components(const int32_t *tetIds, int32_t numTets, int32_t numVerts, int32_t *vertComp) {
    if ((!tetIds && numTets > 0) || !vertComp || numTets < 0 || numVerts < 0) return fail(TETSIM_E_INVALID, "null or negative argument");
    for (int64_t c = 0; c < 4 * (int64_t)numTets; c++)
        if (tetIds[c] < 0 || tetIds[c] >= numVerts) return fail(TETSIM_E_INVALID, "vertex id out of range");
    std::vector<int> comp;
    const int n = connected_components(numVerts, numTets, tetIds, comp);
    std::copy(comp.begin(), comp.end(), vertComp);
    return n;
}

int tetsim_plan_partition(const float *verts, int32_t numVerts, const int32_t *tetIds, int32_t numTets,
                          int32_t clusterSize, int32_t reorder, int32_t rank, int32_t worldSize, int32_t counts[4],
                          int32_t *localToCaller, int32_t *localTets) {
    if (!verts || !tetIds || !counts) return fail(TETSIM_E_INVALID, "null argument");
    if (clusterSize < 1 || worldSize < 1 || rank < 0 || rank >= worldSize) return fail(TETSIM_E_INVALID, "bad clusterSize/rank/worldSize");
    for (int64_t c = 0; c < 4 * (int64_t)numTets; c++)
        if (tetIds[c] < 0 || tetIds[c] >= numVerts) return fail(TETSIM_E_INVALID, "vertex id out of range");
    std::vector<int> rankStart;
    std::vector<int> order = solver_order(numVerts, numTets, verts, tetIds, reorder != 0, worldSize, rankStart);
    ClusterPlan P;
    std::string err;
    if (!build_cluster_plan(numVerts, numTets, tetIds, order, rankStart, clusterSize, rank, worldSize, P, err)) return fail(TETSIM_E_INVALID, err);
    counts[0] = P.localTets; counts[1] = P.numInterior; counts[2] = P.numBoundary; counts[3] = P.numClusters;
    if (localToCaller) std::copy(P.localToCaller.begin(), P.localToCaller.end(), localToCaller);
    if (localTets) {
        int n = 0;
        for (int t : P.recordTet) if (t >= 0) localTets[n++] = t;
    }
    return TETSIM_OK;
}

int tetsim_plan_halo(const float *verts, int32_t numVerts, const int32_t *tetIds, int32_t numTets, int32_t clusterSize,
                     int32_t reorder, int32_t rank, int32_t worldSize, int32_t capacity, int32_t *peers,
                     int32_t *segStart, int32_t *remoteOff, int32_t *remoteTotal, int32_t *remoteSlot) {
    if (!verts || !tetIds || !peers || !segStart || !remoteOff || !remoteTotal || !remoteSlot) return fail(TETSIM_E_INVALID, "null argument");
    if (clusterSize < 1 || worldSize < 1 || rank < 0 || rank >= worldSize) return fail(TETSIM_E_INVALID, "bad clusterSize/rank/worldSize");
    for (int64_t c = 0; c < 4 * (int64_t)numTets; c++)
        if (tetIds[c] < 0 || tetIds[c] >= numVerts) return fail(TETSIM_E_INVALID, "vertex id out of range");
    std::vector<int> rankStart;
    std::vector<int> order = solver_order(numVerts, numTets, verts, tetIds, reorder != 0, worldSize, rankStart);
    ClusterPlan P;
    std::string err;
    if (!build_cluster_plan(numVerts, numTets, tetIds, order, rankStart, clusterSize, rank, worldSize, P, err)) return fail(TETSIM_E_INVALID, err);
    if (worldSize > 1 && !P.haloOk) return fail(TETSIM_E_STATE, "neighbour lists need worldSize <= 64");
    const int np = (int)P.hxPeers.size();
    if (np > capacity) return fail(TETSIM_E_INVALID, "capacity too small");
    for (int i = 0; i < np; i++) {
        peers[i] = P.hxPeers[i]; segStart[i] = P.hxSegStart[i];
        remoteOff[i] = P.pxRemoteOff[i]; remoteTotal[i] = P.pxRemoteTotal[i]; remoteSlot[i] = P.pxRemoteSlot[i];
    }
    segStart[np] = np ? P.hxSegStart[np] : 0;
    return np;
}

}  // extern "C"
