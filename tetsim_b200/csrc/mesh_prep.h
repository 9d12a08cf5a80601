// mesh_prep.h -- host-side, once-per-mesh topology pre-pass (pure C++, no CUDA).
// The reference's counterpart is the table builder in SoftBodyGPU.initPhysics
// (src/SoftbodyGPU.js:553-577: forward tet->vertex table and the 9x4 reverse tables); graph
// colouring is a README TODO there (README.md:25).  Everything numeric stays on the device.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tsim {

// vertex -> incident tet corners, encoded 4*tet + corner, ascending; CSR.
struct CornerTable {
    std::vector<int> start;  // numVerts + 1
    std::vector<int> ent;    // 4 * numTets
    int maxValence = 0;
};
CornerTable build_corner_table(int numVerts, int numTets, const int *tetIds);

// The reference's reverse table semantics (src/SoftbodyGPU.js:559-577) as CSR: 36 slots per vertex,
// a slot holding a value <= 0 counts as free, corners beyond the capacity are dropped.
CornerTable build_reference_table(int numVerts, int numTets, const int *tetIds, bool referenceBug, int capacity);

// Connected components over shared vertices.  comp[v] in [0, count); isolated vertices get their own.
int connected_components(int numVerts, int numTets, const int *tetIds, std::vector<int> &vertComp);

// Order-preserving dependency levels of the sequential sweep: level[j] = 1 + max level of any
// earlier tet sharing a vertex (0-based levels returned).  Sweeping level by level is
// bit-identical to the reference's in-order Gauss-Seidel loop (src/Softbody.js:207-208).
int level_schedule(int numVerts, int numTets, const int *tetIds, int *level);

// Greedy colouring in tet order: smallest colour not used by any earlier tet sharing a vertex.
// Returns the number of colours, or -1 if more than 256 would be needed.
int greedy_colors(int numVerts, int numTets, const int *tetIds, int *color);

// Tets sorted along a 3-D Hilbert curve of their centroids (ties by tet index).
std::vector<int> morton_order(int numVerts, int numTets, const float *verts, const int *tetIds);

// Tiling of a tet sequence into CTA tiles for the clustered Jacobi kernel.
struct ClusterPlan {
    int T = 0;                          // tets per tile
    int numClusters = 0;                // tiles on this rank
    int numBoundaryTiles = 0;           // the first numBoundaryTiles tiles touch rank-shared vertices
    // Neighbour ("halo") exchange of the boundary sums, the alternative to an all-reduce over all ranks:
    // rank r and q exchange their partial sums of the boundary vertices BOTH touch, then every sharer
    // adds the contributions of all sharers in ascending rank order (identical on every sharer).
    bool haloOk = false;                // false when worldSize > 64 (falls back to the all-reduce)
    std::vector<uint8_t> boundaryActive;  // [numBoundary] 1 = touched by this rank's tets
    std::vector<int> hxPeers;           // neighbour ranks, ascending
    std::vector<int> hxSegStart;        // [peers + 1] segment offsets (in boundary entries) of send AND recv buffers
    std::vector<int> hxSendIdx;         // boundary index of each send entry (same order on both sides of a pair)
    std::vector<int> hxSrcStart;        // [numBoundary + 1] CSR over the sources of each active boundary vertex
    std::vector<int> hxSrc;             // < numBoundary: own partial sum; >= numBoundary: recv entry - numBoundary
    // Peer-memory ("fused") exchange: the same segments, but every rank STORES its boundary sums straight into the
    // receive buffer of each sharer through a mapped peer pointer.  All of it follows from the global rank masks, so
    // every rank derives its peers' layouts locally and nothing but the memory handles has to be exchanged.
    std::vector<int> pxRemoteOff;       // [peers] offset (entries) of MY segment inside peer q's receive buffer
    std::vector<int> pxRemoteTotal;     // [peers] total entries of peer q's receive buffer (its parity stride)
    std::vector<int> pxRemoteSlot;      // [peers] my index in peer q's ascending peer list (which of its flags is mine)
    std::vector<int> pxStart;           // [numBoundary + 1] CSR over the push destinations of each active boundary vertex
    std::vector<int> pxPeer;            //   index into hxPeers
    std::vector<int> pxEntry;           //   entry inside that peer's receive buffer
    // fused form: every TILE pushes its own partial of a boundary vertex (no sum, no ticket).  Partial slots of the
    // boundary tiles come first in part[]; pxSlotIdx[slot] = which of its vertex's partials this slot is (its rank-local
    // order, 0 .. count-1), 0xff for slots of vertices that are not rank-shared.
    std::vector<uint8_t> pxSlotIdx;     // [clVertStart[numBoundaryTiles]]
    int pxMaxPartials = 0;              // most tile partials any boundary vertex has on this rank
    int numLocalVerts = 0;              // interior + all boundary vertices
    int numInterior = 0;
    int numBoundary = 0;                // global count of rank-shared vertices (same on every rank)
    std::vector<int> localToCaller;     // [numLocalVerts]
    std::vector<int> recordTet;         // [numClusters * T] caller tet index, -1 = padding
    // per record: {slot0|slot1<<16, slot2|slot3<<16, dest0|dest1<<16, dest2|dest3<<16}.  slot = byte offset
    // (16 * tile vertex index) of a corner's position in the staged vertex tile; dest = byte offset
    // (16 * entry) of the corner's dx in the staged corner buffer (grouped rows).
    std::vector<uint32_t> recordAux;    // [numClusters * T * 4]
    std::vector<int> clVertStart;       // [numClusters + 1]
    std::vector<int> clVerts;           // local vertex ids per tile, descending tile valence
    // One metadata block per tile (variable size, 16-byte granular), fetched with a single bulk async copy:
    //   [0,16)   int32 {first partial-sum slot, tile vertex count nl, max tile valence, byte offset of ids}
    //   [16, ..) uint16 gbase[colStride]    byte offset (16 * entry) of the row block of vertex group g (8 vertices per group)
    //   then     uint8  val[nlPad]          tile valence of tile vertex j (descending), nlPad = roundup16(nl)
    //   then     int32  ids[nlPad]          handle-local vertex id of tile vertex j
    std::vector<unsigned char> tileMeta;
    std::vector<uint32_t> metaOff;      // [numClusters + 1] block offsets in units of 16 bytes
    int metaStride = 0;                 // largest block, bytes
    int metaValOff = 0;
    int colStride = 0;
    int maxTileEntries = 0;             // sixteen-byte entries of the largest tile's corner buffer (grouped rows, see mesh_prep.cpp)
    int maxTileVerts = 0, maxTileVertsPad = 0;
    std::vector<int> vpStart, vpSlot;   // local vertex -> indices into the partial-sum array
    std::vector<float> invValence;      // [numLocalVerts] 1 / GLOBAL valence
    int localTets = 0;
    int maxValence = 0;
};
// Solver order of the tets: Hilbert-sorted (reorder) and, for worldSize > 1, grouped by a recursive
// coordinate bisection into worldSize spatial parts.  rankStart[r] = position of rank r's first tet.
std::vector<int> solver_order(int numVerts, int numTets, const float *verts, const int *tetIds, bool reorder,
                              int worldSize, std::vector<int> &rankStart);
// order: the global tet sequence (caller tet indices); rank r owns positions [rankStart[r], rankStart[r+1]).
bool build_cluster_plan(int numVerts, int numTets, const int *tetIds, const std::vector<int> &order,
                        const std::vector<int> &rankStart, int T, int rank, int worldSize, ClusterPlan &plan,
                        std::string &err);

}  // namespace tsim
