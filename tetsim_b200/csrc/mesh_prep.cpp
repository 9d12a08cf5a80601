// mesh_prep.cpp -- see mesh_prep.h.
#include "mesh_prep.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

namespace tsim {

CornerTable build_corner_table(int numVerts, int numTets, const int *tetIds) {
    CornerTable t;
    t.start.assign((size_t)numVerts + 1, 0);
    const size_t nc = 4 * (size_t)numTets;
    for (size_t c = 0; c < nc; c++) t.start[(size_t)tetIds[c] + 1]++;
    for (int v = 0; v < numVerts; v++) {
        t.maxValence = std::max(t.maxValence, t.start[v + 1]);
        t.start[v + 1] += t.start[v];
    }
    t.ent.resize(nc);
    std::vector<int> fill(t.start.begin(), t.start.end() - 1);
    for (size_t c = 0; c < nc; c++) t.ent[fill[tetIds[c]]++] = (int)c;  // c = 4*tet + corner, ascending
    return t;
}

CornerTable build_reference_table(int numVerts, int numTets, const int *tetIds, bool referenceBug, int capacity) {
    // Per vertex the reference scans its 36 slots for the first value <= 0.0 (src/SoftbodyGPU.js:566-573).
    // Only the encoded value 0 (tet 0, corner 0) can be non-negative and still "free"; it can sit in
    // exactly one vertex's list, so the general rule reduces to: append, except that vertex's slot
    // holding 0 is reused once.
    std::vector<std::vector<int>> lists((size_t)numVerts);
    for (int e = 0; e < numTets; e++)
        for (int k = 0; k < 4; k++) {
            std::vector<int> &L = lists[tetIds[4 * (size_t)e + k]];
            const int v = 4 * e + k;
            size_t slot = L.size();
            if (referenceBug)
                for (size_t j = 0; j < L.size(); j++)
                    if (L[j] <= 0) { slot = j; break; }
            if (capacity > 0 && slot >= (size_t)capacity) continue;
            if (slot == L.size()) L.push_back(v);
            else L[slot] = v;
        }
    CornerTable t;
    t.start.assign((size_t)numVerts + 1, 0);
    for (int v = 0; v < numVerts; v++) {
        t.start[v + 1] = t.start[v] + (int)lists[v].size();
        t.maxValence = std::max(t.maxValence, (int)lists[v].size());
    }
    t.ent.reserve((size_t)t.start[numVerts]);
    for (int v = 0; v < numVerts; v++) t.ent.insert(t.ent.end(), lists[v].begin(), lists[v].end());
    return t;
}

int connected_components(int numVerts, int numTets, const int *tetIds, std::vector<int> &vertComp) {
    std::vector<int> parent((size_t)numVerts);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int a) {
        while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
        return a;
    };
    for (int e = 0; e < numTets; e++) {
        int r = find(tetIds[4 * (size_t)e]);
        for (int k = 1; k < 4; k++) {
            int s = find(tetIds[4 * (size_t)e + k]);
            if (s != r) { if (s < r) std::swap(s, r); parent[s] = r; }  // smallest vertex id is the root
        }
    }
    vertComp.assign((size_t)numVerts, -1);
    int count = 0;
    for (int v = 0; v < numVerts; v++) {  // roots are met before their members: components numbered by first vertex
        int r = find(v);
        if (vertComp[r] < 0) vertComp[r] = count++;
        vertComp[v] = vertComp[r];
    }
    return count;
}

int level_schedule(int numVerts, int numTets, const int *tetIds, int *level) {
    std::vector<int> last((size_t)numVerts, -1);
    int maxLevel = -1;
    for (int e = 0; e < numTets; e++) {
        const int *t = tetIds + 4 * (size_t)e;
        int lv = 1 + std::max(std::max(last[t[0]], last[t[1]]), std::max(last[t[2]], last[t[3]]));
        level[e] = lv;
        last[t[0]] = last[t[1]] = last[t[2]] = last[t[3]] = lv;
        maxLevel = std::max(maxLevel, lv);
    }
    return maxLevel + 1;
}

int greedy_colors(int numVerts, int numTets, const int *tetIds, int *color) {
    struct Mask { uint64_t w[4]; };
    std::vector<Mask> used((size_t)numVerts, Mask{{0, 0, 0, 0}});
    int numColors = 0;
    for (int e = 0; e < numTets; e++) {
        const int *t = tetIds + 4 * (size_t)e;
        int c = -1;
        for (int w = 0; w < 4 && c < 0; w++) {
            uint64_t m = used[t[0]].w[w] | used[t[1]].w[w] | used[t[2]].w[w] | used[t[3]].w[w];
            if (~m) c = 64 * w + __builtin_ctzll(~m);
        }
        if (c < 0) return -1;
        color[e] = c;
        for (int k = 0; k < 4; k++) used[t[k]].w[c >> 6] |= 1ull << (c & 63);
        numColors = std::max(numColors, c + 1);
    }
    return numColors;
}

// 3-D Hilbert index of a point on a 2^bits grid (Skilling, "Programming the Hilbert curve", 2004):
// unlike a Morton curve it never jumps, so a run of consecutive tets is always a connected blob
// and the vertex footprint of a tile stays close to its average.
static inline uint64_t hilbert3(uint32_t x, uint32_t y, uint32_t z, int bits) {
    uint32_t X[3] = {x, y, z};
    const uint32_t M = 1u << (bits - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {  // inverse undo
        const uint32_t P = Q - 1;
        for (int i = 0; i < 3; i++) {
            if (X[i] & Q) X[0] ^= P;
            else { uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    for (int i = 1; i < 3; i++) X[i] ^= X[i - 1];  // Gray encode
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    for (int i = 0; i < 3; i++) X[i] ^= t;
    uint64_t h = 0;  // interleave, X[0] most significant within each bit triple
    for (int b = bits - 1; b >= 0; b--)
        for (int i = 0; i < 3; i++) h = (h << 1) | ((X[i] >> b) & 1u);
    return h;
}

std::vector<int> morton_order(int numVerts, int numTets, const float *verts, const int *tetIds) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int v = 0; v < numVerts; v++)
        for (int c = 0; c < 3; c++) {
            float x = verts[3 * (size_t)v + c];
            if (x == x) { lo[c] = std::min(lo[c], x); hi[c] = std::max(hi[c], x); }
        }
    // one cubic cell size for all axes so the curve follows the shape of the domain
    double ext = 0.0;
    for (int c = 0; c < 3; c++) ext = std::max(ext, (double)hi[c] - (double)lo[c]);
    const int bits = 20;
    const double scale = ext > 0.0 ? (double)((1 << bits) - 1) / ext : 0.0;
    std::vector<std::pair<uint64_t, int>> keys((size_t)numTets);
    for (int e = 0; e < numTets; e++) {
        const int *t = tetIds + 4 * (size_t)e;
        uint32_t q[3];
        for (int c = 0; c < 3; c++) {
            double m = 0.25 * ((double)verts[3 * (size_t)t[0] + c] + (double)verts[3 * (size_t)t[1] + c] +
                               (double)verts[3 * (size_t)t[2] + c] + (double)verts[3 * (size_t)t[3] + c]);
            double g = (m - (double)lo[c]) * scale;
            q[c] = (g == g && g > 0.0) ? (uint32_t)std::min(g, (double)((1 << bits) - 1)) : 0u;
        }
        keys[e] = {hilbert3(q[0], q[1], q[2], bits), e};
    }
    std::sort(keys.begin(), keys.end());  // ties broken by tet index
    std::vector<int> order((size_t)numTets);
    for (int e = 0; e < numTets; e++) order[e] = keys[e].second;
    return order;
}

// Recursive coordinate bisection of the tets into `parts` pieces of (nearly) equal tet count: split
// the longest axis of the piece's bounding box at the proportional quantile of the tets' LOWEST
// vertex coordinate, ties all going to the upper piece.  On structured meshes (beam: every tet of a
// cell layer has the same lowest x) the cut then falls between two cell layers, so only ONE vertex
// plane is shared; on unstructured meshes it is an ordinary median split.
static void rcb(std::vector<int> &idx, int lo, int hi, int parts, int firstPart, const std::vector<float> &key,
                std::vector<int> &partOf) {
    if (parts <= 1 || hi - lo <= 1) {
        for (int i = lo; i < hi; i++) partOf[idx[i]] = firstPart;
        return;
    }
    float bl[3] = {INFINITY, INFINITY, INFINITY}, bh[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = lo; i < hi; i++)
        for (int c = 0; c < 3; c++) {
            float x = key[3 * (size_t)idx[i] + c];
            bl[c] = std::min(bl[c], x); bh[c] = std::max(bh[c], x);
        }
    int ax = 0;
    for (int c = 1; c < 3; c++) if (bh[c] - bl[c] > bh[ax] - bl[ax]) ax = c;
    const int pl = parts / 2;
    int mid = lo + (int)((int64_t)(hi - lo) * pl / parts);
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) {
        float xa = key[3 * (size_t)a + ax], xb = key[3 * (size_t)b + ax];
        return xa != xb ? xa < xb : a < b;
    });
    // move the cut to the nearer end of the run of equal keys around the quantile
    const float m = key[3 * (size_t)idx[mid] + ax];
    auto below = std::partition(idx.begin() + lo, idx.begin() + hi, [&](int a) { return key[3 * (size_t)a + ax] < m; });
    auto notAbove = std::partition(below, idx.begin() + hi, [&](int a) { return key[3 * (size_t)a + ax] <= m; });
    const int cutLo = (int)(below - idx.begin()), cutHi = (int)(notAbove - idx.begin());
    int cut = (mid - cutLo <= cutHi - mid) ? cutLo : cutHi;
    if (cut <= lo || cut >= hi) cut = (cutLo > lo && cutLo < hi) ? cutLo : ((cutHi > lo && cutHi < hi) ? cutHi : mid);
    rcb(idx, lo, cut, pl, firstPart, key, partOf);
    rcb(idx, cut, hi, parts - pl, firstPart + pl, key, partOf);
}

std::vector<int> solver_order(int numVerts, int numTets, const float *verts, const int *tetIds, bool reorder,
                              int worldSize, std::vector<int> &rankStart) {
    std::vector<int> order;
    if (reorder) order = morton_order(numVerts, numTets, verts, tetIds);
    else { order.resize((size_t)numTets); std::iota(order.begin(), order.end(), 0); }
    rankStart.assign((size_t)worldSize + 1, 0);
    if (worldSize <= 1 || !reorder) {  // caller's order: contiguous equal chunks
        for (int r = 0; r <= worldSize; r++) rankStart[r] = (int)((int64_t)numTets * r / worldSize);
        return order;
    }
    std::vector<float> cent(3 * (size_t)numTets);
    for (int e = 0; e < numTets; e++) {
        const int *t = tetIds + 4 * (size_t)e;
        for (int c = 0; c < 3; c++)
            cent[3 * (size_t)e + c] = std::min(std::min(verts[3 * (size_t)t[0] + c], verts[3 * (size_t)t[1] + c]),
                                               std::min(verts[3 * (size_t)t[2] + c], verts[3 * (size_t)t[3] + c]));
    }
    std::vector<int> idx((size_t)numTets), partOf((size_t)numTets, 0);
    std::iota(idx.begin(), idx.end(), 0);
    rcb(idx, 0, numTets, worldSize, 0, cent, partOf);
    // keep the Hilbert order inside each part
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return partOf[a] < partOf[b]; });
    int pos = 0;
    for (int r = 0; r < worldSize; r++) {
        rankStart[r] = pos;
        while (pos < numTets && partOf[order[pos]] == r) pos++;
    }
    rankStart[worldSize] = numTets;
    return order;
}

bool build_cluster_plan(int numVerts, int numTets, const int *tetIds, const std::vector<int> &order,
                        const std::vector<int> &rankStart, int T, int rank, int worldSize, ClusterPlan &P,
                        std::string &err) {
    P = ClusterPlan();
    P.T = T;
    // ---- cut the tet sequence into tiles: at most T tets and at most `vertCap` distinct vertices ----
    // A space-filling curve occasionally jumps, and a tile straddling a jump touches 2-3x the average
    // number of vertices; since the kernel sizes its shared-memory vertex buffers by the LARGEST tile,
    // such tiles are closed early (their remaining record slots become padding).
    std::vector<int> tileStart;  // position in `order` of each tile's first tet, + sentinel
    {
        std::vector<int> stampG((size_t)numVerts, -1);
        auto cut = [&](int cap, double *avgOut) {
            tileStart.clear();
            std::fill(stampG.begin(), stampG.end(), -1);
            int tile = -1, nt = T, nv = 0, nextRank = 0;
            bool forced = false;
            long long sumV = 0;
            for (int pos = 0; pos < numTets; pos++) {
                const int *t = tetIds + 4 * (size_t)order[pos];
                int fresh = 0;
                if (tile >= 0) for (int k = 0; k < 4; k++) fresh += stampG[t[k]] != tile;
                while (nextRank <= worldSize && rankStart[nextRank] <= pos) { if (rankStart[nextRank] == pos) forced = true; nextRank++; }
                if (nt == T || forced || (cap > 0 && nv + fresh > cap)) {
                    forced = false;
                    sumV += nv;
                    tile++; nt = 0; nv = 0;
                    tileStart.push_back(pos);
                }
                for (int k = 0; k < 4; k++)
                    if (stampG[t[k]] != tile) { stampG[t[k]] = tile; nv++; }
                nt++;
            }
            sumV += nv;
            tileStart.push_back(numTets);
            if (avgOut) *avgOut = tile >= 0 ? (double)sumV / (tile + 1) : 0.0;
        };
        double avg = 0.0;
        cut(0, &avg);
        int cap = (((int)(1.4 * avg) + 15) / 16) * 16;
        if (const char *e = getenv("TETSIM_TILE_VERT_CAP")) cap = atoi(e);
        if (cap > 0 && cap < 16) cap = 16;
        if (cap > 0) cut(cap, nullptr);
    }
    const int totalClusters = (int)tileStart.size() - 1;
    std::vector<int> rankFirstTile((size_t)worldSize + 1, totalClusters);
    for (int r = 0; r <= worldSize; r++)
        rankFirstTile[r] = (int)(std::lower_bound(tileStart.begin(), tileStart.end() - 1, rankStart[r]) - tileStart.begin());
    auto firstClusterOf = [&](int r) { return rankFirstTile[r]; };
    const int c0 = firstClusterOf(rank), c1 = firstClusterOf(rank + 1);
    P.numClusters = c1 - c0;

    // global valence and rank-shared (boundary) vertices
    std::vector<int> valence((size_t)numVerts, 0);
    std::vector<int> rmin, rmax;
    std::vector<uint64_t> rmask;  // ranks touching each vertex (worldSize <= 64)
    if (worldSize > 1) { rmin.assign((size_t)numVerts, worldSize); rmax.assign((size_t)numVerts, -1); }
    if (worldSize > 1 && worldSize <= 64) rmask.assign((size_t)numVerts, 0ull);
    {
        int r = 0, cl = 0;
        for (int pos = 0; pos < numTets; pos++) {
            if (worldSize > 1) {
                while (pos >= tileStart[cl + 1]) cl++;
                while (cl >= firstClusterOf(r + 1)) r++;
            }
            const int *t = tetIds + 4 * (size_t)order[pos];
            for (int k = 0; k < 4; k++) {
                valence[t[k]]++;
                if (worldSize > 1) { rmin[t[k]] = std::min(rmin[t[k]], r); rmax[t[k]] = std::max(rmax[t[k]], r); }
                if (!rmask.empty()) rmask[t[k]] |= 1ull << r;
            }
        }
    }
    for (int v = 0; v < numVerts; v++) P.maxValence = std::max(P.maxValence, valence[v]);

    // local numbering: interior vertices by first touch along the tile sequence, then every boundary vertex
    std::vector<int> local((size_t)numVerts, -1);
    std::vector<int> boundary;
    if (worldSize > 1)
        for (int v = 0; v < numVerts; v++)
            if (rmax[v] > rmin[v]) boundary.push_back(v);
    P.numBoundary = (int)boundary.size();
    if (!rmask.empty()) {  // neighbour exchange lists
        P.haloOk = true;
        const int nB = P.numBoundary;
        P.boundaryActive.assign((size_t)nB, 0);
        std::vector<std::vector<int>> shared((size_t)worldSize);  // per peer: boundary indices shared with it, ascending
        for (int b = 0; b < nB; b++) {
            const uint64_t m = rmask[boundary[b]];
            if (!(m >> rank & 1ull)) continue;
            P.boundaryActive[b] = 1;
            for (int q = 0; q < worldSize; q++)
                if (q != rank && (m >> q & 1ull)) shared[q].push_back(b);
        }
        std::vector<int> segOf((size_t)worldSize, -1);
        P.hxSegStart.assign(1, 0);
        for (int q = 0; q < worldSize; q++)
            if (!shared[q].empty()) {
                segOf[q] = (int)P.hxPeers.size();
                P.hxPeers.push_back(q);
                P.hxSendIdx.insert(P.hxSendIdx.end(), shared[q].begin(), shared[q].end());
                P.hxSegStart.push_back((int)P.hxSendIdx.size());
            }
        // sources of each active boundary vertex in ascending rank order
        std::vector<int> cursor((size_t)worldSize, 0);
        P.hxSrcStart.assign((size_t)nB + 1, 0);
        for (int b = 0; b < nB; b++) {
            P.hxSrcStart[b] = (int)P.hxSrc.size();
            if (!P.boundaryActive[b]) continue;
            const uint64_t m = rmask[boundary[b]];
            for (int q = 0; q < worldSize; q++) {
                if (!(m >> q & 1ull)) continue;
                if (q == rank) P.hxSrc.push_back(b);
                else P.hxSrc.push_back(nB + P.hxSegStart[segOf[q]] + cursor[q]++);  // shared[q] is ascending in b
            }
        }
        P.hxSrcStart[nB] = (int)P.hxSrc.size();
        // peer-memory exchange: where my entries land in each sharer's receive buffer.  pair[q][p] = number of
        // boundary vertices ranks q and p both touch = length of the segment they exchange.
        const size_t W = (size_t)worldSize;
        std::vector<int> pair(W * W, 0);
        for (int b = 0; b < nB; b++) {
            const uint64_t m = rmask[boundary[b]];
            for (int q = 0; q < worldSize; q++) {
                if (!(m >> q & 1ull)) continue;
                for (int p2 = 0; p2 < worldSize; p2++)
                    if (p2 != q && (m >> p2 & 1ull)) pair[(size_t)q * W + p2]++;
            }
        }
        for (int q : P.hxPeers) {
            int off = 0, total = 0, slot = 0;
            for (int p2 = 0; p2 < worldSize; p2++) {
                const int n = pair[(size_t)q * W + p2];
                if (p2 < rank) { off += n; slot += n > 0; }
                total += n;
            }
            P.pxRemoteOff.push_back(off);
            P.pxRemoteTotal.push_back(total);
            P.pxRemoteSlot.push_back(slot);
        }
        std::fill(cursor.begin(), cursor.end(), 0);
        P.pxStart.assign((size_t)nB + 1, 0);
        for (int b = 0; b < nB; b++) {
            P.pxStart[b] = (int)P.pxPeer.size();
            if (!P.boundaryActive[b]) continue;
            const uint64_t m = rmask[boundary[b]];
            for (int q = 0; q < worldSize; q++) {
                if (q == rank || !(m >> q & 1ull)) continue;
                P.pxPeer.push_back(segOf[q]);
                P.pxEntry.push_back(P.pxRemoteOff[segOf[q]] + cursor[q]++);  // both sides list the shared set ascending in b
            }
        }
        P.pxStart[nB] = (int)P.pxPeer.size();
    }
    P.localTets = std::max(0, tileStart[c1] - tileStart[c0]);
    for (int v : boundary) local[v] = -2;
    // Tiles that touch a rank-shared vertex go FIRST: a multi-GPU handle launches them alone, starts
    // the all-reduce of the boundary sums, and hides it behind the remaining (interior) tiles.
    std::vector<int> tileOrder;  // local tile c -> global tile index
    {
        std::vector<int> inner;
        for (int g = c0; g < c1; g++) {
            bool touches = false;
            for (int pos = tileStart[g]; pos < tileStart[g + 1] && !touches; pos++) {
                const int *t = tetIds + 4 * (size_t)order[pos];
                touches = local[t[0]] == -2 || local[t[1]] == -2 || local[t[2]] == -2 || local[t[3]] == -2;
            }
            (touches ? tileOrder : inner).push_back(g);
        }
        P.numBoundaryTiles = (int)tileOrder.size();
        tileOrder.insert(tileOrder.end(), inner.begin(), inner.end());
    }
    int nI = 0;
    for (int g : tileOrder)
        for (int pos = tileStart[g]; pos < tileStart[g + 1]; pos++) {
            const int *t = tetIds + 4 * (size_t)order[pos];
            for (int k = 0; k < 4; k++)
                if (local[t[k]] == -1) { local[t[k]] = nI++; P.localToCaller.push_back(t[k]); }
        }
    // vertices no tet references still fall and collide in the reference (src/Softbody.js:198-202 has
    // no valence test): keep them resident (rank 0 of a multi-GPU job), with no correction to apply
    if (rank == 0)
        for (int v = 0; v < numVerts; v++)
            if (valence[v] == 0) { local[v] = nI++; P.localToCaller.push_back(v); }
    P.numInterior = nI;
    for (size_t b = 0; b < boundary.size(); b++) { local[boundary[b]] = nI + (int)b; P.localToCaller.push_back(boundary[b]); }
    P.numLocalVerts = nI + P.numBoundary;
    P.invValence.resize((size_t)P.numLocalVerts);
    for (int i = 0; i < P.numLocalVerts; i++) {
        int val = valence[P.localToCaller[i]];
        P.invValence[i] = val > 0 ? 1.0f / (float)val : 0.0f;
    }

    // tiles
    const size_t nRec = (size_t)P.numClusters * T;
    P.recordTet.assign(nRec, -1);
    P.recordAux.assign(4 * nRec, 0u);
    P.clVertStart.assign((size_t)P.numClusters + 1, 0);
    std::vector<int> stamp((size_t)P.numLocalVerts, -1), tileIdx((size_t)P.numLocalVerts, 0);
    std::vector<int> tileVerts, tileVal, perm, rank_of;
    std::vector<std::vector<uint16_t>> colOffs((size_t)P.numClusters);
    constexpr int kGroup = 8, kRowStride = 9;
    int maxTileEntries = 0;
    std::vector<uint8_t> allVal;
    std::vector<int> cornerStart, cornerFill, cornerList, cornerDiag, loadG, loadS, sorted;
    std::vector<char> posTaken, diagFree;
    int maxTileVal = 0;
    for (int c = 0; c < P.numClusters; c++) {
        const int pb = tileStart[tileOrder[c]], pe = tileStart[tileOrder[c] + 1];
        tileVerts.clear(); tileVal.clear();
        for (int pos = pb; pos < pe; pos++) {
            const int *t = tetIds + 4 * (size_t)order[pos];
            P.recordTet[(size_t)c * T + (pos - pb)] = order[pos];
            for (int k = 0; k < 4; k++) {
                int lv = local[t[k]];
                if (stamp[lv] != c) { stamp[lv] = c; tileIdx[lv] = (int)tileVerts.size(); tileVerts.push_back(lv); tileVal.push_back(0); }
                tileVal[tileIdx[lv]]++;
            }
        }
        const int nl = (int)tileVerts.size();
        const int ntile = pe - pb;
        // sort tile vertices by descending tile valence (stable)
        perm.resize(nl);
        std::iota(perm.begin(), perm.end(), 0);
        std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return tileVal[a] > tileVal[b]; });
        const int tv = nl ? tileVal[perm[0]] : 0;
        if (tv > 255) { err = "a vertex has more than 255 tet corners inside one tile"; return false; }
        maxTileVal = std::max(maxTileVal, tv);
        // grouped rows: the valence-sorted tile vertices are cut into groups of kGroup = 8; group g owns a block of
        // (largest valence in g) rows of kRowStride = 9 sixteen-byte entries, entry (i, j) = i-th corner of tile vertex j at
        // gbase[j / 8] + 9 i + j % 8.  The thread that sums vertex j then reads base + 144 i: immediate offsets, no
        // per-entry index load (the jagged diagonals of round 1 cost a LDS.U16 + two LEA per entry), and a quarter-warp
        // reads 128 contiguous bytes (conflict-free).  The odd row stride keeps the row choice a degree of freedom for
        // the bank placement of the scatter: entry (i, j) lies in bank group (i + j) mod 8.
        std::vector<uint16_t> &co = colOffs[c];  // gbase per group, in entries
        const int ngroups = (nl + kGroup - 1) / kGroup;
        co.assign((size_t)ngroups + 1, 0);
        {
            int off = 0;
            for (int g = 0; g < ngroups; g++) {
                co[g] = (uint16_t)off;
                off += tileVal[perm[g * kGroup]] * kRowStride;  // perm is descending: the group's first vertex has its largest valence
            }
            co[ngroups] = (uint16_t)off;
            if (off > 4000) { err = "tile too large for 16-bit byte offsets"; return false; }
            maxTileEntries = std::max(maxTileEntries, off);
        }
        auto entryOf = [&](int i, int j) { return (int)co[j / kGroup] + kRowStride * i + j % kGroup; };
        // ---- shared-memory bank placement (round-1 ncu: 37 % of this kernel's shared-memory wavefronts
        // were bank-conflict replays of the 16-byte gathers and scatters) ----
        // A 128-bit warp access is served a quarter-warp (8 lanes) at a time and is conflict-free when
        // the 8 sixteen-byte chunks fall into 8 different bank groups (byte offset / 16 mod 8).  The
        // lanes of one quarter-warp executing corner k form a "conflict set".  Two degrees of freedom
        // cost nothing at run time: the order of vertices inside a run of equal tile valence, and
        // which of a vertex's diagonals each of its corners is parked in.  Both are chosen greedily.
        const int nSets = ((T + 31) / 32) * 16;
        auto setOf = [&](int tl, int k) { return ((tl >> 5) * 4 + k) * 4 + ((tl & 31) >> 3); };
        cornerStart.assign((size_t)nl + 1, 0);   // corners of each tile vertex (pre-sort index), appearance order
        for (int v = 0; v < nl; v++) cornerStart[v + 1] = cornerStart[v] + tileVal[v];
        cornerFill.assign(cornerStart.begin(), cornerStart.end() - 1);
        cornerList.resize((size_t)cornerStart[nl]);
        for (int tl = 0; tl < ntile; tl++) {
            const int *t = tetIds + 4 * (size_t)order[pb + tl];
            for (int k = 0; k < 4; k++) cornerList[cornerFill[tileIdx[local[t[k]]]]++] = 4 * tl + k;
        }
        // (a) position of each vertex inside its equal-valence run
        loadG.assign((size_t)nSets * 8, 0);
        rank_of.assign(nl, -1);
        posTaken.assign(nl, 0);
        for (int a0 = 0; a0 < nl;) {
            int b0 = a0;
            while (b0 < nl && tileVal[perm[b0]] == tileVal[perm[a0]]) b0++;
            int nextFree[8];
            for (int r = 0; r < 8; r++) { int p = a0 + ((r - a0) % 8 + 8) % 8; nextFree[r] = p < b0 ? p : -1; }
            for (int q = a0; q < b0; q++) {
                const int v = perm[q];
                int best = -1;
                long bestCost = 0;
                for (int r = 0; r < 8; r++) {
                    if (nextFree[r] < 0) continue;
                    long cost = 0;
                    int lastSet = -1;
                    for (int e = cornerStart[v]; e < cornerStart[v + 1]; e++) {
                        const int sid = setOf(cornerList[e] >> 2, cornerList[e] & 3);
                        if (sid == lastSet) continue;  // same vertex twice in a set is a broadcast, not a conflict
                        lastSet = sid;
                        cost += loadG[(size_t)sid * 8 + r];
                    }
                    if (best < 0 || cost < bestCost || (cost == bestCost && nextFree[r] < nextFree[best])) { best = r; bestCost = cost; }
                }
                const int pos = nextFree[best];
                rank_of[v] = pos;
                posTaken[pos] = 1;
                nextFree[best] = pos + 8 < b0 ? pos + 8 : -1;
                int lastSet = -1;
                for (int e = cornerStart[v]; e < cornerStart[v + 1]; e++) {
                    const int sid = setOf(cornerList[e] >> 2, cornerList[e] & 3);
                    if (sid == lastSet) continue;
                    lastSet = sid;
                    loadG[(size_t)sid * 8 + best]++;
                }
            }
            a0 = b0;
        }
        sorted.assign(nl, 0);
        for (int v = 0; v < nl; v++) sorted[rank_of[v]] = v;
        P.clVertStart[c + 1] = P.clVertStart[c] + nl;
        for (int j = 0; j < nl; j++) { P.clVerts.push_back(tileVerts[sorted[j]]); allVal.push_back((uint8_t)tileVal[sorted[j]]); }
        P.maxTileVerts = std::max(P.maxTileVerts, nl);
        // (b) diagonal of each corner
        loadS.assign((size_t)nSets * 8, 0);
        cornerDiag.assign(4 * (size_t)T, 0);
        for (int j = 0; j < nl; j++) {
            const int v = sorted[j], val = tileVal[v];
            diagFree.assign(val, 1);
            for (int e = cornerStart[v]; e < cornerStart[v + 1]; e++) {
                const int cn = cornerList[e], sid = setOf(cn >> 2, cn & 3);
                int best = -1, bestLoad = 0;
                for (int i = 0; i < val; i++) {
                    if (!diagFree[i]) continue;
                    const int ld = loadS[(size_t)sid * 8 + (entryOf(i, j) & 7)];
                    if (best < 0 || ld < bestLoad) { best = i; bestLoad = ld; if (ld == 0) break; }
                }
                diagFree[best] = 0;
                cornerDiag[cn] = best;
                loadS[(size_t)sid * 8 + (entryOf(best, j) & 7)]++;
            }
        }
        for (int tl = 0; tl < ntile; tl++) {
            const int *t = tetIds + 4 * (size_t)order[pb + tl];
            uint32_t sl[4], ds[4];
            for (int k = 0; k < 4; k++) {
                const int j = rank_of[tileIdx[local[t[k]]]];
                sl[k] = 16u * (uint32_t)j;
                ds[k] = 16u * (uint32_t)entryOf(cornerDiag[4 * tl + k], j);
            }
            size_t r = (size_t)c * T + tl;
            P.recordAux[4 * r + 0] = sl[0] | sl[1] << 16;
            P.recordAux[4 * r + 1] = sl[2] | sl[3] << 16;
            P.recordAux[4 * r + 2] = ds[0] | ds[1] << 16;
            P.recordAux[4 * r + 3] = ds[2] | ds[3] << 16;
        }
    }
    // padding records of the last tile: read tile vertex 0, park their (zero) dx in the spare entry after the largest tile
    P.maxTileEntries = maxTileEntries;
    for (size_t r = 0; r < nRec; r++)
        if (P.recordTet[r] < 0) {
            const uint32_t spare = 16u * (uint32_t)maxTileEntries;
            P.recordAux[4 * r + 2] = spare | spare << 16;
            P.recordAux[4 * r + 3] = spare | spare << 16;
        }
    if (P.maxTileVerts < 1) P.maxTileVerts = 1;
    if (16 * P.maxTileVerts > 65535 || 16 * (maxTileEntries + 1) > 65535) { err = "tile too large for 16-bit byte offsets"; return false; }
    P.colStride = (((P.maxTileVerts + kGroup - 1) / kGroup + 1 + 7) / 8) * 8;  // entries of the per-tile group-base table
    P.maxTileVertsPad = ((P.maxTileVerts + 15) / 16) * 16;
    P.metaValOff = 16 + 2 * P.colStride;
    P.metaStride = P.metaValOff + 5 * P.maxTileVertsPad;  // largest block (shared-memory slot size)
    P.metaOff.assign((size_t)P.numClusters + 1, 0);       // in units of 16 bytes
    for (int c = 0; c < P.numClusters; c++) {
        const int nl = P.clVertStart[c + 1] - P.clVertStart[c];
        const int nlPad = ((nl + 15) / 16) * 16;
        P.metaOff[c + 1] = P.metaOff[c] + (uint32_t)((P.metaValOff + 5 * nlPad) / 16);
    }
    P.tileMeta.assign((size_t)P.metaOff[P.numClusters] * 16, 0);
    for (int c = 0; c < P.numClusters; c++) {
        unsigned char *m = P.tileMeta.data() + (size_t)P.metaOff[c] * 16;
        const int v0 = P.clVertStart[c], nl = P.clVertStart[c + 1] - v0;
        const int nlPad = ((nl + 15) / 16) * 16;
        int hdr[4] = {v0, nl, (int)colOffs[c].size() - 1 /* groups */, P.metaValOff + nlPad /* byte offset of ids */};
        memcpy(m, hdr, 16);
        uint16_t *co = reinterpret_cast<uint16_t *>(m + 16);
        for (size_t i = 0; i < colOffs[c].size(); i++) co[i] = (uint16_t)(16u * colOffs[c][i]);
        memcpy(m + P.metaValOff, allVal.data() + v0, (size_t)nl);
        memcpy(m + P.metaValOff + nlPad, P.clVerts.data() + v0, 4 * (size_t)nl);
    }

    // vertex -> partial-sum slots, ascending tile order
    P.vpStart.assign((size_t)P.numLocalVerts + 1, 0);
    for (int v : P.clVerts) P.vpStart[(size_t)v + 1]++;
    for (int i = 0; i < P.numLocalVerts; i++) P.vpStart[i + 1] += P.vpStart[i];
    P.vpSlot.resize(P.clVerts.size());
    std::vector<int> fill(P.vpStart.begin(), P.vpStart.end() - 1);
    for (size_t s = 0; s < P.clVerts.size(); s++) P.vpSlot[fill[P.clVerts[s]]++] = (int)s;
    if (worldSize > 1 && P.haloOk) {  // fused peer exchange: partial index of every boundary-tile slot
        const size_t nSlots = (size_t)P.clVertStart[P.numBoundaryTiles];
        P.pxSlotIdx.assign(nSlots, 0xff);
        for (int b = 0; b < P.numBoundary; b++) {
            const int id = P.numInterior + b;
            const int n = P.vpStart[id + 1] - P.vpStart[id];
            P.pxMaxPartials = std::max(P.pxMaxPartials, n);
            for (int i = 0; i < n; i++) {
                const size_t slot = (size_t)P.vpSlot[P.vpStart[id] + i];
                if (slot >= nSlots) { err = "internal: a rank-shared vertex has a partial outside the boundary tiles"; return false; }
                P.pxSlotIdx[slot] = (uint8_t)std::min(i, 254);
            }
        }
    }
    return true;
}

}  // namespace tsim
