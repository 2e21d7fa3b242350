// Host-side integer artefacts of the product: node->element / node->node adjacency
// (PointNeighbor::getEsup/getPsup, pointNeighbor.f90:5-91), the Laplacian CSR pattern
// (Mlaplace::initialize, mLaplace.f90:60-94) and the per-node boundary-condition tables the
// fused node kernel uses.  Serial counting-sort code, run once per context; results are
// bit-exact against the oracle (tests/test_topology.py).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace topo {

using std::vector;

// esup2: npoin+1 zero-based offsets; esup1: 1-based element ids, ascending per node;
// eslot (optional): 3*(e-1)+local for the same entries (where the node sits inside the element)
inline void build_esup(const int32_t* inpoel, int nelem, int npoin, vector<int32_t>& esup1, vector<int32_t>& esup2,
                       vector<int32_t>* eslot) {
    esup2.assign((size_t)npoin + 1, 0);
    for (size_t k = 0; k < 3 * (size_t)nelem; ++k) esup2[inpoel[k]]++;  // inpoel is 1-based: bucket n lands at n
    for (int n = 1; n <= npoin; ++n) esup2[n] += esup2[n - 1];
    esup1.assign(esup2[npoin], 0);
    if (eslot) eslot->assign(esup2[npoin], 0);
    vector<int32_t> cursor(esup2.begin(), esup2.end() - 1);
    for (int e = 0; e < nelem; ++e)
        for (int i = 0; i < 3; ++i) {
            int n = inpoel[3 * (size_t)e + i] - 1;
            int at = cursor[n]++;
            esup1[at] = e + 1;
            if (eslot) (*eslot)[at] = 3 * e + i;
        }
}

// psup in first-encounter order walking esup then local nodes 1..3 (marker array lpoin)
inline void build_psup(const int32_t* inpoel, int npoin, const vector<int32_t>& esup1, const vector<int32_t>& esup2,
                       vector<int32_t>& psup1, vector<int32_t>& psup2) {
    vector<int32_t> mark((size_t)npoin, 0);
    psup2.assign((size_t)npoin + 1, 0);
    psup1.clear();
    psup1.reserve(esup1.size() * 2 + 16);
    for (int n = 1; n <= npoin; ++n) {
        for (int k = esup2[n - 1]; k < esup2[n]; ++k) {
            const int32_t* el = inpoel + 3 * (size_t)(esup1[k] - 1);
            for (int i = 0; i < 3; ++i) {
                int j = el[i];
                if (j != n && mark[j - 1] != n) {
                    mark[j - 1] = n;
                    psup1.push_back(j);
                }
            }
        }
        psup2[n] = (int32_t)psup1.size();
    }
}

// CSR pattern: row n = [n, psup(n)...]; rowptr 0-based, idx 1-based
inline void build_lap_pattern(int npoin, const vector<int32_t>& psup1, const vector<int32_t>& psup2, vector<int32_t>& idx,
                              vector<int32_t>& rowptr) {
    rowptr.resize((size_t)npoin + 1);
    for (int n = 0; n <= npoin; ++n) rowptr[n] = psup2[n] + n;
    idx.resize(psup1.size() + npoin);
    for (int n = 0; n < npoin; ++n) {
        int at = rowptr[n];
        idx[at++] = n + 1;
        for (int k = psup2[n]; k < psup2[n + 1]; ++k) idx[at++] = psup1[k];
    }
}

// lpos[3*k+j] = position inside row(n) of local node j of esup entry k (n = the node owning entry k)
inline void build_lap_pos(const int32_t* inpoel, int npoin, const vector<int32_t>& esup1, const vector<int32_t>& esup2,
                          const vector<int32_t>& idx, const vector<int32_t>& rowptr, vector<uint8_t>& lpos, int& maxrow) {
    lpos.assign(3 * esup1.size(), 0);
    maxrow = 0;
    for (int n = 0; n < npoin; ++n) {
        int r0 = rowptr[n], r1 = rowptr[n + 1];
        maxrow = std::max(maxrow, r1 - r0);
        for (int k = esup2[n]; k < esup2[n + 1]; ++k) {
            const int32_t* el = inpoel + 3 * (size_t)(esup1[k] - 1);
            for (int j = 0; j < 3; ++j) {
                int pos = 0;
                for (int t = r0; t < r1; ++t)
                    if (idx[t] == el[j]) { pos = t - r0; break; }
                lpos[3 * (size_t)k + j] = (uint8_t)pos;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tiling for the fused RK stage (kernels.cuh: stage_fused).
//
// Element order.  The library keeps its element arrays in an INTERNAL order: recursive coordinate bisection of the element
// centroids into tiles of TE consecutive elements (rcb_split below).  A tile is then a compact patch of the
// mesh whatever numbering the mesh file uses, and everything the stage kernel streams per tile (connectivity, geometry,
// stabilisation parameters) is one contiguous run per array -- a handful of bulk copies.  i2e[pos] = original element at
// internal position pos, e2i = inverse.  The C ABI keeps speaking the file's numbering (cfdb.cu permutes at get/set).
//
// Nodes keep their numbering.  A node is INTERIOR to tile t when every element touching it lies in t (~85 % of the nodes at
// TE = 384): its ordered sum and its nodal update are finished inside the tile's CTA from shared memory.  All other nodes
// (shared by two or more tiles, or touched by no element) are tile-BOUNDARY nodes: their contributions go through the
// staging buffer EC and node_update runs over the list `bnodes`.
//
// Per tile one static block of tb_bytes (TileLayout gives the offsets; every section starts on a 16-byte boundary):
//   header  int32[4]           ne (elements in the tile), ntn (nodes touched), nint (interior nodes), 0
//   lnode   uint16[3][TE]      tile-local index of each element vertex
//   tnode   int32[ntn_max]     node id of each tile-local index: the nint interior nodes first (ascending), then the others
//   nptr    uint16[nint_max+1] CSR over `slots` for the interior nodes
//   slots   uint16[nslot_max]  contributions of each interior node in ascending ORIGINAL element order (the reference's
//                              1-thread summation order); value = (4*local_vertex)*TE + position in tile, i.e. the index
//                              of equation 0 of that contribution in the shared-memory array C[12][TE]
//   bcf     uint8[nint_max]    bcflag of the interior nodes
//   brank   uint8[3][TE]       for an element vertex that is a tile-boundary node: the position of this element in the node's
//                              element list (ascending ORIGINAL element id)
//   bbase   uint32[nbd_max]    for each tile-boundary node of the tile (local index nint + jb): the first record of that node in
//                              the boundary staging buffer ECB.  An element warp stores the contribution of (element, vertex)
//                              at record bbase[lnode - nint] + brank, so the records of one node lie side by side in summation
//                              order and the boundary pass (kernels.cuh: boundary_update) reads them as one contiguous run.
struct TileLayout {
    int TE = 0, ntn_max = 0, nint_max = 0, nslot_max = 0, nbd_max = 0;
    int off_lnode = 0, off_tnode = 0, off_nptr = 0, off_slots = 0, off_bcf = 0, off_brank = 0, off_bbase = 0, tb_bytes = 0;
};
struct Tiling {
    TileLayout L;
    int ntiles = 0;
    vector<int32_t> i2e, e2i;      // 0-based
    vector<uint8_t> blocks;        // ntiles * L.tb_bytes
    vector<int32_t> bnodes;        // tile-boundary nodes, ascending
    vector<int32_t> bn_ptr;        // CSR over the boundary staging buffer: contributions of bnodes[i] are records bn_ptr[i] .. bn_ptr[i+1]
    vector<int32_t> orphans;       // nodes touched by no element (never reached by the stage kernel), ascending
    double interior_fraction = 0;
    bool rank_overflow = false;    // a tile-boundary node with more than 256 elements: the tiling cannot be used
};
inline uint32_t morton16(uint32_t x, uint32_t y) {
    auto spread = [](uint32_t v) {
        v &= 0xffffu;
        v = (v | (v << 8)) & 0x00ff00ffu;
        v = (v | (v << 4)) & 0x0f0f0f0fu;
        v = (v | (v << 2)) & 0x33333333u;
        v = (v | (v << 1)) & 0x55555555u;
        return v;
    };
    return spread(x) | (spread(y) << 1);
}
// Recursive coordinate bisection of the element centroids into runs of TE elements: n elements = m tiles are cut at
// floor(m/2)*TE along the longer side of their bounding box (nth_element; ties by element id), recursively, so every tile
// but the last is full and is a near-square patch whatever the mesh numbering or grading.  On a structured triangulation
// 86 % of the nodes end up interior to one tile at TE = 384 (Morton runs of 384: 55 %; a Z-curve run whose length is not a
// power of four is a ragged shape).  Tiles come out in kd-tree order (neighbouring tiles are neighbours in memory); inside a
// tile the elements are in ascending original id.
struct RcbPoint { double x, y; int32_t e; };
inline void rcb_split(RcbPoint* c, size_t n, size_t TE) {
    while (n > TE) {
        const size_t m = (n + TE - 1) / TE, left = (m / 2) * TE;
        double x0 = c[0].x, x1 = c[0].x, y0 = c[0].y, y1 = c[0].y;
        for (size_t i = 1; i < n; ++i) {
            x0 = std::min(x0, c[i].x); x1 = std::max(x1, c[i].x);
            y0 = std::min(y0, c[i].y); y1 = std::max(y1, c[i].y);
        }
        if (x1 - x0 >= y1 - y0)
            std::nth_element(c, c + left, c + n, [](const RcbPoint& a, const RcbPoint& b) { return a.x < b.x || (a.x == b.x && a.e < b.e); });
        else
            std::nth_element(c, c + left, c + n, [](const RcbPoint& a, const RcbPoint& b) { return a.y < b.y || (a.y == b.y && a.e < b.e); });
        if (left >= ((size_t)1 << 15)) {
#pragma omp task default(none) firstprivate(c, left, TE)
            rcb_split(c, left, TE);
        } else {
            rcb_split(c, left, TE);
        }
        c += left;
        n -= left;
    }
    std::sort(c, c + n, [](const RcbPoint& a, const RcbPoint& b) { return a.e < b.e; });
}
enum TileOrder { ORDER_FILE = 0, ORDER_MORTON = 1, ORDER_RCB = 2 };
inline void element_order(const int32_t* inpoel, int nelem, int npoin, const double* X, const double* Y, int TE, int order, vector<int32_t>& i2e) {
    const size_t E = nelem;
    i2e.resize(E);
    if (order == ORDER_FILE || E == 0) {
        for (size_t p = 0; p < E; ++p) i2e[p] = (int32_t)p;
        return;
    }
    if (order == ORDER_RCB) {
        vector<RcbPoint> c(E);
        for (size_t e = 0; e < E; ++e) {
            const int32_t* t = inpoel + 3 * e;
            c[e].x = (X[t[0] - 1] + X[t[1] - 1] + X[t[2] - 1]) / 3.0;
            c[e].y = (Y[t[0] - 1] + Y[t[1] - 1] + Y[t[2] - 1]) / 3.0;
            c[e].e = (int32_t)e;
        }
#pragma omp parallel
#pragma omp single
        rcb_split(c.data(), E, (size_t)TE);
        for (size_t p = 0; p < E; ++p) i2e[p] = c[p].e;
        return;
    }
    double x0 = X[0], x1 = X[0], y0 = Y[0], y1 = Y[0];
    for (int n = 1; n < npoin; ++n) {
        x0 = std::min(x0, X[n]); x1 = std::max(x1, X[n]);
        y0 = std::min(y0, Y[n]); y1 = std::max(y1, Y[n]);
    }
    const double span = std::max(x1 - x0, y1 - y0);
    const double q = span > 0 ? 65535.0 / span : 0.0;   // one scale for both axes: square cells
    vector<uint64_t> key(E);
    for (size_t e = 0; e < E; ++e) {
        const int32_t* t = inpoel + 3 * e;
        double cx = (X[t[0] - 1] + X[t[1] - 1] + X[t[2] - 1]) / 3.0, cy = (Y[t[0] - 1] + Y[t[1] - 1] + Y[t[2] - 1]) / 3.0;
        uint32_t ix = (uint32_t)std::min(65535.0, std::max(0.0, (cx - x0) * q)), iy = (uint32_t)std::min(65535.0, std::max(0.0, (cy - y0) * q));
        key[e] = ((uint64_t)morton16(ix, iy) << 32) | (uint64_t)e;
    }
    std::sort(key.begin(), key.end());
    for (size_t p = 0; p < E; ++p) i2e[p] = (int32_t)(key[p] & 0xffffffffu);
}
// inpoel: (3,nelem) 1-based, original order; esup1 (1-based original element ids) / esup2 / eslot (3*(e-1)+local) in the
// reference's order (ascending element id per node); bcflag[npoin]
inline void build_tiling(const int32_t* inpoel, int nelem, int npoin, const double* X, const double* Y,
                         const vector<int32_t>& esup1, const vector<int32_t>& esup2, const vector<int32_t>& eslot,
                         const vector<uint8_t>& bcflag, int TE, int order, Tiling& T) {
    const size_t E = nelem;
    T.i2e.resize(E);
    T.e2i.resize(E);
    element_order(inpoel, nelem, npoin, X, Y, TE, order, T.i2e);
    for (size_t p = 0; p < E; ++p) T.e2i[T.i2e[p]] = (int32_t)p;
    const int nt = (int)((E + TE - 1) / TE);
    T.ntiles = nt;
    // interior tile of every node (-1: boundary)
    vector<int32_t> owner((size_t)npoin, -1);
    long ninterior = 0;
#pragma omp parallel for schedule(static) reduction(+ : ninterior)
    for (int n = 0; n < npoin; ++n) {
        int k0 = esup2[n], k1 = esup2[n + 1];
        if (k0 == k1) continue;
        int t0 = T.e2i[esup1[k0] - 1] / TE;
        bool same = true;
        for (int k = k0 + 1; k < k1 && same; ++k) same = T.e2i[esup1[k] - 1] / TE == t0;
        if (same) { owner[n] = t0; ++ninterior; }
    }
    T.bnodes.clear();
    T.orphans.clear();
    vector<int32_t> bpos((size_t)npoin, -1);
    T.bn_ptr.assign(1, 0);
    for (int n = 0; n < npoin; ++n)
        if (owner[n] < 0) {
            bpos[n] = (int32_t)T.bnodes.size();
            T.bnodes.push_back(n);
            if (esup2[n] == esup2[n + 1]) T.orphans.push_back(n);
            T.bn_ptr.push_back(T.bn_ptr.back() + (esup2[n + 1] - esup2[n]));
        }
    T.interior_fraction = npoin ? (double)ninterior / npoin : 0.0;
    // The tiles are independent of each other: both passes run over them in parallel (OpenMP).  A tile's node lists -- its
    // interior nodes ascending, then the others ascending -- come from sorting its own <= 3*TE vertex ids, so no mesh-sized
    // scratch array is shared between tiles.
    auto tile_nodes = [&](int t, vector<int32_t>& a, vector<int32_t>& b, vector<int32_t>& all) {
        size_t p0 = (size_t)t * TE, p1 = std::min(E, p0 + TE);
        all.clear();
        for (size_t p = p0; p < p1; ++p) {
            const int32_t* el = inpoel + 3 * (size_t)T.i2e[p];
            all.push_back(el[0] - 1); all.push_back(el[1] - 1); all.push_back(el[2] - 1);
        }
        std::sort(all.begin(), all.end());
        all.erase(std::unique(all.begin(), all.end()), all.end());
        a.clear(); b.clear();
        for (int n : all) (owner[n] == t ? a : b).push_back(n);
    };
    // pass 1: the section sizes
    int ntn_max = 0, nint_max = 0, nslot_max = 0, nbd_max = 0;
#pragma omp parallel
    {
        vector<int32_t> a, b, all;
#pragma omp for schedule(dynamic, 64) reduction(max : ntn_max, nint_max, nslot_max, nbd_max)
        for (int t = 0; t < nt; ++t) {
            tile_nodes(t, a, b, all);
            int ns = 0;
            for (int n : a) ns += esup2[n + 1] - esup2[n];
            ntn_max = std::max(ntn_max, (int)(a.size() + b.size()));
            nint_max = std::max(nint_max, (int)a.size());
            nslot_max = std::max(nslot_max, ns);
            nbd_max = std::max(nbd_max, (int)b.size());
        }
    }
    auto up16 = [](int v) { return (v + 15) & ~15; };
    TileLayout& L = T.L;
    L.TE = TE;
    L.ntn_max = ntn_max; L.nint_max = nint_max; L.nslot_max = nslot_max; L.nbd_max = nbd_max;
    L.off_lnode = 16;
    L.off_tnode = up16(L.off_lnode + 3 * TE * 2);
    L.off_nptr = up16(L.off_tnode + ntn_max * 4);
    L.off_slots = up16(L.off_nptr + (nint_max + 1) * 2);
    L.off_bcf = up16(L.off_slots + nslot_max * 2);
    L.off_brank = up16(L.off_bcf + nint_max);
    L.off_bbase = up16(L.off_brank + 3 * TE);
    L.tb_bytes = up16(L.off_bbase + nbd_max * 4);
    // pass 2: fill the blocks
    T.blocks.assign((size_t)nt * L.tb_bytes, 0);
    bool overflow = false;
#pragma omp parallel
    {
        vector<int32_t> a, b, all;
#pragma omp for schedule(dynamic, 64) reduction(|| : overflow)
        for (int t = 0; t < nt; ++t) {
            tile_nodes(t, a, b, all);
            uint8_t* blk = T.blocks.data() + (size_t)t * L.tb_bytes;
            int32_t* hdr = reinterpret_cast<int32_t*>(blk);
            uint16_t* lnode = reinterpret_cast<uint16_t*>(blk + L.off_lnode);
            int32_t* tnode = reinterpret_cast<int32_t*>(blk + L.off_tnode);
            uint16_t* nptr = reinterpret_cast<uint16_t*>(blk + L.off_nptr);
            uint16_t* slots = reinterpret_cast<uint16_t*>(blk + L.off_slots);
            uint8_t* bcf = blk + L.off_bcf;
            uint8_t* brank = blk + L.off_brank;
            uint32_t* bbase = reinterpret_cast<uint32_t*>(blk + L.off_bbase);
            size_t p0 = (size_t)t * TE, p1 = std::min(E, p0 + TE);
            const int nint = (int)a.size(), ntn = nint + (int)b.size();
            hdr[0] = (int32_t)(p1 - p0); hdr[1] = ntn; hdr[2] = nint; hdr[3] = 0;
            for (int j = 0; j < nint; ++j) tnode[j] = a[j];
            for (int j = nint; j < ntn; ++j) tnode[j] = b[j - nint];
            auto local = [&](int n) -> int {      // tile-local index of node n: position in a (interior) or nint + position in b
                if (owner[n] == t) return (int)(std::lower_bound(a.begin(), a.end(), n) - a.begin());
                return nint + (int)(std::lower_bound(b.begin(), b.end(), n) - b.begin());
            };
            for (size_t p = p0; p < p1; ++p) {
                const int32_t* el = inpoel + 3 * (size_t)T.i2e[p];
                for (int i = 0; i < 3; ++i) {
                    const int n = el[i] - 1;
                    lnode[(size_t)i * TE + (p - p0)] = (uint16_t)local(n);
                    if (bpos[n] >= 0) {   // rank of this element among the node's elements (the list is in ascending original id)
                        int k = esup2[n];
                        while (eslot[k] != 3 * T.i2e[p] + i) ++k;
                        if (k - esup2[n] > 255) overflow = true;
                        brank[(size_t)i * TE + (p - p0)] = (uint8_t)(k - esup2[n]);
                    }
                }
            }
            int q = 0;
            for (int j = 0; j < nint; ++j) {
                int n = tnode[j];
                nptr[j] = (uint16_t)q;
                for (int k = esup2[n]; k < esup2[n + 1]; ++k) {
                    int pos = T.e2i[esup1[k] - 1] - (int)p0, ln = eslot[k] % 3;
                    slots[q++] = (uint16_t)(4 * ln * TE + pos);
                }
                bcf[j] = bcflag[n];
            }
            nptr[nint] = (uint16_t)q;
            for (int jb = 0; jb < ntn - nint; ++jb) bbase[jb] = (uint32_t)T.bn_ptr[bpos[tnode[nint + jb]]];
        }
    }
    T.rank_overflow = overflow;
}

// Independent check of a tiling against the mesh it was built from (host code, used by cfdb_tile_elements and the CPU tests):
// every element sits in exactly one tile position; lnode names the element's own vertices; an interior node's slot list is
// its element list in ascending ORIGINAL element id with the right local vertex; the boundary records bbase + brank map the
// contributions of every tile-boundary node one-to-one onto its run bn_ptr[i] .. bn_ptr[i+1], in ascending original element id.
// Returns the number of violations (0 = consistent).
inline long check_tiling(const int32_t* inpoel, int nelem, int npoin, const vector<int32_t>& esup1, const vector<int32_t>& esup2,
                         const vector<int32_t>& eslot, const Tiling& T) {
    const TileLayout& L = T.L;
    const int TE = L.TE;
    long bad = 0;
    vector<uint8_t> seen((size_t)nelem, 0);
    for (size_t p = 0; p < (size_t)nelem; ++p) {
        if (T.i2e[p] < 0 || T.i2e[p] >= nelem || seen[T.i2e[p]]++) ++bad;
        else if (T.e2i[T.i2e[p]] != (int32_t)p) ++bad;
    }
    vector<int32_t> bpos((size_t)npoin, -1);
    for (size_t i = 0; i < T.bnodes.size(); ++i) bpos[T.bnodes[i]] = (int32_t)i;
    vector<uint8_t> rec_hit(T.bn_ptr.empty() ? 0 : (size_t)T.bn_ptr.back(), 0);
    vector<uint8_t> node_done((size_t)npoin, 0);
    for (int t = 0; t < T.ntiles; ++t) {
        const uint8_t* blk = T.blocks.data() + (size_t)t * L.tb_bytes;
        const int32_t* hdr = reinterpret_cast<const int32_t*>(blk);
        const uint16_t* lnode = reinterpret_cast<const uint16_t*>(blk + L.off_lnode);
        const int32_t* tnode = reinterpret_cast<const int32_t*>(blk + L.off_tnode);
        const uint16_t* nptr = reinterpret_cast<const uint16_t*>(blk + L.off_nptr);
        const uint16_t* slots = reinterpret_cast<const uint16_t*>(blk + L.off_slots);
        const uint8_t* brank = blk + L.off_brank;
        const uint32_t* bbase = reinterpret_cast<const uint32_t*>(blk + L.off_bbase);
        const size_t p0 = (size_t)t * TE;
        const int ne = hdr[0], ntn = hdr[1], nint = hdr[2];
        if (ne != (int)std::min((size_t)TE, (size_t)nelem - p0) || nint > ntn || ntn > L.ntn_max) ++bad;
        for (int k = 0; k < ne; ++k)
            for (int i = 0; i < 3; ++i) {
                const int ln = lnode[(size_t)i * TE + k];
                const int n = inpoel[3 * (size_t)T.i2e[p0 + k] + i] - 1;
                if (ln >= ntn || tnode[ln] != n) { ++bad; continue; }
                if (ln >= nint) {   // tile-boundary node: its record
                    if (bpos[n] < 0) { ++bad; continue; }
                    const long rec = (long)bbase[ln - nint] + brank[(size_t)i * TE + k];
                    const int k0 = T.bn_ptr[bpos[n]], k1 = T.bn_ptr[bpos[n] + 1];
                    if (bbase[ln - nint] != (uint32_t)k0 || rec < k0 || rec >= k1 || rec_hit[rec]++) { ++bad; continue; }
                    // record position = position of this (element, vertex) in the node's list
                    if (eslot[esup2[n] + (rec - k0)] != 3 * T.i2e[p0 + k] + i) ++bad;
                }
            }
        for (int j = 0; j < nint; ++j) {
            const int n = tnode[j];
            if (bpos[n] >= 0 || node_done[n]++) ++bad;
            if (nptr[j + 1] - nptr[j] != esup2[n + 1] - esup2[n]) { ++bad; continue; }
            for (int q = nptr[j], k = esup2[n]; q < nptr[j + 1]; ++q, ++k) {
                const int ln = slots[q] / (4 * TE), pos = slots[q] % (4 * TE);
                if (pos >= ne || ln > 2) { ++bad; continue; }
                if (T.i2e[p0 + pos] != esup1[k] - 1 || eslot[k] != 3 * (esup1[k] - 1) + ln) ++bad;
            }
        }
    }
    for (uint8_t h : rec_hit) bad += h != 1;
    for (int n = 0; n < npoin; ++n)
        if ((bpos[n] >= 0) == (node_done[n] != 0)) ++bad;     // every node is interior to one tile XOR in the boundary list
    return bad;
}

// last[i] = index of the last entry of list[] naming the same node as entry i ("last entry wins",
// SURVEY.md B.2, for list-driven writes whose OpenMP order is undefined in the reference)
inline void last_wins(const int32_t* list, int m, int npoin, vector<int32_t>& last) {
    vector<int32_t> pos((size_t)npoin + 1, -1);
    for (int i = 0; i < m; ++i) pos[list[i]] = i;
    last.resize(m);
    for (int i = 0; i < m; ++i) last[i] = pos[list[i]];
}

}  // namespace topo
