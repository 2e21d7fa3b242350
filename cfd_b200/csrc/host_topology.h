// Host-side integer artefacts of the product: node->element / node->node adjacency
// (PointNeighbor::getEsup/getPsup, pointNeighbor.f90:5-91), the Laplacian CSR pattern
// (Mlaplace::initialize, mLaplace.f90:60-94) and the per-node boundary-condition tables the
// fused node kernel uses.  Serial counting-sort code, run once per context; results are
// bit-exact against the oracle (tests/test_topology.py).
#pragma once
#include <algorithm>
#include <cstdint>
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

// Tiling for the fused RK stage: elements are grouped into spatially compact tiles of `tile` elements (Morton order
// of the centroids, cut every `tile` elements; ids ascending inside a tile so a warp reads runs of consecutive
// elements).  A node is INTERIOR to a tile when every element touching it belongs to that tile; all other nodes
// (and nodes without elements) are tile-boundary nodes.
struct Tiling {
    int ntiles = 0;
    vector<int32_t> tile_elems;    // ntiles*tile, 0-based element ids, -1 padding
    vector<uint8_t> ebmask;        // per tile position: bit n set when local node n is a tile-boundary node
    vector<int32_t> tnode_ptr;     // ntiles+1
    vector<int32_t> tnodes;        // interior nodes of each tile, ascending
    vector<uint16_t> tslot;        // per esup entry (original order): 3*position-in-tile + local node
    vector<int32_t> bnodes;        // tile-boundary nodes, ascending
    double interior_fraction = 0;
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
inline void build_tiling(const int32_t* inpoel, int nelem, int npoin, const double* X, const double* Y,
                         const vector<int32_t>& esup1, const vector<int32_t>& esup2, const vector<int32_t>& eslot, int tile,
                         Tiling& T) {
    double x0 = X[0], x1 = X[0], y0 = Y[0], y1 = Y[0];
    for (int n = 1; n < npoin; ++n) {
        x0 = std::min(x0, X[n]); x1 = std::max(x1, X[n]);
        y0 = std::min(y0, Y[n]); y1 = std::max(y1, Y[n]);
    }
    (void)x0; (void)x1; (void)y0; (void)y1;
    // k-d bisection of the element centroids (longer side of the bounding box, median split) down to <= tile elements:
    // compact, balanced tiles for any mesh density
    vector<double> cx((size_t)nelem), cy((size_t)nelem);
    for (int e = 0; e < nelem; ++e) {
        const int32_t* t = inpoel + 3 * (size_t)e;
        cx[e] = (X[t[0] - 1] + X[t[1] - 1] + X[t[2] - 1]) / 3.0;
        cy[e] = (Y[t[0] - 1] + Y[t[1] - 1] + Y[t[2] - 1]) / 3.0;
    }
    vector<int32_t> ids((size_t)nelem);
    for (int e = 0; e < nelem; ++e) ids[e] = e;
    vector<std::pair<size_t, size_t>> leaves, stack;
    stack.push_back({0, (size_t)nelem});
    while (!stack.empty()) {
        auto [lo, hi] = stack.back();
        stack.pop_back();
        if (hi - lo <= (size_t)tile) { leaves.push_back({lo, hi}); continue; }
        double ax0 = cx[ids[lo]], ax1 = ax0, ay0 = cy[ids[lo]], ay1 = ay0;
        for (size_t k = lo + 1; k < hi; ++k) {
            double a = cx[ids[k]], b = cy[ids[k]];
            ax0 = std::min(ax0, a); ax1 = std::max(ax1, a); ay0 = std::min(ay0, b); ay1 = std::max(ay1, b);
        }
        const bool by_x = (ax1 - ax0) >= (ay1 - ay0);
        // left part = half of the leaves this range needs, so leaves come out as full as possible
        size_t nleaf = (hi - lo + tile - 1) / tile;
        size_t mid = lo + (hi - lo) * (nleaf / 2) / nleaf;
        const vector<double>& c = by_x ? cx : cy;
        std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int32_t a, int32_t b) {
            return c[a] < c[b] || (c[a] == c[b] && a < b);
        });
        stack.push_back({mid, hi});
        stack.push_back({lo, mid});
    }
    std::sort(leaves.begin(), leaves.end());
    T.ntiles = (int)leaves.size();
    T.tile_elems.assign((size_t)T.ntiles * tile, -1);
    vector<int32_t> etile((size_t)nelem), epos((size_t)nelem);
    for (int t = 0; t < T.ntiles; ++t) {
        auto [lo, hi] = leaves[t];
        std::sort(ids.begin() + lo, ids.begin() + hi);
        for (size_t k = lo; k < hi; ++k) {
            T.tile_elems[(size_t)t * tile + (k - lo)] = ids[k];
            etile[ids[k]] = t;
            epos[ids[k]] = (int32_t)(k - lo);
        }
    }
    vector<int32_t> owner((size_t)npoin, -1);  // tile of an interior node, -1 for boundary nodes
    vector<int32_t> count((size_t)T.ntiles + 1, 0);
    for (int n = 0; n < npoin; ++n) {
        int k0 = esup2[n], k1 = esup2[n + 1];
        if (k0 == k1) continue;
        int t0 = etile[esup1[k0] - 1];
        bool same = true;
        for (int k = k0 + 1; k < k1 && same; ++k) same = etile[esup1[k] - 1] == t0;
        if (same) { owner[n] = t0; count[t0 + 1]++; }
    }
    for (int t = 0; t < T.ntiles; ++t) count[t + 1] += count[t];
    T.tnode_ptr = count;
    T.tnodes.assign(count[T.ntiles], 0);
    vector<int32_t> cur(count.begin(), count.end() - 1);
    T.bnodes.clear();
    for (int n = 0; n < npoin; ++n) {
        if (owner[n] >= 0) T.tnodes[cur[owner[n]]++] = n;
        else T.bnodes.push_back(n);
    }
    T.tslot.assign(esup1.size(), 0);
    for (size_t k = 0; k < esup1.size(); ++k) T.tslot[k] = (uint16_t)(epos[esup1[k] - 1] * 3 + eslot[k] % 3);
    T.ebmask.assign(T.tile_elems.size(), 0);
    for (size_t k = 0; k < T.tile_elems.size(); ++k) {
        int e = T.tile_elems[k];
        if (e < 0) continue;
        uint8_t m = 0;
        for (int i = 0; i < 3; ++i)
            if (owner[inpoel[3 * (size_t)e + i] - 1] < 0) m |= (uint8_t)(1u << i);
        T.ebmask[k] = m;
    }
    T.interior_fraction = npoin ? (double)T.tnodes.size() / npoin : 0.0;
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
