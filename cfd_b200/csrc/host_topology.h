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

// last[i] = index of the last entry of list[] naming the same node as entry i ("last entry wins",
// SURVEY.md B.2, for list-driven writes whose OpenMP order is undefined in the reference)
inline void last_wins(const int32_t* list, int m, int npoin, vector<int32_t>& last) {
    vector<int32_t> pos((size_t)npoin + 1, -1);
    for (int i = 0; i < m; ++i) pos[list[i]] = i;
    last.resize(m);
    for (int i = 0; i < m; ++i) last[i] = pos[list[i]];
}

}  // namespace topo
