// Device-side construction of the integer artefacts cfdb_create needs: PointNeighbor::getEsup/getPsup
// (pointNeighbor.f90:5-91: "histogram + scan + fill", as the reference's own comments at :13,:21 put it),
// Mlaplace::initialize's CSR pattern (mLaplace.f90:60-94) and the row positions the laplace kernel uses.
//
// Bit-exact against the serial host code (host_topology.h) and the oracle:
//   * esup: slots s = 3e+i are generated in ascending order and sorted by node with a STABLE radix sort, so every node's
//     elements come out in ascending element order -- the order the reference's cursor fill produces;
//   * psup: one thread per node walks its esup entries and local nodes 1..3 and keeps first encounters, which is what the
//     reference's lpoin marker does;
//   * pattern: row n = [n, psup(n)...], rowptr(n) = psup2(n) + n.
// Library code used: cub::DeviceRadixSort / DeviceScan (set-up only, not on the time-step path).
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdint>

namespace topogpu {

constexpr int kMaxNbr = 31;  // row length <= 32 (the laplace kernel's limit, checked by the caller)

__global__ void slot_keys(int nelem, const int* __restrict__ inp, int* __restrict__ keys, int* __restrict__ vals,
                          int* __restrict__ counts) {
    long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= 3L * nelem) return;
    int e = (int)(s / 3), i = (int)(s - 3L * e);
    int n = inp[(size_t)i * nelem + e];
    keys[s] = n;
    vals[s] = (int)s;
    atomicAdd(&counts[n + 1], 1);
}

// FILL == false: cnt[n+1] = number of distinct neighbours, *maxrow = max row length
// FILL == true : psup1 (1-based), CSR pattern idx0 (0-based, diagonal first), lpos[3k+j] = position inside row(n) of local
//                node j of esup entry k
template <bool FILL>
__global__ void psup_pass(int npoin, int nelem, const int* __restrict__ esup2, const int* __restrict__ eslot,
                          const int* __restrict__ inp, int* __restrict__ cnt, const int* __restrict__ psup2,
                          int* __restrict__ psup1, int* __restrict__ rowptr, int* __restrict__ idx0,
                          unsigned char* __restrict__ lpos, int* __restrict__ maxrow) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= npoin) return;
    int lst[kMaxNbr];
    int m = 0;
    for (int k = esup2[n]; k < esup2[n + 1]; ++k) {
        int e = eslot[k] / 3;
        int el[3] = {inp[e], inp[(size_t)nelem + e], inp[2 * (size_t)nelem + e]};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            int j = el[i];
            int pos = 0;
            if (j != n) {
                int t = 0;
                while (t < m && t < kMaxNbr && lst[t] != j) ++t;
                if (t == m || t >= kMaxNbr) {
                    if (m < kMaxNbr) lst[m] = j;
                    t = m;
                    ++m;
                }
                pos = t + 1;
            }
            if (FILL) lpos[3 * (size_t)k + i] = (unsigned char)pos;
        }
    }
    if (!FILL) {
        cnt[n + 1] = m;
        atomicMax(maxrow, m + 1);
    } else {
        int p0 = psup2[n], r0 = p0 + n;
        rowptr[n] = r0;
        if (n == npoin - 1) rowptr[npoin] = psup2[npoin] + npoin;
        idx0[r0] = n;
        for (int t = 0; t < m && t < kMaxNbr; ++t) {
            psup1[p0 + t] = lst[t] + 1;
            idx0[r0 + 1 + t] = lst[t];
        }
    }
}

#define TG(x)                         \
    do {                              \
        cudaError_t _e = (x);         \
        if (_e != cudaSuccess) { rc = (int)_e; goto done; } \
    } while (0)

// inp: device, SoA [3][nelem], 0-based.  Outputs are device arrays the caller allocated: esup2[npoin+1], eslot[3*nelem],
// psup2[npoin+1], rowptr[npoin+1], lpos[9*nelem]; psup1 / idx0 are allocated here once their sizes are known (cudaMalloc,
// ownership passes to the caller).  Returns 0 or a cudaError_t.
int build(cudaStream_t st, const int* inp, int nelem, int npoin, int* esup2, int* eslot, int* psup2, int** psup1_out,
          int* npsup_out, int* rowptr, int** idx0_out, unsigned char* lpos, int* maxrow_out) {
    int rc = 0;
    const size_t S = 3 * (size_t)nelem;
    int *keys = nullptr, *keys2 = nullptr, *vals = nullptr, *counts = nullptr, *d_max = nullptr, *psup1 = nullptr, *idx0 = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0, need = 0;
    int bits = 1, npsup = 0, maxrow = 0;
    while ((1L << bits) < npoin) ++bits;
    TG(cudaMalloc(&keys, S * sizeof(int)));
    TG(cudaMalloc(&keys2, S * sizeof(int)));
    TG(cudaMalloc(&vals, S * sizeof(int)));
    TG(cudaMalloc(&counts, ((size_t)npoin + 1) * sizeof(int)));
    TG(cudaMalloc(&d_max, sizeof(int)));
    TG(cudaMemsetAsync(counts, 0, ((size_t)npoin + 1) * sizeof(int), st));
    TG(cudaMemsetAsync(d_max, 0, sizeof(int), st));
    if (S) slot_keys<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(nelem, inp, keys, vals, counts);
    TG(cudaGetLastError());
    TG(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, eslot, (long)S, 0, bits, st));
    TG(cub::DeviceScan::InclusiveSum(nullptr, need, counts, esup2, npoin + 1, st));
    if (need > tmp_bytes) tmp_bytes = need;
    TG(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    TG(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, eslot, (long)S, 0, bits, st));
    TG(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, counts, esup2, npoin + 1, st));
    // psup: count, scan, fill
    TG(cudaMemsetAsync(counts, 0, ((size_t)npoin + 1) * sizeof(int), st));
    if (npoin)
        psup_pass<false><<<(npoin + 127) / 128, 128, 0, st>>>(npoin, nelem, esup2, eslot, inp, counts, nullptr, nullptr, nullptr,
                                                             nullptr, nullptr, d_max);
    TG(cudaGetLastError());
    TG(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, counts, psup2, npoin + 1, st));
    TG(cudaMemcpyAsync(&npsup, psup2 + npoin, sizeof(int), cudaMemcpyDeviceToHost, st));
    TG(cudaMemcpyAsync(&maxrow, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    TG(cudaStreamSynchronize(st));
    *maxrow_out = maxrow;
    *npsup_out = npsup;
    if (maxrow > kMaxNbr + 1) goto done;  // caller reports "valence not supported"
    TG(cudaMalloc(&psup1, ((size_t)npsup + 1) * sizeof(int)));
    TG(cudaMalloc(&idx0, ((size_t)npsup + npoin + 1) * sizeof(int)));
    if (npoin)
        psup_pass<true><<<(npoin + 127) / 128, 128, 0, st>>>(npoin, nelem, esup2, eslot, inp, nullptr, psup2, psup1, rowptr, idx0,
                                                            lpos, nullptr);
    TG(cudaGetLastError());
    TG(cudaStreamSynchronize(st));
    *psup1_out = psup1;
    *idx0_out = idx0;
    psup1 = nullptr;
    idx0 = nullptr;
done:
    cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(counts); cudaFree(d_max); cudaFree(tmp);
    cudaFree(psup1); cudaFree(idx0);
    return rc;
}

}  // namespace topogpu
