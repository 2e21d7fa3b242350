// Fused RK stage for fixed meshes (included inside namespace CFDB_KNS by kernels.cuh).
//
// One RK stage of the reference is  RHS = 0 ; calcRHS ; U1 = U - rk/M*RHS ; primitives ; fixvel/normalvel/FIX ; conservative
// (subrutinas.f90:667-826, calcRHS.f90:36-151).  The two-kernel stage (calcrhs_elem + node_update) pays for the reference's
// summation order with a 96 B/element staging buffer written and read back through HBM.  This kernel keeps that buffer in
// shared memory for ~80 % of the nodes:
//
//   * the elements are stored in tile order (host_topology.h: build_tiling): tile t = TE consecutive internal elements, a
//     compact patch of the mesh; a node all of whose elements lie in one tile is INTERIOR to it;
//   * a persistent CTA (one per SM) walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...;
//   * warp specialisation: ONE loader warp brings everything a tile needs into a two-stage shared-memory ring while the
//     NCW compute warps work on the previous tile -- the tile's static block and its element stream (geometry,
//     stabilisation parameters) by cp.async.bulk (TMA bulk copies completing on an mbarrier's transaction count), the
//     nodal state of the tile's nodes by cp.async gathers (one 32-byte sector per node) tracked by the same mbarrier.
//     Compute warps therefore never wait for HBM: their inputs are in shared memory when full[stage] completes;
//   * compute warps: thread = element; gradients and the 12 contributions in registers (calcrhs_body, unchanged
//     arithmetic), results to the shared-memory array C[12][TE]; contributions to tile-boundary nodes ALSO go to the global
//     staging buffer EC.  After a CTA barrier the same threads become node threads: interior node j sums its
//     contributions from C in ascending ORIGINAL element order (the `slots` list of the static block) and runs the nodal
//     chain (node_finish_v) -- same operations, same order, same bits as node_update;
//   * setmaxnreg moves registers from the loader's warpgroup to the compute warpgroups.
// Tile-boundary nodes (~20 %) are finished by node_update over the list `bnodes` right after this kernel.
//
// Roofline: HBM traffic per element-stage falls from 404 B (220 algorithmic + 192 staging, measured 6.46 GB per stage on the
// 16 M-triangle mesh) to ~230 B; the kernel is then bound by the fp64 pipe (~812 non-FMA fp64 instructions per element).
#pragma once

namespace ptx {
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// Bounded wait: a protocol bug must end in a trap (sticky error the host sees), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s at 2 GHz
    }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
}  // namespace ptx

// offsets of the static tile block (mirror of topo::TileLayout) and of one ring stage
struct TileGeom {
    int TE, ntn_max, nint_max, nslot_max;
    int off_lnode, off_tnode, off_nptr, off_slots, off_bcf, tb_bytes;
    // one stage of the ring, in bytes from the stage base
    int st_static, st_stream, st_u, st_t, st_m, st_g, stage_bytes;
    int nfields;       // doubles per element in the stream: 11, or 12 with a local time step array
    int off_c;         // C[12][TE] from the shared-memory base (after the barriers)
    int off_stage0;
};
struct StageArgs {
    int ntiles;
    const unsigned char* TB;                       // static tile blocks
    const double* geo;                             // [7][Epad]: dNx(3), dNy(3), area in tile order, Epad a multiple of TE
    long Epad;
    const double *shoc, *ts1, *ts2, *ts3, *dtl_arr; // [Epad] each (dtl_arr may be null)
    const double* dtl_sc;
    const double* Usrc;                            // state calcRHS is evaluated at (U; U1 for true_rk stages 2..4)
    const double* U;                               // state the update starts from
    const double* T;
    const double *M, *GAMM, *WXa, *WYa;
    BcTab bc;
    double rk_fact, FR;
    Gas g;
    double *EC, *U1, *RHS, *RHO, *VELX, *VELY, *Ea, *Pa, *Ta, *RMACH;
};

// everything a compute thread reads for its element comes from the ring stage `sb` (shared memory)
template <bool VISC, bool NB>
__device__ __forceinline__ unsigned fused_elem(const TileGeom& G, const StageArgs& A, const unsigned char* sb, double* C,
                                               int k, int nint, long e_glob, double dtl_uniform) {
    const int TE = G.TE;
    const unsigned short* lnode = reinterpret_cast<const unsigned short*>(sb + G.st_static + G.off_lnode);
    const int ln[3] = {lnode[k], lnode[TE + k], lnode[2 * TE + k]};
    const double* ut = reinterpret_cast<const double*>(sb + G.st_u);
    double Un[3][4], Th[3][4], Tn[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double2* q = reinterpret_cast<const double2*>(ut + 4 * ln[j]);
        double2 a = q[0], b = q[1];
        Un[j][0] = a.x; Un[j][1] = a.y; Un[j][2] = b.x; Un[j][3] = b.y;
    }
    if (VISC) {
        const double* tt = reinterpret_cast<const double*>(sb + G.st_t);
        Tn[0] = tt[ln[0]]; Tn[1] = tt[ln[1]]; Tn[2] = tt[ln[2]];
    }
    const double* sd = reinterpret_cast<const double*>(sb + G.st_stream);
    double Nx[3] = {sd[k], sd[TE + k], sd[2 * TE + k]};
    double Ny[3] = {sd[3 * TE + k], sd[4 * TE + k], sd[5 * TE + k]};
    const double ar = sd[6 * TE + k];
    const double shoc_e = sd[7 * TE + k];
    const double tau[3] = {sd[8 * TE + k], sd[9 * TE + k], sd[10 * TE + k]};
    const double dtl = G.nfields == 12 ? sd[11 * TE + k] : dtl_uniform;
    double Ux[4], Uy[4], rt[3][4];
    unsigned bad = 0;
    calcrhs_body<VISC, false, NB>(A.g, Un, Th, Tn, Nx, Ny, tau, shoc_e, Ux, Uy, rt, &bad);
#pragma unroll
    for (int n = 0; n < 3; ++n) {
        double v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i] = NB ? ex::div3_nb(rt[n][i] * ar * dtl, bad) : ex::div3(rt[n][i] * ar * dtl);
            C[(4 * n + i) * TE + k] = v[i];
        }
        if (ln[n] >= nint) st4(A.EC + 12 * e_glob + 4 * n, v);   // tile-boundary node: finished by node_update afterwards
    }
    return bad;
}
template <bool VISC>
__device__ __noinline__ void fused_elem_plain(const TileGeom& G, const StageArgs& A, const unsigned char* sb, double* C, int k,
                                              int nint, long e_glob, double dtl_uniform) {
    fused_elem<VISC, false>(G, A, sb, C, k, nint, e_glob, dtl_uniform);
}

// NCW compute warps (a multiple of 4) + one auxiliary warpgroup whose warp 0 is the loader; TE = 32*NCW elements per tile
template <bool VISC, int NCW, int RC, int RA>
__global__ void __launch_bounds__((NCW + 4) * 32, 1) stage_fused(const __grid_constant__ TileGeom G, const __grid_constant__ StageArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NSTAGE = 2;
    constexpr int NCT = NCW * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // barriers: full[s] (33 arrivals: the loader's expect_tx + one cp.async arrival per loader lane, plus the bulk bytes),
    // fstat[s] (static block landed; 1 arrival + bytes), empty[s] (NCW arrivals: one per compute warp)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);
    const unsigned full0 = ptx::smem_u32(bars), fstat0 = full0 + 8 * NSTAGE, empty0 = fstat0 + 8 * NSTAGE;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            ptx::mbar_init(full0 + 8 * s, 33);
            ptx::mbar_init(fstat0 + 8 * s, 1);
            ptx::mbar_init(empty0 + 8 * s, NCW);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();
    double* C = reinterpret_cast<double*>(smem + G.off_c);
    const int TE = G.TE;
    if (warp >= NCW) {
        // ---------------- auxiliary warpgroup: give registers away; warp NCW is the loader -----------------------
        ptx::reg_dec<RA>();
        if (warp != NCW) return;
        const unsigned stream_bytes = (unsigned)(G.nfields * TE * 8);
        int it = 0;
        for (int t = blockIdx.x; t < A.ntiles; t += gridDim.x, ++it) {
            const int s = it & 1;
            const unsigned ph = (it >> 1) & 1;
            unsigned char* sb = smem + G.off_stage0 + (size_t)s * G.stage_bytes;
            const unsigned sbu = ptx::smem_u32(sb);
            ptx::mbar_wait(empty0 + 8 * s, ph ^ 1);
            if (lane == 0) {
                ptx::mbar_arrive_expect_tx(fstat0 + 8 * s, (unsigned)G.tb_bytes);
                ptx::bulk_g2s(sbu + G.st_static, A.TB + (size_t)t * G.tb_bytes, (unsigned)G.tb_bytes, fstat0 + 8 * s);
                ptx::mbar_arrive_expect_tx(full0 + 8 * s, stream_bytes);
                const size_t e0 = (size_t)t * TE;
                const unsigned fb = (unsigned)(TE * 8), dst = sbu + G.st_stream;
#pragma unroll 1
                for (int f = 0; f < 7; ++f) ptx::bulk_g2s(dst + f * fb, A.geo + (size_t)f * A.Epad + e0, fb, full0 + 8 * s);
                ptx::bulk_g2s(dst + 7 * fb, A.shoc + e0, fb, full0 + 8 * s);
                ptx::bulk_g2s(dst + 8 * fb, A.ts1 + e0, fb, full0 + 8 * s);
                ptx::bulk_g2s(dst + 9 * fb, A.ts2 + e0, fb, full0 + 8 * s);
                ptx::bulk_g2s(dst + 10 * fb, A.ts3 + e0, fb, full0 + 8 * s);
                if (G.nfields == 12) ptx::bulk_g2s(dst + 11 * fb, A.dtl_arr + e0, fb, full0 + 8 * s);
            }
            ptx::mbar_wait(fstat0 + 8 * s, ph);
            const int* hdr = reinterpret_cast<const int*>(sb + G.st_static);
            const int ntn = hdr[1], nint = hdr[2];
            const int* tnode = reinterpret_cast<const int*>(sb + G.st_static + G.off_tnode);
            for (int j = lane; j < ntn; j += 32) {
                const int n = tnode[j];
                const double* u = A.Usrc + 4 * (size_t)n;
                ptx::cp_async16(sbu + G.st_u + 32 * j, u);
                ptx::cp_async16(sbu + G.st_u + 32 * j + 16, u + 2);
                if (VISC) ptx::cp_async8(sbu + G.st_t + 8 * j, A.T + n);
            }
            for (int j = lane; j < nint; j += 32) {
                const int n = tnode[j];
                ptx::cp_async8(sbu + G.st_m + 8 * j, A.M + n);
                ptx::cp_async8(sbu + G.st_g + 8 * j, A.GAMM + n);
            }
            ptx::cp_async_arrive_noinc(full0 + 8 * s);
        }
        return;
    }
    // ---------------- compute warpgroups --------------------------------------------------------------------------
    ptx::reg_inc<RC>();
    const double dtl_uniform = G.nfields == 12 ? 0.0 : *A.dtl_sc;
    const int k = threadIdx.x;   // element position in the tile, then interior-node index
    int it = 0;
    for (int t = blockIdx.x; t < A.ntiles; t += gridDim.x, ++it) {
        const int s = it & 1;
        const unsigned ph = (it >> 1) & 1;
        const unsigned char* sb = smem + G.off_stage0 + (size_t)s * G.stage_bytes;
        ptx::mbar_wait(fstat0 + 8 * s, ph);
        ptx::mbar_wait(full0 + 8 * s, ph);
        const int* hdr = reinterpret_cast<const int*>(sb + G.st_static);
        const int ne = hdr[0], nint = hdr[2];
        if (k < ne) {
            const long e_glob = (long)t * TE + k;
            if (VISC) {
                if (fused_elem<VISC, true>(G, A, sb, C, k, nint, e_glob, dtl_uniform))
                    fused_elem_plain<VISC>(G, A, sb, C, k, nint, e_glob, dtl_uniform);
            } else {
                fused_elem<VISC, false>(G, A, sb, C, k, nint, e_glob, dtl_uniform);
            }
        }
        ptx::named_bar_sync(1, NCT);   // C complete
        for (int j = k; j < nint; j += NCT) {
            const int* tnode = reinterpret_cast<const int*>(sb + G.st_static + G.off_tnode);
            const unsigned short* nptr = reinterpret_cast<const unsigned short*>(sb + G.st_static + G.off_nptr);
            const unsigned short* slots = reinterpret_cast<const unsigned short*>(sb + G.st_static + G.off_slots);
            const int n = tnode[j];
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const int q1 = nptr[j + 1];
            for (int q = nptr[j]; q < q1; ++q) {
                const int sl = slots[q];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = acc[i] + C[sl + i * TE];
            }
            st4(A.RHS + 4 * (size_t)n, acc);
            double u[4];
            if (A.U == A.Usrc) {   // the tile's copy of the state is the state the update starts from
                const double2* q = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(sb + G.st_u) + 4 * j);
                double2 a = q[0], b = q[1];
                u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y;
            } else {
                ld4(A.U + 4 * (size_t)n, u);
            }
            const double m = reinterpret_cast<const double*>(sb + G.st_m)[j];
            const double gam = reinterpret_cast<const double*>(sb + G.st_g)[j];
            const unsigned fl = (sb + G.st_static + G.off_bcf)[j];
            node_finish_v(n, acc, u, m, gam, fl, A.WXa, A.WYa, A.bc, A.rk_fact, A.FR, A.U1, A.RHO, A.VELX, A.VELY, A.Ea, A.Pa,
                          A.Ta, A.RMACH);
        }
        ptx::named_bar_sync(1, NCT);   // C and the stage may be overwritten
        if (lane == 0) ptx::mbar_arrive(empty0 + 8 * s);
    }
}

// geo[7][Epad] = dNx(3), dNy(3), area: the element stream of stage_fused with a component stride that keeps every tile's
// run 16-byte aligned (cp.async.bulk), whatever nelem is
__global__ void __launch_bounds__(256) pack_geo(long nelem, long Epad, const double* __restrict__ dNx, const double* __restrict__ dNy,
                                                 const double* __restrict__ area, double* __restrict__ geo) {
    long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= nelem) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        geo[c * Epad + e] = dNx[c * nelem + e];
        geo[(3 + c) * Epad + e] = dNy[c * nelem + e];
    }
    geo[6 * Epad + e] = area[e];
}

// element-array permutation between the file's numbering and the internal (tile) order, w doubles/ints per element:
// dst[i][q] = src[map[i]][q]
template <class T>
__global__ void __launch_bounds__(256) perm_rows(long n, int w, const int* __restrict__ map, const T* __restrict__ src, T* __restrict__ dst) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n * w) return;
    long r = i / w;
    int q = (int)(i - r * w);
    dst[i] = src[(size_t)map[r] * w + q];
}
// (3,E) interleaved host layout in file order <-> [3][E] internal order
__global__ void aos3_to_soa_perm(long nelem, const int* __restrict__ i2e, const double* __restrict__ in, double* __restrict__ out) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= 3 * nelem) return;
    long p = i / 3, c = i - 3 * p;
    out[c * nelem + p] = in[3 * (size_t)i2e[p] + c];
}
__global__ void soa_to_aos3_perm(long nelem, const int* __restrict__ i2e, const double* __restrict__ in, double* __restrict__ out) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= 3 * nelem) return;
    long p = i / 3, c = i - 3 * p;
    out[3 * (size_t)i2e[p] + c] = in[c * nelem + p];
}
