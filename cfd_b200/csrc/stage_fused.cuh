// Fused RK stage for fixed meshes (included inside namespace CFDB_KNS by kernels.cuh).
//
// One RK stage of the reference is  RHS = 0 ; calcRHS ; U1 = U - rk/M*RHS ; primitives ; fixvel/normalvel/FIX ; conservative
// (subrutinas.f90:667-826, calcRHS.f90:36-151).  The two-kernel stage (calcrhs_elem + node_update) pays for the reference's
// summation order with a 96 B/element staging buffer written and read back through HBM.  This kernel keeps that buffer in
// shared memory for ~82 % of the nodes:
//
//   * the elements are stored in tile order (host_topology.h: build_tiling): tile t = TE consecutive internal elements, a
//     compact patch of the mesh (recursive coordinate bisection); a node all of whose elements lie in one tile is INTERIOR
//     to it;
//   * a persistent CTA (one per SM) walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...;
//   * warp specialisation (see stage_fused below): ONE loader warp brings everything a tile needs into shared-memory rings
//     -- the tile's static block and its element stream (geometry, stabilisation parameters) by cp.async.bulk (TMA bulk
//     copies completing on an mbarrier's transaction count), the nodal state of the tile's nodes by cp.async gathers (one
//     32-byte sector per node) tracked by an mbarrier -- twelve ELEMENT warps do the element arithmetic (thread = element;
//     calcrhs_body, unchanged arithmetic) into the shared-memory array C[12][TE] (double-buffered), three NODE warps sum
//     each interior node's contributions from C in ascending ORIGINAL element order (the `slots` list of the static
//     block) and run the nodal chain -- same operations, same order, same bits as node_update;
//   * contributions to tile-boundary nodes (~18 % of the nodes) go to the boundary staging buffer ECB, laid out so that
//     the records of one node are contiguous and in summation order; boundary_update finishes those nodes right after
//     this kernel, reading each node's records as one run.
//
// Roofline: HBM traffic per element-stage falls from 404 B (220 algorithmic + 192 staging, measured 6.46 GB per stage on the
// 16 M-triangle mesh) to ~290 B; the kernel is then bound by the fp64 pipe and the issue port together (819 fp64 + 762 other
// instructions per element warp and tile: 2 x 819 pipe cycles against 1581 issue slots, profiles/r2_stage_mix.txt).
#pragma once
// build-time switches of the stage kernel (A/B builds: make ab ABFLAGS=-D...)
#ifndef CFDB_FB_COUNT
#define CFDB_FB_COUNT 1        // count the plain-form recomputations of the stage kernel (kernels.cuh: g_fallbacks)
#endif
#ifndef CFDB_NODE_NB
#define CFDB_NODE_NB 1        // node warps: branch-free divisions / square root in the nodal chain
#endif
#ifndef CFDB_STAGE_RE
#define CFDB_STAGE_RE 144     // registers per thread of the element warps; the auxiliary warps get 512 - 3 x this
#endif

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
__device__ __forceinline__ unsigned mbar_try_wait_hint(unsigned bar, unsigned parity, unsigned ns) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return ok;
}
// Bounded wait: a protocol bug must end in a trap (sticky error the host sees), never in a hung GPU.  A waiting warp shares
// its scheduler with three element warps: the retry loop is four instructions (the clock is read every 1024th retry only)
// and every attempt lets the hardware park the warp (suspend-time hint).
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = 0;
    for (unsigned spins = 1;; ++spins) {
        if (mbar_try_wait_hint(bar, parity, 2000u)) return;
        if ((spins & 1023u) == 0) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s at 2 GHz
        }
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

// offsets of the static tile block (mirror of topo::TileLayout) and of the two shared-memory rings
struct TileGeom {
    int TE, ntn_max, nint_max, nslot_max;
    int off_lnode, off_tnode, off_nptr, off_slots, off_bcf, off_brank, off_bbase, tb_bytes;
    // A ring slot (static block + gathered nodal data), bytes from the slot base
    int a_static, a_u, a_t, a_m, a_g, a_bytes;
    int b_bytes;       // B ring slot: the element stream, nfields x TE doubles
    int nfields;       // doubles per element in the stream: 11, or 12 with a local time step array
    int off_c;         // C[2][12][TE] from the shared-memory base (after the barriers)
    int off_a, off_b;  // ring bases
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
    CF T, GAMM;                                    // columns of the nodal record array NR1 (one 32-byte sector per node)
    const double *M, *WXa, *WYa;
    BcTab bc;
    double rk_fact, FR;
    Gas g;
    double* ECB;                                   // boundary staging buffer: 32-byte records, the records of one node side by side
    double *U1, *RHS;
    WF RHO, VELX, VELY, Ea, Pa, Ta, RMACH;
    unsigned long long* stats;                     // optional (CFDB_STAGE_STATS): cycle counters, see stage_fused
    int stat_warp;                                 // the element warp that reports (CFDB_STAGE_STATS=k: warp k-1)
};
enum { ST_E = 0, ST_N, ST_WAIT_IN, ST_WAIT_CE, ST_WAIT_CF, ST_LD_WB, ST_LD_WS, ST_LD_WA, ST_TILES, ST_LOOP_CYC, ST_LOOP_NS, ST_ARRIVE, ST_COUNT };

// One element: inputs from shared memory -- `sa` is the tile's A slot (connectivity + nodal state), `sbm` its B slot (element
// stream) -- the twelve contributions to C (and to the global staging buffer EC for tile-boundary nodes, finished by
// node_update afterwards).  Stores happen as the values become ready; when the fast-path flag comes back raised the caller
// runs the plain form, whose stores (same thread, same addresses, program order) replace these.
template <bool VISC, bool NB, int TE>
__device__ __forceinline__ unsigned fused_elem(const TileGeom& G, const StageArgs& A, const unsigned char* sa, const unsigned char* sbm,
                                               double* __restrict__ C, int k, int nint, double dtl_uniform) {
    // TE is a compile-time constant (32 x element warps): every stream operand is then base + immediate
    const unsigned short* lnode = reinterpret_cast<const unsigned short*>(sa + G.a_static + G.off_lnode);
    const int ln[3] = {lnode[k], lnode[TE + k], lnode[2 * TE + k]};
    const double* ut = reinterpret_cast<const double*>(sa + G.a_u);
    double Un[3][4], Th[3][4], Tn[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double2* q = reinterpret_cast<const double2*>(ut + 4 * ln[j]);
        double2 a = q[0], b = q[1];
        Un[j][0] = a.x; Un[j][1] = a.y; Un[j][2] = b.x; Un[j][3] = b.y;
    }
    if (VISC) {
        const double* tt = reinterpret_cast<const double*>(sa + G.a_t);
        Tn[0] = tt[ln[0]]; Tn[1] = tt[ln[1]]; Tn[2] = tt[ln[2]];
    }
    const double* sd = reinterpret_cast<const double*>(sbm);
    double Nx[3] = {sd[k], sd[TE + k], sd[2 * TE + k]};
    double Ny[3] = {sd[3 * TE + k], sd[4 * TE + k], sd[5 * TE + k]};
    const double ar = sd[6 * TE + k];
    const double shoc_e = sd[7 * TE + k];
    const double tau[3] = {sd[8 * TE + k], sd[9 * TE + k], sd[10 * TE + k]};
    const double dtl = G.nfields == 12 ? sd[11 * TE + k] : dtl_uniform;
    const unsigned bmask = (ln[0] >= nint ? 1u : 0u) | (ln[1] >= nint ? 2u : 0u) | (ln[2] >= nint ? 4u : 0u);
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
        if (bmask & (1u << n)) {   // tile-boundary node: record bbase[node] + (rank of this element in the node's list)
            const unsigned rec = reinterpret_cast<const unsigned*>(sa + G.a_static + G.off_bbase)[ln[n] - nint] +
                                 (sa + G.a_static + G.off_brank)[n * TE + k];
            st4(A.ECB + 4 * (size_t)rec, v);
        }
    }
    return bad;
}
template <bool VISC, int TE>
__device__ __noinline__ void fused_elem_plain(const TileGeom& G, const StageArgs& A, const unsigned char* sa, const unsigned char* sbm,
                                              double* __restrict__ C, int k, int nint, double dtl_uniform) {
    fused_elem<VISC, false, TE>(G, A, sa, sbm, C, k, nint, dtl_uniform);
}

// Warp roles.  NCW element warps (a multiple of 4) + one auxiliary warpgroup: warp NCW is the loader, warps NCW+1..NCW+3 are
// node warps; TE = 32*NCW elements per tile; one CTA per SM.
//
// Registers.  The register file is split over the four SM sub-partitions (16384 registers each, warps dealt round-robin), so
// with 16 warps per CTA every sub-partition holds three element warps and one auxiliary warp, launched at 128 registers per
// thread.  setmaxnreg then moves registers from the auxiliary warpgroup (80 each: eight contributions of a node in flight) to the element
// warpgroups (144 each: the branch-free form of the element arithmetic runs without spills and its single-warp schedule is
// within 40 % of the fp64 issue time, tools/sass_stalls.py).  3 x 144 + 80 = 4 x 128: what the element warps claim is exactly
// what the auxiliary warpgroup released (the registers a warpgroup may claim come from its own CTA's pool).
//
// Pipeline (all hand-overs are mbarriers; no CTA-wide barrier in the steady state):
//   loader       : stream(i) -> B ring [NBR slots] (completing on tile i's inputs-landed barrier); gathers(i) (U, T, M, GAMM of the tile's nodes, addresses from the static
//                  block that landed earlier) -> A ring [NA slots]; static block of tile i+1 -> A ring, one tile ahead of its
//                  gathers so that the two dependent round trips to HBM never sit on an element warp's path;
//   element warps: E(i): 32 elements each from A(i), B(i) -> registers; wait until C[i&1] is free (node phase i-2 done: long
//                  ago); C[i&1] <- contributions (+ the ECB records of tile-boundary nodes); release B(i), their share of A(i).  They do
//                  nothing else: with three of them per sub-partition the fp64 pipe sees a pure arithmetic stream;
//   node warps   : N(i): wait C[i&1] full; interior node j sums its contributions from C in ascending ORIGINAL element order
//                  and runs the nodal chain; release C[i&1] and A(i).  Gather chains, four divisions and a square root per
//                  node, scattered stores: latency-bound work that now runs beside the element arithmetic instead of
//                  interrupting it.
template <bool VISC, int NCW, int NA, int NBR, bool STATS = false>
__global__ void __launch_bounds__((NCW + 4) * 32, 1) stage_fused(const __grid_constant__ TileGeom G, const __grid_constant__ StageArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NNW = 3;                       // node warps
    constexpr int NNT = NNW * 32;
    constexpr int RE = (NCW == 12) ? CFDB_STAGE_RE : 112;   // registers per thread of an element warp after the hand-over
    constexpr int RAUX = (NCW == 12) ? 4 * 128 - 3 * CFDB_STAGE_RE : 32;  //                  ... of an auxiliary warp      (NCW*RE + 4*RAUX == (NCW+4) * launch registers)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // barriers (8 bytes each): astat[NA] static block landed (1 arrival + bytes); afull[NA] the tile's inputs landed (32 cp.async
    // arrivals of the gathers + 1 arrival with the stream's bytes); aempty[NA] (NCW + NNW arrivals); bfull[NBR] (unused: the
    // stream completes on afull); bempty[NBR] (NCW); cfull[2] (NCW); cempty[2] (NNW)
    const unsigned bar0 = ptx::smem_u32(smem);
    const unsigned astat0 = bar0, afull0 = astat0 + 8 * NA, aempty0 = afull0 + 8 * NA, bfull0 = aempty0 + 8 * NA,
                   bempty0 = bfull0 + 8 * NBR, cfull0 = bempty0 + 8 * NBR, cempty0 = cfull0 + 16;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) {
            ptx::mbar_init(astat0 + 8 * s, 1);
            ptx::mbar_init(afull0 + 8 * s, 33);   // 32 cp.async arrivals (gathers) + the stream's expect-tx arrival
            ptx::mbar_init(aempty0 + 8 * s, NCW + NNW);
        }
        for (int s = 0; s < NBR; ++s) {
            ptx::mbar_init(bfull0 + 8 * s, 1);
            ptx::mbar_init(bempty0 + 8 * s, NCW);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(cfull0 + 8 * s, NCW);
            ptx::mbar_init(cempty0 + 8 * s, NNW);
        }
        ptx::fence_barrier_init();
    }
    constexpr int TE = 32 * NCW;   // == G.TE (checked on the host)
    // optional cycle counters (CFDB_STAGE_STATS): element warp 0, node warp 0 and the loader report; kept in shared memory so
    // that they cost no registers
    unsigned long long* st = reinterpret_cast<unsigned long long*>(smem + 176);   // barriers end at byte 160 (NA = 4), C starts at 384
    static_assert(176 + 8 * ST_COUNT <= 384, "the cycle counters must end before C");
    if (STATS && threadIdx.x < ST_COUNT) st[threadIdx.x] = 0;
    __syncthreads();
    // STATS is a template parameter: the production instantiation carries none of the counter code (a run-time flag cost 3 %)
    const bool stat = STATS && A.stats != nullptr && lane == 0 && (warp == A.stat_warp || warp == NCW || warp == NCW + 1);
    auto waitc = [&](unsigned bar, unsigned par, int slot) {   // wait, with the cycles charged to st[slot] on the reporting lanes
        if (!STATS) { ptx::mbar_wait(bar, par); return; }
        const long long t0 = clock64();                         // before the first attempt: try_wait itself may park the warp
        ptx::mbar_wait(bar, par);
        if (stat) st[slot] += (unsigned long long)(clock64() - t0);
    };
    const int my_tiles = A.ntiles > (int)blockIdx.x ? (A.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    double* const Cbase = reinterpret_cast<double*>(smem + G.off_c);
    if (warp >= NCW) {
        ptx::reg_dec<RAUX>();
        if (warp == NCW) {
            // ---------------- loader warp ------------------------------------------------------------------------
            const unsigned stream_bytes = (unsigned)G.b_bytes;
            auto issue_static = [&](int it2, int t2) {
                const int sa = it2 % NA;
                waitc(aempty0 + 8 * sa, ((it2 / NA) & 1) ^ 1, ST_LD_WA);
                if (lane == 0) {
                    ptx::mbar_arrive_expect_tx(astat0 + 8 * sa, (unsigned)G.tb_bytes);
                    ptx::bulk_g2s(ptx::smem_u32(smem + G.off_a + (size_t)sa * G.a_bytes + G.a_static), A.TB + (size_t)t2 * G.tb_bytes,
                                  (unsigned)G.tb_bytes, astat0 + 8 * sa);
                }
            };
            int it = 0, t = blockIdx.x;
            if (t < A.ntiles) issue_static(0, t);
            for (; t < A.ntiles; t += gridDim.x, ++it) {
                // element stream of tile it
                const int sb = it % NBR;
                waitc(bempty0 + 8 * sb, ((it / NBR) & 1) ^ 1, ST_LD_WB);
                if (lane == 0) {
                    ptx::mbar_arrive_expect_tx(afull0 + 8 * (it % NA), stream_bytes);   // the stream completes on the tile's inputs-landed barrier too
                    const size_t e0 = (size_t)t * TE;
                    const unsigned fb = (unsigned)(TE * 8), dst = ptx::smem_u32(smem + G.off_b + (size_t)sb * G.b_bytes), bar = afull0 + 8 * (it % NA);
#pragma unroll 1
                    for (int f = 0; f < 7; ++f) ptx::bulk_g2s(dst + f * fb, A.geo + (size_t)f * A.Epad + e0, fb, bar);
                    ptx::bulk_g2s(dst + 7 * fb, A.shoc + e0, fb, bar);
                    ptx::bulk_g2s(dst + 8 * fb, A.ts1 + e0, fb, bar);
                    ptx::bulk_g2s(dst + 9 * fb, A.ts2 + e0, fb, bar);
                    ptx::bulk_g2s(dst + 10 * fb, A.ts3 + e0, fb, bar);
                    if (G.nfields == 12) ptx::bulk_g2s(dst + 11 * fb, A.dtl_arr + e0, fb, bar);
                }
                // nodal state of tile it (its static block was requested one tile ago)
                const int sa = it % NA;
                unsigned char* ab = smem + G.off_a + (size_t)sa * G.a_bytes;
                const unsigned abu = ptx::smem_u32(ab);
                waitc(astat0 + 8 * sa, (it / NA) & 1, ST_LD_WS);
                const int* hdr = reinterpret_cast<const int*>(ab + G.a_static);
                const int ntn = hdr[1], nint = hdr[2];
                const int* tnode = reinterpret_cast<const int*>(ab + G.a_static + G.off_tnode);
#pragma unroll 1
                for (int j = lane; j < ntn; j += 32) {
                    const int n = tnode[j];
                    const double* u = A.Usrc + 4 * (size_t)n;
                    ptx::cp_async16(abu + G.a_u + 32 * j, u);
                    ptx::cp_async16(abu + G.a_u + 32 * j + 16, u + 2);
                    if (VISC) ptx::cp_async8(abu + G.a_t + 8 * j, A.T.p + NREC * (size_t)n);
                }
#pragma unroll 1
                for (int j = lane; j < nint; j += 32) {
                    const int n = tnode[j];
                    ptx::cp_async8(abu + G.a_m + 8 * j, A.M + n);
                    ptx::cp_async8(abu + G.a_g + 8 * j, A.GAMM.p + NREC * (size_t)n);
                }
                ptx::cp_async_arrive_noinc(afull0 + 8 * sa);
                // static block of the next tile
                if (t + (int)gridDim.x < A.ntiles) issue_static(it + 1, t + gridDim.x);
            }
            if (stat) {
                atomicAdd(A.stats + ST_LD_WB, st[ST_LD_WB]);
                atomicAdd(A.stats + ST_LD_WS, st[ST_LD_WS]);
                atomicAdd(A.stats + ST_LD_WA, st[ST_LD_WA]);
            }
            return;
        }
        // ---------------- node warps -----------------------------------------------------------------------------
        const int nt = (warp - NCW - 1) * 32 + lane;   // 0 .. NNT-1
        for (int it = 0; it < my_tiles; ++it) {
            const int cj = it & 1, sa = it % NA;
            const unsigned char* ab = smem + G.off_a + (size_t)sa * G.a_bytes;
            const double* C = Cbase + (size_t)cj * 12 * TE;
            waitc(cfull0 + 8 * cj, (it >> 1) & 1, ST_WAIT_CF);
            const long long t0 = stat ? clock64() : 0;
            const int* hdr = reinterpret_cast<const int*>(ab + G.a_static);
            const int nint = hdr[2];
            const int* tnode = reinterpret_cast<const int*>(ab + G.a_static + G.off_tnode);
            const unsigned short* nptr = reinterpret_cast<const unsigned short*>(ab + G.a_static + G.off_nptr);
            const unsigned short* slots = reinterpret_cast<const unsigned short*>(ab + G.a_static + G.off_slots);
#pragma unroll 1
            for (int j = nt; j < nint; j += NNT) {
                const int n = tnode[j];
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                const int q1 = nptr[j + 1];
                // eight contributions at a time: all slot indices, then all 32 values, are requested before the first add, so
                // the shared-memory latencies overlap and only the eight dependent additions per component remain in sequence
                // (the order of the additions is the list's: ascending original element id)
#pragma unroll 1
                for (int q = nptr[j]; q < q1; q += 8) {
                    int sl[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) sl[r] = slots[min(q + r, q1 - 1)];
                    double cv[8][4];
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int i = 0; i < 4; ++i) cv[r][i] = C[sl[r] + i * TE];
#pragma unroll
                    for (int r = 0; r < 8; ++r)
                        if (q + r < q1) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[i] = acc[i] + cv[r][i];
                        }
                }
                st4(A.RHS + 4 * (size_t)n, acc);
                double u[4];
                if (A.U == A.Usrc) {   // the tile's copy of the state is the state the update starts from
                    const double2* q = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(ab + G.a_u) + 4 * j);
                    double2 a = q[0], b = q[1];
                    u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y;
                } else {
                    ld4(A.U + 4 * (size_t)n, u);
                }
                const double m = reinterpret_cast<const double*>(ab + G.a_m)[j];
                const double gam = reinterpret_cast<const double*>(ab + G.a_g)[j];
                const unsigned fl = (ab + G.a_static + G.off_bcf)[j];
#if CFDB_NODE_NB
                if (node_finish_nb(n, acc, u, m, gam, fl, A.WXa, A.WYa, A.bc, A.rk_fact, A.FR, A.U1, A.RHO, A.VELX, A.VELY, A.Ea, A.Pa,
                                   A.Ta, A.RMACH)) {
                    if (CFDB_FB_COUNT) atomicAdd(&g_fallbacks[FB_STAGE_NODE], 1ull);
                    node_finish_plain(n, acc, u, m, gam, fl, A.WXa, A.WYa, A.bc, A.rk_fact, A.FR, A.U1, A.RHO, A.VELX, A.VELY, A.Ea,
                                      A.Pa, A.Ta, A.RMACH);
                }
#else
                node_finish_v(n, acc, u, m, gam, fl, A.WXa, A.WYa, A.bc, A.rk_fact, A.FR, A.U1, A.RHO, A.VELX, A.VELY, A.Ea, A.Pa,
                              A.Ta, A.RMACH);
#endif
            }
            __syncwarp();
            if (lane < 2) ptx::mbar_arrive(lane == 0 ? cempty0 + 8 * cj : aempty0 + 8 * sa);
            if (stat) st[ST_N] += (unsigned long long)(clock64() - t0);
        }
        if (stat) {
            atomicAdd(A.stats + ST_N, st[ST_N]);
            atomicAdd(A.stats + ST_WAIT_CF, st[ST_WAIT_CF]);
        }
        return;
    }
    // ---------------- element warps -------------------------------------------------------------------------------
    ptx::reg_inc<RE>();
    const double dtl_uniform = G.nfields == 12 ? 0.0 : *A.dtl_sc;
    const int k = threadIdx.x;   // element position in the tile
    int t = blockIdx.x;
    if (stat) {   // start stamps live in the shared-memory counters, not in registers the element arithmetic needs
        unsigned long long n0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
        st[ST_LOOP_CYC] = (unsigned long long)clock64();
        st[ST_LOOP_NS] = n0;
    }
    for (int it = 0; it < my_tiles; ++it, t += gridDim.x) {
        const int sa = it % NA, sb = it % NBR, c = it & 1;
        const unsigned char* ab = smem + G.off_a + (size_t)sa * G.a_bytes;
        const unsigned char* bb = smem + G.off_b + (size_t)sb * G.b_bytes;
        waitc(afull0 + 8 * sa, (it / NA) & 1, ST_WAIT_IN);
        waitc(cempty0 + 8 * c, ((it >> 1) & 1) ^ 1, ST_WAIT_CE);   // node phase it-2 is done: C[c] is free
        const long long t0 = stat ? clock64() : 0;
        const int* hdr = reinterpret_cast<const int*>(ab + G.a_static);
        const int ne = hdr[0], nint = hdr[2];
        if (k < ne) {
            double* C = Cbase + (size_t)c * 12 * TE;
            // branch-free divisions (exact.cuh) first; the plain form only if an operand left their range
            if (fused_elem<VISC, true, TE>(G, A, ab, bb, C, k, nint, dtl_uniform)) {
                if (CFDB_FB_COUNT) atomicAdd(&g_fallbacks[FB_STAGE_ELEM], 1ull);
                fused_elem_plain<VISC, TE>(G, A, ab, bb, C, k, nint, dtl_uniform);
            }
        }
        const long long t1 = stat ? clock64() : 0;
        if (stat) st[ST_E] += (unsigned long long)(t1 - t0);
        __syncwarp();
        if (lane < 3)   // three arrivals as ONE instruction: lanes 0..2 each on their own barrier
            ptx::mbar_arrive(lane == 0 ? cfull0 + 8 * c : lane == 1 ? bempty0 + 8 * sb : aempty0 + 8 * sa);
        if (stat) st[ST_ARRIVE] += (unsigned long long)(clock64() - t1);
    }
    if (stat) {
        unsigned long long n1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
        atomicAdd(A.stats + ST_LOOP_CYC, (unsigned long long)clock64() - st[ST_LOOP_CYC]);   // the reporting warp's whole tile loop: cycles and ns,
        atomicAdd(A.stats + ST_LOOP_NS, n1 - st[ST_LOOP_NS]);                                 // i.e. the SM clock the kernel actually ran at
        atomicAdd(A.stats + ST_E, st[ST_E]);
        atomicAdd(A.stats + ST_ARRIVE, st[ST_ARRIVE]);
        atomicAdd(A.stats + ST_WAIT_IN, st[ST_WAIT_IN]);
        atomicAdd(A.stats + ST_WAIT_CE, st[ST_WAIT_CE]);
        atomicAdd(A.stats + ST_TILES, (unsigned long long)my_tiles);
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
