// libcfdb200.so — context, time loop and C ABI (include/cfdb.h) on top of kernels.cuh.
// Compiled for sm_100a only, with -fmad=false (see exact.cuh).  No CPU fallback anywhere.
#include "../../include/cfdb.h"
#include "host_topology.h"
#include "../../host/mesh_smoothing.h"
#include "../../host/fortran_format.h"
#include "kernels.cuh"

#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using std::vector;

namespace fastmode {  // fast.cu: the relaxed stage, compiled with FMA contraction
int launch_calcrhs_scatter(bool visc, bool ale, cudaStream_t st, int nelem, const int* inp, const double* U, const double* T,
                           const double* WX, const double* WY, const double* dNx, const double* dNy, const double* area,
                           const double* shoc, const double* dtl_arr, const double* dtl_sc, const double* ts1, const double* ts2,
                           const double* ts3, double Cv, double lambda_ref, double mu_ref, double gamma0, double T_inf, double cte,
                           double* RHS);
int launch_node_update_rhs(cudaStream_t st, int npoin, const double* RHS, const double* U, const double* M, const double* GAMM,
                           const double* WX, const double* WY, const unsigned char* bcflag, int nb, const int* bnode,
                           const int* bkind, const double* bvx, const double* bvy, const double* brho, const double* bT,
                           const int* bwslot, const double* wnx, const double* wny, const int* wnvalid, double rk_fact, double FR,
                           double* U1, double* RHO, double* VX, double* VY, double* E, double* P, double* T, double* RMACH);
int launch_calcrhs_staged_fma(bool visc, cudaStream_t st, int nelem, const int* inp, const double* U, const double* T,
                              const double* dNx, const double* dNy, const double* area, const double* shoc, const double* dtl_arr,
                              const double* dtl_sc, const double* ts1, const double* ts2, const double* ts3, double Cv,
                              double lambda_ref, double mu_ref, double gamma0, double T_inf, double cte, double* EC);
}  // namespace fastmode

namespace topogpu {
int build(cudaStream_t st, const int* inp, int nelem, int npoin, int* esup2, int* eslot, int* psup2, int** psup1_out,
          int* npsup_out, int* rowptr, int** idx0_out, unsigned char* lpos, int* maxrow_out);
}

namespace {
// eslot entries 3*e+local: file-order element ids -> internal positions
__global__ void remap_slots(long n, const int* __restrict__ e2i, int* __restrict__ eslot) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = eslot[i], e = s / 3;
    eslot[i] = 3 * e2i[e] + (s - 3 * e);
}
__global__ void remap_ids(int n, const int* __restrict__ e2i, int* __restrict__ ids) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && ids[i] >= 0) ids[i] = e2i[ids[i]];
}
}  // namespace

static thread_local std::string g_err;
static int fail(const std::string& m) {
    g_err = m;
    return 1;
}
#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return fail(std::string(#call) + ": " + cudaGetErrorString(_e) + " @" + std::to_string(__LINE__)); \
    } while (0)
#define NK(call)                                                                                        \
    do {                                                                                                \
        ncclResult_t _e = (call);                                                                       \
        if (_e != ncclSuccess)                                                                          \
            return fail(std::string(#call) + ": " + ncclGetErrorString(_e) + " @" + std::to_string(__LINE__)); \
    } while (0)
#define TRY(call)                \
    do {                         \
        int _r = (call);         \
        if (_r) return _r;       \
    } while (0)

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        if (count <= n && p) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        if (count == 0) return 0;
        CK(cudaMalloc(&p, count * sizeof(T)));
        n = count;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

enum KernelId {
    K_DERIV, K_MASAS, K_NORMALES, K_DELTAT, K_DTLOGIC, K_DTL, K_ESTAB, K_CALCRHS, K_NODE, K_DOT, K_NORMS, K_SPMV,
    K_VEC, K_FIXROWS, K_SCALAR, K_LAPLACE, K_TRANSF, K_MOVE, K_FORCES, K_GCL, K_LAYOUT, K_FILL, K_HALO, K_STAGE, K_COUNT
};
static const char* kKernelNames[K_COUNT] = {"deriv", "masas", "normales", "deltat", "dt_logic", "dtl", "estab",
                                            "calcrhs_elem", "node_update", "dot", "norms", "spmv", "vec", "fixrows",
                                            "scalar", "laplace", "transf", "move_apply", "forces", "gcl", "layout", "fill", "halo", "stage_fused"};

// one column of a nodal record array (kernels.cuh: CF / WF): col = &record[0][column], entries k::NREC doubles apart
struct NodeField { k::Col col; };

struct cfdb_ctx {
    int device = 0;
    cudaStream_t st = nullptr;
    const double* Usrc = nullptr;  // state calcRHS/FUENTE are evaluated at (U, or U1 for true_rk stages 2..4)
    cfdb_params par{};
    int npoin = 0, nelem = 0;
    // host copies of integer artefacts (API layout: 1-based)
    vector<int32_t> h_inpoel, esup1, esup2, psup1, psup2, lap_idx, lap_rowptr, h_eslot;   // host mirrors: see ensure_host_topology
    bool host_topo_valid = true;
    int npsup = 0;
    vector<int32_t> h_wall, h_wn_node, h_ilaux, h_fixidx_last;
    vector<unsigned char> h_bcflag;
    int nnz = 0, maxrow = 0, nwn = 0, nb = 0, nmove = 0, nnmove = 0, nset = 0, nse = 0;
    bool ale = false;  // mesh can move (body sets present or W set by the caller; multi-rank: on ANY rank, see agree_on_ale)
    bool ale_agreed = false;
    // CUDA graphs of the fixed-mesh step, one per parity of the U/U1 buffer swap (step_graph)
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    uint64_t epoch = 0, gepoch[2] = {0, 0};
    const double* gU[2] = {nullptr, nullptr};
    int64_t glaunches[2] = {0, 0}, graph_replays = 0;
    // device: mesh
    DBuf<int> inp, d_esup2, eslot, d_lap_idx, d_lap_rowptr, wall, wn_node, wn_ptr, wn_edge, wn_valid;
    DBuf<unsigned char> lpos, bcflag;
    DBuf<double> X, Y, X1, Y1, area, HH, HHX, HHY, dNx, dNy, M;
    // device: state
    DBuf<double> U, U1, RHS, RHS1, RHS2, RHS3, UN, W_X, W_Y;
    // nodal records (kernels.cuh): NR1[n] = {T, GAMM, VEL_X, VEL_Y}, NR2[n] = {RHO, E, P, RMACH}; the named fields are columns
    DBuf<double> NR1, NR2, ntmp;
    NodeField T, GAMM, VEL_X, VEL_Y, RHO, E, P, RMACH;
    DBuf<double> SHOC, TS1, TS2, TS3, DT, DTL, EC, FC;
    // device: BC tables
    DBuf<int> bc_node, bc_kind, bc_wslot;
    DBuf<double> bc_vx, bc_vy, bc_rho, bc_T, wn_x, wn_y;
    // device: laplace / bicg / mesh motion
    DBuf<unsigned char> isfix;
    DBuf<double> bp2;
    DBuf<double> lap_sparse, lap_diag, by, bp, br, bz, bb, xpos, ypos, dxpos, dypos, pos_aux, xref, yref;
    DBuf<double> by2, pos_aux2;  // second solve's A*x and Dirichlet values: both first SpMVs of fluidStructure run as one pass
    DBuf<int> ilaux, ilaux_last, se_node, se_set, set_ptr, set_n1, set_n2, set_el, d_psup1, d_psup2;
    DBuf<double> fvisc, skin;   // FORCE_VISC: F_VX(10) F_VY(10); SKIN.DAT columns [3][nedges]
    int nedges = 0;
    // gcl
    DBuf<double> W_x_old, W_y_old, area_old;
    // reductions / scalars
    DBuf<double> redA, redB, tmpA, tmpB, tmpC;
    DBuf<int> flags;
    k::Scal* sc = nullptr;
    k::Scal* h_sc = nullptr;  // pinned
    // host mirror of loop scalars
    int h_iter = 0;
    int iterprint = 0;
    double DISN[2] = {0, 0};
    int bicg_iters[2] = {0, 0};
    bool theta_nonzero = false;
    int use_cuarto = 0, true_rk = 0;  // "next" rows N1 / N2, default off (reference behaviour)
    int adamsb = 0, nestab = 1;       // "next" row N2: ADAMSB replaces RK when BANDERA > 4 (ns2DComp.ALE.f90:134,174-178)
    int colored = 0;                  // relaxed stage with the deterministic coloured scatter (opt-in, NOT bit-exact vs the reference)
    vector<int> color_ptr;            // per colour: range of color_list
    DBuf<int> color_list;             // internal element positions, colour by colour, ascending original id inside a colour
    bool dtl_force = false;           // call-site RK: the local time step array is the caller's dtl whatever ITLOCAL says
    int fast = 0;                     // relaxed stage (FMA + atomic scatter), opt-in, NOT bit-exact
    bool u1_is_u = false;  // after U = U1 (ns2DComp.ALE.f90:277-281) the two arrays are one buffer
    // fused RK stage (kernels: stage_fused.cuh; tiling: host_topology.h build_tiling).  The element arrays of the context
    // are kept in INTERNAL (tile) order: i2e[p] = file-order element at internal position p, e2i its inverse (null: identity)
    bool tiles_ok = false, perm_on = false, geo_dirty = true;
    int ntiles = 0, nbnodes = 0, tile_ncw = 0, tile_na = 3;
    unsigned long long* stage_stats = nullptr;   // CFDB_STAGE_STATS=1: cycle counters of stage_fused, printed by cfdb_sync
    double tile_interior = 0.0;
    long Epad = 0;
    k::TileGeom tgeom{};
    size_t stage_smem = 0;
    vector<int32_t> h_i2e, h_e2i;
    DBuf<int> i2e, e2i, bnodes, bn_ptr, orphans;
    int norphans = 0;
    DBuf<unsigned char> TB;
    DBuf<double> geo;
    // cfdb_step_streamed: copy streams, device-side staging on both sides, and the events that order them
    cudaStream_t st_in = nullptr, st_out = nullptr;
    // two sets of device staging (call k uses set k & 1): neither the upload of call k+1 nor the staging of step k's results
    // ever waits for the transfer of the neighbouring call
    cudaEvent_t ev_in_done = nullptr, ev_step_done = nullptr, ev_in_used[2] = {nullptr, nullptr}, ev_out_done[2] = {nullptr, nullptr};
    long streamed_calls = 0;
    bool skip_stage_halo = false;   // set by run_rk around stages 1-3 when their ghost refresh is not needed (see run_rk)
    DBuf<double> sin[2], sout[2];   // [U(4P) | T(P) | VEL_X(P) | VEL_Y(P)] (+ 8 norms on the way out)
    // multi-GPU (one rank per context)
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int n_owned = 0;           // reductions run over local nodes [0,n_owned)
    // chunk-aligned ownership (cfdb_set_reduction_layout): the owned nodes are the global range [red_gid0, red_gid0+n_owned),
    // red_gid0 % 4096 == 0, so this rank's first-level chunk sums ARE chunk sums of the global canonical order
    bool red_aligned = false;
    long red_chunk0 = 0, red_nchunk_global = 0;
    DBuf<double> redG;
    vector<int> nb_rank, send_ptr, recv_ptr;
    DBuf<int> send_idx, recv_idx;
    DBuf<double> sendbuf, recvbuf;
    // profiling
    bool prof = false;
    int64_t launches = 0;
    struct Pending { int id; cudaEvent_t a, b; };
    vector<Pending> pending;
    vector<cudaEvent_t> evpool;
    double prof_ms[K_COUNT] = {0};
    int64_t prof_n[K_COUNT] = {0};
};

static inline int grid_for(long n, int block) { return (int)((n + block - 1) / block); }

static int prof_begin(cfdb_ctx* c, cudaStream_t st, int id, cudaEvent_t* a, cudaEvent_t* b) {
    c->launches++;
    if (!c->prof) return 0;
    for (int i = 0; i < 2; ++i) {
        cudaEvent_t e;
        if (c->evpool.empty()) {
            CK(cudaEventCreate(&e));
        } else {
            e = c->evpool.back();
            c->evpool.pop_back();
        }
        (i ? *b : *a) = e;
    }
    CK(cudaEventRecord(*a, st));
    (void)id;
    return 0;
}
static int prof_end(cfdb_ctx* c, cudaStream_t st, int id, cudaEvent_t a, cudaEvent_t b) {
    if (!c->prof) return 0;
    CK(cudaEventRecord(b, st));
    c->pending.push_back({id, a, b});
    return 0;
}
static int prof_resolve(cfdb_ctx* c) {
    if (c->pending.empty()) return 0;
    CK(cudaStreamSynchronize(c->st));
    for (auto& p : c->pending) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, p.a, p.b));
        c->prof_ms[p.id] += ms;
        c->prof_n[p.id] += 1;
        c->evpool.push_back(p.a);
        c->evpool.push_back(p.b);
    }
    c->pending.clear();
    return 0;
}
#define LAUNCH_ON(stream, id, kernel, grid, block, ...)                       \
    do {                                                                      \
        cudaEvent_t _a = nullptr, _b = nullptr;                               \
        TRY(prof_begin(c, stream, id, &_a, &_b));                             \
        kernel<<<(grid), (block), 0, stream>>>(__VA_ARGS__);                  \
        CK(cudaGetLastError());                                               \
        TRY(prof_end(c, stream, id, _a, _b));                                 \
    } while (0)
#define LAUNCH(id, kernel, grid, block, ...) LAUNCH_ON(c->st, id, kernel, grid, block, __VA_ARGS__)

static int up_plain(cfdb_ctx* c, double* dev, const double* h, size_t n) {
    if (n) CK(cudaMemcpyAsync(dev, h, n * sizeof(double), cudaMemcpyHostToDevice, c->st));
    return 0;
}
static int down_plain(cfdb_ctx* c, const double* dev, double* h, size_t n) {
    if (n) CK(cudaMemcpyAsync(h, dev, n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    return 0;
}
// Columns of the nodal records cross the ABI as plain arrays: through the device scratch ntmp (stream-ordered, so
// back-to-back calls may share it).
static int col_from_plain(cfdb_ctx* c, NodeField f, const double* dev_plain, size_t n) {
    if (n) LAUNCH(K_FILL, k::col_scatter, grid_for((long)n, 256), 256, (long)n, dev_plain, f.col);
    return 0;
}
static int col_to_plain(cfdb_ctx* c, NodeField f, double* dev_plain, size_t n) {
    if (n) LAUNCH(K_FILL, k::col_gather, grid_for((long)n, 256), 256, (long)n, f.col, dev_plain);
    return 0;
}
static int up_node(cfdb_ctx* c, NodeField f, const double* h, size_t n) {
    TRY(up_plain(c, c->ntmp.p, h, n));
    return col_from_plain(c, f, c->ntmp.p, n);
}
static int down_node(cfdb_ctx* c, NodeField f, double* h, size_t n) {
    TRY(col_to_plain(c, f, c->ntmp.p, n));
    return down_plain(c, c->ntmp.p, h, n);
}

template <class T>
static int upload(cfdb_ctx* c, DBuf<T>& d, const T* h, size_t n) {
    TRY(d.alloc(n));
    if (n) CK(cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, c->st));
    return 0;
}
template <class T>
static int upload(cfdb_ctx* c, DBuf<T>& d, const vector<T>& h) {
    return upload(c, d, h.data(), h.size());
}
template <class T>
static int zero(cfdb_ctx* c, DBuf<T>& d, size_t n) {
    TRY(d.alloc(n));
    if (n) CK(cudaMemsetAsync(d.p, 0, n * sizeof(T), c->st));
    return 0;
}

// ---------------------------------------------------------------------------------------------
extern "C" const char* cfdb_last_error(void) { return g_err.c_str(); }
extern "C" int cfdb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int cfdb_get_esup(const int32_t* inpoel, int32_t nelem, int32_t npoin, int32_t* esup1, int32_t* esup2) {
    vector<int32_t> e1, e2;
    topo::build_esup(inpoel, nelem, npoin, e1, e2, nullptr);
    std::copy(e1.begin(), e1.end(), esup1);
    std::copy(e2.begin(), e2.end(), esup2);
    return 0;
}
extern "C" int cfdb_get_psup(const int32_t* inpoel, int32_t nelem, int32_t npoin, int32_t* psup1, int32_t cap,
                             int32_t* psup2, int32_t* count) {
    vector<int32_t> e1, e2, p1, p2;
    topo::build_esup(inpoel, nelem, npoin, e1, e2, nullptr);
    topo::build_psup(inpoel, npoin, e1, e2, p1, p2);
    *count = (int32_t)p1.size();
    std::copy(p2.begin(), p2.end(), psup2);
    if ((int)p1.size() > cap) return fail("cfdb_get_psup: capacity too small");
    std::copy(p1.begin(), p1.end(), psup1);
    return 0;
}

extern "C" int cfdb_smoothing(double* X, double* Y, const int32_t* inpoel, const unsigned char* fixed, int32_t npoin,
                              int32_t nelem, int32_t* sweeps) {
    if (npoin < 1 || nelem < 1) return fail("cfdb_smoothing: empty mesh");
    for (size_t k = 0; k < 3 * (size_t)nelem; ++k)
        if (inpoel[k] < 1 || inpoel[k] > npoin) return fail("cfdb_smoothing: inpoel entry out of range");
    host::MeshSmoother sm(X, Y, inpoel, npoin, nelem);
    *sweeps = sm.run(fixed);
    return 0;
}
extern "C" int cfdb_smoothing_colored(double* X, double* Y, const int32_t* inpoel, const unsigned char* fixed, int32_t npoin,
                                      int32_t nelem, int32_t* sweeps) {
    if (npoin < 1 || nelem < 1) return fail("cfdb_smoothing_colored: empty mesh");
    for (size_t k = 0; k < 3 * (size_t)nelem; ++k)
        if (inpoel[k] < 1 || inpoel[k] > npoin) return fail("cfdb_smoothing_colored: inpoel entry out of range");
    host::MeshSmoother sm(X, Y, inpoel, npoin, nelem);
    *sweeps = sm.run_colored(fixed);
    return 0;
}

static int build_bc_tables(cfdb_ctx* c, const cfdb_bc* bc) {
    const int P = c->npoin;
    static const cfdb_bc empty{};
    if (!bc) bc = &empty;
    // wall nodes: unique sorted; CSR of wall edges per node in ascending edge order
    c->h_wall.assign(bc->wall, bc->wall + 2 * (size_t)bc->nwall);
    vector<int32_t> wcount((size_t)P + 1, 0);
    for (int i = 0; i < bc->nwall; ++i) {
        int a = bc->wall[2 * i], b = bc->wall[2 * i + 1];
        if (a < 1 || a > P || b < 1 || b > P) return fail("wall edge node out of range");
        wcount[a]++;
        if (b != a) wcount[b]++;
    }
    vector<int32_t> wn_slot((size_t)P + 1, -1), wn_node, wn_ptr(1, 0);
    for (int n = 1; n <= P; ++n)
        if (wcount[n]) {
            wn_slot[n] = (int)wn_node.size();
            wn_node.push_back(n - 1);
            wn_ptr.push_back(wn_ptr.back() + wcount[n]);
        }
    vector<int32_t> wn_edge(wn_ptr.back()), cur(wn_ptr.begin(), wn_ptr.end() - 1);
    for (int i = 0; i < bc->nwall; ++i) {
        int a = bc->wall[2 * i], b = bc->wall[2 * i + 1];
        wn_edge[cur[wn_slot[a]]++] = i;
        if (b != a) wn_edge[cur[wn_slot[b]]++] = i;
    }
    c->nwn = (int)wn_node.size();
    c->h_wn_node = wn_node;
    vector<int32_t> wall0(c->h_wall);
    for (auto& v : wall0) v -= 1;
    TRY(upload(c, c->wall, wall0));
    TRY(upload(c, c->wn_node, wn_node));
    TRY(upload(c, c->wn_ptr, wn_ptr));
    TRY(upload(c, c->wn_edge, wn_edge));
    TRY(zero(c, c->wn_x, (size_t)c->nwn));
    TRY(zero(c, c->wn_y, (size_t)c->nwn));
    TRY(zero(c, c->wn_valid, (size_t)c->nwn));
    // per-node BC table, lists applied in order so the last entry wins
    vector<unsigned char> flag((size_t)P, 0);
    vector<double> vx((size_t)P, 0), vy((size_t)P, 0), rho((size_t)P, 0), tf((size_t)P, 0);
    auto chk = [&](int n) { return n >= 1 && n <= P; };
    for (int i = 0; i < bc->nfixv; ++i) {
        int n = bc->ifixv_node[i];
        if (!chk(n)) return fail("ifixv_node out of range");
        flag[n - 1] |= 1; vx[n - 1] = bc->rfixv_valuex[i]; vy[n - 1] = bc->rfixv_valuey[i];
    }
    for (int n : wn_node) flag[n] |= 2;
    for (int i = 0; i < bc->nfixrho; ++i) {
        int n = bc->ifixrho_node[i];
        if (!chk(n)) return fail("ifixrho_node out of range");
        flag[n - 1] |= 4; rho[n - 1] = bc->rfixrho_value[i];
    }
    for (int i = 0; i < bc->nfixt; ++i) {
        int n = bc->ifixt_node[i];
        if (!chk(n)) return fail("ifixt_node out of range");
        flag[n - 1] |= 8; tf[n - 1] = bc->rfixt_value[i];
    }
    vector<int32_t> bnode, bkind, bw;
    vector<double> bvx, bvy, brho, bT;
    for (int n = 0; n < P; ++n)
        if (flag[n]) {
            bnode.push_back(n); bkind.push_back(flag[n]); bw.push_back(wn_slot[n + 1]);
            bvx.push_back(vx[n]); bvy.push_back(vy[n]); brho.push_back(rho[n]); bT.push_back(tf[n]);
        }
    c->nb = (int)bnode.size();
    c->h_bcflag = flag;
    TRY(upload(c, c->bcflag, flag));
    TRY(upload(c, c->bc_node, bnode)); TRY(upload(c, c->bc_kind, bkind)); TRY(upload(c, c->bc_wslot, bw));
    TRY(upload(c, c->bc_vx, bvx)); TRY(upload(c, c->bc_vy, bvy)); TRY(upload(c, c->bc_rho, brho)); TRY(upload(c, c->bc_T, bT));
    // mesh-motion lists: ilaux = [I_M; IFM] (dataLoader.f90:258-268), body sets grouped by id (:208-215)
    c->nmove = bc->nmove;
    c->nnmove = bc->nmove + bc->nfix_move;
    c->h_ilaux.assign(bc->i_m, bc->i_m + bc->nmove);
    c->h_ilaux.insert(c->h_ilaux.end(), bc->ifm, bc->ifm + bc->nfix_move);
    for (int n : c->h_ilaux)
        if (!chk(n)) return fail("I_M/IFM node out of range");
    vector<int32_t> il0(c->h_ilaux), last;
    topo::last_wins(c->h_ilaux.data(), c->nnmove, P, last);
    for (auto& v : il0) v -= 1;
    TRY(upload(c, c->ilaux, il0));
    TRY(upload(c, c->ilaux_last, last));
    int nset = 0;
    for (int i = 0; i < bc->nsets; ++i) {
        if (bc->iset_id[i] < 1 || bc->iset_id[i] > 10) return fail("set id out of range (1..10)");
        nset = std::max(nset, bc->iset_id[i]);
    }
    c->nset = nset;
    vector<int32_t> sptr((size_t)nset + 1, 0), n1, n2, el, se_node, se_set;
    for (int s = 1; s <= nset; ++s) {
        for (int i = 0; i < bc->nsets; ++i)
            if (bc->iset_id[i] == s) {
                n1.push_back(bc->iset_n1[i] - 1); n2.push_back(bc->iset_n2[i] - 1);
                if (bc->iset_elem[i] < 0 || bc->iset_elem[i] > c->nelem) return fail("ISET element out of range");
                el.push_back(bc->iset_elem[i] - 1);   // 0 in a rank-local deck = element not held by this rank -> -1
                se_node.push_back(bc->iset_n1[i] - 1); se_set.push_back(s - 1);
                se_node.push_back(bc->iset_n2[i] - 1); se_set.push_back(s - 1);
            }
        sptr[s] = (int)n1.size();
    }
    c->nse = (int)se_node.size();
    c->ale = c->nse > 0;
    TRY(upload(c, c->set_ptr, sptr)); TRY(upload(c, c->set_n1, n1)); TRY(upload(c, c->set_n2, n2)); TRY(upload(c, c->set_el, el));
    c->nedges = (int)n1.size();
    TRY(c->fvisc.alloc(20)); TRY(c->skin.alloc(3 * (size_t)std::max(c->nedges, 1)));
    CK(cudaMemsetAsync(c->fvisc.p, 0, 20 * sizeof(double), c->st));
    CK(cudaMemsetAsync(c->skin.p, 0, 3 * (size_t)std::max(c->nedges, 1) * sizeof(double), c->st));
    TRY(upload(c, c->se_node, se_node)); TRY(upload(c, c->se_set, se_set));
    return 0;
}

// The device-built topology (topo_gpu.cu) is mirrored into the host vectors only when something asks for it: cfdb_get of
// esup1/esup2/psup1/psup2/lap_idx/lap_rowptr, the stage tiler.
static int ensure_host_topology(cfdb_ctx* c) {
    if (c->host_topo_valid) return 0;
    const size_t P = c->npoin, E = c->nelem;
    c->esup2.resize(P + 1); c->h_eslot.resize(3 * E); c->psup2.resize(P + 1); c->psup1.resize(c->npsup);
    c->lap_rowptr.resize(P + 1); c->lap_idx.resize(c->nnz);
    CK(cudaMemcpyAsync(c->esup2.data(), c->d_esup2.p, (P + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(c->h_eslot.data(), c->eslot.p, 3 * E * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(c->psup2.data(), c->d_psup2.p, (P + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    if (c->npsup) CK(cudaMemcpyAsync(c->psup1.data(), c->d_psup1.p, (size_t)c->npsup * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(c->lap_rowptr.data(), c->d_lap_rowptr.p, (P + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(c->lap_idx.data(), c->d_lap_idx.p, (size_t)c->nnz * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    c->esup1.resize(3 * E);
    if (c->perm_on)   // the device list holds internal element positions: back to the file's numbering
        for (size_t k = 0; k < 3 * E; ++k) c->h_eslot[k] = 3 * c->h_i2e[c->h_eslot[k] / 3] + c->h_eslot[k] % 3;
    for (size_t k = 0; k < 3 * E; ++k) c->esup1[k] = c->h_eslot[k] / 3 + 1;
    for (auto& v : c->lap_idx) v += 1;   // host copy is 1-based like the reference's lap_idx
    c->host_topo_valid = true;
    return 0;
}

// Tiles of the fused RK stage and the internal element order (host_topology.h: build_tiling).  Called at the end of
// cfdb_create: the device topology has been built from the file-order connectivity (so every per-node list is in ascending
// FILE-order element id, the reference's summation order); from here on the element arrays live in tile order.
//   CFDB_NO_PERM=1   keep the file's element order (tiles = runs of consecutive elements)
//   CFDB_TILE_TE=384|512 tile size (384 default)
static int build_stage_tiles(cfdb_ctx* c, const int32_t* inpoel, const double* X, const double* Y) {
    const size_t E = c->nelem, P = c->npoin;
    // TE = 32 x element warps (stage_fused.cuh): 12 element warps at 144 registers + 4 auxiliary warps at 80; 16 + 4 at
    // 112 / 32 is kept for experiments but needs more shared memory than an SM has once the mesh is large
    int TE = getenv("CFDB_TILE_TE") ? atoi(getenv("CFDB_TILE_TE")) : 384;
    if (TE != 384 && TE != 512) return fail("CFDB_TILE_TE must be 384 or 512");
    // CFDB_TILE_ORDER=morton: runs of a Z-curve instead of recursive coordinate bisection, kept for A/B timing
    int order = getenv("CFDB_NO_PERM") ? topo::ORDER_FILE : topo::ORDER_RCB;
    if (order != topo::ORDER_FILE && getenv("CFDB_TILE_ORDER") && !strcmp(getenv("CFDB_TILE_ORDER"), "morton")) order = topo::ORDER_MORTON;
    const bool permute = order != topo::ORDER_FILE;
    // host copies of esup2 / eslot (file order)
    vector<int32_t> esup2(P + 1), eslot(3 * E), esup1(3 * E);
    if (c->host_topo_valid) {
        esup2 = c->esup2; eslot = c->h_eslot; esup1 = c->esup1;
    } else {
        CK(cudaMemcpyAsync(esup2.data(), c->d_esup2.p, (P + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(eslot.data(), c->eslot.p, 3 * E * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        for (size_t k = 0; k < 3 * E; ++k) esup1[k] = eslot[k] / 3 + 1;
    }
    topo::Tiling T;
    topo::build_tiling(inpoel, c->nelem, c->npoin, X, Y, esup1, esup2, eslot, c->h_bcflag, TE, order, T);
    if (T.L.nint_max > 65535 || 12 * TE > 65535) return fail("tile too large for 16-bit slots");
    if (T.rank_overflow) { c->tiles_ok = false; return 0; }   // a node with more than 256 elements: the two-kernel stage
    if (getenv("CFDB_CHECK_TILES")) {   // independent check of the summation schedule against the mesh (host_topology.h)
        const long bad = topo::check_tiling(inpoel, c->nelem, c->npoin, esup1, esup2, eslot, T);
        if (bad) return fail("cfdb_create: the tiling is inconsistent with the mesh (" + std::to_string(bad) + " violations)");
    }
    c->ntiles = T.ntiles;
    c->tile_ncw = TE / 32;
    c->tile_interior = T.interior_fraction;
    c->nbnodes = (int)T.bnodes.size();
    c->Epad = (long)T.ntiles * TE;
    TRY(upload(c, c->TB, T.blocks));
    TRY(upload(c, c->bnodes, T.bnodes));
    TRY(upload(c, c->bn_ptr, T.bn_ptr));
    c->norphans = (int)T.orphans.size();
    TRY(upload(c, c->orphans, T.orphans));
    TRY(zero(c, c->geo, 7 * (size_t)c->Epad));
    if (permute) {
        c->h_i2e = T.i2e;
        c->h_e2i = T.e2i;
        TRY(upload(c, c->i2e, T.i2e));
        TRY(upload(c, c->e2i, T.e2i));
        // connectivity into tile order; per-node element lists keep their (file-order ascending) sequence but name internal positions
        DBuf<int> tmp;
        TRY(tmp.alloc(3 * E));
        for (int cpt = 0; cpt < 3; ++cpt)
            LAUNCH(K_LAYOUT, k::perm_rows<int>, grid_for((long)E, 256), 256, (long)E, 1, c->i2e.p, c->inp.p + cpt * E, tmp.p + cpt * E);
        CK(cudaMemcpyAsync(c->inp.p, tmp.p, 3 * E * sizeof(int), cudaMemcpyDeviceToDevice, c->st));
        LAUNCH(K_LAYOUT, remap_slots, grid_for(3 * (long)E, 256), 256, 3 * (long)E, c->e2i.p, c->eslot.p);
        if (c->nedges) LAUNCH(K_LAYOUT, remap_ids, grid_for(c->nedges, 128), 128, c->nedges, c->e2i.p, c->set_el.p);
        CK(cudaStreamSynchronize(c->st));
        tmp.release();
        c->perm_on = true;
    }
    // shared-memory layout of the stage kernel: barriers + counters, C[2][12][TE], the A ring (static block + gathered
    // nodal data of a tile), the B ring (its element stream)
    const topo::TileLayout& L = T.L;
    k::TileGeom& G = c->tgeom;
    auto up16 = [](int v) { return (v + 15) & ~15; };
    auto up128 = [](int v) { return (v + 127) & ~127; };
    G.TE = TE; G.ntn_max = L.ntn_max; G.nint_max = L.nint_max; G.nslot_max = L.nslot_max;
    G.off_lnode = L.off_lnode; G.off_tnode = L.off_tnode; G.off_nptr = L.off_nptr; G.off_slots = L.off_slots; G.off_bcf = L.off_bcf;
    G.off_brank = L.off_brank; G.off_bbase = L.off_bbase;
    G.tb_bytes = L.tb_bytes;
    G.nfields = c->par.ITLOCAL != 0 ? 12 : 11;
    G.a_static = 0;
    G.a_u = up128(L.tb_bytes);
    G.a_t = G.a_u + L.ntn_max * 32;
    G.a_m = G.a_t + up16(L.ntn_max * 8);
    G.a_g = G.a_m + up16(L.nint_max * 8);
    G.a_bytes = up128(G.a_g + up16(L.nint_max * 8));
    G.b_bytes = up128(G.nfields * TE * 8);
    G.off_c = 384;   // barriers (<= 160 B) + the optional cycle counters (ST_COUNT x 8 B from byte 176)
    G.off_a = up128(G.off_c + 2 * 12 * TE * 8);
    c->tile_na = 3;
    G.off_b = G.off_a + c->tile_na * G.a_bytes;
    c->stage_smem = (size_t)G.off_b + 2 * (size_t)G.b_bytes;
    int smem_max = 0;
    CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    c->tiles_ok = c->stage_smem <= (size_t)smem_max;   // else: the two-kernel stage (very high valence / odd meshes)
    c->geo_dirty = true;
    if (getenv("CFDB_STAGE_STATS") && !c->stage_stats) {
        CK(cudaMalloc(&c->stage_stats, k::ST_COUNT * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(c->stage_stats, 0, k::ST_COUNT * sizeof(unsigned long long), c->st));
    }
    if (getenv("CFDB_VERBOSE"))
        fprintf(stderr, "[cfdb_create] tiles: TE %d, %d tiles, interior nodes %.3f, nodes/tile <= %d (interior <= %d), block %d B, smem %zu B%s\n",
                TE, c->ntiles, c->tile_interior, L.ntn_max, L.nint_max, L.tb_bytes, c->stage_smem, c->tiles_ok ? "" : " (too large: fused stage off)");
    return 0;
}

extern "C" int cfdb_create(cfdb_ctx** out, const cfdb_params* par, int32_t npoin, int32_t nelem, const double* X,
                           const double* Y, const int32_t* inpoel, const cfdb_bc* bc, int device) {
    *out = nullptr;
    int ndev = 0;
    const bool verbose = getenv("CFDB_VERBOSE") != nullptr;   // phase times of the set-up to stderr
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!verbose) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[cfdb_create] %-28s %8.3f s\n", what, std::chrono::duration<double>(now - t_prev).count());
        t_prev = now;
    };
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("cfdb_create: no CUDA device available (libcfdb200 has no CPU path)");
    if (device < 0 || device >= ndev) return fail("cfdb_create: bad device index");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("cfdb_create: device is not sm_100 (this library is built for sm_100a only)");
    if (npoin < 1 || nelem < 1) return fail("cfdb_create: empty mesh");
    // NGAS /= 0 selects the equilibrium-air TGAS branch of the reference (subrutinas.f90:706-741, ns2DComp.ALE.f90:78,456),
    // which this library does not implement: refuse the deck instead of silently running ideal gas
    if (par->NGAS != 0) return fail("cfdb_create: NGAS /= 0 (equilibrium-air TGAS) is not implemented; only the ideal-gas path NGAS = 0");
    for (size_t k = 0; k < 3 * (size_t)nelem; ++k)
        if (inpoel[k] < 1 || inpoel[k] > npoin) return fail("cfdb_create: inpoel entry out of range");
    lap("validate inpoel");
    CK(cudaSetDevice(device));
    CK(cudaFree(nullptr));
    lap("CUDA context");
    cfdb_ctx* c = new cfdb_ctx();
    c->device = device;
    c->par = *par;
    c->npoin = npoin;
    c->nelem = nelem;
    c->n_owned = npoin;
    auto bail = [&](int r) { cfdb_destroy(c); return r; };
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) return bail(fail("stream create failed"));
    const size_t P = npoin, E = nelem;
#define B(x)                    \
    do {                        \
        int _r = (x);           \
        if (_r) return bail(_r); \
    } while (0)
    c->h_inpoel.assign(inpoel, inpoel + 3 * E);
    vector<int32_t> inp_soa(3 * E);
    for (size_t e = 0; e < E; ++e)
        for (int i = 0; i < 3; ++i) inp_soa[i * E + e] = inpoel[3 * e + i] - 1;
    B(upload(c, c->inp, inp_soa));
    if (verbose) cudaStreamSynchronize(c->st);
    lap("connectivity SoA + upload");
    // topology: getEsup/getPsup, the Laplacian pattern and the row positions -- built on the device (topo_gpu.cu,
    // SURVEY.md 8f N4) and mirrored to the host vectors; CFDB_HOST_TOPO=1 keeps the serial host code (bit-identical,
    // tests/test_topology.py compares the two)
    vector<uint8_t> lpos;
    const bool host_topo = getenv("CFDB_HOST_TOPO") != nullptr || E == 0 || P == 0;
    if (host_topo) {
        topo::build_esup(inpoel, nelem, npoin, c->esup1, c->esup2, &c->h_eslot);
        topo::build_psup(inpoel, npoin, c->esup1, c->esup2, c->psup1, c->psup2);
        topo::build_lap_pattern(npoin, c->psup1, c->psup2, c->lap_idx, c->lap_rowptr);
        topo::build_lap_pos(inpoel, npoin, c->esup1, c->esup2, c->lap_idx, c->lap_rowptr, lpos, c->maxrow);
    } else {
        DBuf<int> d_psup2;
        B(c->d_esup2.alloc(P + 1));
        B(c->eslot.alloc(3 * E));
        B(d_psup2.alloc(P + 1));
        B(c->d_lap_rowptr.alloc(P + 1));
        B(c->lpos.alloc(9 * E));
        int *d_psup1 = nullptr, *d_idx0 = nullptr, npsup = 0;
        int rc = topogpu::build(c->st, c->inp.p, nelem, npoin, c->d_esup2.p, c->eslot.p, d_psup2.p, &d_psup1, &npsup,
                                c->d_lap_rowptr.p, &d_idx0, c->lpos.p, &c->maxrow);
        if (rc) { d_psup2.release(); return bail(fail(std::string("device topology build failed: ") + cudaGetErrorString((cudaError_t)rc))); }
        if (c->maxrow > 32) { d_psup2.release(); return bail(fail("node valence above 31 is not supported")); }
        lap("device topology");
        c->nnz = npsup + npoin;
        c->d_lap_idx.p = d_idx0;     // ownership of the two late-sized arrays passes to the context
        c->d_lap_idx.n = (size_t)c->nnz + 1;
        // host mirrors (fields esup1/esup2/psup1/psup2/lap_idx/lap_rowptr, the chunk/tile planners) are downloaded on
        // first use: ensure_host_topology
        c->npsup = npsup;
        c->d_psup1.p = d_psup1; c->d_psup1.n = (size_t)npsup + 1;
        c->d_psup2.p = d_psup2.p; c->d_psup2.n = d_psup2.n;
        d_psup2.p = nullptr; d_psup2.n = 0;
        c->host_topo_valid = false;
    }
    if (host_topo) c->nnz = c->lap_rowptr[npoin];
    if (host_topo) lap("host topology");
    if (c->maxrow > 32) return bail(fail("node valence above 31 is not supported"));
    if (host_topo) {
        vector<int32_t> idx0(c->lap_idx);
        for (auto& v : idx0) v -= 1;
        B(upload(c, c->d_esup2, c->esup2));
        B(upload(c, c->eslot, c->h_eslot));
        B(upload(c, c->d_lap_idx, idx0));
        B(upload(c, c->d_lap_rowptr, c->lap_rowptr));
        B(upload(c, c->lpos, lpos));
    }
    B(upload(c, c->X, X, P));
    B(upload(c, c->Y, Y, P));
    B(zero(c, c->NR1, k::NREC * P));
    B(zero(c, c->NR2, k::NREC * P));
    B(zero(c, c->ntmp, P));
    c->T.col.q = c->NR1.p + k::NR1_T; c->GAMM.col.q = c->NR1.p + k::NR1_GAMM; c->VEL_X.col.q = c->NR1.p + k::NR1_VX; c->VEL_Y.col.q = c->NR1.p + k::NR1_VY;
    c->RHO.col.q = c->NR2.p + k::NR2_RHO; c->E.col.q = c->NR2.p + k::NR2_E; c->P.col.q = c->NR2.p + k::NR2_P; c->RMACH.col.q = c->NR2.p + k::NR2_RMACH;
    for (auto* d : {&c->X1, &c->Y1, &c->M, &c->W_X, &c->W_Y, &c->lap_diag, &c->by, &c->bp, &c->bp2, &c->br, &c->bz, &c->bb, &c->xpos, &c->ypos, &c->dxpos,
                    &c->dypos, &c->pos_aux, &c->W_x_old, &c->W_y_old, &c->tmpA, &c->tmpB, &c->tmpC, &c->by2, &c->pos_aux2})
        B(zero(c, *d, P));
    for (auto* d : {&c->U, &c->U1, &c->RHS, &c->RHS1, &c->RHS2, &c->RHS3, &c->UN}) B(zero(c, *d, 4 * P));
    // +1024: the fused stage streams whole tiles (<= 512 elements) of SHOC, T_SUGN1-3 and DTL with bulk copies
    for (auto* d : {&c->area, &c->HH, &c->HHX, &c->HHY, &c->SHOC, &c->TS1, &c->TS2, &c->TS3, &c->DT, &c->DTL, &c->area_old})
        B(zero(c, *d, E + 1024));
    B(zero(c, c->dNx, 3 * E));
    B(zero(c, c->dNy, 3 * E));
    B(zero(c, c->EC, 12 * E));
    B(zero(c, c->lap_sparse, (size_t)c->nnz));
    size_t nchunk = (std::max(P, E) + 4095) / 4096;
    B(zero(c, c->redA, 8 * nchunk + 8));
    B(zero(c, c->redB, 64 * 8 * ((nchunk + 4095) / 4096) + 8));   // x64: room for the global chunk count of a multi-rank tree
    B(zero(c, c->flags, 4));
    c->par.XREF[1] = 1.4;  // meshMove.f90:58 overwrites set 2's reference point before its first use
    c->par.YREF[1] = 0.0;
    B(upload(c, c->xref, c->par.XREF, 10));
    B(upload(c, c->yref, c->par.YREF, 10));
    if (verbose) cudaStreamSynchronize(c->st);
    lap("allocate + zero + upload");
    B(build_bc_tables(c, bc));
    if (verbose) cudaStreamSynchronize(c->st);
    lap("BC tables");
    if (c->ale) B(zero(c, c->FC, 12 * E));
    B(build_stage_tiles(c, inpoel, X, Y));
    if (verbose) cudaStreamSynchronize(c->st);
    lap("stage tiles + element order");
    if (cudaMalloc(&c->sc, sizeof(k::Scal)) != cudaSuccess) return bail(fail("cudaMalloc Scal failed"));
    if (cudaMemsetAsync(c->sc, 0, sizeof(k::Scal), c->st) != cudaSuccess) return bail(fail("memset Scal failed"));
    if (cudaMallocHost(&c->h_sc, sizeof(k::Scal)) != cudaSuccess) return bail(fail("cudaMallocHost failed"));
    if (cudaStreamSynchronize(c->st) != cudaSuccess) return bail(fail("sync after create failed"));
    lap("tail");
#undef B
    *out = c;
    return 0;
}

extern "C" void cfdb_destroy(cfdb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    for (int i = 0; i < 2; ++i) if (c->gexec[i]) cudaGraphExecDestroy(c->gexec[i]);
    if (c->st_in) { cudaStreamSynchronize(c->st_in); cudaStreamDestroy(c->st_in); }
    if (c->st_out) { cudaStreamSynchronize(c->st_out); cudaStreamDestroy(c->st_out); }
    for (auto e : {c->ev_in_done, c->ev_step_done, c->ev_in_used[0], c->ev_in_used[1], c->ev_out_done[0], c->ev_out_done[1]}) if (e) cudaEventDestroy(e);
    for (auto* d : {&c->sin[0], &c->sin[1], &c->sout[0], &c->sout[1]}) d->release();
    if (c->comm) ncclCommDestroy(c->comm);
    c->send_idx.release(); c->recv_idx.release(); c->sendbuf.release(); c->recvbuf.release(); c->redG.release();
    for (auto* d : {&c->inp, &c->d_esup2, &c->eslot, &c->d_lap_idx, &c->d_lap_rowptr, &c->wall, &c->wn_node, &c->wn_ptr,
                    &c->wn_edge, &c->wn_valid, &c->bc_node, &c->bc_kind, &c->bc_wslot, &c->ilaux, &c->ilaux_last, &c->se_node,
                    &c->se_set, &c->set_ptr, &c->set_n1, &c->set_n2, &c->set_el, &c->d_psup1, &c->d_psup2, &c->flags})
        d->release();
    c->lpos.release();
    c->bcflag.release();
    c->color_list.release();
    c->i2e.release(); c->e2i.release(); c->bnodes.release(); c->orphans.release(); c->bn_ptr.release(); c->TB.release(); c->geo.release();
    if (c->stage_stats) cudaFree(c->stage_stats);
    c->isfix.release();
    c->bp2.release(); c->by2.release(); c->pos_aux2.release();
    for (auto* d : {&c->X, &c->Y, &c->X1, &c->Y1, &c->area, &c->HH, &c->HHX, &c->HHY, &c->dNx, &c->dNy, &c->M, &c->U, &c->U1,
                    &c->RHS, &c->RHS1, &c->RHS2, &c->RHS3, &c->UN, &c->NR1, &c->NR2, &c->ntmp, &c->W_X, &c->W_Y,
                    &c->SHOC, &c->TS1, &c->TS2, &c->TS3, &c->DT, &c->DTL, &c->EC, &c->FC,
                    &c->bc_vx, &c->bc_vy, &c->bc_rho, &c->bc_T, &c->wn_x, &c->wn_y, &c->lap_sparse, &c->lap_diag, &c->by,
                    &c->bp, &c->br, &c->bz, &c->bb, &c->xpos, &c->ypos, &c->dxpos, &c->dypos, &c->pos_aux, &c->xref, &c->yref,
                    &c->W_x_old, &c->W_y_old, &c->area_old, &c->redA, &c->redB, &c->tmpA, &c->tmpB, &c->tmpC})
        d->release();
    for (auto& p : c->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : c->evpool) cudaEventDestroy(e);
    if (c->sc) cudaFree(c->sc);
    if (c->h_sc) cudaFreeHost(c->h_sc);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU plumbing: ghost refresh (packed ncclSend/ncclRecv per neighbour) and tiny all-reduces
extern "C" int cfdb_nccl_unique_id(void* out128) {
    ncclUniqueId id;
    NK(ncclGetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(out128, &id, 128);
    return 0;
}
extern "C" int cfdb_comm_init(cfdb_ctx* c, const void* uid128, int32_t rank, int32_t nranks) {
    CK(cudaSetDevice(c->device));
    if (c->comm) return fail("cfdb_comm_init: communicator already initialised");
    ncclUniqueId id;
    memcpy(&id, uid128, 128);
    NK(ncclCommInitRank(&c->comm, nranks, id, rank));
    c->rank = rank;
    c->nranks = nranks;
    c->epoch++;
    return 0;
}
extern "C" int cfdb_set_halo(cfdb_ctx* c, int32_t n_owned, int32_t nneigh, const int32_t* neigh_rank,
                             const int32_t* send_ptr, const int32_t* send_idx, const int32_t* recv_ptr,
                             const int32_t* recv_idx) {
    CK(cudaSetDevice(c->device));
    if (n_owned < 0 || n_owned > c->npoin) return fail("cfdb_set_halo: n_owned out of range");
    c->n_owned = n_owned;
    c->nb_rank.assign(neigh_rank, neigh_rank + nneigh);
    c->send_ptr.assign(send_ptr, send_ptr + nneigh + 1);
    c->recv_ptr.assign(recv_ptr, recv_ptr + nneigh + 1);
    for (int i = 0; i < c->send_ptr[nneigh]; ++i)
        if (send_idx[i] < 0 || send_idx[i] >= n_owned) return fail("cfdb_set_halo: send list must name owned nodes");
    for (int i = 0; i < c->recv_ptr[nneigh]; ++i)
        if (recv_idx[i] < n_owned || recv_idx[i] >= c->npoin) return fail("cfdb_set_halo: receive list must name ghost nodes");
    TRY(upload(c, c->send_idx, send_idx, (size_t)c->send_ptr[nneigh]));
    TRY(upload(c, c->recv_idx, recv_idx, (size_t)c->recv_ptr[nneigh]));
    TRY(c->sendbuf.alloc(k::HALO_W * (size_t)c->send_ptr[nneigh] + 8));
    TRY(c->recvbuf.alloc(k::HALO_W * (size_t)c->recv_ptr[nneigh] + 8));
    CK(cudaStreamSynchronize(c->st));
    c->epoch++;
    return 0;
}
// chunk-aligned ownership: this rank's owned nodes are the global nodes [gid0, gid0 + n_owned)
extern "C" int cfdb_set_reduction_layout(cfdb_ctx* c, int64_t gid0, int64_t npoin_global) {
    CK(cudaSetDevice(c->device));
    if (gid0 < 0 || gid0 % 4096 != 0) return fail("cfdb_set_reduction_layout: gid0 must be a non-negative multiple of 4096");
    if (gid0 + c->n_owned > npoin_global) return fail("cfdb_set_reduction_layout: owned range exceeds the global node count");
    if (c->n_owned % 4096 != 0 && gid0 + c->n_owned != npoin_global)
        return fail("cfdb_set_reduction_layout: only the last rank may own a partial chunk");
    c->red_chunk0 = gid0 / 4096;
    c->red_nchunk_global = (npoin_global + 4095) / 4096;
    TRY(c->redG.alloc(8 * (size_t)c->red_nchunk_global + 8));
    // the upper tree levels run over the global chunk count: redA/redB must hold 8 vectors of ceil(M/4096) sums
    const size_t need = 8 * (((size_t)c->red_nchunk_global + 4095) / 4096) + 8;
    if (c->redA.n < need || c->redB.n < need) return fail("cfdb_set_reduction_layout: reduction scratch too small for the global chunk count");
    c->red_aligned = true;
    c->epoch++;
    CK(cudaStreamSynchronize(c->st));
    return 0;
}
// exchange `w` doubles per node between sendbuf and recvbuf (already packed / to be unpacked by the caller)
static int halo_sendrecv(cfdb_ctx* c, int w, cudaStream_t st) {
    const int nn = (int)c->nb_rank.size();
    if (!nn) return 0;
    if (!c->comm) return fail("halo exchange requested but cfdb_comm_init was not called");
    NK(ncclGroupStart());
    for (int k = 0; k < nn; ++k) {
        int ns = c->send_ptr[k + 1] - c->send_ptr[k], nr = c->recv_ptr[k + 1] - c->recv_ptr[k];
        if (ns) NK(ncclSend(c->sendbuf.p + (size_t)w * c->send_ptr[k], (size_t)w * ns, ncclDouble, c->nb_rank[k], c->comm, st));
        if (nr) NK(ncclRecv(c->recvbuf.p + (size_t)w * c->recv_ptr[k], (size_t)w * nr, ncclDouble, c->nb_rank[k], c->comm, st));
    }
    NK(ncclGroupEnd());
    c->launches += 1;
    return 0;
}
static int halo_vec(cfdb_ctx* c, double* v, int w) {
    const int nn = (int)c->nb_rank.size();
    if (!nn) return 0;
    int ms = c->send_ptr[nn], mr = c->recv_ptr[nn];
    if (ms) LAUNCH(K_HALO, k::halo_pack, grid_for((long)ms * w, 128), 128, ms, w, c->send_idx.p, v, c->sendbuf.p);
    TRY(halo_sendrecv(c, w, c->st));
    if (mr) LAUNCH(K_HALO, k::halo_unpack, grid_for((long)mr * w, 128), 128, mr, w, c->recv_idx.p, c->recvbuf.p, v);
    return 0;
}
// ghosts of U1, T, VEL_X, VEL_Y, RHO, E, P, RMACH in one message per neighbour (after every RK stage)
static int halo_state(cfdb_ctx* c, cudaStream_t st = nullptr) {
    const int nn = (int)c->nb_rank.size();
    if (!nn) return 0;
    if (!st) st = c->st;
    int ms = c->send_ptr[nn], mr = c->recv_ptr[nn];
    if (ms) LAUNCH_ON(st, K_HALO, k::halo_pack_state, grid_for(ms, 128), 128, ms, c->send_idx.p, c->U1.p, c->T.col, c->VEL_X.col, c->VEL_Y.col, c->E.col, c->P.col, c->RMACH.col, c->sendbuf.p);
    TRY(halo_sendrecv(c, k::HALO_W, st));
    if (mr) LAUNCH_ON(st, K_HALO, k::halo_unpack_state, grid_for(mr, 128), 128, mr, c->recv_idx.p, c->recvbuf.p, c->U1.p, c->T.col, c->VEL_X.col, c->VEL_Y.col,
                      c->RHO.col, c->E.col, c->P.col, c->RMACH.col);
    return 0;
}
static int allreduce(cfdb_ctx* c, double* dev, int count, ncclRedOp_t op) {
    if (c->nranks <= 1) return 0;
    NK(ncclAllReduce(dev, dev, count, ncclDouble, op, c->comm, c->st));
    c->launches += 1;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// canonical reduction of nv vectors whose first-level chunk sums are in redA laid out [v][m]
// Multi-rank: with chunk-aligned ownership (cfd_b200/partition.py) every rank places its chunk sums at their global chunk
// positions in a zero-filled array, one ncclAllReduce(sum) assembles it (exactly one rank contributes a non-zero value per
// entry, and x + 0 + ... + 0 = x bit for bit: chunk sums start from +0.0, so they are never -0), and every rank runs the
// upper levels of the tree on the same numbers: the result has the bits of the single-GPU reduction.  Without the alignment
// (owned sets not contiguous in the global numbering) the per-rank canonical sums are added by ncclAllReduce: round-off level.
static int reduce_levels(cfdb_ctx* c, int nv, long m, int slot) {
    double* in = c->redA.p;
    double *out = c->redB.p, *spare = c->redA.p;   // outputs alternate; the level-1 sums in redA are dead once level 2 is done
    const bool global_tree = c->nranks > 1 && c->red_aligned;
    if (global_tree) {
        const long M = c->red_nchunk_global;
        CK(cudaMemsetAsync(c->redG.p, 0, (size_t)nv * M * sizeof(double), c->st));
        for (int v = 0; v < nv; ++v)
            CK(cudaMemcpyAsync(c->redG.p + v * M + c->red_chunk0, in + v * m, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
        NK(ncclAllReduce(c->redG.p, c->redG.p, (size_t)nv * M, ncclDouble, ncclSum, c->comm, c->st));
        c->launches += 1;
        in = c->redG.p;
        out = c->redA.p;
        spare = c->redB.p;
        m = M;
    }
    while (m > 1) {
        long m2 = (m + 4095) / 4096;
        for (int v = 0; v < nv; ++v)
            LAUNCH(K_DOT, k::dot_chunks, (int)std::min<long>(m2, 1024), 256, m, in + v * m, (const double*)nullptr, out + v * m2);
        in = out;
        std::swap(out, spare);
        m = m2;
    }
    CK(cudaMemcpyAsync(&c->sc->red[slot], in, nv * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    if (!global_tree) TRY(allreduce(c, &c->sc->red[slot], nv, ncclSum));  // multi-rank, not aligned: sum of the per-rank canonical sums
    return 0;
}
static int dev_dot(cfdb_ctx* c, long n, const double* x, const double* y, int slot) {
    long m = (n + 4095) / 4096;
    LAUNCH(K_DOT, k::dot_chunks, (int)std::min<long>(m, 148 * 8), 256, n, x, y, c->redA.p);
    return reduce_levels(c, 1, m, slot);
}

static int read_scal(cfdb_ctx* c) {
    CK(cudaMemcpyAsync(c->h_sc, c->sc, sizeof(k::Scal), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

// ---------------------------------------------------------------------------------------------
static int run_normales(cfdb_ctx* c) {
    if (c->nwn)
        LAUNCH(K_NORMALES, k::normales, grid_for(c->nwn, 128), 128, c->nwn, c->wn_node.p, c->wn_ptr.p, c->wn_edge.p,
               c->wall.p, c->X.p, c->Y.p, c->wn_x.p, c->wn_y.p, c->wn_valid.p);
    return 0;
}
static int run_deriv(cfdb_ctx* c) {
    LAUNCH(K_SCALAR, k::set_double, 1, 1, &c->sc->HMIN, (double)INFINITY);
    LAUNCH(K_DERIV, k::deriv, grid_for(c->nelem, 256), 256, c->nelem, c->inp.p, c->X.p, c->Y.p, c->area.p, c->HH.p,
           c->HHX.p, c->HHY.p, c->dNx.p, c->dNy.p, c->sc);
    TRY(allreduce(c, &c->sc->HMIN, 1, ncclMin));
    return 0;
}
static int run_masas(cfdb_ctx* c) {
    LAUNCH(K_MASAS, k::masas, grid_for(c->npoin, 256), 256, c->npoin, c->d_esup2.p, c->eslot.p, c->area.p, c->M.p);
    return 0;
}
static int run_laplace(cfdb_ctx* c) {
    if (c->maxrow <= 12)
        LAUNCH(K_LAPLACE, k::laplace<12>, grid_for(c->npoin, 128), 128, c->npoin, c->nelem, c->d_esup2.p, c->eslot.p,
               c->lpos.p, c->inp.p, c->dNx.p, c->dNy.p, c->X.p, c->Y.p, c->d_lap_rowptr.p, c->lap_sparse.p,
               c->lap_diag.p, c->flags.p);
    else
        LAUNCH(K_LAPLACE, k::laplace<32>, grid_for(c->npoin, 128), 128, c->npoin, c->nelem, c->d_esup2.p, c->eslot.p,
               c->lpos.p, c->inp.p, c->dNx.p, c->dNy.p, c->X.p, c->Y.p, c->d_lap_rowptr.p, c->lap_sparse.p,
               c->lap_diag.p, c->flags.p);
    return 0;
}
static int run_gcl(cfdb_ctx* c, const double* dt_dev, double dt_const) {
    LAUNCH(K_GCL, k::gcl, grid_for(c->npoin, 256), 256, c->npoin, c->nelem, c->d_esup2.p, c->eslot.p, c->inp.p, c->dNx.p,
           c->dNy.p, c->area.p, c->area_old.p, c->W_X.p, c->W_x_old.p, dt_const, dt_dev, c->M.p);
    return 0;
}

extern "C" int cfdb_geometry(cfdb_ctx* c, int32_t moving_step) {
    CK(cudaSetDevice(c->device));
    const size_t P = c->npoin, E = c->nelem;
    bool gcl = moving_step && c->par.use_gcl;
    if (gcl) CK(cudaMemcpyAsync(c->area_old.p, c->area.p, E * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    TRY(run_normales(c));
    TRY(run_deriv(c));
    c->geo_dirty = true;
    TRY(run_masas(c));
    if (gcl) {
        TRY(run_gcl(c, &c->sc->DTMIN, 0.0));
        CK(cudaMemcpyAsync(c->W_x_old.p, c->W_X.p, P * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
        CK(cudaMemcpyAsync(c->W_y_old.p, c->W_Y.p, P * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    }
    TRY(run_laplace(c));
    return 0;
}

extern "C" int cfdb_init(cfdb_ctx* c) {
    CK(cudaSetDevice(c->device));
    const cfdb_params& p = c->par;
    const size_t P = c->npoin;
    // GAMM = GAMA (ns2DComp.ALE.f90:59); RESTART free-stream branch (:408-420), evaluated on the host in
    // the reference's order — every node gets the same five numbers
    double RHOAMB = p.RHO_inf, TAMB = p.T_inf, UAMB = p.U_inf, VAMB = p.V_inf, PAMB = p.RHO_inf * p.FR * p.T_inf;
    double ENERGIA = PAMB / ((p.GAMA - 1.0) * RHOAMB) + .5 * (UAMB * UAMB + VAMB * VAMB);
    vector<double> u(4 * P);
    for (size_t i = 0; i < P; ++i) {
        u[4 * i] = RHOAMB; u[4 * i + 1] = RHOAMB * UAMB; u[4 * i + 2] = RHOAMB * VAMB; u[4 * i + 3] = ENERGIA * RHOAMB;
    }
    TRY(upload(c, c->U, u));
    LAUNCH(K_FILL, k::fill_col, grid_for(P, 256), 256, (long)P, c->GAMM.col, p.GAMA);
    LAUNCH(K_FILL, k::fill_col, grid_for(P, 256), 256, (long)P, c->VEL_X.col, UAMB);
    LAUNCH(K_FILL, k::fill_col, grid_for(P, 256), 256, (long)P, c->VEL_Y.col, VAMB);
    LAUNCH(K_FILL, k::fill_col, grid_for(P, 256), 256, (long)P, c->T.col, TAMB);
    LAUNCH(K_FILL, k::fill_const, grid_for(P, 256), 256, (long)P, c->W_X.p, -0.0);  // :121
    LAUNCH(K_FILL, k::fill_const, grid_for(P, 256), 256, (long)P, c->W_Y.p, 0.0);
    CK(cudaMemsetAsync(c->sc, 0, sizeof(k::Scal), c->st));
    LAUNCH(K_SCALAR, k::bandera_inc, 1, 1, c->sc);  // BANDERA = 1 (ns2DComp.ALE.f90:133)
    TRY(cfdb_geometry(c, 0));
    CK(cudaMemcpyAsync(c->W_x_old.p, c->W_X.p, P * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemcpyAsync(c->W_y_old.p, c->W_Y.p, P * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemcpyAsync(c->area_old.p, c->area.p, c->nelem * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    c->h_iter = 0;
    c->nestab = 1;   // ns2DComp.ALE.f90:134
    c->iterprint = 0;
    c->DISN[0] = c->DISN[1] = 0.0;
    c->theta_nonzero = false;
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

// ---------------------------------------------------------------------------------------------
static k::BcTab bctab(cfdb_ctx* c) {
    k::BcTab b;
    b.nb = c->nb; b.node = c->bc_node.p; b.kind = c->bc_kind.p; b.vx = c->bc_vx.p; b.vy = c->bc_vy.p;
    b.rho = c->bc_rho.p; b.Tfix = c->bc_T.p; b.wslot = c->bc_wslot.p; b.wn_x = c->wn_x.p; b.wn_y = c->wn_y.p;
    b.wn_valid = c->wn_valid.p;
    return b;
}

static int run_estab(cfdb_ctx* c, const double* dtmin_dev) {
    const cfdb_params& p = c->par;
    // branch-free estab_fast: 64 regs without spills (MINB 4) 0.617 ms, 48 regs (5) 0.676, 40 regs (6) 0.809; the plain form was 1.06
    static int minb = getenv("CFDB_ESTAB_MINB") ? atoi(getenv("CFDB_ESTAB_MINB")) : 4;
    auto kern = k::estab<3, true>;
    if (c->ale || (c->nranks > 1 && !c->ale_agreed)) {  // multi-rank: once the ranks have agreed on the flag (agree_on_ale); before that no rank assumes W = 0 from its own lists alone
        if (minb == 4) kern = k::estab<4, true>;
        else if (minb == 5) kern = k::estab<5, true>;
        else if (minb == 6) kern = k::estab<6, true>;
        else if (minb == 2) kern = k::estab<2, true>;
    } else {  // fixed mesh: W_X = W_Y = +0 (cfdb_set of either turns c->ale on)
        if (minb == 4) kern = k::estab<4, false>;
        else if (minb == 5) kern = k::estab<5, false>;
        else if (minb == 6) kern = k::estab<6, false>;
        else if (minb == 2) kern = k::estab<2, false>;
        else kern = k::estab<3, false>;
    }
    LAUNCH(K_ESTAB, kern, grid_for(c->nelem, 256), 256, c->nelem, c->inp.p, c->U.p, c->T.col, c->VEL_X.col, c->VEL_Y.col,
           c->W_X.p, c->W_Y.p, c->GAMM.col, c->dNx.p, c->dNy.p, p.FR, dtmin_dev, p.RHO_inf, p.T_inf, c->SHOC.p, c->TS1.p,
           c->TS2.p, c->TS3.p);
    return 0;
}

// calcRHS (+FUENTE when `ale`) into the staging buffers
// Launch shapes from the A/B runs of round 1 (profiles/r1_experiments.md): Euler flow runs the plain divisions at 4 CTAs/SM
// (128 registers, 16 warps/SM: 1.10 ms per launch on the 16 M-triangle mesh against 1.23 at 3 and 1.33 at 5 CTAs), viscous
// flow the branch-free forms with the Gauss loop rolled at 3 CTAs/SM (1.96 -> 1.67 ms).
static int run_calcrhs_elem(cfdb_ctx* c, const k::Gas& g, bool theta, bool ale, const double* dtl_arr, const double* dtl_sc) {
    const bool visc = g.mu_ref > 2.2250738585072014e-308;  // tiny(0d0), calcRHS.f90:119
    const int e0 = 0, e1 = c->nelem;
    const int B = 128, G = grid_for(e1 - e0, B);
    decltype(&k::calcrhs_elem<false, false, false, 4>) kern = nullptr;
    switch ((visc ? 4 : 0) | (theta ? 2 : 0) | (ale ? 1 : 0)) {
        case 0: kern = k::calcrhs_elem<false, false, false, 4>; break;
        case 1: kern = k::calcrhs_elem<false, false, true, 4>; break;
        case 2: kern = k::calcrhs_elem<false, true, false, 4>; break;
        case 3: kern = k::calcrhs_elem<false, true, true, 4>; break;
        case 4: kern = k::calcrhs_elem<true, false, false, 3, 128, true>; break;
        case 5: kern = k::calcrhs_elem<true, false, true, 3, 128, true>; break;
        case 6: kern = k::calcrhs_elem<true, true, false, 3, 128, true>; break;
        default: kern = k::calcrhs_elem<true, true, true, 3, 128, true>; break;
    }
    LAUNCH(K_CALCRHS, kern, G, B, e0, e1, c->nelem, c->inp.p, (c->Usrc ? c->Usrc : c->U.p), c->UN.p, c->T.col, c->W_X.p, c->W_Y.p,
           c->dNx.p, c->dNy.p, c->area.p, c->SHOC.p, dtl_arr, dtl_sc, c->TS1.p, c->TS2.p, c->TS3.p, g, c->EC.p, c->FC.p);
    return 0;
}

static int run_node(cfdb_ctx* c, cudaStream_t st, bool ale, bool update, double rk_fact, int n0 = 0, int n1 = -1,
                    const int* nlist = nullptr) {
    if (n1 < 0) n1 = c->npoin;
    if (n1 <= n0) return 0;
    const int B = CFDB_NODE_BS, G = grid_for(n1 - n0, B);
#define ARGS n0, n1, nlist, c->d_esup2.p, c->eslot.p, c->EC.p, c->FC.p, c->U.p, c->M.p, c->GAMM.col, c->W_X.p, c->W_Y.p, \
             c->bcflag.p, bctab(c), rk_fact, c->par.FR, c->U1.p, c->RHS.p, c->RHO.col, c->VEL_X.col, c->VEL_Y.col, c->E.col, \
             c->P.col, c->T.col, c->RMACH.col
    auto kern = k::node_update<false, true>;
    if (ale && update) kern = k::node_update<true, true>;
    else if (ale) kern = k::node_update<true, false>;
    else if (!update) kern = k::node_update<false, false>;
    LAUNCH_ON(st, K_NODE, kern, G, B, ARGS);
#undef ARGS
    return 0;
}

// geo[7][Epad] (the fused stage's copy of dNx, dNy, area with a tile-aligned component stride) follows the geometry
static int refresh_geo(cfdb_ctx* c) {
    if (!c->tiles_ok || !c->geo_dirty) return 0;
    LAUNCH(K_LAYOUT, k::pack_geo, grid_for(c->nelem, 256), 256, (long)c->nelem, c->Epad, c->dNx.p, c->dNy.p, c->area.p, c->geo.p);
    c->geo_dirty = false;
    return 0;
}

// fused tile stage (stage_fused.cuh) + node_update over the tile-boundary nodes
static bool fused_eligible(const cfdb_ctx* c) {
    static const bool off = getenv("CFDB_NO_FUSED") != nullptr;
    // viscous flow: the element arithmetic needs ~200 registers to run without spills (tools/sass_stalls.py); at the 144 the
    // stage kernel's element warps get it measured 2.26 ms + 0.35 ms per stage on the 16 M-triangle mesh against 1.67 + 0.55 ms
    // for the two-kernel stage, so viscous flow keeps the two kernels unless CFDB_FUSED_VISC=1
    static const bool fused_visc = getenv("CFDB_FUSED_VISC") != nullptr;
    if (c->par.FMU > 2.2250738585072014e-308 && !fused_visc) return false;
    return !off && c->tiles_ok && !c->ale && !c->use_cuarto && !c->fast && !c->colored && !c->theta_nonzero && !c->dtl_force;
}
static int run_stage_fused(cfdb_ctx* c, const k::Gas& g, const double* dtl_arr, const double* dtl_sc, double rk_fact) {
    const bool visc = g.mu_ref > 2.2250738585072014e-308;
    k::StageArgs A{};
    A.ntiles = c->ntiles; A.TB = c->TB.p; A.geo = c->geo.p; A.Epad = c->Epad;
    A.shoc = c->SHOC.p; A.ts1 = c->TS1.p; A.ts2 = c->TS2.p; A.ts3 = c->TS3.p; A.dtl_arr = dtl_arr; A.dtl_sc = dtl_sc;
    A.Usrc = c->Usrc ? c->Usrc : c->U.p; A.U = c->U.p; A.T = c->T.col; A.M = c->M.p; A.GAMM = c->GAMM.col; A.WXa = c->W_X.p; A.WYa = c->W_Y.p;
    A.bc = bctab(c); A.rk_fact = rk_fact; A.FR = c->par.FR; A.g = g;
    A.ECB = c->EC.p; A.U1 = c->U1.p; A.RHS = c->RHS.p; A.RHO = c->RHO.col; A.VELX = c->VEL_X.col; A.VELY = c->VEL_Y.col; A.Ea = c->E.col;
    A.Pa = c->P.col; A.Ta = c->T.col; A.RMACH = c->RMACH.col;
    k::TileGeom G = c->tgeom;
    G.nfields = dtl_arr ? 12 : 11;
    if (G.nfields > c->tgeom.nfields) return fail("run_stage_fused: stage layout was sized without a local time step array");
    void (*kern)(const k::TileGeom, const k::StageArgs) = nullptr;
    if (c->tile_ncw == 16) kern = visc ? k::stage_fused<true, 16, 3, 2> : k::stage_fused<false, 16, 3, 2>;
    else if (c->stage_stats) kern = visc ? k::stage_fused<true, 12, 3, 2, true> : k::stage_fused<false, 12, 3, 2, true>;   // CFDB_STAGE_STATS
    else kern = visc ? k::stage_fused<true, 12, 3, 2> : k::stage_fused<false, 12, 3, 2>;
    // tile-boundary nodes: their contributions go to the boundary records (the otherwise idle staging buffer EC), finished by
    // boundary_update below.  (Finishing them inside the stage kernel -- last contributing tile, atomics + fences -- measured
    // 3.08 ms against 1.16 + 0.35 ms on the 16 M-triangle mesh and was removed, profiles/r2_experiments.md.)
    A.stats = c->stage_stats;
    static const int stat_warp = getenv("CFDB_STAGE_STATS") ? std::max(0, atoi(getenv("CFDB_STAGE_STATS")) - 1) % 12 : 0;
    A.stat_warp = stat_warp;
    static std::map<const void*, size_t> attr_done;   // per kernel: the largest dynamic shared-memory size opted into so far
    if (attr_done[(const void*)kern] < c->stage_smem) {
        CK(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->stage_smem));
        attr_done[(const void*)kern] = c->stage_smem;
    }
    static int nsm = 0;
    if (!nsm) CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device));
    const int grid = std::min(c->ntiles, nsm), block = (c->tile_ncw + 4) * 32;
    cudaEvent_t _a = nullptr, _b = nullptr;
    TRY(prof_begin(c, c->st, K_STAGE, &_a, &_b));
    kern<<<grid, block, c->stage_smem, c->st>>>(G, A);
    CK(cudaGetLastError());
    TRY(prof_end(c, c->st, K_STAGE, _a, _b));
    // tile-boundary nodes (and nodes no element touches: an empty run of records)
    if (c->nbnodes)
        LAUNCH(K_NODE, k::boundary_update, grid_for(c->nbnodes, CFDB_NODE_BS), CFDB_NODE_BS, c->nbnodes, c->bnodes.p, c->bn_ptr.p, c->EC.p, c->U.p,
               c->M.p, c->GAMM.col, c->W_X.p, c->W_Y.p, c->bcflag.p, bctab(c), rk_fact, c->par.FR, c->U1.p, c->RHS.p, c->RHO.col, c->VEL_X.col,
               c->VEL_Y.col, c->E.col, c->P.col, c->T.col, c->RMACH.col);
    return 0;
}

extern "C" int cfdb_rk_stage(cfdb_ctx* c, int32_t irk) {
    CK(cudaSetDevice(c->device));
    const cfdb_params& p = c->par;
    const int NRK = 4;
    if (irk < 1 || irk > NRK) return fail("cfdb_rk_stage: irk must be 1..4");
    double RK_FACT = 1.0 / (NRK + 1 - irk);
    if (irk == 1) {
        // cuarto_orden's projection is discarded by UN = 0.0 (subrutinas.f90:673-674) unless use_cuarto
        if (c->use_cuarto) {
            // cuarto_orden(U1, UN, ...): the loop has just copied U1 = U (ns2DComp.ALE.f90:168-172; the copy itself is
            // elided here because every stage rewrites U1), so the projection is evaluated at U
            LAUNCH(K_CALCRHS, k::cuarto_elem, grid_for(c->nelem, 128), 128, c->nelem, c->inp.p, c->U.p, c->GAMM.col, c->dNx.p,
                   c->dNy.p, c->area.p, c->EC.p);
            LAUNCH(K_NODE, k::cuarto_node, grid_for(c->npoin, 256), 256, c->npoin, c->d_esup2.p, c->eslot.p, c->EC.p, c->M.p, c->UN.p);
            TRY(halo_vec(c, c->UN.p, 4));
            c->theta_nonzero = true;
        } else if (c->theta_nonzero) {
            CK(cudaMemsetAsync(c->UN.p, 0, 4 * (size_t)c->npoin * sizeof(double), c->st));
            c->theta_nonzero = false;
        }
        TRY(run_estab(c, &c->sc->DTMIN));
    }
    c->u1_is_u = false;
    c->Usrc = (c->true_rk && irk > 1) ? c->U1.p : nullptr;
    struct UsrcReset { cfdb_ctx* c; ~UsrcReset() { c->Usrc = nullptr; } } usrc_reset{c};
    k::Gas g{p.FCv, p.FK, p.FMU, p.GAMA, p.T_inf, p.CTE};
    const double* dtl_arr = (p.ITLOCAL != 0 || c->dtl_force) ? c->DTL.p : nullptr;
    if (c->fast == 2 && !c->ale && !c->use_cuarto && !c->true_rk) {
        // measurement only: the staged element kernel compiled with FMA contraction + the exact ordered node kernel
        const bool visc = g.mu_ref > 2.2250738585072014e-308;
        cudaEvent_t _a = nullptr, _b = nullptr;
        TRY(prof_begin(c, c->st, K_CALCRHS, &_a, &_b));
        if (fastmode::launch_calcrhs_staged_fma(visc, c->st, c->nelem, c->inp.p, c->U.p, c->T.col.q, c->dNx.p, c->dNy.p, c->area.p,
                                                c->SHOC.p, dtl_arr, &c->sc->DTMIN, c->TS1.p, c->TS2.p, c->TS3.p, g.Cv, g.lambda_ref,
                                                g.mu_ref, g.gamma0, g.T_inf, g.cte, c->EC.p))
            return fail("calcrhs_staged_fma launch failed");
        TRY(prof_end(c, c->st, K_CALCRHS, _a, _b));
        TRY(run_node(c, c->st, false, true, RK_FACT));
        TRY(halo_state(c));
        return 0;
    }
    if (c->colored && !c->use_cuarto && !c->true_rk) {
        // coloured deterministic scatter: RHS = 0, one launch per colour (no two elements of a colour share a node), nodal chain
        // from RHS; exact arithmetic, fixed summation order (by colour) -- reproducible, not the reference's order
        const bool visc = g.mu_ref > 2.2250738585072014e-308;
        CK(cudaMemsetAsync(c->RHS.p, 0, 4 * (size_t)c->npoin * sizeof(double), c->st));
        auto kc = visc ? (c->ale ? k::calcrhs_colored<true, true> : k::calcrhs_colored<true, false>)
                       : (c->ale ? k::calcrhs_colored<false, true> : k::calcrhs_colored<false, false>);
        for (size_t col = 0; col + 1 < c->color_ptr.size(); ++col) {
            const int n0 = c->color_ptr[col], nc = c->color_ptr[col + 1] - n0;
            if (nc)
                LAUNCH(K_CALCRHS, kc, grid_for(nc, 128), 128, nc, c->color_list.p + n0, c->nelem, c->inp.p, c->U.p, c->T.col, c->W_X.p, c->W_Y.p,
                       c->dNx.p, c->dNy.p, c->area.p, c->SHOC.p, dtl_arr, &c->sc->DTMIN, c->TS1.p, c->TS2.p, c->TS3.p, g, c->RHS.p);
        }
        LAUNCH(K_NODE, k::node_update_rhs, grid_for(c->npoin, 256), 256, c->npoin, c->RHS.p, c->U.p, c->M.p, c->GAMM.col, c->W_X.p, c->W_Y.p,
               c->bcflag.p, bctab(c), RK_FACT, c->par.FR, c->U1.p, c->RHO.col, c->VEL_X.col, c->VEL_Y.col, c->E.col, c->P.col, c->T.col, c->RMACH.col);
        TRY(halo_state(c));
        return 0;
    }
    if (c->fast && !c->use_cuarto && !c->true_rk) {
        // relaxed stage: RHS = 0, scatter-add with red.global.add.f64, nodal chain from RHS
        const bool visc = g.mu_ref > 2.2250738585072014e-308;
        CK(cudaMemsetAsync(c->RHS.p, 0, 4 * (size_t)c->npoin * sizeof(double), c->st));
        cudaEvent_t _a = nullptr, _b = nullptr;
        TRY(prof_begin(c, c->st, K_CALCRHS, &_a, &_b));
        if (fastmode::launch_calcrhs_scatter(visc, c->ale, c->st, c->nelem, c->inp.p, c->U.p, c->T.col.q, c->W_X.p, c->W_Y.p, c->dNx.p,
                                             c->dNy.p, c->area.p, c->SHOC.p, dtl_arr, &c->sc->DTMIN, c->TS1.p, c->TS2.p, c->TS3.p,
                                             g.Cv, g.lambda_ref, g.mu_ref, g.gamma0, g.T_inf, g.cte, c->RHS.p))
            return fail("calcrhs_scatter launch failed");
        TRY(prof_end(c, c->st, K_CALCRHS, _a, _b));
        TRY(prof_begin(c, c->st, K_NODE, &_a, &_b));
        k::BcTab b = bctab(c);
        if (fastmode::launch_node_update_rhs(c->st, c->npoin, c->RHS.p, c->U.p, c->M.p, c->GAMM.col.q, c->W_X.p, c->W_Y.p, c->bcflag.p,
                                             b.nb, b.node, b.kind, b.vx, b.vy, b.rho, b.Tfix, b.wslot, b.wn_x, b.wn_y, b.wn_valid,
                                             RK_FACT, c->par.FR, c->U1.p, c->RHO.col.q, c->VEL_X.col.q, c->VEL_Y.col.q, c->E.col.q, c->P.col.q, c->T.col.q,
                                             c->RMACH.col.q))
            return fail("node_update_rhs launch failed");
        TRY(prof_end(c, c->st, K_NODE, _a, _b));
        TRY(halo_state(c));
        return 0;
    }
    if (fused_eligible(c)) {
        TRY(refresh_geo(c));
        TRY(run_stage_fused(c, g, dtl_arr, &c->sc->DTMIN, RK_FACT));
        if (!c->skip_stage_halo) TRY(halo_state(c));
        return 0;
    }
    TRY(run_calcrhs_elem(c, g, c->use_cuarto != 0, c->ale, dtl_arr, &c->sc->DTMIN));
    TRY(run_node(c, c->st, c->ale, true, RK_FACT));
    TRY(halo_state(c));
    return 0;
}

// ADAMSB(DTMIN, NESTAB, GAMM, dtl), subrutinas.f90:851-1034, on the resident state (option "adamsb"): CUARTO_ORDEN's projection
// and ESTAB on every third call (NESTAB), ONE calcRHS + FUENTE with theta = UN, the Adams-Bashforth update from the RHS history.
static int run_adamsb(cfdb_ctx* c) {
    const cfdb_params& p = c->par;
    if (c->nestab == 4) c->nestab = 1;
    if (c->nestab == 2) {
        // cuarto_orden(U1, UN, ...) with U1 = U (ns2DComp.ALE.f90:168-172); no UN = 0.0 afterwards in this routine
        LAUNCH(K_CALCRHS, k::cuarto_elem, grid_for(c->nelem, 128), 128, c->nelem, c->inp.p, c->U.p, c->GAMM.col, c->dNx.p, c->dNy.p,
               c->area.p, c->EC.p);
        LAUNCH(K_NODE, k::cuarto_node, grid_for(c->npoin, 256), 256, c->npoin, c->d_esup2.p, c->eslot.p, c->EC.p, c->M.p, c->UN.p);
        TRY(halo_vec(c, c->UN.p, 4));
        c->theta_nonzero = true;
        TRY(run_estab(c, &c->sc->DTMIN));
    }
    c->nestab += 1;
    c->u1_is_u = false;
    k::Gas g{p.FCv, p.FK, p.FMU, p.GAMA, p.T_inf, p.CTE};
    const double* dtl_arr = p.ITLOCAL != 0 ? c->DTL.p : nullptr;
    TRY(run_calcrhs_elem(c, g, true, c->ale, dtl_arr, &c->sc->DTMIN));
    auto kern = c->ale ? k::node_update_adamsb<true> : k::node_update_adamsb<false>;
    LAUNCH(K_NODE, kern, grid_for(c->npoin, 128), 128, c->npoin, c->d_esup2.p, c->eslot.p, c->EC.p, c->FC.p, c->U.p, c->M.p, c->GAMM.col,
           c->W_X.p, c->W_Y.p, c->bcflag.p, bctab(c), p.FR, c->U1.p, c->RHS.p, c->RHS1.p, c->RHS2.p, c->RHS3.p, c->RHO.col, c->VEL_X.col,
           c->VEL_Y.col, c->E.col, c->P.col, c->T.col, c->RMACH.col);
    TRY(halo_state(c));
    // RHS1-3 at ghost nodes are never read (the history enters the update of owned nodes only)
    return 0;
}

// RK (subrutinas.f90:645-849) inside the time loop
static int run_rk(cfdb_ctx* c) {
    // Multi-rank, fixed mesh, Euler flow, reference options: every stage evaluates calcRHS at U and reads
    // nothing a stage writes (T only enters the viscous terms), so the ghost copies of U1, T, ... are first needed after the
    // LAST stage.  The ghost refresh after stages 1-3 -- communication this implementation added, not reference work --
    // is then skipped; the one after stage 4 delivers the same ghost values as before.
    const bool defer = c->nranks > 1 && fused_eligible(c) && !c->true_rk;
    for (int irk = 1; irk <= 4; ++irk) {
        c->skip_stage_halo = defer && irk < 4;
        int rc = cfdb_rk_stage(c, irk);
        c->skip_stage_halo = false;
        if (rc) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// biCG on device arrays (biconjGrad.f90:8-62).  The early return (:35), the loop condition (:47) and every scalar of the
// algorithm live on the device; the host only enqueues.  The prologue (:37-45) is enqueued unconditionally -- if r.r < tol
// its results are work-array garbage that nothing applies -- and the iterations follow in batches of 1, 2, 4, 8, 16, 16, ...
// each closed by a flush of the pending x update and ONE read-back of the loop state (most mesh solves need a handful of
// iterations because the tolerance is absolute).  static_zero: the caller knows that x, x_fix and b are all zero (a fixed
// mesh: r.r = 0 < tol, the reference returns at :35 after its SpMV and inner product): nothing after r.r is enqueued and
// nothing is read back.
static int bicg_dev(cfdb_ctx* c, const double* A, const int* idx, const int* rowptr, const double* diag, double* x,
                    const double* b, const double* x_fix, const int* fixIdx, const int* fixLast, int npoin, int nfix,
                    int* iters, double* y_pre = nullptr, bool static_zero = false) {
    const int B = 256, G = grid_for(npoin, B), GF = grid_for(std::max(nfix, 1), 128);
    // y_pre: x already carries its Dirichlet values and y_pre = A*x (fluid_structure computes both solves' first
    // products in one pass over the matrix); it then serves as this solve's y work array
    double *y = y_pre ? y_pre : c->by.p, *p = c->bp.p, *r = c->br.p, *z = c->bz.p;
    const int nred = (c->nranks > 1 && npoin == c->npoin) ? c->n_owned : npoin;  // inner products over owned nodes
    if (!y_pre) {
        if (nfix) LAUNCH(K_FIXROWS, k::copy1, GF, 128, nfix, fixIdx, fixLast, 1.0, x_fix, x);
        TRY(halo_vec(c, x, 1));
        LAUNCH(K_SPMV, k::spmv, G, B, npoin, A, idx, rowptr, x, y);
    }
    if (nfix) LAUNCH(K_FIXROWS, k::copy2, GF, 128, nfix, fixIdx, 1.e30, x, y);
    LAUNCH(K_VEC, k::vecsum, G, B, npoin, (int)k::ALFA_CONST, -1.0, c->sc, y, b, r);
    if (nfix) LAUNCH(K_FIXROWS, k::assign2, GF, 128, nfix, fixIdx, 0.0, r);
    TRY(dev_dot(c, nred, r, r, 0));
    LAUNCH(K_SCALAR, k::bicg_scalar, 1, 1, c->sc, (int)k::SC_RR, 0);
    if (static_zero) { *iters = -1; return 0; }   // :35 is known to return; the ghosts of x (all zero) need no refresh
    LAUNCH(K_VEC, k::vecdiv, G, B, npoin, r, diag, p);
    TRY(dev_dot(c, nred, r, p, 0));
    LAUNCH(K_SCALAR, k::bicg_scalar, 1, 1, c->sc, (int)k::SC_ERRNEW, 0);
    TRY(halo_vec(c, p, 1));
    LAUNCH(K_SPMV, k::spmv, G, B, npoin, A, idx, rowptr, p, y);
    if (nfix) LAUNCH(K_FIXROWS, k::copy2, GF, 128, nfix, fixIdx, 1.e30, p, y);
    TRY(dev_dot(c, nred, p, y, 1));
    LAUNCH(K_SCALAR, k::bicg_scalar, 1, 1, c->sc, (int)k::SC_PY_ALFA, 1);
    // while loop (:47-61): fused iterations; x = alfa*p + x (:44, :59) is applied by the next bicg_k1 or by bicg_flush
    TRY(c->isfix.alloc((size_t)c->npoin > (size_t)npoin ? c->npoin : npoin));
    CK(cudaMemsetAsync(c->isfix.p, 0, npoin, c->st));
    if (nfix) LAUNCH(K_FIXROWS, k::mark_fixed, GF, 128, nfix, fixIdx, c->isfix.p);
    LAUNCH(K_SCALAR, k::bicg_fused_scalar, 1, 1, c->sc, (int)k::SCF_START, 0);
    const long nch = ((long)npoin + 4095) / 4096, mred = ((long)nred + 4095) / 4096;
    const int GB = (int)std::min<long>(nch, 148 * 8);
    double* pa = p;         // current p
    double* pb = c->bp2.p;  // next p
    // the first read-back follows the prologue alone: most steps of a slowly moving mesh end there
    for (int done = 0, batch = 0; done <= 1000; batch = batch ? std::min(2 * batch, 16) : 1) {
        for (int it = 0; it < batch; ++it, ++done) {
            LAUNCH(K_VEC, k::bicg_k1, GB, 256, npoin, nred, c->sc, y, diag, pa, x, r, z, c->redA.p);
            TRY(reduce_levels(c, 1, mred, 0));
            LAUNCH(K_SCALAR, k::bicg_fused_scalar, 1, 1, c->sc, (int)k::SCF_BETA, 0);
            TRY(halo_vec(c, z, 1));
            LAUNCH(K_SPMV, k::bicg_k2, GB, 256, npoin, nred, c->sc, A, idx, rowptr, c->isfix.p, pa, z, pb, y, c->redA.p);
            TRY(reduce_levels(c, 1, mred, 1));
            LAUNCH(K_SCALAR, k::bicg_fused_scalar, 1, 1, c->sc, (int)k::SCF_ALFA, 1);
            std::swap(pa, pb);
        }
        LAUNCH(K_VEC, k::bicg_flush, GB, 256, npoin, c->sc, pa, x);
        LAUNCH(K_SCALAR, k::bicg_flushed, 1, 1, c->sc);
        TRY(read_scal(c));
        if (!c->h_sc->bicg_state) break;
    }
    *iters = c->h_sc->rr < 1.e-10 ? -1 : c->h_sc->bicg_k;
    TRY(halo_vec(c, x, 1));  // ghosts of the solution (the caller moves ghost nodes with it)
    return 0;
}

extern "C" int cfdb_fluid_structure(cfdb_ctx* c, double dtmin, double time) {
    CK(cudaSetDevice(c->device));
    const int P = c->npoin;
    (void)dtmin;  // W = XPOS/DTMIN uses the device-resident DTMIN (same value)
    if (c->nse) {  // DXPOS = DYPOS = 0 (meshMove.f90:51-52): without body sets nothing ever writes them (zeroed at create)
        CK(cudaMemsetAsync(c->dxpos.p, 0, P * sizeof(double), c->st));
        CK(cudaMemsetAsync(c->dypos.p, 0, P * sizeof(double), c->st));
    }
    // XREF(2)=1.4, YREF(2)=0 are overwritten on every call (meshMove.f90:58): applied once in cfdb_create
    if (c->nset || (c->nranks > 1 && c->ale)) {  // every rank of a moving-mesh run joins the all-reduce, with or without body edges of its own
        CK(cudaMemsetAsync(c->sc->FX, 0, 30 * sizeof(double), c->st));
        if (c->nset) LAUNCH(K_FORCES, k::forces, c->nset, 256, c->nset, c->n_owned, c->set_ptr.p, c->set_n1.p, c->set_n2.p, c->X.p, c->Y.p, c->P.col,
               c->xref.p, c->yref.p, c->sc);
        TRY(allreduce(c, c->sc->FX, 30, ncclSum));  // FX,FY,RM are contiguous in Scal
    }
    double PI = std::acos(-1.0);
    double AMPLI = PI / 8.0;
    double ALPHAV = c->DISN[1], YPOSRV = c->DISN[0];
    c->DISN[1] = AMPLI * std::sin(10.0 * time);  // :70
    double ALPHA = c->DISN[1] - ALPHAV, YPOSR = c->DISN[0] - YPOSRV;
    if (c->nse)
        LAUNCH(K_TRANSF, k::transf, grid_for(c->nse, 128), 128, c->nse, c->se_node.p, c->se_set.p, std::cos(ALPHA),
               std::sin(ALPHA), YPOSR, c->xref.p, c->yref.p, c->X.p, c->Y.p, c->dxpos.p, c->dypos.p);
    // The two solves (meshMove.f90:97, :119) are independent until their results are applied: the y-solve's Dirichlet
    // values, warm start and first product A*ypos do not depend on the x-solve.  Both first products are therefore taken
    // in ONE pass over the matrix (k::spmv2: same row-sequential sums, 12 B/nnz read once instead of twice).
    // Fixed mesh (no body sets on any rank, xpos/ypos/W never set by the caller): DXPOS = 0, the warm start is 0, so both
    // solves return at biconjGrad.f90:35 -- the host knows it and neither enqueues the rest of biCG nor reads anything back;
    // what the reference executes on that path (Dirichlet rows, SpMV, residual, r.r, the move with XPOS = 0) still runs.
    const bool static_zero = !c->ale;
    const int GF = grid_for(std::max(c->nnmove, 1), 128);
    if (c->nnmove) {
        LAUNCH(K_MOVE, k::pos_aux_fill, GF, 128, c->nmove, c->nnmove, c->ilaux.p, c->dxpos.p, c->pos_aux.p);
        LAUNCH(K_MOVE, k::pos_aux_fill, GF, 128, c->nmove, c->nnmove, c->ilaux.p, c->dypos.p, c->pos_aux2.p);
        LAUNCH(K_FIXROWS, k::copy1, GF, 128, c->nnmove, c->ilaux.p, c->ilaux_last.p, 1.0, c->pos_aux.p, c->xpos.p);
        LAUNCH(K_FIXROWS, k::copy1, GF, 128, c->nnmove, c->ilaux.p, c->ilaux_last.p, 1.0, c->pos_aux2.p, c->ypos.p);
    }
    if (!static_zero) {
        TRY(halo_vec(c, c->xpos.p, 1));
        TRY(halo_vec(c, c->ypos.p, 1));
    }
    LAUNCH(K_SPMV, k::spmv2, grid_for(P, 256), 256, P, c->lap_sparse.p, c->d_lap_idx.p, c->d_lap_rowptr.p, c->xpos.p, c->ypos.p,
           c->by.p, c->by2.p);
    for (int dir = 0; dir < 2; ++dir) {
        double* pos = dir ? c->ypos.p : c->xpos.p;
        double* paux = dir ? c->pos_aux2.p : c->pos_aux.p;
        // B = 0 (meshMove.f90:91-95, :113-117): bb is zeroed at create and never written by anything else
        TRY(bicg_dev(c, c->lap_sparse.p, c->d_lap_idx.p, c->d_lap_rowptr.p, c->lap_diag.p, pos, c->bb.p, paux,
                     c->ilaux.p, c->ilaux_last.p, P, c->nnmove, &c->bicg_iters[dir], dir ? c->by2.p : c->by.p, static_zero));
        LAUNCH(K_MOVE, k::move_apply, grid_for(P, 256), 256, P, pos, &c->sc->DTMIN, dir ? c->Y.p : c->X.p,
               dir ? c->Y1.p : c->X1.p, dir ? c->W_Y.p : c->W_X.p);
    }
    return 0;
}

static int run_norms(cfdb_ctx* c) {
    long m = ((long)c->n_owned + 4095) / 4096;
    LAUNCH(K_NORMS, k::norm_chunks, (int)std::min<long>(m, 148 * 8), 256, (long)c->n_owned, c->U.p,
           c->u1_is_u ? c->U.p : c->U1.p, c->redA.p);
    TRY(reduce_levels(c, 8, m, 0));
    CK(cudaMemcpyAsync(c->sc->ER, c->sc->red, 8 * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    return 0;
}
extern "C" int cfdb_residual_norms(cfdb_ctx* c, double er[4], double err[4]) {
    CK(cudaSetDevice(c->device));
    TRY(run_norms(c));
    TRY(read_scal(c));
    for (int i = 0; i < 4; ++i) { er[i] = c->h_sc->ER[i]; err[i] = c->h_sc->ERR[i]; }
    return 0;
}

extern "C" int cfdb_selftest(cfdb_ctx* c, int32_t which, int64_t n, uint64_t seed, int64_t* mismatches) {
    CK(cudaSetDevice(c->device));
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), c->st));
    LAUNCH(K_FILL, k::selftest, 148 * 8, 256, (int)which, (long)n, (unsigned long long)seed, d);
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    cudaFree(d);
    *mismatches = (int64_t)h;
    return 0;
}

// FORCE_VISC (ns2DComp.ALE.f90:819-893) on the current device state; results in the fields F_VX, F_VY, skin, skin_x, skin_p
extern "C" int cfdb_force_visc(cfdb_ctx* c) {
    CK(cudaSetDevice(c->device));
    const cfdb_params& p = c->par;
    CK(cudaMemsetAsync(c->fvisc.p, 0, 20 * sizeof(double), c->st));  // F_VX = 0; F_VY = 0 (:832)
    if (c->nset)
        LAUNCH(K_FORCES, k::force_visc, 1, 32, c->nset, c->n_owned, c->set_ptr.p, c->set_n1.p, c->set_n2.p, c->set_el.p, c->nelem,
               c->inp.p, c->X.p, c->Y.p, c->P.col, c->T.col, c->VEL_X.col, c->VEL_Y.col, c->dNx.p, c->dNy.p, p.U_inf, p.V_inf, p.RHO_inf,
               p.T_inf, c->fvisc.p, c->skin.p, c->nedges);
    if (c->nranks > 1) TRY(allreduce(c, c->fvisc.p, 20, ncclSum));
    return 0;
}

// one real in Fortran Ew.d ('E') or Fw.d ('F') layout (host/fortran_format.h), for tests of the output formats
extern "C" int cfdb_format_real(int32_t kind, double v, int32_t w, int32_t d, char* buf, int32_t buflen) {
    std::string s = kind == 'L' ? ffmt::list_r8(v) : kind == 'F' ? ffmt::F(v, w, d) : ffmt::E(v, w, d);   // 'L': list-directed REAL(8)
    if ((int)s.size() + 1 > buflen) return fail("cfdb_format_real: buffer too small");
    memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}
extern "C" int cfdb_format_cnv(int32_t iter, double time, const double r[4], char* buf, int32_t buflen) {
    std::string s = ffmt::cnv_record(iter, time, r);
    if ((int)s.size() + 1 > buflen) return fail("cfdb_format_cnv: buffer too small");
    memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}

// PRINTFLAVIA (ns2DComp.ALE.f90:701-817) with the arguments of its call site (:225-226): velocities relative to the mesh,
// X1/Y1 as positions.  flags: RHO, VEL2, MACH, PRES, TEMP, ENER, POS ('.si.' in <name>-1.dat).
extern "C" int cfdb_printflavia(cfdb_ctx* c, const char* path, int32_t iter, const int32_t flags[7], int32_t append) {
    CK(cudaSetDevice(c->device));
    const size_t P = c->npoin;
    vector<double> rho(P), vx(P), vy(P), wx(P), wy(P), pr(P), tt(P), en(P), gm(P), x1(P), y1(P);
    struct { double* h; const double* d; } cp[] = {{wx.data(), c->W_X.p}, {wy.data(), c->W_Y.p}, {x1.data(), c->X1.p}, {y1.data(), c->Y1.p}};
    for (auto& q : cp) CK(cudaMemcpyAsync(q.h, q.d, P * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    struct { double* h; NodeField f; } cn[] = {{rho.data(), c->RHO}, {vx.data(), c->VEL_X}, {vy.data(), c->VEL_Y}, {pr.data(), c->P},
        {tt.data(), c->T}, {en.data(), c->E}, {gm.data(), c->GAMM}};
    for (auto& q : cn) TRY(down_node(c, q.f, q.h, P));
    CK(cudaStreamSynchronize(c->st));
    for (size_t i = 0; i < P; ++i) { vx[i] = vx[i] - wx[i]; vy[i] = vy[i] - wy[i]; }   // VEL_X - W_X, VEL_Y - W_Y (:225)
    std::FILE* f = std::fopen(path, append ? "a" : "w");
    if (!f) return fail(std::string("cfdb_printflavia: cannot open ") + path);
    auto header = [&](const char* name, int ncomp) {
        std::string h = ffmt::A(name, 15);
        const int v[5] = {2, iter, ncomp, 1, 1};
        for (int k = 0; k < 5; ++k) h += ffmt::I(v[k], 8) + "  ";
        std::fprintf(f, "%s\n", h.c_str());
    };
    if (flags[1]) {
        header("VELOCITY", 2);
        std::fprintf(f, "VEL_X\nVEL_Y\n");
        for (size_t i = 0; i < P; ++i)
            std::fprintf(f, "%s%s%s\n", ffmt::I((long)i + 1, 8).c_str(), ffmt::E(vx[i], 13, 4).c_str(), ffmt::E(vy[i], 13, 4).c_str());
    }
    if (flags[6]) {
        header("POSITION", 2);
        std::fprintf(f, "X\nY\n");
        for (size_t i = 0; i < P; ++i)
            std::fprintf(f, "%s%s%s\n", ffmt::I((long)i + 1, 8).c_str(), ffmt::E(x1[i], 16, 6).c_str(), ffmt::E(y1[i], 16, 6).c_str());
    }
    auto scalar = [&](const char* name, const vector<double>& a, int w, int d) {
        header(name, 1);
        std::fprintf(f, "%s\n", name);
        for (size_t i = 0; i < P; ++i) std::fprintf(f, "%s%s\n", ffmt::I((long)i + 1, 8).c_str(), ffmt::E(a[i], w, d).c_str());
    };
    if (flags[0]) scalar("DENSITY", rho, 13, 4);
    if (flags[3]) scalar("PRESSURE", pr, 16, 3);
    if (flags[4]) scalar("TEMPERATURE", tt, 13, 3);
    if (flags[2]) {
        header("Mach_Number", 1);
        std::fprintf(f, "Mach_Number\n");
        for (size_t i = 0; i < P; ++i) {
            double VEL = std::sqrt(vx[i] * vx[i] + vy[i] * vy[i]);
            double VC = std::sqrt(gm[i] * c->par.FR * tt[i]);
            std::fprintf(f, "%s%s\n", ffmt::I((long)i + 1, 8).c_str(), ffmt::F(VEL / VC, 11, 2).c_str());
        }
    }
    if (flags[5]) scalar("Internal_Energy", en, 16, 8);
    std::fclose(f);
    return 0;
}

extern "C" int cfdb_step_norms(cfdb_ctx* c, double er[4], double err[4]) {
    CK(cudaSetDevice(c->device));
    TRY(read_scal(c));
    for (int i = 0; i < 4; ++i) { er[i] = c->h_sc->ER[i]; err[i] = c->h_sc->ERR[i]; }
    return 0;
}

// Multi-rank: whether the mesh can move is a GLOBAL fact (the body sets usually sit on one rank, the mesh solve moves every
// rank's nodes).  The ranks agree on it once, before the first step: after that `ale` is the same everywhere, and a
// fixed-mesh multi-GPU run uses the same W = 0 kernels as a single-GPU one.
static int agree_on_ale(cfdb_ctx* c) {
    if (c->nranks <= 1 || c->ale_agreed) return 0;
    LAUNCH(K_SCALAR, k::set_double, 1, 1, &c->sc->red[14], c->ale ? 1.0 : 0.0);
    TRY(allreduce(c, &c->sc->red[14], 1, ncclMax));
    TRY(read_scal(c));
    if (c->h_sc->red[14] != 0.0 && !c->ale) { c->ale = true; TRY(zero(c, c->FC, 12 * (size_t)c->nelem)); }
    c->ale_agreed = true;
    c->epoch++;
    return 0;
}

// everything one pass of ns2DComp.ALE.f90:138-282 enqueues, except the print-step work and the U = U1 pointer swap
static int step_body(cfdb_ctx* c) {
    const cfdb_params& p = c->par;
    const int E = c->nelem;
    const size_t P = c->npoin;
    LAUNCH(K_DTLOGIC, k::step_begin, 1, 1, c->sc);
    {
        const bool moving = c->ale;  // as in run_estab
        auto kdt = p.ITLOCAL != 0 ? (moving ? k::deltat<true, true> : k::deltat<true, false>)
                                  : (moving ? k::deltat<false, true> : k::deltat<false, false>);
        LAUNCH(K_DELTAT, kdt, grid_for(E, 256), 256, E, c->inp.p, c->area.p, c->T.col, c->VEL_X.col, c->VEL_Y.col,
               c->W_X.p, c->W_Y.p, p.FSAFE, p.T_inf, c->DT.p, c->sc);
    }
    TRY(allreduce(c, &c->sc->dtmin_acc, 1, ncclMin));
    LAUNCH(K_DTLOGIC, k::dt_logic, 1, 1, c->sc);
    if (p.ITLOCAL != 0) {
        double DTFACT = 1.0 - std::exp(-c->h_iter * 4.6 / p.ITLOCAL);  // :160
        LAUNCH(K_SCALAR, k::set_double, 1, 1, &c->sc->dtfact, DTFACT);
        LAUNCH(K_DTL, k::dtl_blend, grid_for(E, 256), 256, E, c->DT.p, c->DTL.p, c->sc, 1);
    }
    // U1 = U (:168-172) is dead: every RK stage overwrites U1 from U (subrutinas.f90:697)
    bool adams = false;
    if (c->adamsb) {   // `if (BANDERA.LE.4) RK else ADAMSB` (:174-178 as its comments intend); BANDERA lives on the device
        TRY(read_scal(c));
        adams = c->h_sc->BANDERA > 4;
    }
    if (adams) {
        TRY(run_adamsb(c));
    } else {
        TRY(run_rk(c));
        // RHS history copies for BANDERA 2..4 (subrutinas.f90:830-848); BANDERA lives on the device, the kernel
        // exits at once for any other value
        LAUNCH(K_FILL, k::rhs_history, (int)std::min<long>(grid_for(4 * (long)P, 256), 148 * 8), 256, 4 * (long)P, c->sc, c->RHS.p, c->RHS1.p,
               c->RHS2.p, c->RHS3.p);
    }
    double dtmin = 0.0, time = 0.0;
    if (c->nse) {  // pitching law needs TIME on the host (meshMove.f90:70)
        TRY(read_scal(c));
        dtmin = c->h_sc->DTMIN;
        time = c->h_sc->TIME;
    }
    TRY(cfdb_fluid_structure(c, dtmin, time));
    return 0;
}

// A fixed-mesh step enqueues ~35 small-to-large launches and never needs the host (no body sets: no TIME read-back, the
// mesh solve is known to return early; uniform time step: no host-computed blend factor).  It is captured ONCE per state
// buffer parity (U and U1 swap roles every step) into a CUDA graph and replayed: one launch per step instead of ~35, which
// removes the launch gaps between the kernels (0.5 ms of an 8.2 ms step in round 1).  NCCL's ghost refresh and all-reduces
// are captured with it.  CFDB_NO_GRAPH=1 keeps stream launches (tests compare the two).
static bool graph_eligible(const cfdb_ctx* c) {
    static const bool off = getenv("CFDB_NO_GRAPH") != nullptr;
    const cfdb_params& p = c->par;
    return !off && !c->prof && !c->ale && c->nse == 0 && p.ITLOCAL == 0 && p.MOVING != 1 && !c->theta_nonzero && !c->use_cuarto && !c->adamsb;
}
static int step_graph(cfdb_ctx* c) {
    int par = -1;
    for (int i = 0; i < 2; ++i)
        if (c->gexec[i] && c->gepoch[i] == c->epoch && c->gU[i] == c->U.p) par = i;
    if (par < 0) {
        par = (c->gexec[0] && c->gepoch[0] == c->epoch) ? 1 : 0;
        if (c->gexec[par]) { cudaGraphExecDestroy(c->gexec[par]); c->gexec[par] = nullptr; }
        const int64_t l0 = c->launches;
        CK(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
        int rc = step_body(c);
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->st, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
        e = cudaGraphInstantiate(&c->gexec[par], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        c->glaunches[par] = c->launches - l0;
        c->launches = l0;
        c->gepoch[par] = c->epoch;
        c->gU[par] = c->U.p;
    }
    CK(cudaGraphLaunch(c->gexec[par], c->st));
    c->launches += c->glaunches[par];
    c->graph_replays++;
    return 0;
}

// one pass of ns2DComp.ALE.f90:138-282
static int step_once(cfdb_ctx* c) {
    const cfdb_params& p = c->par;
    TRY(agree_on_ale(c));
    if (fused_eligible(c)) TRY(refresh_geo(c));   // outside the captured step
    c->h_iter += 1;
    if (graph_eligible(c)) TRY(step_graph(c));
    else TRY(step_body(c));
    c->u1_is_u = false;   // the RK stages have rewritten U1 (a replayed graph does not run cfdb_rk_stage's host code)
    c->iterprint += 1;
    if (c->iterprint == p.IPRINT || c->h_iter == p.MAXITER) {  // :186-197
        TRY(run_norms(c));
        if (p.FMU != 0.0) TRY(cfdb_force_visc(c));  // :228-233
        c->iterprint = 0;
    }
    LAUNCH(K_SCALAR, k::bandera_inc, 1, 1, c->sc);
    if (p.MOVING == 1) TRY(cfdb_geometry(c, 1));
    // U = U1 (:277-281): swap the buffers instead of copying
    std::swap(c->U.p, c->U1.p);
    c->u1_is_u = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Streamed stepping: the state of every step comes from host arrays and its results go back to host arrays, as a literal
// "the host program owns the arrays" integration needs, but pipelined.  Three streams: st_in uploads the inputs of call
// k+1 into device staging while the compute stream runs step k; st_out downloads the results of step k from a second
// staging area while step k+1 runs.  Events order the hand-overs; one staging area per direction is enough because the
// copy out of (into) staging is a device-to-device copy at the very start (end) of the step.  PCIe carries both
// directions at once, so a step costs max(upload, step, download) instead of their sum.
static int streamed_setup(cfdb_ctx* c) {
    if (c->st_in) return 0;
    const size_t P = c->npoin;
    CK(cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking));
    for (auto* e : {&c->ev_in_done, &c->ev_step_done, &c->ev_in_used[0], &c->ev_in_used[1], &c->ev_out_done[0], &c->ev_out_done[1]})
        CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (int s = 0; s < 2; ++s) {
        TRY(c->sin[s].alloc(7 * P));
        TRY(c->sout[s].alloc(7 * P + 8));
    }
    return 0;
}
extern "C" int cfdb_step_streamed(cfdb_ctx* c, const double* in_U, const double* in_T, const double* in_VEL_X, const double* in_VEL_Y,
                                  double* out_U, double* out_T, double* out_VEL_X, double* out_VEL_Y, double* out_norms) {
    CK(cudaSetDevice(c->device));
    TRY(streamed_setup(c));
    const size_t P = c->npoin, B = sizeof(double);
    const int set = (int)(c->streamed_calls & 1);
    const bool reuse = c->streamed_calls >= 2;     // this set was used two calls ago
    double *si = c->sin[set].p, *so = c->sout[set].p;
    // 1. upload into this call's staging set (waits until the call before last has consumed it)
    if (reuse) CK(cudaStreamWaitEvent(c->st_in, c->ev_in_used[set], 0));
    if (in_U) CK(cudaMemcpyAsync(si, in_U, 4 * P * B, cudaMemcpyHostToDevice, c->st_in));
    if (in_T) CK(cudaMemcpyAsync(si + 4 * P, in_T, P * B, cudaMemcpyHostToDevice, c->st_in));
    if (in_VEL_X) CK(cudaMemcpyAsync(si + 5 * P, in_VEL_X, P * B, cudaMemcpyHostToDevice, c->st_in));
    if (in_VEL_Y) CK(cudaMemcpyAsync(si + 6 * P, in_VEL_Y, P * B, cudaMemcpyHostToDevice, c->st_in));
    CK(cudaEventRecord(c->ev_in_done, c->st_in));
    // 2. compute stream: staging -> state, the step, state -> staging
    CK(cudaStreamWaitEvent(c->st, c->ev_in_done, 0));
    if (in_U) CK(cudaMemcpyAsync(c->U.p, si, 4 * P * B, cudaMemcpyDeviceToDevice, c->st));
    if (in_T) TRY(col_from_plain(c, c->T, si + 4 * P, P));
    if (in_VEL_X) TRY(col_from_plain(c, c->VEL_X, si + 5 * P, P));
    if (in_VEL_Y) TRY(col_from_plain(c, c->VEL_Y, si + 6 * P, P));
    CK(cudaEventRecord(c->ev_in_used[set], c->st));
    TRY(step_once(c));
    if (reuse) CK(cudaStreamWaitEvent(c->st, c->ev_out_done[set], 0));   // the download of the call before last has left this set
    if (out_norms) {   // ER, ERR of this step (ns2DComp.ALE.f90:191-197): after the swap U1.p holds the state the step started from
        long m = ((long)c->n_owned + 4095) / 4096;
        LAUNCH(K_NORMS, k::norm_chunks, (int)std::min<long>(m, 148 * 8), 256, (long)c->n_owned, c->U1.p, c->U.p, c->redA.p);
        TRY(reduce_levels(c, 8, m, 0));
        CK(cudaMemcpyAsync(so + 7 * P, c->sc->red, 8 * B, cudaMemcpyDeviceToDevice, c->st));
    }
    if (out_U) CK(cudaMemcpyAsync(so, c->U.p, 4 * P * B, cudaMemcpyDeviceToDevice, c->st));
    if (out_T) TRY(col_to_plain(c, c->T, so + 4 * P, P));
    if (out_VEL_X) TRY(col_to_plain(c, c->VEL_X, so + 5 * P, P));
    if (out_VEL_Y) TRY(col_to_plain(c, c->VEL_Y, so + 6 * P, P));
    CK(cudaEventRecord(c->ev_step_done, c->st));
    // 3. download
    CK(cudaStreamWaitEvent(c->st_out, c->ev_step_done, 0));
    if (out_U) CK(cudaMemcpyAsync(out_U, so, 4 * P * B, cudaMemcpyDeviceToHost, c->st_out));
    if (out_T) CK(cudaMemcpyAsync(out_T, so + 4 * P, P * B, cudaMemcpyDeviceToHost, c->st_out));
    if (out_VEL_X) CK(cudaMemcpyAsync(out_VEL_X, so + 5 * P, P * B, cudaMemcpyDeviceToHost, c->st_out));
    if (out_VEL_Y) CK(cudaMemcpyAsync(out_VEL_Y, so + 6 * P, P * B, cudaMemcpyDeviceToHost, c->st_out));
    if (out_norms) CK(cudaMemcpyAsync(out_norms, so + 7 * P, 8 * B, cudaMemcpyDeviceToHost, c->st_out));
    CK(cudaEventRecord(c->ev_out_done[set], c->st_out));
    c->streamed_calls++;
    return 0;
}
extern "C" int cfdb_streamed_wait(cfdb_ctx* c) {
    CK(cudaSetDevice(c->device));
    if (c->st_in) CK(cudaStreamSynchronize(c->st_in));
    CK(cudaStreamSynchronize(c->st));
    if (c->st_out) CK(cudaStreamSynchronize(c->st_out));
    return 0;
}

extern "C" int cfdb_step(cfdb_ctx* c, int32_t nsteps) {
    CK(cudaSetDevice(c->device));
    for (int i = 0; i < nsteps; ++i) TRY(step_once(c));
    return 0;
}
extern "C" int cfdb_sync(cfdb_ctx* c) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st));
    TRY(prof_resolve(c));
    static const bool verbose_fb = getenv("CFDB_VERBOSE") != nullptr;
    if (verbose_fb) {   // elements / nodes recomputed in the plain form since the last call (kernels.cuh: g_fallbacks)
        unsigned long long fb[k::FB_COUNT] = {}, zero[k::FB_COUNT] = {};
        CK(cudaMemcpyFromSymbol(fb, k::g_fallbacks, sizeof fb));
        CK(cudaMemcpyToSymbol(k::g_fallbacks, zero, sizeof zero));
        if (fb[0] | fb[1] | fb[2] | fb[3] | fb[4])
            fprintf(stderr, "[fallbacks] plain-form recomputations since the last sync: estab %llu, deltat %llu, stage elements %llu, stage nodes %llu, calcrhs_elem %llu\n",
                    fb[k::FB_ESTAB], fb[k::FB_DELTAT], fb[k::FB_STAGE_ELEM], fb[k::FB_STAGE_NODE], fb[k::FB_CALCRHS]);
    }
    if (c->stage_stats) {   // CFDB_STAGE_STATS: where the cycles of the fused stage went (sums over the 148 CTAs)
        unsigned long long h[k::ST_COUNT];
        CK(cudaMemcpy(h, c->stage_stats, sizeof h, cudaMemcpyDeviceToHost));
        CK(cudaMemset(c->stage_stats, 0, sizeof h));
        const double tl = h[k::ST_TILES] ? (double)h[k::ST_TILES] : 1.0;
        fprintf(stderr, "[stage_fused] cycles per tile (compute warp 0): element %.0f  node %.0f  wait inputs %.0f  wait C free %.0f  wait C full %.0f"
                        " | loader waits: stream slot %.0f  static landed %.0f  A slot %.0f  (tiles %.0f)\n",
                h[k::ST_E] / tl, h[k::ST_N] / tl, h[k::ST_WAIT_IN] / tl, h[k::ST_WAIT_CE] / tl, h[k::ST_WAIT_CF] / tl, h[k::ST_LD_WB] / tl,
                h[k::ST_LD_WS] / tl, h[k::ST_LD_WA] / tl, tl);
        if (h[k::ST_LOOP_NS])
            fprintf(stderr, "[stage_fused] the reporting element warp's tile loop: %.0f cycles per tile (hand-over after the arithmetic: %.0f), SM clock %.3f GHz while the kernel ran\n",
                    h[k::ST_LOOP_CYC] / tl, h[k::ST_ARRIVE] / tl, (double)h[k::ST_LOOP_CYC] / (double)h[k::ST_LOOP_NS]);
    }
    return 0;
}
extern "C" int cfdb_set_option(cfdb_ctx* c, const char* name, int32_t value) {
    std::string n(name);
    if (n == "fast") c->fast = value;
    else if (n == "use_cuarto") c->use_cuarto = value;
    else if (n == "true_rk") c->true_rk = value;
    else if (n == "adamsb") c->adamsb = value;
    else if (n == "colored") {
        if (value && c->color_ptr.empty()) {
            // greedy first-fit colours of the element list in the file's order (cfdb_color_elements)
            vector<int32_t> col((size_t)c->nelem);
            int32_t nc = 0;
            TRY(cfdb_color_elements(c->h_inpoel.data(), c->nelem, c->npoin, col.data(), &nc));
            c->color_ptr.assign((size_t)nc + 1, 0);
            for (int e = 0; e < c->nelem; ++e) c->color_ptr[col[e] + 1]++;
            for (int k = 0; k < nc; ++k) c->color_ptr[k + 1] += c->color_ptr[k];
            vector<int> cur(c->color_ptr.begin(), c->color_ptr.end() - 1), list((size_t)c->nelem);
            for (int e = 0; e < c->nelem; ++e) list[cur[col[e]]++] = c->perm_on ? c->h_e2i[e] : e;
            TRY(upload(c, c->color_list, list));
            CK(cudaStreamSynchronize(c->st));
        }
        c->colored = value;
    }
    else if (n == "ale") {
        // The mesh moves although this context holds no body set of its own: a rank of a multi-GPU run whose sub-domain
        // does not touch the body still receives W_X, W_Y from the global mesh solve (cfd_b200/partition.py sets this).
        if (value && !c->ale) { c->ale = true; TRY(zero(c, c->FC, 12 * (size_t)c->nelem)); }
    }
    else return fail("cfdb_set_option: unknown option " + n);
    c->epoch++;
    return 0;
}
extern "C" void* cfdb_stream(cfdb_ctx* c) { return (void*)c->st; }
extern "C" int cfdb_profile_enable(cfdb_ctx* c, int32_t on) {
    TRY(prof_resolve(c));
    c->prof = on != 0;
    c->epoch++;
    if (on) {
        for (int i = 0; i < K_COUNT; ++i) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
    }
    return 0;
}
extern "C" int cfdb_profile_get(cfdb_ctx* c, const char* kernel, double* total_ms, int64_t* launches) {
    TRY(prof_resolve(c));
    for (int i = 0; i < K_COUNT; ++i)
        if (!strcmp(kernel, kKernelNames[i])) {
            *total_ms = c->prof_ms[i];
            *launches = c->prof_n[i];
            return 0;
        }
    return fail(std::string("unknown kernel name ") + kernel);
}
extern "C" int64_t cfdb_launch_count(cfdb_ctx* c) { return c->launches; }

// ---------------------------------------------------------------------------------------------
// host <-> device transfers of nodal and element arrays
// Element arrays cross the ABI in the file's element order and live on the device in the internal (tile) order.
// scratch: EC is free between calls (12 doubles per element); an array that IS EC goes through a temporary.
static int up_elem(cfdb_ctx* c, double* dev, const double* h, int w = 1) {   // dev[p][q] = h[i2e[p]][q]
    const long E = c->nelem;
    if (!c->perm_on) return up_plain(c, dev, h, (size_t)w * E);
    double* stage = c->EC.p;
    DBuf<double> tmp;
    if (dev == c->EC.p || w > 12) { TRY(tmp.alloc((size_t)w * E)); stage = tmp.p; }
    CK(cudaMemcpyAsync(stage, h, (size_t)w * E * sizeof(double), cudaMemcpyHostToDevice, c->st));
    LAUNCH(K_LAYOUT, k::perm_rows<double>, grid_for(w * E, 256), 256, E, w, c->i2e.p, stage, dev);
    if (tmp.p) { CK(cudaStreamSynchronize(c->st)); tmp.release(); }
    return 0;
}
static int down_elem(cfdb_ctx* c, const double* dev, double* h, int w = 1) {   // h[e][q] = dev[e2i[e]][q]
    const long E = c->nelem;
    if (!c->perm_on) return down_plain(c, dev, h, (size_t)w * E);
    double* stage = c->EC.p;
    DBuf<double> tmp;
    if (dev == c->EC.p || w > 12) { TRY(tmp.alloc((size_t)w * E)); stage = tmp.p; }
    LAUNCH(K_LAYOUT, k::perm_rows<double>, grid_for(w * E, 256), 256, E, w, c->e2i.p, dev, stage);
    CK(cudaMemcpyAsync(h, stage, (size_t)w * E * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));   // the scratch buffer is reused by the next transfer
    if (tmp.p) tmp.release();
    return 0;
}
static int up_soa3(cfdb_ctx* c, double* dev, const double* h) {  // (3,E) host, file order -> [3][E] device, internal order
    const long E = c->nelem;
    TRY(c->EC.alloc(12 * (size_t)E));
    double* stage = c->EC.p;
    CK(cudaMemcpyAsync(stage, h, 3 * E * sizeof(double), cudaMemcpyHostToDevice, c->st));
    if (c->perm_on) LAUNCH(K_LAYOUT, k::aos3_to_soa_perm, grid_for(3 * E, 256), 256, E, c->i2e.p, stage, dev);
    else LAUNCH(K_LAYOUT, k::aos3_to_soa, grid_for(3 * E, 256), 256, E, stage, dev);
    c->geo_dirty = true;   // dNx / dNy are the only (3,E) arrays
    return 0;
}
static int down_soa3(cfdb_ctx* c, const double* dev, double* h) {
    const long E = c->nelem;
    double* stage = c->EC.p;
    if (c->perm_on) LAUNCH(K_LAYOUT, k::soa_to_aos3_perm, grid_for(3 * E, 256), 256, E, c->i2e.p, dev, stage);
    else LAUNCH(K_LAYOUT, k::soa_to_aos3, grid_for(3 * E, 256), 256, E, dev, stage);
    CK(cudaMemcpyAsync(h, stage, 3 * E * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// field access
struct Field;
static bool find_field(cfdb_ctx* c, const std::string& n, Field& f);
struct Field {
    void* dev = nullptr;          // device pointer (doubles or ints)
    const void* host = nullptr;   // or host-resident integer artefact
    int64_t count = 0;
    int kind = 0;                 // 0 f64 plain, 1 f64 (3,E) stored [3][E], 2 int32 host (read-only), 3 computed, 4 column of a nodal record array
    int ew = 0;                   // > 0: element-indexed, ew doubles per element, stored in the internal element order
};
static bool find_field(cfdb_ctx* c, const std::string& n, Field& f) {
    const int64_t P = c->npoin, E = c->nelem;
#define FD(name, buf, cnt) if (n == name) { f.dev = c->buf.p; f.count = (cnt); f.kind = 0; return true; }
#define FE(name, buf, w) if (n == name) { f.dev = c->buf.p; f.count = (w) * E; f.kind = 0; f.ew = (w); return true; }
#define FS(name, buf) if (n == name) { f.dev = c->buf.p; f.count = 3 * E; f.kind = 1; return true; }
#define FH(name, vec) if (n == name) { f.host = c->vec.data(); f.count = (int64_t)c->vec.size(); f.kind = 2; return true; }
    FD("X", X, P) FD("Y", Y, P) FD("X1", X1, P) FD("Y1", Y1, P) FD("M", M, P) FE("area", area, 1) FE("HH", HH, 1)
    FE("HHX", HHX, 1) FE("HHY", HHY, 1) FS("dNx", dNx) FS("dNy", dNy)
    FD("U", U, 4 * P) FD("U1", U1, 4 * P) FD("RHS", RHS, 4 * P) FD("UN", UN, 4 * P)
    FD("RHS1", RHS1, 4 * P) FD("RHS2", RHS2, 4 * P) FD("RHS3", RHS3, 4 * P)
#define FN(name, fld) if (n == name) { f.dev = c->fld.col.q; f.count = P; f.kind = 4; return true; }
    FN("VEL_X", VEL_X) FN("VEL_Y", VEL_Y) FD("W_X", W_X, P) FD("W_Y", W_Y, P) FN("P", P) FN("T", T)
    FN("RHO", RHO) FN("E", E) FN("RMACH", RMACH) FN("GAMM", GAMM)
#undef FN
    FE("SHOC", SHOC, 1) FE("T_SUGN1", TS1, 1) FE("T_SUGN2", TS2, 1) FE("T_SUGN3", TS3, 1) FE("DT", DT, 1)
    FD("lap_sparse", lap_sparse, c->nnz) FD("lap_diag", lap_diag, P) FD("xpos", xpos, P) FD("ypos", ypos, P)
    FD("dxpos", dxpos, P) FD("dypos", dypos, P) FD("W_x_old", W_x_old, P) FD("W_y_old", W_y_old, P)
    FE("area_old", area_old, 1) FE("EC", EC, 12)
    if (n == "esup1" || n == "esup2" || n == "psup1" || n == "psup2" || n == "lap_idx" || n == "lap_rowptr")
        if (ensure_host_topology(c)) return false;
    FH("inpoel", h_inpoel) FH("esup1", esup1) FH("esup2", esup2) FH("psup1", psup1) FH("psup2", psup2)
    FH("lap_idx", lap_idx) FH("lap_rowptr", lap_rowptr) FH("ilaux", h_ilaux)
#undef FD
#undef FE
#undef FS
#undef FH
    if (n == "FX") { f.dev = c->sc->FX; f.count = 10; f.kind = 0; return true; }
    if (n == "FY") { f.dev = c->sc->FY; f.count = 10; f.kind = 0; return true; }
    if (n == "RM") { f.dev = c->sc->RM; f.count = 10; f.kind = 0; return true; }
    if (n == "F_VX") { f.dev = c->fvisc.p; f.count = 10; f.kind = 0; return true; }
    if (n == "F_VY") { f.dev = c->fvisc.p + 10; f.count = 10; f.kind = 0; return true; }
    if (n == "skin") { f.dev = c->skin.p; f.count = c->nedges; f.kind = 0; return true; }
    if (n == "skin_x") { f.dev = c->skin.p + c->nedges; f.count = c->nedges; f.kind = 0; return true; }
    if (n == "skin_p") { f.dev = c->skin.p + 2 * (size_t)c->nedges; f.count = c->nedges; f.kind = 0; return true; }
    if (n == "DTL") { f.count = E; f.kind = 3; return true; }
    if (n == "n_ipoin" || n == "n_x" || n == "n_y") { f.count = c->nwn; f.kind = 3; return true; }
    return false;
}
extern "C" int cfdb_halo_exchange(cfdb_ctx* c, const char* name) {
    CK(cudaSetDevice(c->device));
    Field f;
    std::string n(name);
    if (!find_field(c, n, f) || (f.kind != 0 && f.kind != 4)) return fail("cfdb_halo_exchange: not a nodal float64 field: " + n);
    if (f.kind == 4) {   // a column of a nodal record array: through the plain scratch
        TRY(col_to_plain(c, NodeField{k::Col{(double*)f.dev}}, c->ntmp.p, (size_t)c->npoin));
        TRY(halo_vec(c, c->ntmp.p, 1));
        TRY(col_from_plain(c, NodeField{k::Col{(double*)f.dev}}, c->ntmp.p, (size_t)c->npoin));
        CK(cudaStreamSynchronize(c->st));
        return 0;
    }
    int w = f.count == 4 * (int64_t)c->npoin ? 4 : (f.count == c->npoin ? 1 : 0);
    if (!w) return fail("cfdb_halo_exchange: not a nodal field: " + n);
    TRY(halo_vec(c, (double*)f.dev, w));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}
extern "C" int64_t cfdb_field_size(cfdb_ctx* c, const char* name) {
    Field f;
    if (!find_field(c, name, f)) return -1;
    return f.count;
}
extern "C" int cfdb_get(cfdb_ctx* c, const char* name, void* host, int64_t count) {
    CK(cudaSetDevice(c->device));
    Field f;
    std::string n(name);
    if (!find_field(c, n, f)) return fail("cfdb_get: unknown field " + n);
    if (n == "U1" && c->u1_is_u) f.dev = c->U.p;
    if (count < f.count && f.kind != 3) return fail("cfdb_get: host buffer too small for " + n);
    if (f.kind == 2) {
        memcpy(host, f.host, f.count * sizeof(int32_t));
        return 0;
    }
    if (f.kind == 0 && f.ew > 0) {
        TRY(down_elem(c, (const double*)f.dev, (double*)host, f.ew));
        CK(cudaStreamSynchronize(c->st));
        return 0;
    }
    if (f.kind == 0) {
        if (f.count) CK(cudaMemcpyAsync(host, f.dev, f.count * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        return 0;
    }
    if (f.kind == 1) return down_soa3(c, (const double*)f.dev, (double*)host);
    if (f.kind == 4) {
        TRY(down_node(c, NodeField{k::Col{(double*)f.dev}}, (double*)host, (size_t)f.count));
        CK(cudaStreamSynchronize(c->st));
        return 0;
    }
    if (n == "DTL") {  // ns2DComp.ALE.f90:159-164
        if (count < c->nelem) return fail("cfdb_get: host buffer too small for DTL");
        if (c->par.ITLOCAL != 0) {
            TRY(down_elem(c, c->DTL.p, (double*)host));
            CK(cudaStreamSynchronize(c->st));
        } else {
            TRY(read_scal(c));
            for (int e = 0; e < c->nelem; ++e) ((double*)host)[e] = c->h_sc->DTMIN;
        }
        return 0;
    }
    // compacted normals list in ascending node order (subrutinas.f90:51-63); count returned via n_m scalar
    vector<double> wx(c->nwn), wy(c->nwn);
    vector<int> wv(c->nwn);
    if (c->nwn) {
        CK(cudaMemcpyAsync(wx.data(), c->wn_x.p, c->nwn * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(wy.data(), c->wn_y.p, c->nwn * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(wv.data(), c->wn_valid.p, c->nwn * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    }
    CK(cudaStreamSynchronize(c->st));
    int m = 0;
    for (int j = 0; j < c->nwn; ++j)
        if (wv[j]) {
            if (m >= count) return fail("cfdb_get: host buffer too small for " + n);
            if (n == "n_ipoin") ((int32_t*)host)[m] = c->h_wn_node[j] + 1;
            else if (n == "n_x") ((double*)host)[m] = wx[j];
            else ((double*)host)[m] = wy[j];
            ++m;
        }
    return 0;
}
extern "C" int cfdb_set(cfdb_ctx* c, const char* name, const void* host, int64_t count) {
    CK(cudaSetDevice(c->device));
    Field f;
    std::string n(name);
    if (!find_field(c, n, f)) return fail("cfdb_set: unknown field " + n);
    if (f.kind == 2 || f.kind == 3 || n == "dxpos" || n == "dypos") return fail("cfdb_set: field " + n + " is read-only");
    if (count != f.count) return fail("cfdb_set: wrong element count for " + n);
    if (f.kind == 4) {
        TRY(up_node(c, NodeField{k::Col{(double*)f.dev}}, (const double*)host, (size_t)f.count));
        CK(cudaStreamSynchronize(c->st));
        return 0;
    }
    if (f.kind == 0 && f.ew > 0) {
        TRY(up_elem(c, (double*)f.dev, (const double*)host, f.ew));
        CK(cudaStreamSynchronize(c->st));
        if (n == "area") c->geo_dirty = true;
        return 0;
    }
    if (f.kind == 0) {
        if (f.count) CK(cudaMemcpyAsync(f.dev, host, f.count * sizeof(double), cudaMemcpyHostToDevice, c->st));
        CK(cudaStreamSynchronize(c->st));
        if (n == "UN") { c->theta_nonzero = true; c->epoch++; }
        if (n == "U" && c->u1_is_u) {  // keep U1 == U as in the reference
            CK(cudaMemcpyAsync(c->U1.p, c->U.p, f.count * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            CK(cudaStreamSynchronize(c->st));
        }
        // anything that can make the mesh velocity non-zero on a context without body sets: from here on the context
        // computes FUENTE and the mesh-velocity terms of ESTAB/deltat
        if (n == "W_X" || n == "W_Y" || n == "xpos" || n == "ypos") {
            if (!c->ale) { c->ale = true; c->epoch++; TRY(zero(c, c->FC, 12 * (size_t)c->nelem)); }
        }
        return 0;
    }
    TRY(up_soa3(c, (double*)f.dev, (const double*)host));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}
extern "C" int cfdb_get_scalar(cfdb_ctx* c, const char* name, double* v) {
    CK(cudaSetDevice(c->device));
    TRY(read_scal(c));
    std::string n(name);
    const k::Scal& s = *c->h_sc;
    if (n == "TIME") *v = s.TIME;
    else if (n == "DTMIN") *v = s.DTMIN;
    else if (n == "DTMIN1") *v = s.DTMIN1;
    else if (n == "HMIN") *v = s.HMIN > 1.e10 ? 1.e10 : s.HMIN;  // subrutinas.f90:125
    else if (n == "ITER") *v = s.ITER;
    else if (n == "BANDERA") *v = s.BANDERA;
    else if (n == "bicg_x") *v = c->bicg_iters[0];
    else if (n == "bicg_y") *v = c->bicg_iters[1];
    else if (n == "FX1") *v = s.FX[0];
    else if (n == "FY1") *v = s.FY[0];
    else if (n == "RM1") *v = s.RM[0];
    else if (n == "tile_interior") *v = c->tiles_ok ? c->tile_interior : 0.0;
    else if (n == "graph_replays") *v = (double)c->graph_replays;
    else if (n == "n_m") {
        vector<int> wv(c->nwn);
        if (c->nwn) CK(cudaMemcpy(wv.data(), c->wn_valid.p, c->nwn * sizeof(int), cudaMemcpyDeviceToHost));
        int m = 0;
        for (int x : wv) m += x;
        *v = m;
    } else return fail("cfdb_get_scalar: unknown scalar " + n);
    return 0;
}
extern "C" int cfdb_set_scalar(cfdb_ctx* c, const char* name, double v) {
    CK(cudaSetDevice(c->device));
    TRY(read_scal(c));
    std::string n(name);
    k::Scal& s = *c->h_sc;
    if (n == "TIME") s.TIME = v;
    else if (n == "DTMIN") s.DTMIN = v;
    else if (n == "DTMIN1") s.DTMIN1 = v;
    else if (n == "ITER") { s.ITER = (int)v; c->h_iter = (int)v; }
    else if (n == "BANDERA") s.BANDERA = (int)v;
    else return fail("cfdb_set_scalar: unknown scalar " + n);
    CK(cudaMemcpyAsync(c->sc, c->h_sc, sizeof(k::Scal), cudaMemcpyHostToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// (i) call-site mode
static int check_mesh(cfdb_ctx* c, int32_t nelem, int32_t npoin, const char* who) {
    if (nelem != c->nelem || npoin != c->npoin)
        return fail(std::string(who) + ": nelem/npoin differ from the connectivity this context was created with");
    return 0;
}
extern "C" int cfdb_calcrhs(cfdb_ctx* c, double* rhs, const double* U, const double* theta, const double* T,
                            const double* dNx, const double* dNy, const double* area, const double* shoc,
                            const double* dtl, const double* ts1, const double* ts2, const double* ts3,
                            const int32_t* inpoel, int32_t nelem, int32_t npoin, double Cv, double lambda_ref,
                            double mu_ref, double gamma0, double T_inf, double cte) {
    CK(cudaSetDevice(c->device));
    TRY(check_mesh(c, nelem, npoin, "cfdb_calcrhs"));
    (void)inpoel;
    const size_t P = npoin, E = nelem;
    TRY(up_soa3(c, c->dNx.p, dNx));
    TRY(up_soa3(c, c->dNy.p, dNy));
    TRY(up_plain(c, c->U.p, U, 4 * P));
    TRY(up_plain(c, c->UN.p, theta, 4 * P));
    TRY(up_node(c, c->T, T, P));
    TRY(up_elem(c, c->area.p, area)); c->geo_dirty = true;
    TRY(up_elem(c, c->SHOC.p, shoc));
    TRY(up_elem(c, c->DTL.p, dtl));
    TRY(up_elem(c, c->TS1.p, ts1));
    TRY(up_elem(c, c->TS2.p, ts2));
    TRY(up_elem(c, c->TS3.p, ts3));
    c->theta_nonzero = true;
    c->epoch++;
    k::Gas g{Cv, lambda_ref, mu_ref, gamma0, T_inf, cte};
    TRY(run_calcrhs_elem(c, g, true, false, c->DTL.p, nullptr));
    // rhs is inout: the reference adds onto the caller's array in element order, (((rhs+a1)+a2)+...)
    TRY(up_plain(c, c->RHS1.p, rhs, 4 * P));
    LAUNCH(K_NODE, k::node_accumulate, grid_for(npoin, 256), 256, npoin, c->d_esup2.p, c->eslot.p, c->EC.p, c->RHS1.p, c->RHS.p);
    TRY(down_plain(c, c->RHS.p, rhs, 4 * P));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int cfdb_fuente(cfdb_ctx* c, double* rhs, const double* U, const double* w_x, const double* w_y,
                           const double* dNx, const double* dNy, const double* area, const double* dtl,
                           const int32_t* inpoel, int32_t nelem, int32_t npoin) {
    CK(cudaSetDevice(c->device));
    TRY(check_mesh(c, nelem, npoin, "cfdb_fuente"));
    (void)inpoel;
    const size_t P = npoin, E = nelem;
    TRY(c->FC.alloc(12 * E));
    c->ale = true;   // the caller's W replaces the resident one: a later cfdb_step must not assume W = 0
    c->epoch++;
    TRY(up_soa3(c, c->dNx.p, dNx));
    TRY(up_soa3(c, c->dNy.p, dNy));
    TRY(up_plain(c, c->U.p, U, 4 * P));
    TRY(up_plain(c, c->W_X.p, w_x, P));
    TRY(up_plain(c, c->W_Y.p, w_y, P));
    TRY(up_elem(c, c->area.p, area)); c->geo_dirty = true;
    TRY(up_elem(c, c->DTL.p, dtl));
    k::Gas g{1.0, 0.0, 0.0, 1.4, 1.0, 1.0};
    // the ALE instantiation writes FC (FUENTE) next to EC (calcRHS, ignored here)
    TRY(run_calcrhs_elem(c, g, false, true, c->DTL.p, nullptr));
    TRY(up_plain(c, c->RHS1.p, rhs, 4 * P));
    LAUNCH(K_NODE, k::node_accumulate, grid_for(npoin, 256), 256, npoin, c->d_esup2.p, c->eslot.p, c->FC.p, c->RHS1.p, c->RHS.p);
    TRY(down_plain(c, c->RHS.p, rhs, 4 * P));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int cfdb_deltat(cfdb_ctx* c, double* dtmin, double* dt, const int32_t* inpoel, const double* area,
                           const double* T, const double* vel_x, const double* vel_y, const double* w_x,
                           const double* w_y, int32_t nelem, int32_t npoin, double FSAFE, double FR, double GAMA,
                           double T_inf) {
    CK(cudaSetDevice(c->device));
    TRY(check_mesh(c, nelem, npoin, "cfdb_deltat"));
    (void)inpoel; (void)FR; (void)GAMA;  // VC = sqrt(GAMA*FR*T) is dead in the reference (subrutinas.f90:179)
    const size_t P = npoin, E = nelem;
    TRY(up_elem(c, c->area.p, area)); c->geo_dirty = true;
    TRY(up_node(c, c->T, T, P));
    TRY(up_node(c, c->VEL_X, vel_x, P));
    TRY(up_node(c, c->VEL_Y, vel_y, P));
    TRY(up_plain(c, c->W_X.p, w_x, P));
    TRY(up_plain(c, c->W_Y.p, w_y, P));
    if (!c->ale) { c->ale = true; c->epoch++; TRY(zero(c, c->FC, 12 * E)); }   // caller's W is resident now (see cfdb_fuente)
    LAUNCH(K_SCALAR, k::set_double, 1, 1, &c->sc->dtmin_acc, 1.e20);
    LAUNCH(K_DELTAT, k::deltat<true>, grid_for(nelem, 256), 256, nelem, c->inp.p, c->area.p, c->T.col, c->VEL_X.col,
           c->VEL_Y.col, c->W_X.p, c->W_Y.p, FSAFE, T_inf, c->DT.p, c->sc);
    LAUNCH(K_DTL, k::dtl_blend, grid_for(nelem, 256), 256, nelem, c->DT.p, c->DTL.p, c->sc, 0);
    TRY(down_elem(c, c->DT.p, dt));
    TRY(read_scal(c));
    *dtmin = c->h_sc->dtmin_acc;
    return 0;
}

extern "C" int cfdb_estab(cfdb_ctx* c, const double* U, const double* T, const double* vel_x, const double* vel_y,
                          const double* w_x, const double* w_y, const double* GAMM, const double* dNx,
                          const double* dNy, const int32_t* inpoel, int32_t nelem, int32_t npoin, double FR,
                          double DTMIN, double RHOINF, double TINF, double* shoc, double* ts1, double* ts2, double* ts3) {
    CK(cudaSetDevice(c->device));
    TRY(check_mesh(c, nelem, npoin, "cfdb_estab"));
    (void)inpoel;
    const size_t P = npoin, E = nelem;
    TRY(up_soa3(c, c->dNx.p, dNx));
    TRY(up_soa3(c, c->dNy.p, dNy));
    TRY(up_plain(c, c->U.p, U, 4 * P));
    TRY(up_node(c, c->T, T, P));
    TRY(up_node(c, c->VEL_X, vel_x, P));
    TRY(up_node(c, c->VEL_Y, vel_y, P));
    TRY(up_plain(c, c->W_X.p, w_x, P));
    TRY(up_plain(c, c->W_Y.p, w_y, P));
    TRY(up_node(c, c->GAMM, GAMM, P));
    if (!c->ale) { c->ale = true; c->epoch++; TRY(zero(c, c->FC, 12 * E)); }   // caller's W is resident now (see cfdb_fuente)
    LAUNCH(K_SCALAR, k::set_double, 1, 1, &c->sc->red[15], DTMIN);
    LAUNCH(K_ESTAB, k::estab<3>, grid_for(nelem, 256), 256, nelem, c->inp.p, c->U.p, c->T.col, c->VEL_X.col, c->VEL_Y.col, c->W_X.p,
           c->W_Y.p, c->GAMM.col, c->dNx.p, c->dNy.p, FR, &c->sc->red[15], RHOINF, TINF, c->SHOC.p, c->TS1.p, c->TS2.p,
           c->TS3.p);
    TRY(down_elem(c, c->SHOC.p, shoc));
    TRY(down_elem(c, c->TS1.p, ts1));
    TRY(down_elem(c, c->TS2.p, ts2));
    TRY(down_elem(c, c->TS3.p, ts3));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int cfdb_deriv(cfdb_ctx* c, const double* X, const double* Y, const int32_t* inpoel, int32_t nelem,
                          int32_t npoin, double* area, double* HH, double* HHX, double* HHY, double* dNx, double* dNy,
                          double* hmin) {
    CK(cudaSetDevice(c->device));
    TRY(check_mesh(c, nelem, npoin, "cfdb_deriv"));
    (void)inpoel;
    const size_t P = npoin, E = nelem;
    TRY(up_plain(c, c->X.p, X, P));
    TRY(up_plain(c, c->Y.p, Y, P));
    TRY(run_deriv(c));
    TRY(down_elem(c, c->area.p, area));
    TRY(down_elem(c, c->HH.p, HH));
    TRY(down_elem(c, c->HHX.p, HHX));
    TRY(down_elem(c, c->HHY.p, HHY));
    TRY(down_soa3(c, c->dNx.p, dNx));
    CK(cudaStreamSynchronize(c->st));
    TRY(down_soa3(c, c->dNy.p, dNy));
    TRY(read_scal(c));
    *hmin = c->h_sc->HMIN > 1.e10 ? 1.e10 : c->h_sc->HMIN;
    return 0;
}

extern "C" int cfdb_masas(cfdb_ctx* c, const double* area, const int32_t* inpoel, int32_t nelem, int32_t npoin, double* M) {
    CK(cudaSetDevice(c->device));
    TRY(check_mesh(c, nelem, npoin, "cfdb_masas"));
    (void)inpoel;
    TRY(up_elem(c, c->area.p, area)); c->geo_dirty = true;
    TRY(run_masas(c));
    TRY(down_plain(c, c->M.p, M, npoin));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int cfdb_normales(cfdb_ctx* c, const int32_t* wall, int32_t nwall, const double* X, const double* Y,
                             int32_t npoin, int32_t* m_out, int32_t* n_ipoin, double* n_x, double* n_y) {
    CK(cudaSetDevice(c->device));
    if (npoin != c->npoin) return fail("cfdb_normales: npoin differs from the context");
    if (nwall != (int)c->h_wall.size() / 2 || memcmp(wall, c->h_wall.data(), c->h_wall.size() * sizeof(int32_t)))
        return fail("cfdb_normales: wall list differs from the one the context was created with");
    TRY(up_plain(c, c->X.p, X, npoin));
    TRY(up_plain(c, c->Y.p, Y, npoin));
    TRY(run_normales(c));
    double m = 0;
    TRY(cfdb_get_scalar(c, "n_m", &m));
    *m_out = (int32_t)m;
    TRY(cfdb_get(c, "n_ipoin", n_ipoin, *m_out));
    TRY(cfdb_get(c, "n_x", n_x, *m_out));
    TRY(cfdb_get(c, "n_y", n_y, *m_out));
    return 0;
}

extern "C" int cfdb_laplace(cfdb_ctx* c, const int32_t* inpoel, const double* area, const double* dNx, const double* dNy,
                            const double* X, const double* Y, int32_t nelem, int32_t npoin, double* lap_sparse,
                            double* lap_diag) {
    CK(cudaSetDevice(c->device));
    TRY(check_mesh(c, nelem, npoin, "cfdb_laplace"));
    (void)inpoel; (void)area;  // `a = area(ielem)` is assigned and never used (mLaplace.f90:37)
    TRY(up_soa3(c, c->dNx.p, dNx));
    TRY(up_soa3(c, c->dNy.p, dNy));
    TRY(up_plain(c, c->X.p, X, npoin));
    TRY(up_plain(c, c->Y.p, Y, npoin));
    TRY(run_laplace(c));
    TRY(down_plain(c, c->lap_sparse.p, lap_sparse, c->nnz));
    TRY(down_plain(c, c->lap_diag.p, lap_diag, npoin));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// call-site entries for the list-driven boundary conditions and for RK / fluidStructure as whole routines
namespace {
__global__ void k_fixvel(int m, const int* __restrict__ idx, const int* __restrict__ last, const double* __restrict__ vx,
                         const double* __restrict__ vy, double* __restrict__ VX, double* __restrict__ VY) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    VX[idx[i]] = vx[last[i]];   // duplicates: every copy writes the value of the LAST list entry
    VY[idx[i]] = vy[last[i]];
}
__global__ void k_normalvel(int m, const int* __restrict__ ip, const double* __restrict__ nx, const double* __restrict__ ny,
                            double* __restrict__ VX, double* __restrict__ VY, const double* __restrict__ WX,
                            const double* __restrict__ WY) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int n = ip[i];
    double vx = VX[n], vy = VY[n], wx = WX[n], wy = WY[n];
    double p = -ny[i] * (vx - wx) + nx[i] * (vy - wy);   // subrutinas.f90:78-80
    VX[n] = -ny[i] * p + wx;
    VY[n] = nx[i] * p + wy;
}
__global__ void k_fix_rho(int m, const int* __restrict__ idx, const int* __restrict__ last, const double* __restrict__ val,
                          double* __restrict__ RHO) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) RHO[idx[i]] = val[last[i]];
}
__global__ void k_fix_T(int m, const int* __restrict__ idx, const int* __restrict__ last, const double* __restrict__ val, double FR,
                        const double* __restrict__ GAMM, const double* __restrict__ VX, const double* __restrict__ VY,
                        double* __restrict__ T, double* __restrict__ E) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int j = idx[i];
    double GM = GAMM[j] - 1.0;
    double t = val[last[i]];
    T[j] = t;
    E[j] = t * FR / GM + .5 * (VX[j] * VX[j] + VY[j] * VY[j]);   // subrutinas.f90:636-638
}
}  // namespace

// 1-based node list -> device (0-based) + the last-wins table
static int up_list(cfdb_ctx* c, const int32_t* list, int m, int npoin, DBuf<int>& d_idx, DBuf<int>& d_last, const char* who) {
    vector<int32_t> i0(list, list + m), last;
    for (int v : i0)
        if (v < 1 || v > npoin) return fail(std::string(who) + ": node id out of range");
    topo::last_wins(list, m, npoin, last);
    for (auto& v : i0) v -= 1;
    TRY(upload(c, d_idx, i0));
    TRY(upload(c, d_last, last));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int cfdb_fixvel(cfdb_ctx* c, int32_t nfixv, const int32_t* ifixv_node, const double* rfixv_valuex,
                           const double* rfixv_valuey, double* vel_x, double* vel_y, int32_t npoin) {
    CK(cudaSetDevice(c->device));
    if (npoin != c->npoin) return fail("cfdb_fixvel: npoin differs from the context");
    DBuf<int> idx, last;
    DBuf<double> vx, vy;
    auto body = [&]() -> int {
        TRY(up_list(c, ifixv_node, nfixv, npoin, idx, last, "cfdb_fixvel"));
        TRY(upload(c, vx, rfixv_valuex, (size_t)nfixv));
        TRY(upload(c, vy, rfixv_valuey, (size_t)nfixv));
        TRY(up_plain(c, c->tmpA.p, vel_x, npoin));
        TRY(up_plain(c, c->tmpB.p, vel_y, npoin));
        if (nfixv) LAUNCH(K_FIXROWS, k_fixvel, grid_for(nfixv, 128), 128, nfixv, idx.p, last.p, vx.p, vy.p, c->tmpA.p, c->tmpB.p);
        TRY(down_plain(c, c->tmpA.p, vel_x, npoin));
        TRY(down_plain(c, c->tmpB.p, vel_y, npoin));
        CK(cudaStreamSynchronize(c->st));
        return 0;
    };
    int r = body();
    idx.release(); last.release(); vx.release(); vy.release();
    return r;
}

extern "C" int cfdb_normalvel(cfdb_ctx* c, int32_t m, const int32_t* n_ipoin, const double* n_x, const double* n_y, double* vel_x,
                              double* vel_y, const double* w_x, const double* w_y, int32_t npoin) {
    CK(cudaSetDevice(c->device));
    if (npoin != c->npoin) return fail("cfdb_normalvel: npoin differs from the context");
    DBuf<int> idx;
    DBuf<double> nx, ny;
    auto body = [&]() -> int {
        vector<int32_t> i0(n_ipoin, n_ipoin + m);
        for (auto& v : i0) {
            if (v < 1 || v > npoin) return fail("cfdb_normalvel: n_ipoin out of range");
            v -= 1;
        }
        TRY(upload(c, idx, i0));
        TRY(upload(c, nx, n_x, (size_t)m));
        TRY(upload(c, ny, n_y, (size_t)m));
        TRY(up_plain(c, c->tmpA.p, vel_x, npoin));
        TRY(up_plain(c, c->tmpB.p, vel_y, npoin));
        TRY(up_plain(c, c->tmpC.p, w_x, npoin));
        TRY(up_plain(c, c->by.p, w_y, npoin));
        if (m) LAUNCH(K_FIXROWS, k_normalvel, grid_for(m, 128), 128, m, idx.p, nx.p, ny.p, c->tmpA.p, c->tmpB.p, c->tmpC.p, c->by.p);
        TRY(down_plain(c, c->tmpA.p, vel_x, npoin));
        TRY(down_plain(c, c->tmpB.p, vel_y, npoin));
        CK(cudaStreamSynchronize(c->st));
        return 0;
    };
    int r = body();
    idx.release(); nx.release(); ny.release();
    return r;
}

extern "C" int cfdb_fix(cfdb_ctx* c, double FR, const double* GAMM, int32_t nfixrho, const int32_t* ifixrho_node,
                        const double* rfixrho_value, int32_t nfixt, const int32_t* ifixt_node, const double* rfixt_value,
                        const double* vel_x, const double* vel_y, double* rho, double* T, double* E, int32_t npoin) {
    CK(cudaSetDevice(c->device));
    if (npoin != c->npoin) return fail("cfdb_fix: npoin differs from the context");
    DBuf<int> ri, rl, ti, tl;
    DBuf<double> rv, tv, drho, dT, dE;
    auto body = [&]() -> int {
        TRY(up_list(c, ifixrho_node, nfixrho, npoin, ri, rl, "cfdb_fix"));
        TRY(up_list(c, ifixt_node, nfixt, npoin, ti, tl, "cfdb_fix"));
        TRY(upload(c, rv, rfixrho_value, (size_t)nfixrho));
        TRY(upload(c, tv, rfixt_value, (size_t)nfixt));
        TRY(upload(c, drho, rho, (size_t)npoin));
        TRY(upload(c, dT, T, (size_t)npoin));
        TRY(upload(c, dE, E, (size_t)npoin));
        TRY(up_plain(c, c->tmpA.p, vel_x, npoin));
        TRY(up_plain(c, c->tmpB.p, vel_y, npoin));
        TRY(up_plain(c, c->tmpC.p, GAMM, npoin));
        if (nfixrho) LAUNCH(K_FIXROWS, k_fix_rho, grid_for(nfixrho, 128), 128, nfixrho, ri.p, rl.p, rv.p, drho.p);
        if (nfixt) LAUNCH(K_FIXROWS, k_fix_T, grid_for(nfixt, 128), 128, nfixt, ti.p, tl.p, tv.p, FR, c->tmpC.p, c->tmpA.p, c->tmpB.p, dT.p, dE.p);
        TRY(down_plain(c, drho.p, rho, npoin));
        TRY(down_plain(c, dT.p, T, npoin));
        TRY(down_plain(c, dE.p, E, npoin));
        CK(cudaStreamSynchronize(c->st));
        return 0;
    };
    int r = body();
    for (auto* d : {&ri, &rl, &ti, &tl}) d->release();
    for (auto* d : {&rv, &tv, &drho, &dT, &dE}) d->release();
    return r;
}

// RK as a whole routine with host arrays (subrutinas.f90:645-849): the module arrays it reads (U, T, VEL_X, VEL_Y, W, GAMM,
// the RHS history) come from the caller, those it writes go back.  Geometry, M and the BC lists are the context's.
extern "C" int cfdb_rk(cfdb_ctx* c, double DTMIN, int32_t NRK, int32_t BANDERA, const double* GAMM, const double* dtl, const double* U,
                       double* U1, double* RHS, double* RHS1, double* RHS2, double* RHS3, double* T, double* P, double* RHO, double* E,
                       double* RMACH, double* VEL_X, double* VEL_Y, const double* W_X, const double* W_Y, double* SHOC, double* T_SUGN1,
                       double* T_SUGN2, double* T_SUGN3, int32_t nelem, int32_t npoin) {
    CK(cudaSetDevice(c->device));
    TRY(check_mesh(c, nelem, npoin, "cfdb_rk"));
    if (NRK != 4) return fail("cfdb_rk: NRK must be 4 (ns2DComp.ALE.f90:111)");
    const size_t P_ = npoin;
    TRY(up_plain(c, c->U.p, U, 4 * P_));
    TRY(up_node(c, c->GAMM, GAMM, P_));
    TRY(up_node(c, c->T, T, P_));
    TRY(up_node(c, c->VEL_X, VEL_X, P_));
    TRY(up_node(c, c->VEL_Y, VEL_Y, P_));
    TRY(up_plain(c, c->W_X.p, W_X, P_));
    TRY(up_plain(c, c->W_Y.p, W_Y, P_));
    TRY(up_plain(c, c->RHS1.p, RHS1, 4 * P_));
    TRY(up_plain(c, c->RHS2.p, RHS2, 4 * P_));
    TRY(up_plain(c, c->RHS3.p, RHS3, 4 * P_));
    TRY(up_elem(c, c->DTL.p, dtl));
    if (!c->ale) { c->ale = true; TRY(zero(c, c->FC, 12 * (size_t)nelem)); }   // the caller's W is resident now
    c->epoch++;
    c->u1_is_u = false;
    TRY(read_scal(c));
    c->h_sc->DTMIN = DTMIN;
    c->h_sc->BANDERA = BANDERA;
    CK(cudaMemcpyAsync(c->sc, c->h_sc, sizeof(k::Scal), cudaMemcpyHostToDevice, c->st));
    c->dtl_force = true;
    int rc = 0;
    for (int irk = 1; irk <= 4 && !rc; ++irk) rc = cfdb_rk_stage(c, irk);
    c->dtl_force = false;
    if (rc) return rc;
    LAUNCH(K_FILL, k::rhs_history, (int)std::min<long>(grid_for(4 * (long)P_, 256), 148 * 8), 256, 4 * (long)P_, c->sc, c->RHS.p, c->RHS1.p,
           c->RHS2.p, c->RHS3.p);
    TRY(down_plain(c, c->U1.p, U1, 4 * P_));
    TRY(down_plain(c, c->RHS.p, RHS, 4 * P_));
    TRY(down_plain(c, c->RHS1.p, RHS1, 4 * P_));
    TRY(down_plain(c, c->RHS2.p, RHS2, 4 * P_));
    TRY(down_plain(c, c->RHS3.p, RHS3, 4 * P_));
    struct { double* h; NodeField f; } outs[] = {{T, c->T}, {P, c->P}, {RHO, c->RHO}, {E, c->E}, {RMACH, c->RMACH}, {VEL_X, c->VEL_X}, {VEL_Y, c->VEL_Y}};
    for (auto& o : outs) TRY(down_node(c, o.f, o.h, P_));
    CK(cudaStreamSynchronize(c->st));
    TRY(down_elem(c, c->SHOC.p, SHOC));
    TRY(down_elem(c, c->TS1.p, T_SUGN1));
    TRY(down_elem(c, c->TS2.p, T_SUGN2));
    TRY(down_elem(c, c->TS3.p, T_SUGN3));
    return 0;
}

// fluidStructure as a whole routine with host arrays (meshMove.f90:28-142)
extern "C" int cfdb_mesh_move(cfdb_ctx* c, double dtmin, double time, double* X, double* Y, double* X1, double* Y1, double* W_X,
                              double* W_Y, const double* P, double* xpos, double* ypos, double fx[10], double fy[10], double rm[10],
                              int32_t npoin) {
    CK(cudaSetDevice(c->device));
    if (npoin != c->npoin) return fail("cfdb_mesh_move: npoin differs from the context");
    const size_t P_ = npoin;
    TRY(up_plain(c, c->X.p, X, P_));
    TRY(up_plain(c, c->Y.p, Y, P_));
    TRY(up_plain(c, c->X1.p, X1, P_));
    TRY(up_plain(c, c->Y1.p, Y1, P_));
    TRY(up_node(c, c->P, P, P_));
    TRY(up_plain(c, c->xpos.p, xpos, P_));
    TRY(up_plain(c, c->ypos.p, ypos, P_));
    if (!c->ale) { c->ale = true; TRY(zero(c, c->FC, 12 * (size_t)c->nelem)); }
    c->epoch++;
    TRY(read_scal(c));
    c->h_sc->DTMIN = dtmin;
    CK(cudaMemcpyAsync(c->sc, c->h_sc, sizeof(k::Scal), cudaMemcpyHostToDevice, c->st));
    TRY(cfdb_fluid_structure(c, dtmin, time));
    struct { double* h; const double* d; } outs[] = {{X, c->X.p}, {Y, c->Y.p}, {X1, c->X1.p}, {Y1, c->Y1.p}, {W_X, c->W_X.p},
                                                     {W_Y, c->W_Y.p}, {xpos, c->xpos.p}, {ypos, c->ypos.p}};
    for (auto& o : outs) TRY(down_plain(c, o.d, o.h, P_));
    TRY(read_scal(c));
    for (int i = 0; i < 10; ++i) { fx[i] = c->h_sc->FX[i]; fy[i] = c->h_sc->FY[i]; rm[i] = c->h_sc->RM[i]; }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// the print step's text files besides the GiD post file
static int get_forces(cfdb_ctx* c, double fx[10], double fy[10], double rm[10], double fvx[10], double fvy[10]) {
    TRY(read_scal(c));
    for (int i = 0; i < 10; ++i) { fx[i] = c->h_sc->FX[i]; fy[i] = c->h_sc->FY[i]; rm[i] = c->h_sc->RM[i]; }
    double fv[20];
    CK(cudaMemcpyAsync(fv, c->fvisc.p, sizeof fv, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    for (int i = 0; i < 10; ++i) { fvx[i] = fv[i]; fvy[i] = fv[10 + i]; }
    return 0;
}
extern "C" int cfdb_write_forces(cfdb_ctx* c, const char* path) {   // ns2DComp.ALE.f90:238-250
    CK(cudaSetDevice(c->device));
    double fx[10], fy[10], rm[10], fvx[10], fvy[10];
    TRY(get_forces(c, fx, fy, rm, fvx, fvy));
    std::FILE* f = std::fopen(path, "w");
    if (!f) return fail(std::string("cfdb_write_forces: cannot open ") + path);
    for (int s = 0; s < c->nset; ++s) {   // NSET_NUMB sets
        std::fprintf(f, "SET NUMERO%s\n", ffmt::I(s + 1, 2).c_str());
        std::fprintf(f, "FUERZA EN X:%s\n", ffmt::E(fx[s], 14, 5).c_str());
        std::fprintf(f, "FUERZA EN Y:%s\n\n", ffmt::E(fy[s], 14, 5).c_str());
        std::fprintf(f, "FUERZA VISCOSA EN X:%s\n", ffmt::E(fvx[s], 14, 5).c_str());
        std::fprintf(f, "FUERZA VISCOSA EN Y:%s\n\n", ffmt::E(fvy[s], 14, 5).c_str());
        std::fprintf(f, "FUERZA TOTAL EN X:%s\n", ffmt::E(fx[s] + fvx[s], 14, 5).c_str());
        std::fprintf(f, "FUERZA TOTAL EN Y:%s\n\n", ffmt::E(fy[s] + fvy[s], 14, 5).c_str());
    }
    std::fclose(f);
    return 0;
}
extern "C" int cfdb_write_desplazamiento(cfdb_ctx* c, const char* path, double time, int32_t append) {   // :237 '(7E13.5)'
    CK(cudaSetDevice(c->device));
    double fx[10], fy[10], rm[10], fvx[10], fvy[10];
    TRY(get_forces(c, fx, fy, rm, fvx, fvy));
    std::FILE* f = std::fopen(path, append ? "a" : "w");
    if (!f) return fail(std::string("cfdb_write_desplazamiento: cannot open ") + path);
    const double v[7] = {time, fvx[0], fvy[0], rm[0], fvx[1], fvy[1], rm[1]};
    std::string s;
    for (double x : v) s += ffmt::E(x, 13, 5);
    std::fprintf(f, "%s\n", s.c_str());
    std::fclose(f);
    return 0;
}
extern "C" int cfdb_write_skin(cfdb_ctx* c, const char* path) {   // :833, :888  write(1, *) SKIN, xmid, press/82713.27
    CK(cudaSetDevice(c->device));
    const int ne = c->nedges;
    vector<double> sk(3 * (size_t)std::max(ne, 1));
    CK(cudaMemcpyAsync(sk.data(), c->skin.p, 3 * (size_t)std::max(ne, 1) * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    std::FILE* f = std::fopen(path, "w");
    if (!f) return fail(std::string("cfdb_write_skin: cannot open ") + path);
    for (int k = 0; k < ne; ++k)
        std::fprintf(f, " %s%s%s\n", ffmt::list_r8(sk[k]).c_str(), ffmt::list_r8(sk[ne + k]).c_str(), ffmt::list_r8(sk[2 * (size_t)ne + k]).c_str());
    std::fclose(f);
    return 0;
}

// host code, no GPU: the tile decomposition of the fused stage for a mesh (host_topology.h), for tests and mesh diagnostics
extern "C" int cfdb_tile_elements(const int32_t* inpoel, int32_t nelem, int32_t npoin, const double* X, const double* Y, int32_t TE,
                                  int32_t order, int32_t* i2e, double* stats) {
    if (TE < 32 || TE % 32 || order < 0 || order > 2) return fail("cfdb_tile_elements: TE must be a multiple of 32, order 0..2");
    for (size_t k = 0; k < 3 * (size_t)nelem; ++k)
        if (inpoel[k] < 1 || inpoel[k] > npoin) return fail("cfdb_tile_elements: inpoel entry out of range");
    vector<int32_t> esup1, esup2, eslot;
    topo::build_esup(inpoel, nelem, npoin, esup1, esup2, &eslot);
    vector<uint8_t> bcf((size_t)npoin, 0);
    topo::Tiling T;
    topo::build_tiling(inpoel, nelem, npoin, X, Y, esup1, esup2, eslot, bcf, TE, order, T);
    if (!T.rank_overflow) {   // independent check of the summation schedule against the mesh (host_topology.h: check_tiling)
        const long bad = topo::check_tiling(inpoel, nelem, npoin, esup1, esup2, eslot, T);
        if (bad) return fail("cfdb_tile_elements: the tiling is inconsistent with the mesh (" + std::to_string(bad) + " violations)");
    }
    if (i2e) std::copy(T.i2e.begin(), T.i2e.end(), i2e);
    if (stats) {
        stats[0] = T.interior_fraction; stats[1] = T.ntiles; stats[2] = T.L.ntn_max; stats[3] = T.L.nint_max;
        stats[4] = T.L.nslot_max; stats[5] = T.L.tb_bytes; stats[6] = (double)T.bnodes.size(); stats[7] = (double)T.orphans.size();
    }
    return 0;
}

// greedy first-fit colouring of the element list, host code: ascending element order, a 64-bit forbidden
// mask per node, colour = lowest bit clear in the union of the three masks
extern "C" int cfdb_color_elements(const int32_t* inpoel, int32_t nelem, int32_t npoin, int32_t* color, int32_t* ncolors) {
    vector<uint64_t> mask((size_t)npoin, 0);
    int nc = 0;
    for (int e = 0; e < nelem; ++e) {
        const int32_t* t = inpoel + 3 * (size_t)e;
        for (int i = 0; i < 3; ++i)
            if (t[i] < 1 || t[i] > npoin) return fail("cfdb_color_elements: inpoel entry out of range");
        uint64_t m = mask[t[0] - 1] | mask[t[1] - 1] | mask[t[2] - 1];
        if (~m == 0) return fail("cfdb_color_elements: more than 64 colours needed (node valence above 63)");
        int col = __builtin_ctzll(~m);
        color[e] = col;
        for (int i = 0; i < 3; ++i) mask[t[i] - 1] |= 1ull << col;
        nc = std::max(nc, col + 1);
    }
    *ncolors = nc;
    return 0;
}

// generic CSR upload for bicg/spmv call-site mode (matrix need not be the context's Laplacian)
static int up_csr(cfdb_ctx* c, const double* A, const int32_t* idx, const int32_t* rowptr, int npoin, DBuf<double>& dA,
                  DBuf<int>& dIdx, DBuf<int>& dPtr) {
    int nnz = rowptr[npoin];
    vector<int32_t> i0(idx, idx + nnz);
    for (auto& v : i0) {
        if (v < 1 || v > npoin) return fail("CSR column index out of range");
        v -= 1;
    }
    TRY(upload(c, dA, A, (size_t)nnz));
    TRY(upload(c, dIdx, i0));
    TRY(upload(c, dPtr, rowptr, (size_t)npoin + 1));
    CK(cudaStreamSynchronize(c->st));  // i0 is a temporary
    return 0;
}

extern "C" int cfdb_spmv(cfdb_ctx* c, const double* A, const int32_t* idx, const int32_t* rowptr, const double* v,
                         double* y, int32_t npoin, int32_t npos) {
    CK(cudaSetDevice(c->device));
    if (npoin > c->npoin) return fail("cfdb_spmv: npoin larger than the context");
    if (npos != rowptr[npoin]) return fail("cfdb_spmv: npos != spRowptr(npoin+1)");
    DBuf<double> dA;
    DBuf<int> dIdx, dPtr;
    int r = up_csr(c, A, idx, rowptr, npoin, dA, dIdx, dPtr);
    if (!r) r = up_plain(c, c->tmpA.p, v, npoin);
    if (!r) {
        auto body = [&]() -> int {
            LAUNCH(K_SPMV, k::spmv, grid_for(npoin, 256), 256, npoin, dA.p, dIdx.p, dPtr.p, c->tmpA.p, c->tmpB.p);
            TRY(down_plain(c, c->tmpB.p, y, npoin));
            CK(cudaStreamSynchronize(c->st));
            return 0;
        };
        r = body();
    }
    dA.release(); dIdx.release(); dPtr.release();
    return r;
}

extern "C" int cfdb_vecdot(cfdb_ctx* c, int32_t n, const double* x, const double* y, double* result) {
    CK(cudaSetDevice(c->device));
    if (n > c->npoin) return fail("cfdb_vecdot: n larger than the context");
    TRY(up_plain(c, c->tmpA.p, x, n));
    TRY(up_plain(c, c->tmpB.p, y, n));
    if (n == 0) { *result = 0.0; return 0; }
    TRY(dev_dot(c, n, c->tmpA.p, c->tmpB.p, 0));
    TRY(read_scal(c));
    *result = c->h_sc->red[0];
    return 0;
}

extern "C" int cfdb_bicg(cfdb_ctx* c, const double* A, const int32_t* idx, const int32_t* rowptr, const double* diag,
                         double* x, const double* b, const double* x_fix, const int32_t* fixIdx, int32_t npoin,
                         int32_t nfix, int32_t* iters) {
    CK(cudaSetDevice(c->device));
    if (npoin != c->npoin) return fail("cfdb_bicg: npoin differs from the context");
    DBuf<double> dA, dfix;
    DBuf<int> dIdx, dPtr, dFixIdx, dFixLast;
    auto body = [&]() -> int {
        TRY(up_csr(c, A, idx, rowptr, npoin, dA, dIdx, dPtr));
        vector<int32_t> f0(fixIdx, fixIdx + nfix), last;
        for (int v : f0)
            if (v < 1 || v > npoin) return fail("cfdb_bicg: x_fixIdx out of range");
        topo::last_wins(fixIdx, nfix, npoin, last);
        for (auto& v : f0) v -= 1;
        TRY(upload(c, dFixIdx, f0));
        TRY(upload(c, dFixLast, last));
        TRY(upload(c, dfix, x_fix, (size_t)nfix));
        TRY(up_plain(c, c->tmpA.p, x, npoin));
        TRY(up_plain(c, c->tmpB.p, b, npoin));
        TRY(up_plain(c, c->tmpC.p, diag, npoin));
        CK(cudaStreamSynchronize(c->st));
        int it = 0;
        TRY(bicg_dev(c, dA.p, dIdx.p, dPtr.p, c->tmpC.p, c->tmpA.p, c->tmpB.p, dfix.p, dFixIdx.p, dFixLast.p, npoin, nfix, &it));
        *iters = it;
        TRY(down_plain(c, c->tmpA.p, x, npoin));
        CK(cudaStreamSynchronize(c->st));
        return 0;
    };
    int r = body();
    dA.release(); dfix.release(); dIdx.release(); dPtr.release(); dFixIdx.release(); dFixLast.release();
    return r;
}

extern "C" int cfdb_gcl_main(cfdb_ctx* c, double* M, const double* W_x, const double* W_y, const double* W_x_old,
                             const double* W_y_old, const double* area_old, const double* dNx, const double* dNy,
                             const double* area, const int32_t* inpoel, int32_t nelem, int32_t npoin, double dt) {
    CK(cudaSetDevice(c->device));
    TRY(check_mesh(c, nelem, npoin, "cfdb_gcl_main"));
    (void)inpoel; (void)W_y; (void)W_y_old;  // gcl.f90:38-41 uses W_x for both terms
    if (!W_x_old || !area_old) return fail("Faltan valores (GCL)");  // gcl.f90:23-25
    TRY(up_soa3(c, c->dNx.p, dNx));
    TRY(up_soa3(c, c->dNy.p, dNy));
    TRY(up_plain(c, c->M.p, M, npoin));
    TRY(up_plain(c, c->W_X.p, W_x, npoin));
    TRY(up_plain(c, c->W_x_old.p, W_x_old, npoin));
    TRY(up_elem(c, c->area.p, area)); c->geo_dirty = true;
    TRY(up_elem(c, c->area_old.p, area_old));
    TRY(run_gcl(c, nullptr, dt));
    TRY(down_plain(c, c->M.p, M, npoin));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}
