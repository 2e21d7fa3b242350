/* cfdb.h — C ABI of libcfdb200.so: the B200 (sm_100a) implementation of the per-timestep hot
 * path of chanshing/cfd.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions (identical to the Fortran side, so a Fortran caller passes its arrays as they are):
 *   - arrays are column-major: U(4,npoin) = 4 consecutive doubles per node, inpoel(3,nelem) and
 *     dNx/dNy(3,nelem) = 3 consecutive values per element;
 *   - node / element ids are 1-based int32;
 *   - CSR (Mlaplace): rowptr holds 0-based offsets (rowptr(1)=0), column ids are 1-based and the
 *     diagonal entry is stored first in each row (mLaplace.f90:82-93);
 *   - every function returns 0 on success, non-zero on error (cfdb_last_error() has the text).
 *     The reference has no error convention: its failures STOP the program (dataLoader.f90:225,284,
 *     gcl.f90:25, ns2DComp.ALE.f90:209); the Fortran shim turns a non-zero return into STOP.
 *
 * Two operating modes:
 *   (ii) resident   cfdb_create / cfdb_init / cfdb_step / cfdb_get — state lives in HBM, the
 *                   time loop of ns2DComp.ALE.f90:138-282 runs on the GPU; performance numbers
 *                   come from here.
 *   (i)  call-site  cfdb_calcrhs, cfdb_fuente, cfdb_deltat, cfdb_estab, cfdb_deriv, cfdb_masas,
 *                   cfdb_normales, cfdb_laplace, cfdb_bicg, cfdb_spmv, cfdb_gcl_main — one
 *                   function per reference subroutine, HOST pointers in and out (H2D, kernels,
 *                   D2H inside the call).  A context is bound to one connectivity, like the
 *                   reference's SAVEd first-call state (mLaplace.f90:20, pointNeighbor.f90).
 *
 * Results are bit-identical to the CPU oracle (oracle/oracle.cpp) for every array: kernels are
 * compiled without FMA contraction, evaluate the source's expression order, and accumulate
 * element->node sums in ascending element order (the reference at one OpenMP thread).
 * There is no CPU fallback: every entry point fails with an error if no sm_100 device is usable.
 */
#ifndef CFDB_H
#define CFDB_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cfdb_ctx cfdb_ctx;

/* module InputData after readInputData (dataLoader.f90:1-65); CTE is already 1/file value (:59) */
typedef struct cfdb_params {
    double FSAFE, U_inf, V_inf, MACH_inf, T_inf, RHO_inf, P_inf, C_inf;
    double FMU, FGX, FGY, QH, FK, FR, FCv, GAMA, CTE;
    double XREF[10], YREF[10];
    int32_t IRESTART, MAXITER, IPRINT, MOVIE, ITLOCAL, MOVING, NGAS;
    int32_t use_gcl; /* 0 = reference behaviour (gcl.f90 is never called) */
} cfdb_params;

/* boundary-condition lists of module MeshData after loadMeshData (dataLoader.f90:79-91,121-268) */
typedef struct cfdb_bc {
    int32_t nfixrho;   const int32_t* ifixrho_node; const double* rfixrho_value;
    int32_t nfixv;     const int32_t* ifixv_node;   const double* rfixv_valuex; const double* rfixv_valuey;
    int32_t nwall;     const int32_t* wall;         /* wall(2,nwall) */
    int32_t nfixt;     const int32_t* ifixt_node;   const double* rfixt_value;
    int32_t nsets;     const int32_t* iset_n1; const int32_t* iset_n2; const int32_t* iset_elem; const int32_t* iset_id;
    int32_t nmove;     const int32_t* i_m;
    int32_t nfix_move; const int32_t* ifm;
} cfdb_bc;

const char* cfdb_last_error(void);
int cfdb_device_count(void);

/* ---- (ii) resident mode ------------------------------------------------------------------ */
/* allocateMeshData (dataLoader.f90:288-321) + upload.  bc may be NULL (no lists). */
int cfdb_create(cfdb_ctx** out, const cfdb_params* par, int32_t npoin, int32_t nelem, const double* X,
                const double* Y, const int32_t* inpoel, const cfdb_bc* bc, int device);
void cfdb_destroy(cfdb_ctx* ctx);
/* ns2DComp.ALE.f90:59-136 minus smoothing: GAMM=GAMA, RESTART free-stream branch (:404-420),
 * getEsup/getPsup, NORMALES, DERIV, MASAS, laplace, W=0, loop scalars. */
int cfdb_init(cfdb_ctx* ctx);
/* nsteps passes of the time loop ns2DComp.ALE.f90:138-282 (asynchronous on the context stream
 * unless the mesh moves, in which case each step reads DTMIN back for the host-side pitching law). */
int cfdb_step(cfdb_ctx* ctx, int32_t nsteps);
int cfdb_sync(cfdb_ctx* ctx);
/* individual pieces of the loop on the resident state (same names as the reference routines) */
int cfdb_rk_stage(cfdb_ctx* ctx, int32_t irk);               /* body of RK's IRK loop, subrutinas.f90:667-828 */
int cfdb_geometry(cfdb_ctx* ctx, int32_t moving_step);       /* NORMALES, DERIV, MASAS[, gcl], laplace */
int cfdb_fluid_structure(cfdb_ctx* ctx, double dtmin, double time); /* meshMove.f90:28-142 */
int cfdb_residual_norms(cfdb_ctx* ctx, double er[4], double err[4]); /* ns2DComp.ALE.f90:191-197, evaluated now */
/* FORCE_VISC (ns2DComp.ALE.f90:819-893; called by the time loop on print steps when FMU /= 0, :228-233): pressure + viscous
 * traction on the ISET body edges from the resident P, T, VEL_X, VEL_Y, X, Y, dNx, dNy.  Results are the fields
 * "F_VX"(10), "F_VY"(10) and the SKIN.DAT columns "skin", "skin_x", "skin_p" (one entry per ISET edge, set by set);
 * FORCES' "FX", "FY", "RM" (10 each, meshMove.f90:153-194) are fields too.  cfdb_step calls it at the same place. */
int cfdb_force_visc(cfdb_ctx* ctx);
/* PRINTFLAVIA (ns2DComp.ALE.f90:701-817, call site :225-226): write the GiD result blocks of the resident state
 * (velocities relative to the mesh, X1/Y1 as positions) with the reference's FORMATs.  flags[7] = RHO, VEL2, MACH, PRES,
 * TEMP, ENER, POS (1 where <name>-1.dat says '.si.'); append = 1 for MOVIE runs (one file, a block per print step). */
int cfdb_printflavia(cfdb_ctx* ctx, const char* path, int32_t iter, const int32_t flags[7], int32_t append);
/* one record of <name>.cnv as '(I7, 5E14.6)' (the reference's '(I7, 4E14.6)' is one slot short) */
int cfdb_format_cnv(int32_t iter, double time, const double r[4], char* buf, int32_t buflen);
/* one real laid out as Fortran Ew.d (kind 'E') or Fw.d (kind 'F') */
int cfdb_format_real(int32_t kind, double v, int32_t w, int32_t d, char* buf, int32_t buflen);
/* the norms cfdb_step evaluated on its last print step (ITERPRINT==IPRINT or ITER==MAXITER, :186), i.e. before U=U1 */
int cfdb_step_norms(cfdb_ctx* ctx, double er[4], double err[4]);
/* field transfer by Fortran variable name ("U","U1","RHS","T","VEL_X","X","inpoel","esup1","lap_idx",...);
 * count = number of elements of the host buffer (checked).  Layout/1-basing as in the header comment. */
int cfdb_get(cfdb_ctx* ctx, const char* name, void* host, int64_t count);
int cfdb_set(cfdb_ctx* ctx, const char* name, const void* host, int64_t count);
int64_t cfdb_field_size(cfdb_ctx* ctx, const char* name);    /* elements; <0 if unknown */
int cfdb_get_scalar(cfdb_ctx* ctx, const char* name, double* value); /* TIME DTMIN DTMIN1 HMIN ITER BANDERA n_m bicg_x bicg_y FX1 FY1 RM1 */
int cfdb_set_scalar(cfdb_ctx* ctx, const char* name, double value);
/* switches beyond the reference's behaviour ("next" rows of SURVEY.md §8f), all default 0:
 *   "use_cuarto" 1: keep CUARTO_ORDEN's projection as theta instead of UN = 0.0 (subrutinas.f90:673-674, F7)
 *   "true_rk"    1: RK stages 2..4 evaluate calcRHS/FUENTE at U1 instead of U (subrutinas.f90:685,697, F6)
 *   "fast"       1: relaxed stage — FMA contraction and red.global.add.f64 scatter straight into RHS, no staging buffer,
 *                   summation order undefined.  Agrees with the default to ~1e-15 per call (meets the 1e-11 per-step
 *                   tolerance) but is NOT bit-exact, so long runs diverge from the reference. */
int cfdb_set_option(cfdb_ctx* ctx, const char* name, int32_t value);
/* CUDA stream the context launches on (cudaStream_t as void*), for event timing by the caller */
void* cfdb_stream(cfdb_ctx* ctx);
/* per-kernel timing: enable, run, then read accumulated device milliseconds and launch counts */
int cfdb_profile_enable(cfdb_ctx* ctx, int32_t on);
int cfdb_profile_get(cfdb_ctx* ctx, const char* kernel, double* total_ms, int64_t* launches);
int64_t cfdb_launch_count(cfdb_ctx* ctx);                    /* kernels launched since create */

/* ---- multi-GPU: one context per rank/GPU on a sub-domain built by cfd_b200/partition.py ------------
 * (new: the reference is single-process.)  Rank r computes every element touching a node it owns, so
 * owned-node sums are complete and bit-identical to the single-GPU run; ghost nodes are refreshed from their
 * owners after every RK stage (ncclSend/ncclRecv, one packed message per neighbour); DTMIN is an
 * ncclAllReduce(min); biCG inner products and the residual norms are canonical sums over the GLOBAL node index when the
 * ownership is chunk-aligned (cfdb_set_reduction_layout: bit-identical to one GPU), else ncclAllReduce(sum) of per-rank
 * canonical sums over owned nodes (round-off level).  Local numbering: owned nodes first. */
int cfdb_nccl_unique_id(void* out128);                       /* ncclGetUniqueId on one rank; broadcast it yourself */
int cfdb_comm_init(cfdb_ctx* ctx, const void* uid128, int32_t rank, int32_t nranks);
int cfdb_set_halo(cfdb_ctx* ctx, int32_t n_owned, int32_t nneigh, const int32_t* neigh_rank,
                  const int32_t* send_ptr, const int32_t* send_idx, const int32_t* recv_ptr,
                  const int32_t* recv_idx);                  /* 0-based local node ids, CSR per neighbour */
int cfdb_halo_exchange(cfdb_ctx* ctx, const char* field);    /* refresh the ghosts of one nodal field ("T", "U", ...) */
/* Chunk-aligned ownership (cfd_b200/partition.py): this rank's owned nodes are the global nodes [gid0, gid0 + n_owned) with
 * gid0 a multiple of 4096, the first-level chunk of the canonical reduction order.  The ranks then exchange chunk sums
 * (one ncclAllReduce over a zero-filled global array: exact) and every rank runs the upper tree levels itself, so biCG's
 * inner products and the residual norms carry the bits of the single-GPU run.  Call after cfdb_set_halo. */
int cfdb_set_reduction_layout(cfdb_ctx* ctx, int64_t gid0, int64_t npoin_global);

/* ---- (i) call-site mode: one entry point per reference subroutine, host pointers ------------ */
/* calcRHS_mod::calcRHS, calcRHS.f90:4 (module inputs FCV,FK,FMU,gama,T_inf,cte and T(:) made explicit) */
int cfdb_calcrhs(cfdb_ctx* ctx, double* rhs, const double* U, const double* theta, const double* T,
                 const double* dNx, const double* dNy, const double* area, const double* shoc,
                 const double* dtl, const double* t_sugn1, const double* t_sugn2, const double* t_sugn3,
                 const int32_t* inpoel, int32_t nelem, int32_t npoin, double Cv, double lambda_ref,
                 double mu_ref, double gamma0, double T_inf, double cte);
/* FUENTE(dtl), subrutinas.f90:1036 (module U, W_X, W_Y, dNx, dNy, area, inpoel, RHS made explicit) */
int cfdb_fuente(cfdb_ctx* ctx, double* rhs, const double* U, const double* w_x, const double* w_y,
                const double* dNx, const double* dNy, const double* area, const double* dtl,
                const int32_t* inpoel, int32_t nelem, int32_t npoin);
/* deltat(dtmin, dt), subrutinas.f90:155 */
int cfdb_deltat(cfdb_ctx* ctx, double* dtmin, double* dt, const int32_t* inpoel, const double* area,
                const double* T, const double* vel_x, const double* vel_y, const double* w_x, const double* w_y,
                int32_t nelem, int32_t npoin, double FSAFE, double FR, double GAMA, double T_inf);
/* ESTAB(U,T,GAMA,FR,RMU,DTMIN,RHOINF,TINF,UINF,VINF,GAMM), subrutinas.f90:332 */
int cfdb_estab(cfdb_ctx* ctx, const double* U, const double* T, const double* vel_x, const double* vel_y,
               const double* w_x, const double* w_y, const double* GAMM, const double* dNx, const double* dNy,
               const int32_t* inpoel, int32_t nelem, int32_t npoin, double FR, double DTMIN, double RHOINF,
               double TINF, double* shoc, double* t_sugn1, double* t_sugn2, double* t_sugn3);
/* deriv(hmin), subrutinas.f90:88 */
int cfdb_deriv(cfdb_ctx* ctx, const double* X, const double* Y, const int32_t* inpoel, int32_t nelem,
               int32_t npoin, double* area, double* HH, double* HHX, double* HHY, double* dNx, double* dNy,
               double* hmin);
/* MASAS(), subrutinas.f90:128 */
int cfdb_masas(cfdb_ctx* ctx, const double* area, const int32_t* inpoel, int32_t nelem, int32_t npoin, double* M);
/* Mnormales::normales, subrutinas.f90:7 — returns m through *m_out */
int cfdb_normales(cfdb_ctx* ctx, const int32_t* wall, int32_t nwall, const double* X, const double* Y,
                  int32_t npoin, int32_t* m_out, int32_t* n_ipoin, double* n_x, double* n_y);
/* Mlaplace::laplace, mLaplace.f90:7 (pattern from cfdb_get "lap_idx"/"lap_rowptr") */
int cfdb_laplace(cfdb_ctx* ctx, const int32_t* inpoel, const double* area, const double* dNx, const double* dNy,
                 const double* X, const double* Y, int32_t nelem, int32_t npoin, double* lap_sparse,
                 double* lap_diag);
/* BiconjGrad::biCG, biconjGrad.f90:8 — *iters = iterations of the while loop, -1 on the early return */
int cfdb_bicg(cfdb_ctx* ctx, const double* spMtx, const int32_t* spIdx, const int32_t* spRowptr,
              const double* diagMtx, double* x, const double* b, const double* x_fix, const int32_t* x_fixIdx,
              int32_t npoin, int32_t nfix, int32_t* iters);
/* BiconjGrad::SpMV, biconjGrad.f90:171 */
int cfdb_spmv(cfdb_ctx* ctx, const double* spMtx, const int32_t* spIdx, const int32_t* spRowptr,
              const double* v, double* y, int32_t npoin, int32_t npos);
/* BiconjGrad::vecdot, biconjGrad.f90:153 (canonical reduction order) */
int cfdb_vecdot(cfdb_ctx* ctx, int32_t n, const double* x, const double* y, double* result);
/* gcl_mod::main / putW / putArea, gcl.f90:8-62 (assumed-shape dummies get explicit extents) */
int cfdb_gcl_main(cfdb_ctx* ctx, double* M, const double* W_x, const double* W_y, const double* W_x_old,
                  const double* W_y_old, const double* area_old, const double* dNx, const double* dNy,
                  const double* area, const int32_t* inpoel, int32_t nelem, int32_t npoin, double dt);

/* smoothing_mod::smoothing(X, Y, inpoel, fixed, npoin, nelem), smoothing.f90:21 — the init-time mesh optimiser the
 * driver applies once before the time loop (ns2DComp.ALE.f90:76).  Host code (serial Gauss-Seidel by construction);
 * X, Y are updated in place, *sweeps returns the number of outer sweeps (0: nothing to smooth). */
int cfdb_smoothing(double* X, double* Y, const int32_t* inpoel, const unsigned char* fixed, int32_t npoin,
                   int32_t nelem, int32_t* sweeps);

/* ---- device self-test of the exact-arithmetic helpers (cfd_b200/csrc/exact.cuh) against the plain IEEE operations:
 * which = 0 shared-reciprocal division, 1 division by three, 2 zero-numerator division, 3 exact scalings by 0, 1/2, 2
 * folded into one fma, 4/5/6 the branch-free division, square root (and the two powers built on it) and x/3 with their
 * fast-path flag; n random operand pairs. */
int cfdb_selftest(cfdb_ctx* ctx, int32_t which, int64_t n, uint64_t seed, int64_t* mismatches);

/* ---- host-side integer artefacts (bit-exact vs the oracle) --------------------------------- */
/* PointNeighbor::getEsup / getPsup, pointNeighbor.f90:5-91.  psup1 needs capacity >= returned count. */
int cfdb_get_esup(const int32_t* inpoel, int32_t nelem, int32_t npoin, int32_t* esup1, int32_t* esup2);
int cfdb_get_psup(const int32_t* inpoel, int32_t nelem, int32_t npoin, int32_t* psup1, int32_t cap,
                  int32_t* psup2, int32_t* count);

#ifdef __cplusplus
}
#endif
#endif /* CFDB_H */
