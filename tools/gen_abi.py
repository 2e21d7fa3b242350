#!/usr/bin/env python
"""ONE table of the C ABI of libcfdb200.so -> three generated files:

    include/cfdb.h            C prototypes (+ the hand-written preamble and struct definitions kept below)
    fortran/cfdb_iface.f90    ISO_C_BINDING interfaces, one per entry point
    cfd_b200/_abi.py          ctypes argtypes / restype, imported by cfd_b200/capi.py

    python tools/gen_abi.py            rewrite the three files
    python tools/gen_abi.py --check    exit 1 if any of them differs from what the table generates (tests/test_abi.py)

Argument kinds (C type / Fortran dummy / ctypes):
    ctx      cfdb_ctx*            type(c_ptr), value                c_void_p
    ctxout   cfdb_ctx**           type(c_ptr)            (by ref)   POINTER(c_void_p)
    par      const cfdb_params*   type(cfdb_params)      (by ref)   POINTER(Params)
    bc       const cfdb_bc*       type(cfdb_bc)          (by ref)   POINTER(BC)
    d i32 i64 u64 int             <kind>, value                      c_double ...
    cd[] d[] ci32[] i32[] cu8[]   assumed-size array (shape given)   ndpointer
    d* i32* i64*                  scalar by reference                POINTER(...) (or ndpointer: 'np')
    cstr     const char*          character(kind=c_char) :: x(*)     c_char_p
    buf      char*                character(kind=c_char) :: x(*)     c_char_p
    cvoid/void  (const) void*     type(c_ptr), value                 c_void_p
"""
import os
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def A(name, kind, shape="(*)", np=False, null=False):
    """null=True: the C side accepts NULL for this array (ctypes: plain c_void_p instead of an ndpointer)"""
    return dict(name=name, kind=kind, shape=shape, np=np, null=null)


def arrs(kind, names, shape="(*)"):
    return [A(n, kind, shape) for n in names.split()]


CTX = A("ctx", "ctx")
SECTIONS = []  # (title comment, [functions])


def sec(title):
    SECTIONS.append((title, []))


def fn(name, ret, args, doc=""):
    SECTIONS[-1][1].append(dict(name=name, ret=ret, args=args, doc=doc))


# =====================================================================================================================
sec(None)
fn("cfdb_last_error", "cstr", [])
fn("cfdb_device_count", "int", [])

sec("---- (ii) resident mode ------------------------------------------------------------------")
fn("cfdb_create", "int", [A("out", "ctxout"), A("par", "par"), A("npoin", "i32"), A("nelem", "i32"), A("X", "cd[]"), A("Y", "cd[]"),
                          A("inpoel", "ci32[]", "(3,*)"), A("bc", "bc"), A("device", "int")],
   "allocateMeshData (dataLoader.f90:288-321) + upload.  bc may be NULL (no lists).  Fails when NGAS /= 0 (the\n"
   "equilibrium-air TGAS branch, subrutinas.f90:706-741, is not implemented).")
fn("cfdb_destroy", "void_", [CTX])
fn("cfdb_init", "int", [CTX],
   "ns2DComp.ALE.f90:59-136 minus smoothing: GAMM=GAMA, RESTART free-stream branch (:404-420),\n"
   "getEsup/getPsup, NORMALES, DERIV, MASAS, laplace, W=0, loop scalars.")
fn("cfdb_step", "int", [CTX, A("nsteps", "i32")],
   "nsteps passes of the time loop ns2DComp.ALE.f90:138-282 (asynchronous on the context stream; a fixed-mesh step is one\n"
   "CUDA-graph launch; with body sets each step reads DTMIN/TIME back for the host-side pitching law).")
fn("cfdb_sync", "int", [CTX])
fn("cfdb_rk_stage", "int", [CTX, A("irk", "i32")], "body of RK's IRK loop on the resident state, subrutinas.f90:667-828")
fn("cfdb_geometry", "int", [CTX, A("moving_step", "i32")], "NORMALES, DERIV, MASAS[, gcl], laplace (ns2DComp.ALE.f90:259-275)")
fn("cfdb_fluid_structure", "int", [CTX, A("dtmin", "d"), A("time", "d")], "meshMove.f90:28-142 on the resident state")
fn("cfdb_residual_norms", "int", [CTX, A("er", "d[]", "(4)"), A("err", "d[]", "(4)")], "ns2DComp.ALE.f90:191-197, evaluated now")
fn("cfdb_force_visc", "int", [CTX],
   "FORCE_VISC (ns2DComp.ALE.f90:819-893; called by the time loop on print steps when FMU /= 0, :228-233): pressure + viscous\n"
   "traction on the ISET body edges from the resident P, T, VEL_X, VEL_Y, X, Y, dNx, dNy.  Results are the fields\n"
   "\"F_VX\"(10), \"F_VY\"(10) and the SKIN.DAT columns \"skin\", \"skin_x\", \"skin_p\" (one entry per ISET edge, set by set);\n"
   "FORCES' \"FX\", \"FY\", \"RM\" (10 each, meshMove.f90:153-194) are fields too.  cfdb_step calls it at the same place.")
fn("cfdb_printflavia", "int", [CTX, A("path", "cstr"), A("iter", "i32"), A("flags", "ci32[]", "(7)"), A("append", "i32")],
   "PRINTFLAVIA (ns2DComp.ALE.f90:701-817, call site :225-226): write the GiD result blocks of the resident state\n"
   "(velocities relative to the mesh, X1/Y1 as positions) with the reference's FORMATs.  flags[7] = RHO, VEL2, MACH, PRES,\n"
   "TEMP, ENER, POS (1 where <name>-1.dat says '.si.'); append = 1 for MOVIE runs (one file, a block per print step).")
fn("cfdb_write_forces", "int", [CTX, A("path", "cstr")],
   "the FORCES file of a print step (ns2DComp.ALE.f90:238-250): per body set 'SET NUMERO' (A, I2) and the pressure, viscous\n"
   "and total forces as (A, E14.5) records, from the resident FX, FY, F_VX, F_VY")
fn("cfdb_write_desplazamiento", "int", [CTX, A("path", "cstr"), A("time", "d"), A("append", "i32")],
   "one '(7E13.5)' record of DESPLAZAMIENTO (ns2DComp.ALE.f90:237): TIME, F_VX(1), F_VY(1), RM(1), F_VX(2), F_VY(2), RM(2)")
fn("cfdb_write_skin", "int", [CTX, A("path", "cstr")],
   "SKIN.DAT (ns2DComp.ALE.f90:833,888): one LIST-DIRECTED record (Cf, edge mid x, p/82713.27) per ISET edge of the last\n"
   "FORCE_VISC, in gfortran's list-directed layout for REAL(8) (1PG25.17E3 semantics: 17 significant digits, width 25, one\n"
   "leading blank per record) -- list-directed output is processor-dependent, this is the layout of the build line the\n"
   "oracle assumes (fortran/README.md)")
fn("cfdb_format_cnv", "int", [A("iter", "i32"), A("time", "d"), A("r", "cd[]", "(4)"), A("buf", "buf"), A("buflen", "i32")],
   "one record of <name>.cnv as '(I7, 5E14.6)' (the reference's '(I7, 4E14.6)' is one slot short, SURVEY.md F14)")
fn("cfdb_format_real", "int", [A("kind", "i32"), A("v", "d"), A("w", "i32"), A("d", "i32"), A("buf", "buf"), A("buflen", "i32")],
   "one real laid out as Fortran Ew.d (kind 'E'), Fw.d (kind 'F') or as a list-directed REAL(8) item (kind 'L'; w, d ignored)")
fn("cfdb_step_norms", "int", [CTX, A("er", "d[]", "(4)"), A("err", "d[]", "(4)")],
   "the norms cfdb_step evaluated on its last print step (ITERPRINT==IPRINT or ITER==MAXITER, :186), i.e. before U=U1")
fn("cfdb_get", "int", [CTX, A("name", "cstr"), A("host", "void"), A("count", "i64")],
   "field transfer by Fortran variable name (\"U\",\"U1\",\"RHS\",\"T\",\"VEL_X\",\"X\",\"inpoel\",\"esup1\",\"lap_idx\",...);\n"
   "count = number of elements of the host buffer (checked).  Layout/1-basing as in the header comment; element arrays in\n"
   "the mesh file's element order (the library's internal tile order never shows).")
fn("cfdb_set", "int", [CTX, A("name", "cstr"), A("host", "cvoid"), A("count", "i64")])
fn("cfdb_field_size", "i64", [CTX, A("name", "cstr")], "elements; <0 if unknown")
fn("cfdb_get_scalar", "int", [CTX, A("name", "cstr"), A("value", "d*")],
   "TIME DTMIN DTMIN1 HMIN ITER BANDERA n_m bicg_x bicg_y FX1 FY1 RM1 tile_interior graph_replays")
fn("cfdb_set_scalar", "int", [CTX, A("name", "cstr"), A("value", "d")])
fn("cfdb_set_option", "int", [CTX, A("name", "cstr"), A("value", "i32")],
   "switches beyond the reference's behaviour (\"next\" rows of SURVEY.md 8f), all default 0:\n"
   "  \"use_cuarto\" 1: keep CUARTO_ORDEN's projection as theta instead of UN = 0.0 (subrutinas.f90:673-674, F7)\n"
   "  \"true_rk\"    1: RK stages 2..4 evaluate calcRHS/FUENTE at U1 instead of U (subrutinas.f90:685,697, F6)\n"
   "  \"adamsb\"     1: ADAMSB (subrutinas.f90:851-1034) replaces RK once three RHS history levels exist, as the commented-out\n"
   "                  call at ns2DComp.ALE.f90:176-178 intends\n"
   "  \"fast\"       1: relaxed stage -- FMA contraction and red.global.add.f64 scatter straight into RHS, no staging buffer,\n"
   "                  summation order undefined.  Agrees with the default to ~1e-15 per call (meets the 1e-11 per-step\n"
   "                  tolerance) but is NOT bit-exact, so long runs diverge from the reference (DESIGN.md section 2).\n"
   "  \"colored\"    1: relaxed stage with a deterministic coloured scatter (SURVEY.md B.3: greedy first-fit colours of the\n"
   "                  element list): reproducible from run to run, not bit-identical to the reference's 1-thread order\n"
   "  \"ale\"        1: the mesh moves although this context holds no body set (ranks of a multi-GPU run)")
fn("cfdb_stream", "void*", [CTX], "CUDA stream the context launches on (cudaStream_t as void*), for event timing by the caller")
fn("cfdb_profile_enable", "int", [CTX, A("on", "i32")], "per-kernel timing: enable, run, then read accumulated device milliseconds and launch counts")
fn("cfdb_profile_get", "int", [CTX, A("kernel", "cstr"), A("total_ms", "d*"), A("launches", "i64*")])
fn("cfdb_launch_count", "i64", [CTX], "kernels launched since create")
fn("cfdb_step_streamed", "int", [CTX, A("in_U", "cd[]", "(4,*)", null=True), A("in_T", "cd[]", null=True), A("in_VEL_X", "cd[]", null=True),
                                 A("in_VEL_Y", "cd[]", null=True), A("out_U", "d[]", "(4,*)", null=True), A("out_T", "d[]", null=True),
                                 A("out_VEL_X", "d[]", null=True), A("out_VEL_Y", "d[]", null=True), A("out_norms", "d[]", "(8)", null=True)],
   "One time step whose state comes from, and goes back to, HOST arrays (pinned memory for full speed), pipelined: the\n"
   "call returns at once; the upload of this call's inputs runs on a copy stream while the previous step computes, the\n"
   "download of its results while the next one does (device-side staging buffers on both sides, full duplex).  The\n"
   "host arrays must stay untouched until cfdb_streamed_wait().  Any of the pointers may be NULL (that array is not\n"
   "transferred).  out_norms = ER(4), ERR(4) of this step.")
fn("cfdb_streamed_wait", "int", [CTX], "wait until every cfdb_step_streamed call so far has delivered its outputs")

sec("---- multi-GPU: one context per rank/GPU on a sub-domain built by cfd_b200/partition.py ------------\n"
    " * (new: the reference is single-process.)  Rank r computes every element touching a node it owns, so\n"
    " * owned-node sums are complete and bit-identical to the single-GPU run; ghost nodes are refreshed from their\n"
    " * owners after every RK stage (one packed message per neighbour); DTMIN is an\n"
    " * ncclAllReduce(min); biCG inner products and the residual norms are canonical sums over the GLOBAL node index when the\n"
    " * ownership is chunk-aligned (cfdb_set_reduction_layout: bit-identical to one GPU), else ncclAllReduce(sum) of per-rank\n"
    " * canonical sums over owned nodes (round-off level).  Local numbering: owned nodes first.")
fn("cfdb_nccl_unique_id", "int", [A("out128", "void")], "ncclGetUniqueId on one rank; broadcast it yourself")
fn("cfdb_comm_init", "int", [CTX, A("uid128", "cvoid"), A("rank", "i32"), A("nranks", "i32")])
fn("cfdb_set_halo", "int", [CTX, A("n_owned", "i32"), A("nneigh", "i32"), A("neigh_rank", "ci32[]"), A("send_ptr", "ci32[]"),
                            A("send_idx", "ci32[]"), A("recv_ptr", "ci32[]"), A("recv_idx", "ci32[]")],
   "0-based local node ids, CSR per neighbour")
fn("cfdb_halo_exchange", "int", [CTX, A("field", "cstr")], "refresh the ghosts of one nodal field (\"T\", \"U\", ...)")
fn("cfdb_set_reduction_layout", "int", [CTX, A("gid0", "i64"), A("npoin_global", "i64")],
   "Chunk-aligned ownership (cfd_b200/partition.py): this rank's owned nodes are the global nodes [gid0, gid0 + n_owned) with\n"
   "gid0 a multiple of 4096, the first-level chunk of the canonical reduction order.  The ranks then exchange chunk sums\n"
   "(one ncclAllReduce over a zero-filled global array: exact) and every rank runs the upper tree levels itself, so biCG's\n"
   "inner products and the residual norms carry the bits of the single-GPU run.  Call after cfdb_set_halo.")

sec("---- (i) call-site mode: one entry point per reference subroutine, host pointers ------------")
E3, P4 = "(3,*)", "(4,*)"
fn("cfdb_calcrhs", "int", [CTX, A("rhs", "d[]", P4), A("U", "cd[]", P4), A("theta", "cd[]", P4), A("T", "cd[]"), A("dNx", "cd[]", E3),
                           A("dNy", "cd[]", E3)] + arrs("cd[]", "area shoc dtl t_sugn1 t_sugn2 t_sugn3") +
   [A("inpoel", "ci32[]", E3), A("nelem", "i32"), A("npoin", "i32")] + [A(n, "d") for n in "Cv lambda_ref mu_ref gamma0 T_inf cte".split()],
   "calcRHS_mod::calcRHS, calcRHS.f90:4 (module inputs FCV,FK,FMU,gama,T_inf,cte and T(:) made explicit)")
fn("cfdb_fuente", "int", [CTX, A("rhs", "d[]", P4), A("U", "cd[]", P4), A("w_x", "cd[]"), A("w_y", "cd[]"), A("dNx", "cd[]", E3),
                          A("dNy", "cd[]", E3), A("area", "cd[]"), A("dtl", "cd[]"), A("inpoel", "ci32[]", E3), A("nelem", "i32"),
                          A("npoin", "i32")],
   "FUENTE(dtl), subrutinas.f90:1036 (module U, W_X, W_Y, dNx, dNy, area, inpoel, RHS made explicit)")
fn("cfdb_deltat", "int", [CTX, A("dtmin", "d*", np=True), A("dt", "d[]"), A("inpoel", "ci32[]", E3)] +
   arrs("cd[]", "area T vel_x vel_y w_x w_y") + [A("nelem", "i32"), A("npoin", "i32")] + [A(n, "d") for n in "FSAFE FR GAMA T_inf".split()],
   "deltat(dtmin, dt), subrutinas.f90:155")
fn("cfdb_estab", "int", [CTX, A("U", "cd[]", P4)] + arrs("cd[]", "T vel_x vel_y w_x w_y GAMM") + [A("dNx", "cd[]", E3), A("dNy", "cd[]", E3),
                         A("inpoel", "ci32[]", E3), A("nelem", "i32"), A("npoin", "i32")] + [A(n, "d") for n in "FR DTMIN RHOINF TINF".split()] +
   arrs("d[]", "shoc t_sugn1 t_sugn2 t_sugn3"),
   "ESTAB(U,T,GAMA,FR,RMU,DTMIN,RHOINF,TINF,UINF,VINF,GAMM), subrutinas.f90:332")
fn("cfdb_deriv", "int", [CTX, A("X", "cd[]"), A("Y", "cd[]"), A("inpoel", "ci32[]", E3), A("nelem", "i32"), A("npoin", "i32")] +
   arrs("d[]", "area HH HHX HHY") + [A("dNx", "d[]", E3), A("dNy", "d[]", E3), A("hmin", "d*", np=True)], "deriv(hmin), subrutinas.f90:88")
fn("cfdb_masas", "int", [CTX, A("area", "cd[]"), A("inpoel", "ci32[]", E3), A("nelem", "i32"), A("npoin", "i32"), A("M", "d[]")],
   "MASAS(), subrutinas.f90:128")
fn("cfdb_normales", "int", [CTX, A("wall", "ci32[]", "(2,*)"), A("nwall", "i32"), A("X", "cd[]"), A("Y", "cd[]"), A("npoin", "i32"),
                            A("m_out", "i32*"), A("n_ipoin", "i32[]"), A("n_x", "d[]"), A("n_y", "d[]")],
   "Mnormales::normales, subrutinas.f90:7 -- returns m through *m_out")
fn("cfdb_normalvel", "int", [CTX, A("m", "i32"), A("n_ipoin", "ci32[]"), A("n_x", "cd[]"), A("n_y", "cd[]"), A("vel_x", "d[]"),
                             A("vel_y", "d[]"), A("w_x", "cd[]"), A("w_y", "cd[]"), A("npoin", "i32")],
   "Mnormales::normalvel, subrutinas.f90:66-85 (the module-private list n_ipoin, n_x, n_y made explicit)")
fn("cfdb_fixvel", "int", [CTX, A("nfixv", "i32"), A("ifixv_node", "ci32[]"), A("rfixv_valuex", "cd[]"), A("rfixv_valuey", "cd[]"),
                          A("vel_x", "d[]"), A("vel_y", "d[]"), A("npoin", "i32")],
   "fixvel, subrutinas.f90:601-616 (duplicate nodes: the last list entry wins)")
fn("cfdb_fix", "int", [CTX, A("FR", "d"), A("GAMM", "cd[]"), A("nfixrho", "i32"), A("ifixrho_node", "ci32[]"), A("rfixrho_value", "cd[]"),
                       A("nfixt", "i32"), A("ifixt_node", "ci32[]"), A("rfixt_value", "cd[]"), A("vel_x", "cd[]"), A("vel_y", "cd[]"),
                       A("rho", "d[]"), A("T", "d[]"), A("E", "d[]"), A("npoin", "i32")],
   "FIX(FR, GAMM), subrutinas.f90:618-643")
fn("cfdb_rk", "int", [CTX, A("DTMIN", "d"), A("NRK", "i32"), A("BANDERA", "i32"), A("GAMM", "cd[]"), A("dtl", "cd[]"), A("U", "cd[]", P4),
                      A("U1", "d[]", P4), A("RHS", "d[]", P4), A("RHS1", "d[]", P4), A("RHS2", "d[]", P4), A("RHS3", "d[]", P4)] +
   arrs("d[]", "T P RHO E RMACH VEL_X VEL_Y") + [A("W_X", "cd[]"), A("W_Y", "cd[]"), A("SHOC", "d[]"), A("T_SUGN1", "d[]"),
                                                 A("T_SUGN2", "d[]"), A("T_SUGN3", "d[]"), A("nelem", "i32"), A("npoin", "i32")],
   "RK(DTMIN, NRK, BANDERA, GAMM, dtl), subrutinas.f90:645-849, with every module array it reads or writes made explicit\n"
   "(geometry and BC lists are the context's)")
fn("cfdb_laplace", "int", [CTX, A("inpoel", "ci32[]", E3), A("area", "cd[]"), A("dNx", "cd[]", E3), A("dNy", "cd[]", E3), A("X", "cd[]"),
                           A("Y", "cd[]"), A("nelem", "i32"), A("npoin", "i32"), A("lap_sparse", "d[]"), A("lap_diag", "d[]")],
   "Mlaplace::laplace, mLaplace.f90:7 (pattern from cfdb_get \"lap_idx\"/\"lap_rowptr\")")
fn("cfdb_bicg", "int", [CTX, A("spMtx", "cd[]"), A("spIdx", "ci32[]"), A("spRowptr", "ci32[]"), A("diagMtx", "cd[]"), A("x", "d[]"),
                        A("b", "cd[]"), A("x_fix", "cd[]"), A("x_fixIdx", "ci32[]"), A("npoin", "i32"), A("nfix", "i32"), A("iters", "i32*")],
   "BiconjGrad::biCG, biconjGrad.f90:8 -- *iters = iterations of the while loop, -1 on the early return")
fn("cfdb_spmv", "int", [CTX, A("spMtx", "cd[]"), A("spIdx", "ci32[]"), A("spRowptr", "ci32[]"), A("v", "cd[]"), A("y", "d[]"),
                        A("npoin", "i32"), A("npos", "i32")], "BiconjGrad::SpMV, biconjGrad.f90:171")
fn("cfdb_vecdot", "int", [CTX, A("n", "i32"), A("x", "cd[]"), A("y", "cd[]"), A("result", "d*")],
   "BiconjGrad::vecdot, biconjGrad.f90:153 (canonical reduction order, see DESIGN.md)")
fn("cfdb_gcl_main", "int", [CTX, A("M", "d[]")] + arrs("cd[]", "W_x W_y W_x_old W_y_old area_old") +
   [A("dNx", "cd[]", E3), A("dNy", "cd[]", E3), A("area", "cd[]"), A("inpoel", "ci32[]", E3), A("nelem", "i32"), A("npoin", "i32"), A("dt", "d")],
   "gcl_mod::main / putW / putArea, gcl.f90:8-62 (assumed-shape dummies get explicit extents)")
fn("cfdb_mesh_move", "int", [CTX, A("dtmin", "d"), A("time", "d"), A("X", "d[]"), A("Y", "d[]"), A("X1", "d[]"), A("Y1", "d[]"),
                             A("W_X", "d[]"), A("W_Y", "d[]"), A("P", "cd[]"), A("xpos", "d[]"), A("ypos", "d[]"),
                             A("fx", "d[]", "(10)"), A("fy", "d[]", "(10)"), A("rm", "d[]", "(10)"), A("npoin", "i32")],
   "MeshMove::fluidStructure(dtmin, time, SMOOTH_FIX, x1, y1), meshMove.f90:28-142, with the module arrays it reads or writes\n"
   "made explicit (the Laplacian is the context's: call cfdb_laplace / cfdb_geometry first when the mesh has moved)")
fn("cfdb_smoothing", "int", [A("X", "d[]"), A("Y", "d[]"), A("inpoel", "ci32[]", E3), A("fixed", "cu8[]"), A("npoin", "i32"),
                             A("nelem", "i32"), A("sweeps", "i32*")],
   "smoothing_mod::smoothing(X, Y, inpoel, fixed, npoin, nelem), smoothing.f90:21 -- the init-time mesh optimiser the\n"
   "driver applies once before the time loop (ns2DComp.ALE.f90:76).  Host code (serial Gauss-Seidel by construction);\n"
   "X, Y are updated in place, *sweeps returns the number of outer sweeps (0: nothing to smooth).")
fn("cfdb_smoothing_colored", "int", [A("X", "d[]"), A("Y", "d[]"), A("inpoel", "ci32[]", E3), A("fixed", "cu8[]"), A("npoin", "i32"),
                                     A("nelem", "i32"), A("sweeps", "i32*")],
   "(new, SURVEY.md N4; opt-in: NOT the reference's results.)  The same optimiser with the nodes visited colour by colour\n"
   "(greedy colouring, nodes sharing an element get different colours) instead of in ascending order: the nodes of one\n"
   "colour are optimised concurrently on the host's cores (OpenMP) with the reference's per-node procedure, and the result\n"
   "does not depend on the number of threads.  Same arguments as cfdb_smoothing.")

sec("---- device self-test of the exact-arithmetic helpers (cfd_b200/csrc/exact.cuh) against the plain IEEE operations:\n"
    " * which = 0 shared-reciprocal division, 1 division by three, 2 zero-numerator division, 3 exact scalings by 0, 1/2, 2\n"
    " * folded into one fma, 4/5/6 the branch-free division, square root (and the two powers built on it) and x/3 with their\n"
    " * fast-path flag; n random operand pairs.")
fn("cfdb_selftest", "int", [CTX, A("which", "i32"), A("n", "i64"), A("seed", "u64"), A("mismatches", "i64*")])

sec("---- host-side integer artefacts (bit-exact vs the oracle) ---------------------------------\n"
    " * PointNeighbor::getEsup / getPsup, pointNeighbor.f90:5-91.  psup1 needs capacity >= returned count.")
fn("cfdb_get_esup", "int", [A("inpoel", "ci32[]", E3), A("nelem", "i32"), A("npoin", "i32"), A("esup1", "i32[]"), A("esup2", "i32[]")])
fn("cfdb_get_psup", "int", [A("inpoel", "ci32[]", E3), A("nelem", "i32"), A("npoin", "i32"), A("psup1", "i32[]"), A("cap", "i32"),
                            A("psup2", "i32[]"), A("count", "i32*")])
fn("cfdb_color_elements", "int", [A("inpoel", "ci32[]", E3), A("nelem", "i32"), A("npoin", "i32"), A("color", "i32[]"), A("ncolors", "i32*")],
   "greedy first-fit colouring of the element list in ascending element order with a 64-bit forbidden mask per node\n"
   "(SURVEY.md B.3): color[e] = lowest colour not yet used at any of the element's three nodes; 0-based colours")
fn("cfdb_tile_elements", "int", [A("inpoel", "ci32[]", E3), A("nelem", "i32"), A("npoin", "i32"), A("X", "cd[]"), A("Y", "cd[]"), A("TE", "i32"),
                                 A("order", "i32"), A("i2e", "i32[]"), A("stats", "d[]")],
   "the internal element order of the fused RK stage (host_topology.h): tiles of TE elements; order 0 = the file's,\n"
   "1 = Morton runs, 2 = recursive coordinate bisection (the default of cfdb_create).  i2e[pos] = 0-based original element\n"
   "at internal position pos; stats[8] = {fraction of nodes interior to one tile, tiles, max nodes touched per tile, max\n"
   "interior nodes, max interior contributions, bytes of a tile's static block, tile-boundary nodes, orphans}.  No GPU.")

FUNCS = [f for _, fs in SECTIONS for f in fs]

# =====================================================================================================================
C_T = {"ctx": "cfdb_ctx* {n}", "ctxout": "cfdb_ctx** {n}", "par": "const cfdb_params* {n}", "bc": "const cfdb_bc* {n}",
       "d": "double {n}", "i32": "int32_t {n}", "i64": "int64_t {n}", "u64": "uint64_t {n}", "int": "int {n}",
       "cd[]": "const double* {n}", "d[]": "double* {n}", "ci32[]": "const int32_t* {n}", "i32[]": "int32_t* {n}",
       "cu8[]": "const unsigned char* {n}", "d*": "double* {n}", "i32*": "int32_t* {n}", "i64*": "int64_t* {n}",
       "cstr": "const char* {n}", "buf": "char* {n}", "cvoid": "const void* {n}", "void": "void* {n}"}
C_RET = {"int": "int", "void_": "void", "cstr": "const char*", "i64": "int64_t", "void*": "void*"}
F_KIND = {"d": "real(c_double)", "i32": "integer(c_int32_t)", "i64": "integer(c_int64_t)", "u64": "integer(c_int64_t)", "int": "integer(c_int)"}
F_ARR = {"cd[]": "real(c_double)", "d[]": "real(c_double)", "ci32[]": "integer(c_int32_t)", "i32[]": "integer(c_int32_t)",
         "cu8[]": "integer(c_int8_t)"}
F_PTR = {"d*": "real(c_double)", "i32*": "integer(c_int32_t)", "i64*": "integer(c_int64_t)"}
F_RET = {"int": "integer(c_int)", "cstr": "type(c_ptr)", "i64": "integer(c_int64_t)", "void*": "type(c_ptr)"}
PY_T = {"ctx": "vp", "ctxout": "C.POINTER(vp)", "par": "C.POINTER(Params)", "bc": "C.POINTER(BC)", "d": "d", "i32": "i32", "i64": "i64",
        "u64": "C.c_uint64", "int": "C.c_int", "cd[]": "_dp", "d[]": "_dp", "ci32[]": "_ip", "i32[]": "_ip", "cu8[]": "_u8p",
        "d*": "C.POINTER(d)", "i32*": "C.POINTER(i32)", "i64*": "C.POINTER(i64)", "cstr": "cp", "buf": "cp", "cvoid": "vp", "void": "vp"}
PY_RET = {"int": "C.c_int", "void_": "None", "cstr": "cp", "i64": "i64", "void*": "vp"}

PREAMBLE = open(os.path.join(ROOT, "tools", "cfdb_h_preamble.txt")).read()


def c_decl(a):
    if a["kind"] in ("cd[]", "d[]") and a["shape"] in ("(4)", "(8)", "(10)"):   # fixed small arrays read better as T x[n]
        return ("const " if a["kind"] == "cd[]" else "") + f"double {a['name']}[{a['shape'][1:-1]}]"
    if a["kind"] == "ci32[]" and a["shape"] == "(7)":
        return f"const int32_t {a['name']}[7]"
    return C_T[a["kind"]].format(n=a["name"])


def gen_header():
    out = [PREAMBLE.rstrip("\n"), ""]
    for title, fs in SECTIONS:
        if title:
            out.append("/* " + title + " */")
        for f in fs:
            if f["doc"]:
                lines = f["doc"].split("\n")
                out.append("/* " + "\n * ".join(lines) + " */")
            decls = [c_decl(a) for a in f["args"]] or ["void"]
            head = f"{C_RET[f['ret']]} {f['name']}("
            indent = " " * len(head)
            line = head
            for i, dcl in enumerate(decls):   # break between arguments only
                piece = dcl + (", " if i + 1 < len(decls) else ");")
                if len(line) + len(piece.rstrip()) > 118 and line.strip() != head.strip():
                    out.append(line.rstrip())
                    line = indent
                line += piece
            out.append(line.rstrip())
        out.append("")
    out += ["#ifdef __cplusplus", "}", "#endif", "#endif /* CFDB_H */", ""]
    return "\n".join(out)


def f_dummy_lines(f):
    groups = {}   # declaration text -> names (keeps first-seen order)
    for a in f["args"]:
        k = a["kind"]
        if k in ("ctx", "void", "cvoid"):
            t, n = "type(c_ptr), value", a["name"]
        elif k == "ctxout":
            t, n = "type(c_ptr)", a["name"]
        elif k == "par":
            t, n = "type(cfdb_params)", a["name"]
        elif k == "bc":
            t, n = "type(cfdb_bc)", a["name"]
        elif k in F_KIND:
            t, n = F_KIND[k] + ", value", a["name"]
        elif k in F_ARR:
            t, n = F_ARR[k], a["name"] + a["shape"]
        elif k in F_PTR:
            t, n = F_PTR[k], a["name"]
        elif k in ("cstr", "buf"):
            t, n = "character(kind=c_char)", a["name"] + "(*)"
        else:
            raise KeyError(k)
        groups.setdefault(t, []).append(n)
    lines = []
    for t, names in groups.items():
        lines.extend(textwrap.wrap(f"{t} :: " + ", ".join(names), 116, subsequent_indent="            ", break_long_words=False))
    # free-form continuation
    fixed = []
    for ln in lines:
        fixed.append(ln)
    out = []
    i = 0
    while i < len(fixed):
        ln = fixed[i]
        while i + 1 < len(fixed) and fixed[i + 1].startswith("            "):
            ln += " &\n" + fixed[i + 1]
            i += 1
        out.append(ln)
        i += 1
    return out


def gen_fortran():
    o = ["! GENERATED by tools/gen_abi.py from the one ABI table -- do not edit; edit the table and run the tool.",
         "! ISO_C_BINDING interfaces to libcfdb200.so (include/cfdb.h).  Source only: no Fortran compiler exists in the build",
         "! image (SURVEY.md F1).  tests/test_abi.py regenerates this file from the table and compares, checks every dummy's",
         "! value/reference attribute and kind against the C prototype, and runs the file through the repository's Fortran",
         "! front end (oracle/f90ref/translate.py) as a syntax check.",
         "module cfdb_iface",
         "  use iso_c_binding",
         "  implicit none",
         "  type(c_ptr), save :: cfdb_ctx = c_null_ptr   ! one context per process, bound to the mesh (like the SAVEd state of Mlaplace)",
         "",
         "  type, bind(C) :: cfdb_params",
         "     real(c_double) :: FSAFE, U_inf, V_inf, MACH_inf, T_inf, RHO_inf, P_inf, C_inf",
         "     real(c_double) :: FMU, FGX, FGY, QH, FK, FR, FCv, GAMA, CTE",
         "     real(c_double) :: XREF(10), YREF(10)",
         "     integer(c_int32_t) :: IRESTART, MAXITER, IPRINT, MOVIE, ITLOCAL, MOVING, NGAS, use_gcl",
         "  end type",
         "",
         "  type, bind(C) :: cfdb_bc",
         "     integer(c_int32_t) :: nfixrho",
         "     type(c_ptr) :: ifixrho_node, rfixrho_value",
         "     integer(c_int32_t) :: nfixv",
         "     type(c_ptr) :: ifixv_node, rfixv_valuex, rfixv_valuey",
         "     integer(c_int32_t) :: nwall",
         "     type(c_ptr) :: wall",
         "     integer(c_int32_t) :: nfixt",
         "     type(c_ptr) :: ifixt_node, rfixt_value",
         "     integer(c_int32_t) :: nsets",
         "     type(c_ptr) :: iset_n1, iset_n2, iset_elem, iset_id",
         "     integer(c_int32_t) :: nmove",
         "     type(c_ptr) :: i_m",
         "     integer(c_int32_t) :: nfix_move",
         "     type(c_ptr) :: ifm",
         "  end type",
         "",
         "  interface"]
    for f in FUNCS:
        names = ", ".join(a["name"] for a in f["args"])
        is_sub = f["ret"] == "void_"
        head = f"{'subroutine' if is_sub else 'function'} {f['name']}({names}) bind(C, name=\"{f['name']}\")" + ("" if is_sub else " result(rc)")
        hl = textwrap.wrap(head, 112, subsequent_indent="         ", break_long_words=False)
        o.append("     " + " &\n     ".join(hl))
        o.append("       import")
        for ln in f_dummy_lines(f):
            o.append("       " + ln.replace("\n", "\n       "))
        if not is_sub:
            o.append(f"       {F_RET[f['ret']]} :: rc")
        o.append("     end " + ("subroutine" if is_sub else "function"))
    o += ["  end interface",
          "contains",
          "  subroutine cfdb_check(rc, who)      ! the reference's error convention is STOP (dataLoader.f90:225,284; gcl.f90:25)",
          "    integer(c_int), intent(in) :: rc",
          "    character(*), intent(in) :: who",
          "    if (rc /= 0) then",
          "       write(*,*) 'libcfdb200 error in ', who",
          "       stop 1",
          "    end if",
          "  end subroutine",
          "end module cfdb_iface", ""]
    return "\n".join(o)


def gen_py():
    o = ['"""GENERATED by tools/gen_abi.py from the one ABI table -- do not edit."""',
         "import ctypes as C", "", "import numpy as np", "",
         "SYMBOLS = [" + ", ".join(f'"{f["name"]}"' for f in FUNCS) + "]", "",
         "# per function: (restype, [(argument name, kind, Fortran shape)])",
         "TABLE = {"]
    for f in FUNCS:
        o.append(f'    "{f["name"]}": ("{f["ret"]}", [' + ", ".join(f'("{a["name"]}", "{a["kind"]}", "{a["shape"]}")' for a in f["args"]) + "]),")
    o += ["}", "", "",
          "def bind(L, Params, BC):",
          '    """set restype / argtypes of every entry point on the loaded library L"""',
          '    _dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")',
          '    _ip = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")',
          '    _u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")',
          "    d, i32, i64, vp, cp = C.c_double, C.c_int32, C.c_int64, C.c_void_p, C.c_char_p"]
    for f in FUNCS:
        at = []
        for a in f["args"]:
            t = PY_T[a["kind"]]
            if a["np"]:
                t = "_dp"
            if a["null"]:
                t = "vp"
            at.append(t)
        o.append(f"    L.{f['name']}.restype = {PY_RET[f['ret']]}")
        ln = f"    L.{f['name']}.argtypes = [" + ", ".join(at) + "]"
        o.extend(textwrap.wrap(ln, 118, subsequent_indent="        ", break_long_words=False))
    o.append("")
    return "\n".join(o)


TARGETS = {"include/cfdb.h": gen_header, "fortran/cfdb_iface.f90": gen_fortran, "cfd_b200/_abi.py": gen_py}


def main():
    check = "--check" in sys.argv
    bad = []
    for rel, g in TARGETS.items():
        path, text = os.path.join(ROOT, rel), g()
        if check:
            if not os.path.exists(path) or open(path).read() != text:
                bad.append(rel)
        else:
            with open(path, "w") as fh:
                fh.write(text)
    if bad:
        print("out of date (run python tools/gen_abi.py):", ", ".join(bad))
        sys.exit(1)


if __name__ == "__main__":
    main()
