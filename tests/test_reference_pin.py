"""The oracle -- and the CUDA path -- against THE REFERENCE'S OWN SOURCES, executed statement by statement.

tests/golden/ref_*.npz were produced by running `PROGRAM NSComp2D` (ns2DComp.ALE.f90) and everything it calls from
/root/reference through oracle/f90ref (a Fortran-subset translator: no Fortran compiler exists in the image).  They hold
the program's module arrays after N passes of its time loop.  The bar is bit-exact for every array.

* test_oracle_matches_reference_program  (CPU, anywhere): the C++ oracle reproduces them;
* test_reference_sources_reproduce_golden (CPU, only where /root/reference exists): the vectors are what the sources give;
* test_gpu_matches_reference_program     (GPU): libcfdb200.so reproduces them through the C ABI, no oracle involved;
* call-site checks of routines the program never reaches (gcl_mod, CUARTO_ORDEN's projection), the deck reader, and the
  independent correctly-rounded pow.
"""
import importlib.util
import os

import numpy as np
import pytest

from conftest import assert_bit_equal

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden_ref", os.path.join(HERE, "golden", "make_golden_ref.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

from cfd_b200 import deck, meshgen  # noqa: E402

from oracle.f90ref import refrun  # noqa: E402

HAVE_REF = refrun.available()       # /root/reference, or its translation oracle/_ref/refprog.py made by build()
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="neither the reference sources nor oracle/_ref/refprog.py are on this machine")
BITEXACT_CASES = [n for n in mg.CASES if n != "ref_ale_seqdot"]


def gold(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def prepare(name, smoother):
    """deck -> LoadedCase, initial bump computed on the un-smoothed mesh (as injected after RESTART), then smoothing"""
    raw = mg.raw_case(name)
    lc = deck.load(raw)
    st = meshgen.density_bump(lc) if mg.CASES[name][2] else None
    smoother(lc)
    return lc, st, mg.CASES[name][1]


def check_state(solver, g, name, skip=()):
    for f in mg.FIELDS:
        if f not in skip and f in g.files:          # the large case keeps a subset of the arrays (make_golden_ref.SUBSET)
            assert_bit_equal(solver.get(f), g[f], f"{name}:{f}")
    for f in mg.INT_FIELDS:
        if f in g.files:
            assert np.array_equal(solver.get(f), g[f]), f"{name}:{f}"
    m = int(g["n_m"][0])
    assert int(solver.scalar("n_m")) == m
    if m:
        assert np.array_equal(solver.get("n_ipoin")[:m], g["n_ipoin"])
        assert_bit_equal(solver.get("n_x")[:m], g["n_x"], "n_x")
        assert_bit_equal(solver.get("n_y")[:m], g["n_y"], "n_y")
    assert solver.scalar("DTMIN") == g["dtmin"][-1] and solver.scalar("TIME") == g["time"][-1], "DTMIN/TIME"
    if "FX" in g.files:   # FORCES (meshMove.f90:153-194) and, on viscous print steps, FORCE_VISC (ns2DComp.ALE.f90:819-893)
        for f in ("FX", "FY", "RM"):
            assert_bit_equal(solver.get(f), g[f], f"{name}:{f}")
        if g["skin"].size:
            for f in ("F_VX", "F_VY"):
                assert_bit_equal(solver.get(f), g[f], f"{name}:{f}")
            for k, f in enumerate(("skin", "skin_x", "skin_p")):
                assert_bit_equal(solver.get(f), g["skin"][:, k], f"{name}:SKIN.DAT column {k + 1}")


def run_and_check(make_solver, name, smoother):
    lc, st, steps = prepare(name, smoother)
    s = make_solver(lc)
    if st is not None:
        for k, v in st.items():
            s.set(k, v)
    g = gold(name)
    every = mg.IPRINT.get(name, 1)
    for it in range(0, steps, every):
        s.step(every)
        # the program's .cnv line: sqrt(ER/ERR) summed sequentially by the reference, canonically here -> round-off
        er, err = s.step_norms() if hasattr(s, "step_norms") else s.norms_last()
        np.testing.assert_allclose(np.sqrt(er / err), g["cnv"][it // every, 2:], rtol=1e-12,
                                   err_msg=f"{name}: residuals of step {it + every}")
        assert s.scalar("DTMIN") == g["dtmin"][it + every - 1], f"{name}: DTMIN of step {it + every}"
    check_state(s, g, name)


def _orc_smooth(lc):
    from oracle import orclib
    orclib.lib().orc_smoothing(lc.X, lc.Y, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem)


class _OrcWithNorms:
    """oracle handle + the residual norms of the last step (evaluated before U = U1, ns2DComp.ALE.f90:191-197)"""

    def __init__(self, lc):
        from oracle.orclib import Oracle
        self.o = Oracle(lc)
        self.get, self.set, self.scalar = self.o.get, self.o.set, self.o.scalar

    def step(self, n):
        for _ in range(n):
            d = self.o.step_part1()
            self.o.step_part2(d)
            for irk in range(1, 5):
                self.o.rk_stage(irk)
            self._norms = self.o.norms()       # U still holds the step-start state here
            self.o.step_part3()

    def norms_last(self):
        return self._norms


@pytest.mark.parametrize("name", BITEXACT_CASES)
def test_oracle_matches_reference_program(name):
    run_and_check(_OrcWithNorms, name, _orc_smooth)


def test_oracle_step_equals_parts():
    """the piecewise stepping used above is the oracle's own orc_step"""
    from oracle.orclib import Oracle
    lc, st, _ = prepare("ref_channel_visc", _orc_smooth)
    a, b = _OrcWithNorms(lc), Oracle(lc)
    for k, v in st.items():
        a.set(k, v)
        b.set(k, v)
    a.step(3)
    b.step(3)
    for f in ("U", "T", "RHS"):
        assert_bit_equal(a.get(f), b.get(f), f)


@needs_ref
@pytest.mark.parametrize("name", ["ref_channel_noslip", "ref_ale", "ref_ale_seqdot"])
def test_reference_sources_reproduce_golden(name):
    """the committed vectors are what the sources under /root/reference compute (re-run live)"""
    out = mg.run_reference(name)
    g = gold(name)
    for k in g.files:
        a, b = np.asarray(out[k]), g[k]
        assert a.shape == b.shape, k
        if a.dtype.kind == "f":
            assert_bit_equal(a, b, f"{name}:{k}")
        else:
            assert np.array_equal(a, b), f"{name}:{k}"


def test_sequential_vecdot_differs_only_by_roundoff_at_first():
    """biCG with the reference's own sequential vecdot vs the canonical order: the mesh after the first moving step agrees
    to round-off (the later divergence is the F9 amplification described in DESIGN.md section 2)."""
    a, b = gold("ref_ale"), gold("ref_ale_seqdot")
    assert np.array_equal(a["dtmin"][:1], b["dtmin"][:1])
    np.testing.assert_allclose(a["cnv"][0, 2:], b["cnv"][0, 2:], rtol=1e-9)
    np.testing.assert_allclose(a["X"], b["X"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(a["Y"], b["Y"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", ["ref_channel_noslip", "ref_ale"])
def test_deck_reader_matches_reference_loader(name):
    """cfd_b200.deck.load == what readInputData/loadMeshData (dataLoader.f90) left in the modules, incl. F11 (real(4) TWALL)"""
    g = gold(name)
    lc = deck.load(mg.raw_case(name))
    assert np.array_equal(lc.ifixv_node, g["in_ifixv_node"])
    assert_bit_equal(lc.rfixv_valuex, g["in_rfixv_valuex"], "rfixv_valuex")
    assert_bit_equal(lc.rfixv_valuey, g["in_rfixv_valuey"], "rfixv_valuey")
    assert np.array_equal(lc.ifixrho_node, g["in_ifixrho_node"])
    assert_bit_equal(lc.rfixrho_value, g["in_rfixrho_value"], "rfixrho_value")
    assert np.array_equal(lc.ifixt_node, g["in_ifixt_node"])
    assert_bit_equal(lc.rfixt_value, g["in_rfixt_value"], "rfixt_value")
    # the reference allocates ILAUX(npoin) and uses the first nnmove = nmove + nfix_move entries (dataLoader.f90:258-268,319)
    assert np.array_equal(lc.ilaux, g["in_ilaux"][:lc.ilaux.size]) and not g["in_ilaux"][lc.ilaux.size:].any()


def test_correctly_rounded_pow_two_independent_ways():
    """oracle/orc_math.h (double-double) against oracle/f90ref/runtime.cr_pow (exact integer arithmetic)"""
    from oracle import orclib
    from oracle.f90ref.runtime import cr_pow
    L = orclib.lib()
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(0.2, 5.0, 4000), 10.0 ** rng.uniform(-12, 12, 4000), [1.0, 4.0, 0.25, 2.0, 1e-100, 1e100]])
    for x in xs:
        assert L.orc_pow15(x) == cr_pow(x, 1.5), x
        assert L.orc_pow05(x) == cr_pow(x, 0.5), x
        assert L.orc_powm05(x) == cr_pow(x, -0.5), x


@needs_ref
def test_gcl_and_cuarto_orden_call_sites():
    """routines PROGRAM NSComp2D never reaches (gcl_mod is an orphan, F5) or whose result it discards (CUARTO_ORDEN, F7):
    call the reference's subroutines directly on a post-run state and compare with the oracle's."""
    from oracle import orclib
    from oracle.f90ref.refrun import Reference
    from oracle.orclib import Oracle

    name = "ref_ale"
    raw = mg.raw_case(name)
    ref = Reference()
    L = orclib.lib()

    def hook(r):
        r.ns["p_biconjgrad__vecdot"] = lambda n, x, y, *a: np.float64(L.orc_vecdot(int(n), np.ascontiguousarray(x), np.ascontiguousarray(y)))

    ref.run_program(raw, maxiter=2, hook=hook)
    md, vel, g = ref.mod("meshdata"), ref.mod("mvelocidades"), ref.mod("mvariabgen")
    npoin, nelem = int(md.npoin), int(md.nelem)
    rng = np.random.default_rng(3)
    # gcl: putW / putArea with "old" values, then main
    W_x_old, W_y_old = rng.normal(size=npoin), rng.normal(size=npoin)
    area_old = md.area * (1 + 0.01 * rng.normal(size=nelem))
    with np.errstate(all="ignore"):
        ref.proc("putw", "gcl_mod")(W_x_old.copy(), W_y_old.copy())
        ref.proc("putarea", "gcl_mod")(area_old.copy())
        M_ref = md.m.copy()
        ref.proc("main", "gcl_mod")(M_ref, vel.w_x, vel.w_y, md.dnx, md.dny, md.area, md.inpoel, np.float64(1e-3))
    M_orc = md.m.copy()
    inp = np.ascontiguousarray(md.inpoel.T)            # repo layout [nelem][3]
    dnx, dny = md.dnx.ravel(order="F"), md.dny.ravel(order="F")              # Fortran memory order [nelem][3]
    L.orc_gcl_main(M_orc, np.ascontiguousarray(vel.w_x), np.ascontiguousarray(vel.w_y), W_x_old, W_y_old, area_old,
                   dnx, dny, np.ascontiguousarray(md.area), inp, nelem, npoin, 1e-3)
    assert_bit_equal(M_orc, M_ref, "gcl main")
    assert np.max(np.abs(M_ref - md.m)) > 0
    # CUARTO_ORDEN(U, U_n, FR, gamm) on the final state
    UN_ref = np.zeros((4, npoin), order="F")
    gamm = np.full(npoin, float(ref.mod("inputdata").gama))
    with np.errstate(all="ignore"):
        ref.proc("cuarto_orden")(g.u, UN_ref, ref.mod("inputdata").fr, gamm)
    lc, st, _ = prepare(name, _orc_smooth)
    o = Oracle(lc)
    o.step(2)
    o.set_scalar("use_cuarto", 1)
    d = o.step_part1()
    o.step_part2(d)
    o.rk_stage(1)
    assert_bit_equal(o.get("UN"), UN_ref.T.ravel(), "CUARTO_ORDEN")
    assert np.max(np.abs(UN_ref)) > 0


# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", BITEXACT_CASES)
def test_gpu_matches_reference_program(name):
    from cfd_b200 import capi
    from cfd_b200.solver import NSComp2D
    run_and_check(lambda lc: NSComp2D(lc), name, lambda lc: capi.smoothing(lc))


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("kind", ["channel_visc", "ale_visc"])
def test_gpu_against_the_reference_program_run_live(kind):
    """PROGRAM NSComp2D interpreted on this machine (from /root/reference, or from the translation __graft_entry__.build()
    left in oracle/_ref/, which travels to the GPU box) next to libcfdb200.so on the same deck: every array bit-identical."""
    from cfd_b200 import capi
    from cfd_b200.solver import NSComp2D

    if kind == "channel_visc":
        raw, bump, steps = meshgen.channel(nx=15, ny=7, FMU=1.8e-5, FK=0.0257, seed=77, jitter=0.3), True, 4
    else:
        raw, bump, steps = meshgen.ale_body(nt=24, nr=6, FMU=1.8e-5, FK=0.0257, seed=5), False, 3
    raw.IPRINT, raw.MAXITER = 1, steps
    st = meshgen.density_bump(deck.load(raw)) if bump else None
    ref = refrun.Reference()
    cnv = ref.run_program(raw, initial_state=st, canonical_vecdot=True)
    lc = deck.load(raw)
    capi.smoothing(lc)
    g = NSComp2D(lc)
    if st is not None:
        for k, v in st.items():
            g.set(k, v)
    g.step(steps)
    gen, var, md, vel, est, lap = (ref.mod(m) for m in ("mvariabgen", "mvariables", "meshdata", "mvelocidades", "mestabilizacion", "mlaplace"))
    for name, a in (("U", gen.u.T.ravel()), ("RHS", gen.rhs.T.ravel()), ("T", var.t), ("P", var.p), ("RMACH", var.rmach),
                    ("VEL_X", vel.vel_x), ("VEL_Y", vel.vel_y), ("W_X", vel.w_x), ("X", md.x), ("Y", md.y), ("M", md.m),
                    ("SHOC", est.shoc), ("T_SUGN2", est.t_sugn2), ("T_SUGN3", est.t_sugn3), ("lap_sparse", lap.lap_sparse),
                    ("F_VX", ref.mod("meshmove").f_vx), ("FX", ref.mod("meshmove").fx)):
        assert_bit_equal(g.get(name), a, f"{kind}:{name}")
    er, err = g.step_norms()
    np.testing.assert_allclose(np.sqrt(er / err), [float(x) for x in cnv[-1][2:]], rtol=1e-12)
