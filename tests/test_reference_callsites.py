"""Call-site pin: the reference's subroutines executed one by one from /root/reference (oracle/f90ref) on RANDOM inputs,
against the oracle's restatement of the same routine -- bit for bit.  Complements tests/test_reference_pin.py (whole-program
runs), whose states never reach some argument combinations: theta /= 0 in calcRHS (the program zeroes UN, F7), a moving
mesh velocity in deltat/ESTAB/FUENTE without a body, biCG with arbitrary right-hand sides and Dirichlet sets.

Runs only where the reference sources exist (this container); the GPU box relies on the oracle these tests pin.
"""
import os

import numpy as np
import pytest

from conftest import assert_bit_equal

from oracle.f90ref import refrun  # noqa: E402

pytestmark = pytest.mark.skipif(not refrun.available(), reason="neither the reference sources nor oracle/_ref/refprog.py are here")


@pytest.fixture(scope="module")
def ctx():
    """a small jittered mesh loaded INTO THE REFERENCE'S MODULES by its own readInputData/loadMeshData + geometry routines"""
    from cfd_b200 import deck, meshgen
    from oracle.f90ref.refrun import Reference

    raw = meshgen.channel(nx=13, ny=7, FMU=1.8e-5, FK=0.0257, jitter=0.3, seed=4)
    raw.IPRINT = 1
    ref = Reference()
    ref.run_program(raw, maxiter=1)        # one pass: modules allocated, geometry and topology built, state non-trivial
    lc = deck.load(raw)
    from oracle import orclib
    orclib.lib().orc_smoothing(lc.X, lc.Y, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem)
    return ref, lc, orclib.lib()


def _state(ref, rng):
    md = ref.mod("meshdata")
    P, E = int(md.npoin), int(md.nelem)
    rho = 1.2 * (1 + 0.2 * rng.random(P))
    u, v = 150 + 40 * rng.normal(size=P), 30 * rng.normal(size=P)
    e = 2.0e5 * (1 + 0.1 * rng.random(P)) + 0.5 * (u * u + v * v)
    U = np.asfortranarray(np.stack([rho, rho * u, rho * v, rho * e]))
    T = 288.0 * (1 + 0.1 * rng.random(P))
    return P, E, U, T, u, v


@pytest.mark.parametrize("visc", [False, True])
def test_calcrhs_with_nonzero_theta(ctx, visc):
    ref, lc, L = ctx
    rng = np.random.default_rng(11 + visc)
    md, inp = ref.mod("meshdata"), ref.mod("inputdata")
    P, E, U, T, _, _ = _state(ref, rng)
    theta = np.asfortranarray(rng.normal(size=(4, P)) * np.array([[1e-2], [1.0], [1.0], [1e3]]))
    shoc, dtl = 1e-3 * rng.random(E), 1e-5 * (1 + rng.random(E))
    ts = [1e-5 * rng.random(E) for _ in range(3)]
    ts[1][::3] = 0.0                                   # the F9 switch: T_SUGN2 exactly zero on some elements
    mu_ref = float(inp.fmu) if visc else 0.0
    old_fmu, old_T = inp.fmu, ref.mod("mvariables").t.copy()
    inp.fmu = np.float64(mu_ref)
    ref.mod("mvariables").t[...] = T
    rhs_ref = np.zeros((4, P), order="F")
    with np.errstate(all="ignore"):
        ref.proc("calcrhs", "calcrhs_mod")(rhs_ref, U, theta, md.dnx, md.dny, md.area, shoc, dtl, ts[0], ts[1], ts[2], md.inpoel, E, P)
    inp.fmu = old_fmu
    ref.mod("mvariables").t[...] = old_T
    rhs = np.zeros(4 * P)
    L.orc_calcrhs(rhs, np.ascontiguousarray(U.T).ravel(), np.ascontiguousarray(theta.T).ravel(), T, md.dnx.ravel(order="F"),
                  md.dny.ravel(order="F"), np.ascontiguousarray(md.area), shoc, dtl, ts[0], ts[1], ts[2],
                  np.ascontiguousarray(md.inpoel.T), E, P, float(inp.fcv), float(inp.fk), mu_ref, float(inp.gama), float(inp.t_inf), float(inp.cte))
    assert_bit_equal(rhs, rhs_ref.T.ravel(), "calcRHS")
    assert np.abs(rhs).max() > 0


def test_deltat_estab_fuente_with_mesh_velocity(ctx):
    ref, lc, L = ctx
    rng = np.random.default_rng(21)
    md, inp, vel, var, gen, est = (ref.mod(m) for m in ("meshdata", "inputdata", "mvelocidades", "mvariables", "mvariabgen", "mestabilizacion"))
    P, E, U, T, u, v = _state(ref, rng)
    wx, wy = 20 * rng.normal(size=P), 20 * rng.normal(size=P)
    gamm = np.full(P, float(inp.gama))
    vel.vel_x[...], vel.vel_y[...], vel.w_x[...], vel.w_y[...] = u, v, wx, wy
    var.t[...] = T
    gen.u[...] = U
    # deltat(dtmin, dt)
    dt_ref = np.zeros(E)
    with np.errstate(all="ignore"):
        dtmin_ref, _ = ref.proc("deltat")(np.float64(0.0), dt_ref)
    dt_o, dtmin_o = np.zeros(E), np.zeros(1)
    inpo, area = np.ascontiguousarray(md.inpoel.T), np.ascontiguousarray(md.area)
    L.orc_deltat(E, inpo, area, T, u, v, wx, wy, float(inp.fsafe), float(inp.fr), float(inp.gama), float(inp.t_inf), dt_o, dtmin_o)
    assert float(dtmin_ref) == dtmin_o[0]
    assert_bit_equal(dt_o, dt_ref, "deltat DT")
    # ESTAB(U,T,GAMA,FR,RMU,DTMIN,RHOINF,TINF,UINF,VINF,GAMM)
    with np.errstate(all="ignore"):
        ref.proc("estab")(gen.u, var.t, inp.gama, inp.fr, np.float64(0.0), dtmin_ref, inp.rho_inf, inp.t_inf, inp.u_inf, inp.v_inf, gamm)
    outs = [np.zeros(E) for _ in range(4)]
    dnx, dny = md.dnx.ravel(order="F"), md.dny.ravel(order="F")
    L.orc_estab(E, inpo, np.ascontiguousarray(U.T).ravel(), T, u, v, wx, wy, gamm, dnx, dny, float(inp.fr), float(dtmin_ref),
                float(inp.rho_inf), float(inp.t_inf), *outs)
    for a, b, n in zip(outs, (est.shoc, est.t_sugn1, est.t_sugn2, est.t_sugn3), ("SHOC", "T_SUGN1", "T_SUGN2", "T_SUGN3")):
        assert_bit_equal(a, b, n)
    # FUENTE(dtl) adds to the module RHS
    dtl = 1e-5 * (1 + rng.random(E))
    gen.rhs[...] = 0.0
    with np.errstate(all="ignore"):
        ref.proc("fuente")(dtl)
    rhs = np.zeros(4 * P)
    L.orc_fuente(rhs, np.ascontiguousarray(U.T).ravel(), wx, wy, dnx, dny, area, dtl, inpo, E)
    assert_bit_equal(rhs, gen.rhs.T.ravel(), "FUENTE")
    assert np.abs(rhs).max() > 0


def test_bicg_spmv_laplace(ctx):
    ref, lc, L = ctx
    rng = np.random.default_rng(31)
    md, lap = ref.mod("meshdata"), ref.mod("mlaplace")
    P, E = int(md.npoin), int(md.nelem)
    A, idx, rowptr, diag = (np.ascontiguousarray(a) for a in (lap.lap_sparse, lap.lap_idx, lap.lap_rowptr, lap.lap_diag))
    # laplace on perturbed coordinates (the q = 1/mu**2 weights use the module X, Y)
    X0, Y0 = md.x.copy(), md.y.copy()
    md.x[...] = X0 + 1e-3 * rng.normal(size=P)
    md.y[...] = Y0 + 1e-3 * rng.normal(size=P)
    with np.errstate(all="ignore"):
        ref.proc("laplace", "mlaplace")(md.inpoel, md.area, md.dnx, md.dny, E, P)
    li, lr, sp_o, dg_o = np.zeros(idx.size, np.int32), np.zeros(P + 1, np.int32), np.zeros(idx.size), np.zeros(P)
    n = L.orc_laplace(np.ascontiguousarray(md.inpoel.T), md.dnx.ravel(order="F"), md.dny.ravel(order="F"), np.ascontiguousarray(md.x),
                      np.ascontiguousarray(md.y), E, P, li, lr, sp_o, dg_o, idx.size)
    assert n == idx.size and np.array_equal(li, lap.lap_idx) and np.array_equal(lr, lap.lap_rowptr)
    assert_bit_equal(sp_o, lap.lap_sparse, "laplace values")
    assert_bit_equal(dg_o, lap.lap_diag, "laplace diagonal")
    md.x[...], md.y[...] = X0, Y0
    A, diag = sp_o.copy(), dg_o.copy()
    # SpMV
    v = rng.normal(size=P)
    y_ref = np.zeros(P)
    ref.proc("spmv", "biconjgrad")(A, idx, rowptr, v, y_ref, P, idx.size)
    y_o = np.zeros(P)
    L.orc_spmv(A, idx, rowptr, v, y_o, P)
    assert_bit_equal(y_o, y_ref, "SpMV")
    # biCG with a Dirichlet set, a non-zero right-hand side and a warm start; vecdot in the build's canonical order
    # (the one order the reference leaves to OpenMP, biconjGrad.f90:162)
    fix = np.ascontiguousarray(rng.choice(P, 9, replace=False).astype(np.int32) + 1)
    xf = rng.normal(size=fix.size)
    b = 1e-3 * rng.normal(size=P)
    x0 = 1e-2 * rng.normal(size=P)
    ref.ns["p_biconjgrad__vecdot"] = lambda n, x, y, *a: np.float64(L.orc_vecdot(int(n), np.ascontiguousarray(x), np.ascontiguousarray(y)))
    x_ref = x0.copy()
    with np.errstate(all="ignore"):
        ref.proc("bicg", "biconjgrad")(A, idx, rowptr, diag, x_ref, b.copy(), xf.copy(), fix.copy(), P, fix.size)
    x_o = x0.copy()
    its = L.orc_bicg(A, idx, rowptr, diag, x_o, b, xf, fix, P, fix.size)
    assert its > 3
    assert_bit_equal(x_o, x_ref, "biCG solution")
    np.testing.assert_allclose(x_o[fix - 1], xf, rtol=1e-12)     # penalty rows hold the prescribed values


def test_restart_files_both_directions(tmp_path):
    """<name>.RST: the reference's PRINTREST output (unformatted sequential records, ns2DComp.ALE.f90:898-917) is byte-identical
    to cfd_b200.deck.write_rst, and its RESTART (IRESTART = 1, :421-432) reads a file deck.write_rst wrote -- U, T and GAMM
    restored, ITER/TIME read and dropped, velocities left as allocated (zero)."""
    from cfd_b200 import deck, meshgen
    from oracle import orclib
    from oracle.f90ref.refrun import Reference
    from oracle.orclib import Oracle

    raw = meshgen.channel(nx=11, ny=5, FMU=1.8e-5, FK=0.0257)
    raw.IPRINT = 2
    ref = Reference()
    ref.run_program(raw, maxiter=4)
    g, v = ref.mod("mvariabgen"), ref.mod("mvariables")
    P = int(ref.mod("meshdata").npoin)
    rst = ref.io.binary[raw.name + ".RST"]
    # the program's locals are gone; TIME is the second item of the first record
    it, time = np.frombuffer(rst[4:8], "<i4")[0], np.frombuffer(rst[8:16], "<f8")[0]
    assert it == 4 and time > 0
    # PRINTREST is called before the loop's U = U1 (ns2DComp.ALE.f90:254 vs :277): the file holds the state the step STARTED
    # from next to the temperature it ENDED with -- a quirk of the reference, reproduced here by feeding write_rst the same
    T4 = v.t.copy()
    ref.run_program(raw, maxiter=3)
    U3 = ref.mod("mvariabgen").u.T.ravel().copy()
    mine = tmp_path / "mine.RST"
    deck.write_rst(str(mine), int(it), float(time), U3, T4, np.full(P, 1.4))
    assert open(mine, "rb").read() == rst
    it2, time2, U2, T2, G2 = deck.read_rst(str(mine), P)
    assert (it2, time2) == (it, time) and np.array_equal(U2.ravel(), U3) and np.array_equal(T2, T4)
    # restart from a perturbed file: one step of the reference program == one step of the oracle from the same state
    rng = np.random.default_rng(2)
    Ur = U2 * (1 + 1e-3 * rng.normal(size=U2.shape))
    Tr = T2 * (1 + 1e-3 * rng.normal(size=P))
    deck.write_rst(str(mine), 77, 0.25, Ur, Tr, G2)
    raw2 = meshgen.channel(nx=11, ny=5, FMU=1.8e-5, FK=0.0257)
    raw2.IPRINT, raw2.IRESTART = 1, 1
    ref.run_program(raw2, maxiter=1, files={raw2.name + ".RST": open(mine, "rb").read()})
    lc = deck.load(raw2)
    orclib.lib().orc_smoothing(lc.X, lc.Y, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem)
    o = Oracle(lc)
    o.set("U", Ur)
    o.set("T", Tr)
    o.set("VEL_X", np.zeros(P))           # RESTART does not restore the velocities (fresh ALLOCATE memory)
    o.set("VEL_Y", np.zeros(P))
    o.step(1)
    # with the velocities unset the reference's first step after a restart produces NaNs around the bump (both sides, same
    # entries); the sign bit of a NaN depends on the instruction that made it, so NaNs are compared by position
    for name, a, b in (("U", o.get("U"), ref.mod("mvariabgen").u.T.ravel()), ("T", o.get("T"), ref.mod("mvariables").t)):
        a, b = np.asarray(a), np.asarray(b)
        assert np.array_equal(np.isnan(a), np.isnan(b)), name
        ok = ~np.isnan(a)
        assert ok.sum() > a.size // 2
        assert_bit_equal(a[ok], b[ok], f"{name} one step after RESTART")
