"""Round-2 additions, GPU side: the call-site entries for fixvel / normalvel / FIX / RK / fluidStructure, streamed stepping,
ADAMSB, the coloured deterministic mode, the print-step text files.  Bar as everywhere: bit-exact against the oracle, except
where a mode is relaxed by definition (coloured scatter: reproducible and within the per-step tolerance)."""
import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

STATE = ["U", "U1", "RHS", "T", "VEL_X", "VEL_Y", "RHO", "E", "P", "RMACH", "SHOC", "T_SUGN1", "T_SUGN2", "T_SUGN3", "W_X", "W_Y", "X", "Y"]


def _pair(lc, use_gcl=0):
    from cfd_b200.solver import NSComp2D
    from oracle.orclib import Oracle

    return NSComp2D(lc, use_gcl=use_gcl), Oracle(lc, use_gcl=use_gcl)


def _perturb(lc, g, o, amp=0.1):
    from cfd_b200.meshgen import density_bump

    for k, v in density_bump(lc, amp=amp).items():
        g.set(k, v)
        if o is not None:
            o.set(k, v)


def test_fixvel_normalvel_fix_call_sites(cases):
    """subrutinas.f90:601-616, :66-85, :618-643 on host arrays, duplicates included (last list entry wins)"""
    lc = cases["channel_noslip"]
    g, _ = _pair(lc)
    P = lc.npoin
    rng = np.random.default_rng(5)
    vx, vy = rng.normal(size=P), rng.normal(size=P)
    # fixvel
    a, b = vx.copy(), vy.copy()
    g.fixvel(lc.ifixv_node, lc.rfixv_valuex, lc.rfixv_valuey, a, b)
    ra, rb = vx.copy(), vy.copy()
    for i, n in enumerate(lc.ifixv_node):
        ra[n - 1], rb[n - 1] = lc.rfixv_valuex[i], lc.rfixv_valuey[i]
    assert_bit_equal(a, ra, "fixvel x")
    assert_bit_equal(b, rb, "fixvel y")
    # normalvel with the context's own wall normals
    m, ip, nx, ny = g.normales(lc.X, lc.Y)
    wx, wy = rng.normal(size=P) * 1e-2, rng.normal(size=P) * 1e-2
    a, b = vx.copy(), vy.copy()
    g.normalvel(ip, nx, ny, a, b, wx, wy)
    ra, rb = vx.copy(), vy.copy()
    for i in range(m):
        n = ip[i] - 1
        p = -ny[i] * (ra[n] - wx[n]) + nx[i] * (rb[n] - wy[n])
        ra[n], rb[n] = -ny[i] * p + wx[n], nx[i] * p + wy[n]
    assert m > 0
    assert_bit_equal(a, ra, "normalvel x")
    assert_bit_equal(b, rb, "normalvel y")
    # FIX
    gam = np.full(P, 1.4)
    rho, T, E = rng.uniform(1, 2, P), rng.uniform(250, 350, P), rng.uniform(1e5, 2e5, P)
    r1, t1, e1 = rho.copy(), T.copy(), E.copy()
    g.fix(287.0, gam, lc.ifixrho_node, lc.rfixrho_value, lc.ifixt_node, lc.rfixt_value, vx, vy, r1, t1, e1)
    r2, t2, e2 = rho.copy(), T.copy(), E.copy()
    for i, n in enumerate(lc.ifixrho_node):
        r2[n - 1] = lc.rfixrho_value[i]
    for i, n in enumerate(lc.ifixt_node):
        j = n - 1
        t2[j] = lc.rfixt_value[i]
        e2[j] = t2[j] * 287.0 / (gam[j] - 1.0) + .5 * (vx[j] * vx[j] + vy[j] * vy[j])
    assert lc.ifixt_node.size > 0
    assert_bit_equal(r1, r2, "FIX rho")
    assert_bit_equal(t1, t2, "FIX T")
    assert_bit_equal(e1, e2, "FIX E")


@pytest.mark.parametrize("name", ["channel_visc", "ale"])
def test_rk_call_site_equals_the_oracles_rk(cases, name):
    """RK(DTMIN, NRK, BANDERA, GAMM, dtl) as one call with host arrays (the shim's entry) against four oracle stages"""
    lc = cases[name]
    g, o = _pair(lc)
    if name != "ale":
        _perturb(lc, g, o)
    g.step(2)
    o.step(2)
    d = o.step_part1()
    o.step_part2(d)
    dtmin, band = o.scalar("DTMIN"), int(o.scalar("BANDERA"))
    P, E = lc.npoin, lc.nelem
    arr = {k: o.get(k) for k in ("U", "T", "VEL_X", "VEL_Y", "W_X", "W_Y", "GAMM", "RHS1", "RHS2", "RHS3", "P")}
    dtl = o.get("DTL")
    out = {k: np.zeros(4 * P) for k in ("U1", "RHS")}
    out.update({k: np.zeros(P) for k in ("RHO", "E", "RMACH")})
    out.update({k: np.zeros(E) for k in ("SHOC", "T_SUGN1", "T_SUGN2", "T_SUGN3")})
    g2, _ = _pair(lc)       # a fresh context: geometry only
    if name == "ale":       # the moved mesh of the oracle
        for k in ("X", "Y"):
            g2.set(k, o.get(k))
        g2.geometry(0)
    g2.rk_callsite(dtmin, 4, band, arr["GAMM"], dtl, arr["U"], out["U1"], out["RHS"], arr["RHS1"], arr["RHS2"], arr["RHS3"], arr["T"],
                   arr["P"], out["RHO"], out["E"], out["RMACH"], arr["VEL_X"], arr["VEL_Y"], arr["W_X"], arr["W_Y"], out["SHOC"],
                   out["T_SUGN1"], out["T_SUGN2"], out["T_SUGN3"])
    for irk in (1, 2, 3, 4):
        o.rk_stage(irk)
    for k in ("U1", "RHS", "RHO", "E", "RMACH", "SHOC", "T_SUGN1", "T_SUGN2", "T_SUGN3"):
        assert_bit_equal(out[k], o.get(k), f"rk call site {k}")
    for k in ("T", "P", "VEL_X", "VEL_Y"):
        assert_bit_equal(arr[k], o.get(k), f"rk call site {k}")


def test_mesh_move_call_site(cases):
    """fluidStructure(dtmin, time, ...) as one call with host arrays against the oracle's"""
    lc = cases["ale"]
    g, o = _pair(lc)
    g.step(3)
    o.step(3)
    dtmin, time = o.scalar("DTMIN"), o.scalar("TIME") + o.scalar("DTMIN")
    a = {k: o.get(k) for k in ("X", "Y", "X1", "Y1", "W_X", "W_Y", "P", "xpos", "ypos")}
    # the stepped context carries the Laplacian of the current mesh and fluidStructure's saved pitch angle
    fx, fy, rm = g.mesh_move(dtmin, time, a["X"], a["Y"], a["X1"], a["Y1"], a["W_X"], a["W_Y"], a["P"], a["xpos"], a["ypos"])
    o.fluid_structure(dtmin, time)
    for k in ("X", "Y", "X1", "Y1", "W_X", "W_Y", "xpos", "ypos"):
        assert_bit_equal(a[k], o.get(k), f"fluidStructure call site {k}")
    assert_bit_equal(fx, o.get("FX"), "FX")
    assert_bit_equal(rm, o.get("RM"), "RM")


def test_streamed_steps_equal_resident_steps(cases):
    """cfdb_step_streamed (state from / to host arrays, pipelined on copy streams) = cfdb_step, norms included"""
    lc = cases["channel_visc"]
    g, o = _pair(lc)
    _perturb(lc, g, o)
    names = ["U", "T", "VEL_X", "VEL_Y"]
    ins = [{n: o.get(n) for n in names} for _ in range(2)]
    outs = [{n: np.zeros_like(ins[0][n]) for n in names} for _ in range(2)]
    norms = [np.zeros(8), np.zeros(8)]
    for k in range(5):
        s = k & 1
        if k:                       # the host program feeds back what it received (after waiting for it)
            g.streamed_wait()
            for n in names:
                ins[s][n][:] = outs[s ^ 1][n]
        g.step_streamed(ins[s], outs[s], norms[s])
        o.set_scalar("norms_every_step", 1)
        o.step(1)
    g.streamed_wait()
    for n in names:
        assert_bit_equal(outs[0][n], o.get(n), f"streamed {n}")
    er, err = o.step_norms()
    assert_bit_equal(norms[0][:4], er, "ER")
    assert_bit_equal(norms[0][4:], err, "ERR")
    assert g.scalar("TIME") == o.scalar("TIME")


@pytest.mark.parametrize("name", ["channel", "channel_visc"])
def test_adamsb_option(cases, name):
    """option "adamsb": RK while BANDERA <= 4, then ADAMSB (subrutinas.f90:851-1034) -- bit-exact against the oracle"""
    lc = cases[name]
    g, o = _pair(lc)
    _perturb(lc, g, o, amp=0.01)
    g.set_option("adamsb", 1)
    o.set_scalar("adamsb", 1)
    used = 0
    for _ in range(12):
        used += int(o.scalar("BANDERA") > 4)
        g.step(1)
        o.step(1)
    assert used >= 3, "the run never reached ADAMSB"
    for n in STATE + ["RHS1", "RHS2", "RHS3", "UN"]:
        assert_bit_equal(g.get(n), o.get(n), f"adamsb {n}")
    assert g.scalar("BANDERA") == o.scalar("BANDERA")


@pytest.mark.parametrize("name", ["channel", "channel_visc", "ale"])
def test_coloured_mode_is_reproducible_and_within_the_per_step_tolerance(cases, name):
    lc = cases[name]
    from cfd_b200.solver import NSComp2D

    runs = []
    for _ in range(2):
        g = NSComp2D(lc)
        if name != "ale":
            _perturb(lc, g, None)
        g.set_option("colored", 1)
        g.step(3)
        runs.append({n: g.get(n) for n in ("U", "T", "RHS")})
    for n in runs[0]:
        assert_bit_equal(runs[0][n], runs[1][n], f"coloured run-to-run {n}")
    g, o = _pair(lc)
    if name != "ale":
        _perturb(lc, g, o)
    g.set_option("colored", 1)
    g.step(1)
    o.step(1)
    U, Uo = g.get("U").reshape(-1, 4), o.get("U").reshape(-1, 4)
    assert np.max(np.abs(U - Uo) / np.abs(Uo).max(0)) <= 1e-11


def test_print_step_text_files(cases, tmp_path):
    """FORCES ('(A, I2)', '(A, E14.5)' records, ns2DComp.ALE.f90:238-250), DESPLAZAMIENTO ('(7E13.5)', :237), SKIN.DAT
    (list-directed, :888) from the resident forces of a viscous moving-body run"""
    from cfd_b200 import deck, meshgen
    from cfd_b200.solver import NSComp2D

    lc = deck.load(meshgen.ale_body(nt=48, nr=14, FMU=1.8e-5, FK=0.0257))
    g = NSComp2D(lc)
    g.step(3)
    fvx, fvy, skin, sx, sp = g.force_visc()
    fx, fy, rm = g.get("FX"), g.get("FY"), g.get("RM")
    g.write_forces(tmp_path / "FORCES")
    g.write_desplazamiento(tmp_path / "DESPLAZAMIENTO", 0.125)
    g.write_desplazamiento(tmp_path / "DESPLAZAMIENTO", 0.25, append=True)
    g.write_skin(tmp_path / "SKIN.DAT")

    def E(v, w, d):
        return NSComp2D.format_real("E", v, w, d)

    lines = open(tmp_path / "FORCES").read().split("\n")
    assert lines[0] == "SET NUMERO 1"
    assert lines[1] == "FUERZA EN X:" + E(fx[0], 14, 5) and lines[2] == "FUERZA EN Y:" + E(fy[0], 14, 5) and lines[3] == ""
    assert lines[4] == "FUERZA VISCOSA EN X:" + E(fvx[0], 14, 5) and lines[7] == "FUERZA TOTAL EN X:" + E(fx[0] + fvx[0], 14, 5)
    d = open(tmp_path / "DESPLAZAMIENTO").read().split("\n")
    assert d[0] == "".join(E(v, 13, 5) for v in (0.125, fvx[0], fvy[0], rm[0], fvx[1], fvy[1], rm[1])) and len(d) == 3
    sk = open(tmp_path / "SKIN.DAT").read().split("\n")
    assert len(sk) == skin.size + 1
    for k in (0, skin.size - 1):
        assert sk[k] == " " + "".join(NSComp2D.format_real("L", v, 0, 0) for v in (skin[k], sx[k], sp[k]))
        assert [float(x) for x in sk[k].split()] == [skin[k], sx[k], sp[k]]     # 17 significant digits round-trip


def test_spmv_row_runs_of_any_length(cases):
    """k::spmv stages the entries of 32 consecutive rows in shared memory when they fit its window and walks the global arrays
    otherwise: rows of 1..60 entries (mixed inside a warp), an empty row, a row count that is not a multiple of 32 — every
    row's sum in the reference's order (biconjGrad.f90:171-190), checked against the oracle bit for bit."""
    from cfd_b200.solver import NSComp2D
    from oracle import orclib

    lc = cases["channel"]
    g = NSComp2D(lc)
    P = lc.npoin - 7
    rng = np.random.default_rng(12)
    width = rng.integers(1, 13, P)
    width[100:164] = rng.integers(30, 61, 64)        # two warps whose runs exceed the window
    width[300:310] = 0
    width[40] = 200
    rowptr = np.zeros(P + 1, np.int32)
    rowptr[1:] = np.cumsum(width)
    nnz = int(rowptr[-1])
    idx = rng.integers(1, P + 1, nnz).astype(np.int32)
    A = rng.standard_normal(nnz)
    v = rng.standard_normal(P)
    y_o = np.zeros(P)
    orclib.lib().orc_spmv(A, idx, rowptr, v, y_o, P)
    assert_bit_equal(g.spmv(A, idx, rowptr, v), y_o, "spmv with mixed row lengths")


def test_forces_over_more_body_edges_than_one_chunk():
    """k::forces evaluates the edge terms of a body set 512 at a time and adds them in list order (meshMove.f90:171-192):
    a body with 1 300 edges (three chunks, the last one partial), two moving-mesh steps, forces and moment bit for bit."""
    from cfd_b200 import deck, meshgen

    lc = deck.load(meshgen.ale_body(nt=1300, nr=4))
    g, o = _pair(lc)
    _perturb(lc, g, o, amp=0.05)
    g.step(2)
    o.step(2)
    for k in ("FX", "FY", "RM"):
        assert_bit_equal(g.get(k), o.get(k), k)
    assert abs(o.get("FX")[0]) > 0
    for k in ("U", "X", "W_X"):
        assert_bit_equal(g.get(k), o.get(k), k)
