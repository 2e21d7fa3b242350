"""Analytic invariants that pin the oracle (the reference ships no tests or golden vectors)."""
import numpy as np
import pytest

from cfd_b200 import deck, meshgen
from oracle import orclib
from oracle.orclib import Oracle


@pytest.mark.parametrize("name", ["channel", "wedge", "ale", "square"])
def test_geometry_invariants(cases, name):
    lc = cases[name]
    o = Oracle(lc)
    area, dNx, dNy, M = o.get("area"), o.get("dNx").reshape(-1, 3), o.get("dNy").reshape(-1, 3), o.get("M")
    assert (area > 0).all()                                        # counter-clockwise triangles
    scale = np.abs(dNx).max(1)
    assert (np.abs(dNx.sum(1)) <= 1e-12 * scale).all() and (np.abs(dNy.sum(1)) <= 1e-12 * np.abs(dNy).max(1)).all()
    assert abs(M.sum() - area.sum()) <= 1e-12 * area.sum()         # sum M = sum area
    # shape-function gradients reproduce a linear field exactly: grad(x) = (1,0), grad(y) = (0,1)
    X, Y, inp = o.get("X"), o.get("Y"), lc.inpoel - 1
    assert np.allclose((dNx * X[inp]).sum(1), 1.0, atol=1e-10) and np.allclose((dNy * X[inp]).sum(1), 0.0, atol=1e-9)
    assert np.allclose((dNy * Y[inp]).sum(1), 1.0, atol=1e-10)
    # Laplacian: rows sum to zero, diagonal first and positive, symmetric pattern
    sp, idx, rp, dg = o.get("lap_sparse"), o.get("lap_idx"), o.get("lap_rowptr"), o.get("lap_diag")
    rows = np.repeat(np.arange(lc.npoin), np.diff(rp))
    rs = np.bincount(rows, sp, lc.npoin)
    assert (np.abs(rs) <= 1e-9 * dg).all() and (dg > 0).all()
    assert np.array_equal(idx[rp[:-1]], np.arange(1, lc.npoin + 1)) and np.array_equal(sp[rp[:-1]], dg)
    pairs = set(zip(rows.tolist(), (idx - 1).tolist()))
    assert all((j, i) in pairs for i, j in pairs)


def test_esup_psup_structure(cases):
    lc = cases["ale"]
    L = orclib.lib()
    e1, e2 = np.zeros(3 * lc.nelem, np.int32), np.zeros(lc.npoin + 1, np.int32)
    L.orc_get_esup(lc.inpoel, lc.nelem, lc.npoin, e1, e2)
    assert e2[0] == 0 and e2[-1] == 3 * lc.nelem
    for n in range(lc.npoin):
        el = e1[e2[n]:e2[n + 1]]
        assert (np.diff(el) > 0).all()                                    # ascending element ids
        assert all((lc.inpoel[e - 1] == n + 1).any() for e in el)
    p1, p2 = np.zeros(8 * lc.nelem, np.int32), np.zeros(lc.npoin + 1, np.int32)
    cnt = L.orc_get_psup(lc.inpoel, lc.nelem, lc.npoin, p1, p1.size, p2)
    nb = [set(p1[p2[n]:p2[n + 1]].tolist()) for n in range(lc.npoin)]
    assert cnt == p2[-1] and all(len(nb[n]) == p2[n + 1] - p2[n] for n in range(lc.npoin))  # no duplicates
    assert all((n + 1) in nb[m - 1] for n in range(lc.npoin) for m in nb[n])                # symmetric
    # first-encounter order: walking esup, local nodes 1..3
    n = 5
    order = []
    for e in e1[e2[n]:e2[n + 1]]:
        for j in lc.inpoel[e - 1]:
            if j != n + 1 and j not in order:
                order.append(j)
    assert order == p1[p2[n]:p2[n + 1]].tolist()


def test_free_stream_is_preserved(cases):
    o = Oracle(cases["square"])
    U0 = o.get("U").reshape(-1, 4)
    o.step(10)
    U = o.get("U").reshape(-1, 4)
    scale = np.array([U0[:, 0].max(), U0[:, 1].max(), U0[:, 1].max(), U0[:, 3].max()])
    assert np.max(np.abs(U - U0) / scale) < 1e-10
    assert o.scalar("bicg_x") == -1 and o.scalar("bicg_y") == -1       # trivial mesh solve returns at once


def test_rk_stages_all_start_from_U(cases):
    """every stage evaluates calcRHS at U; in Euler mode the four RHS are identical."""
    lc = cases["channel"]
    o = Oracle(lc)
    for k, v in meshgen.density_bump(lc).items():
        o.set(k, v)
    o.step(2)
    d = o.step_part1()
    o.step_part2(d)
    rhs = []
    for irk in (1, 2, 3, 4):
        o.rk_stage(irk)
        rhs.append(o.get("RHS"))
    assert all(np.array_equal(rhs[0], r) for r in rhs[1:])


def test_estab_amplifies_one_ulp(cases):
    """T_SUGN2 switches between 0 and ~dt/2 on the exact-zero-ness of a round-off sum."""
    lc = cases["channel"]
    o = Oracle(lc)
    for k, v in meshgen.density_bump(lc).items():
        o.set(k, v)
    o.step(3)
    L = orclib.lib()
    E = lc.nelem
    args = [o.get(k) for k in ("U", "T", "VEL_X", "VEL_Y", "W_X", "W_Y", "GAMM", "dNx", "dNy")]
    p = lc.par

    def run(vx):
        out = [np.zeros(E) for _ in range(4)]
        a = list(args)
        a[2] = vx
        L.orc_estab(E, lc.inpoel, *a, p["FR"], o.scalar("DTMIN"), p["RHO_inf"], p["T_inf"], *out)
        return out

    base = run(args[2])
    pert = run(np.nextafter(args[2], np.inf))
    t2a, t2b = base[2], pert[2]
    assert ((t2a == 0) != (t2b == 0)).sum() > 0.05 * E          # O(1) switches from a one-ulp change
    assert np.allclose(base[1], pert[1], rtol=1e-12)           # while T_SUGN1 moves by round-off only
    frac0 = (t2a == 0).mean()
    assert 0.05 < frac0 < 0.95


def test_bicg_solves_dirichlet_problem(cases):
    lc = cases["ale"]
    o = Oracle(lc)
    L = orclib.lib()
    P = lc.npoin
    sp, idx, rp, dg = o.get("lap_sparse"), o.get("lap_idx"), o.get("lap_rowptr"), o.get("lap_diag")
    fix = lc.ilaux.copy()
    xf = np.concatenate([np.full(lc.i_m.size, 1e-3), np.zeros(lc.ifm.size)])
    x, b = np.zeros(P), np.zeros(P)
    it = L.orc_bicg(sp, idx, rp, dg, x, b, xf, fix, P, fix.size)
    assert 3 < it < 1000
    assert np.allclose(x[lc.i_m - 1], 1e-3, rtol=1e-9) and np.allclose(x[lc.ifm - 1], 0.0, atol=1e-12)
    y = np.zeros(P)
    L.orc_spmv(sp, idx, rp, x, y, P)
    free = np.ones(P, bool)
    free[fix - 1] = False
    assert np.abs(y[free]).max() < 1e-3 * np.abs(sp).max() * 1e-3   # interior rows of A x = 0 to solver tolerance
    assert x.min() >= -1e-9 and x.max() <= 1e-3 + 1e-9              # discrete maximum principle
    x0 = np.zeros(P)
    assert L.orc_bicg(sp, idx, rp, dg, x0, b, np.zeros(fix.size), fix, P, fix.size) == -1 and not x0.any()


def test_gcl_and_ale_smoke(cases):
    lc = cases["ale"]
    o = Oracle(lc, use_gcl=1)
    M0 = o.get("M").copy()
    o.step(3)
    assert o.scalar("bicg_x") > 0 and np.abs(o.get("W_X")).max() > 0      # the body really moves
    assert not np.array_equal(o.get("M"), M0)
    L = orclib.lib()
    P, E = lc.npoin, lc.nelem
    M = M0.copy()
    z = np.zeros(P)
    L.orc_gcl_main(M, z, z, z, z, o.get("area"), o.get("dNx"), o.get("dNy"), o.get("area"), lc.inpoel, E, P, 1e-3)
    assert np.array_equal(M, M0)                                          # W = 0: no correction


def test_smoothing_restatement():
    raw = meshgen.channel(nx=25, ny=9, jitter=0.42, seed=5)
    lc = deck.load(raw)
    L = orclib.lib()

    def mu_min(X, Y):
        inp = lc.inpoel - 1
        x, y = X[inp], Y[inp]
        a2 = x[:, 1] * y[:, 2] + x[:, 2] * y[:, 0] + x[:, 0] * y[:, 1] - (x[:, 1] * y[:, 0] + x[:, 2] * y[:, 1] + x[:, 0] * y[:, 2])
        l = sum((x[:, i] - x[:, j]) ** 2 + (y[:, i] - y[:, j]) ** 2 for i, j in ((2, 1), (0, 2), (1, 0)))
        return (3.46410161513775 * a2 / l).min()

    X, Y = lc.X.copy(), lc.Y.copy()
    before = mu_min(X, Y)
    assert before < 0.85                                   # the optimiser has something to do
    sweeps = L.orc_smoothing(X, Y, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem)
    assert sweeps >= 1 and mu_min(X, Y) > before
    fixed = lc.smooth_fix.astype(bool)
    assert np.array_equal(X[fixed], lc.X[fixed]) and np.array_equal(Y[fixed], lc.Y[fixed])
    X2, Y2 = lc.X.copy(), lc.Y.copy()
    L.orc_smoothing(X2, Y2, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem)
    assert np.array_equal(X, X2) and np.array_equal(Y, Y2)  # deterministic
    # a right-isosceles structured mesh has mu = 0.866 > 0.85 everywhere: smoothing returns at once
    Xs, Ys = np.meshgrid(np.arange(6.0), np.arange(5.0))
    tri = []
    for j in range(4):
        for i in range(5):
            a = j * 6 + i + 1
            tri += [[a, a + 1, a + 7], [a, a + 7, a + 6]]
    tri = np.array(tri, np.int32)
    Xf, Yf = Xs.ravel().copy(), Ys.ravel().copy()
    assert L.orc_smoothing(Xf, Yf, tri, np.zeros(30, np.uint8), 30, len(tri)) == 0
