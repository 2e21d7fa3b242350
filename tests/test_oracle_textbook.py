"""Independent pin of the oracle's transcription of calcRHS.f90 / FUENTE: the sympy-expanded expressions
(calcRHS.f90:73-132) are compared with a from-scratch numpy evaluation of what they stand for — the Euler flux
Jacobians A1, A2 of an ideal gas in conservative variables, the Navier-Stokes viscous flux (Stokes hypothesis,
Fourier conduction) and the SUPG / shock-capturing weak form — written from the textbook definitions, not from the
reference's code.  Agreement to round-off shows the oracle carries no transcription slip (the reference ships no
test of its own that could).  Rounding ORDER is not tested here; that is what the bit-exact GPU
parity tests are for."""
import numpy as np

from cfd_b200 import deck, meshgen
from oracle import orclib
from oracle.orclib import Oracle


def euler_jacobians(u, g):
    rho, m1, m2, E = u
    v1, v2, e = m1 / rho, m2 / rho, E / rho
    q = v1 * v1 + v2 * v2
    A1 = np.array([[0, 1, 0, 0],
                   [0.5 * (g - 1) * q - v1 * v1, (3 - g) * v1, -(g - 1) * v2, g - 1],
                   [-v1 * v2, v2, v1, 0],
                   [v1 * ((g - 1) * q - g * e), g * e - 0.5 * (g - 1) * q - (g - 1) * v1 * v1, -(g - 1) * v1 * v2, g * v1]])
    A2 = np.array([[0, 0, 1, 0],
                   [-v1 * v2, v2, v1, 0],
                   [0.5 * (g - 1) * q - v2 * v2, -(g - 1) * v1, (3 - g) * v2, g - 1],
                   [v2 * ((g - 1) * q - g * e), -(g - 1) * v1 * v2, g * e - 0.5 * (g - 1) * q - (g - 1) * v2 * v2, g * v2]])
    return A1, A2


def viscous_flux(u, Ux, Uy, mu, lam, Cv):
    """F_v^x, F_v^y from conservative gradients: tau_ij with Stokes' hypothesis, q = lam * grad T, T = (e - V^2/2)/Cv."""
    rho, m1, m2, E = u
    v1, v2, e = m1 / rho, m2 / rho, E / rho
    gx = [(Ux[1] - v1 * Ux[0]) / rho, (Ux[2] - v2 * Ux[0]) / rho, (Ux[3] - e * Ux[0]) / rho]   # u_x, v_x, e_x
    gy = [(Uy[1] - v1 * Uy[0]) / rho, (Uy[2] - v2 * Uy[0]) / rho, (Uy[3] - e * Uy[0]) / rho]
    Tx = (gx[2] - v1 * gx[0] - v2 * gx[1]) / Cv
    Ty = (gy[2] - v1 * gy[0] - v2 * gy[1]) / Cv
    txx = mu * (4 / 3 * gx[0] - 2 / 3 * gy[1])
    tyy = mu * (4 / 3 * gy[1] - 2 / 3 * gx[0])
    txy = mu * (gy[0] + gx[1])
    Fx = np.array([0, txx, txy, v1 * txx + v2 * txy + lam * Tx])
    Fy = np.array([0, txy, tyy, v1 * txy + v2 * tyy + lam * Ty])
    return Fx, Fy


def textbook_rhs(lc, U, theta, T, dNx, dNy, area, shoc, dtl, ts, Cv, lam_ref, mu_ref, g, T_inf, cte):
    P = lc.npoin
    rhs = np.zeros((P, 4))
    Ngp = np.array([[0, .5, .5], [.5, 0, .5], [.5, .5, 0]])       # shape functions at the edge mid-points
    for e, tri in enumerate(lc.inpoel - 1):
        Un, Th = U[tri], theta[tri]                                # (3,4)
        Ux, Uy = dNx[e] @ Un, dNy[e] @ Un
        Tavg = T[tri].mean()
        mu = mu_ref * (Tavg / T_inf) ** 1.5 * (T_inf + 110) / (Tavg + 110)
        lam = lam_ref * (Tavg / T_inf) ** 1.5 * (T_inf + 194) / (Tavg + 194)
        acc = np.zeros((3, 4))
        for k in range(3):
            uk, thk = Ngp[k] @ Un, Ngp[k] @ Th
            A1, A2 = euler_jacobians(uk, g)
            adv = A1 @ Ux + A2 @ Uy
            res = adv + thk
            for n in range(3):
                acc[n] += Ngp[k, n] * adv                                           # Galerkin
                acc[n] += ts[n][e] * (dNx[e, n] * (A1 @ res) + dNy[e, n] * (A2 @ res))   # SUPG, tau indexed by node
                acc[n] += shoc[e] * cte * (dNx[e, n] * Ux + dNy[e, n] * Uy)         # shock capturing
                if mu_ref > 0:
                    Fx, Fy = viscous_flux(uk, Ux, Uy, mu, lam, Cv)
                    acc[n] += dNx[e, n] * Fx + dNy[e, n] * Fy                       # viscous (rows 2..4)
        rhs[tri] += acc * area[e] * dtl[e] / 3.0
    return rhs


def test_calcrhs_and_fuente_match_textbook_definitions():
    lc = deck.load(meshgen.channel(nx=17, ny=7, FMU=1.8e-5, FK=0.0257, mach=0.8))
    o = Oracle(lc)
    for k, v in meshgen.density_bump(lc, amp=0.2).items():
        o.set(k, v)
    o.step(8)
    L = orclib.lib()
    P, E = lc.npoin, lc.nelem
    p = lc.par
    rng = np.random.default_rng(4)
    U, T = o.get("U"), o.get("T")
    theta = 1e-2 * rng.standard_normal(4 * P) * np.abs(U)
    dNx, dNy, area = o.get("dNx"), o.get("dNy"), o.get("area")
    shoc = o.get("SHOC") + 1e-3 * rng.random(E)
    ts = [o.get("T_SUGN1") * (1 + 0.3 * rng.random(E)), o.get("T_SUGN2") + 1e-6 * rng.random(E), o.get("T_SUGN3") * (1 + 0.3 * rng.random(E))]
    dtl = o.get("DTL") * (1 + 0.1 * rng.random(E))
    for mu_ref, lam_ref in ((0.0, 0.0), (1.8e-2, 25.7)):      # exaggerated transport coefficients so the viscous part is visible
        rhs = np.zeros(4 * P)
        L.orc_calcrhs(rhs, U, theta, T, dNx, dNy, area, shoc, dtl, ts[0], ts[1], ts[2], lc.inpoel, E, P, p["FCv"], lam_ref, mu_ref,
                      p["GAMA"], p["T_inf"], p["CTE"])
        ref = textbook_rhs(lc, U.reshape(-1, 4), theta.reshape(-1, 4), T, dNx.reshape(-1, 3), dNy.reshape(-1, 3), area, shoc, dtl, ts,
                           p["FCv"], lam_ref, mu_ref, p["GAMA"], p["T_inf"], p["CTE"])
        got = rhs.reshape(-1, 4)
        scale = np.abs(ref).max(0)
        assert np.max(np.abs(got - ref) / scale) < 1e-11, (mu_ref, np.max(np.abs(got - ref) / scale))
        if mu_ref:
            assert np.max(np.abs(got - inviscid) / scale) > 1e-6      # the viscous terms really contribute
        inviscid = got.copy()
    # FUENTE: RHS_i -= area*dtl/3 * sum_k N_ik (w_k . grad U), w interpolated to the same three mid-points
    wx, wy = rng.standard_normal(P), rng.standard_normal(P)
    rhs = np.zeros(4 * P)
    L.orc_fuente(rhs, U, wx, wy, dNx, dNy, area, dtl, lc.inpoel, E)
    ref = np.zeros((P, 4))
    Ngp = np.array([[.5, .5, 0], [0, .5, .5], [.5, 0, .5]])
    Um = U.reshape(-1, 4)
    for e, tri in enumerate(lc.inpoel - 1):
        Ux, Uy = dNx.reshape(-1, 3)[e] @ Um[tri], dNy.reshape(-1, 3)[e] @ Um[tri]
        for k in range(3):
            wk = (Ngp[k] @ wx[tri], Ngp[k] @ wy[tri])
            for n in range(3):
                ref[tri[n]] -= area[e] * dtl[e] / 3.0 * Ngp[k, n] * (Ux * wk[0] + Uy * wk[1])
    assert np.max(np.abs(rhs.reshape(-1, 4) - ref) / np.abs(ref).max(0)) < 1e-12


def test_cuarto_orden_matches_textbook_projection():
    """CUARTO_ORDEN (subrutinas.f90:220-329): theta_i = -(1/M_i) sum_e area/3 sum_k N_ik (A1 U_x + A2 U_y)(U_k)."""
    lc = deck.load(meshgen.channel(nx=17, ny=7, mach=0.8))
    o = Oracle(lc)
    for k, v in meshgen.density_bump(lc, amp=0.2).items():
        o.set(k, v)
    o.step(5)
    o.set_scalar("use_cuarto", 1)
    d = o.step_part1()
    o.step_part2(d)
    o.rk_stage(1)
    got = o.get("UN").reshape(-1, 4)
    U, M, area = o.get("U").reshape(-1, 4), o.get("M"), o.get("area")
    dNx, dNy = o.get("dNx").reshape(-1, 3), o.get("dNy").reshape(-1, 3)
    g = lc.par["GAMA"]
    ref = np.zeros_like(U)
    Ngp = np.array([[.5, .5, 0], [0, .5, .5], [.5, 0, .5]])
    for e, tri in enumerate(lc.inpoel - 1):
        Ux, Uy = dNx[e] @ U[tri], dNy[e] @ U[tri]
        for k in range(3):
            A1, A2 = euler_jacobians(Ngp[k] @ U[tri], g)
            adv = A1 @ Ux + A2 @ Uy
            for n in range(3):
                ref[tri[n]] += Ngp[k, n] * adv * area[e] / 3.0
    ref = -ref / M[:, None]
    assert np.max(np.abs(got - ref) / np.abs(ref).max(0)) < 1e-11


def test_deltat_and_estab_second_transcription():
    """A second, vectorised numpy transcription of deltat (subrutinas.f90:172-215) and ESTAB (:349-443), written
    separately from oracle.cpp: double-entry check of the two ad-hoc routines (agreement to round-off; the exact-zero
    switches of SURVEY.md F9 are compared where both sides are away from them)."""
    lc = deck.load(meshgen.wedge(nx=25, ny=13, mach=2.0))
    o = Oracle(lc)
    for k, v in meshgen.density_bump(lc, amp=0.2).items():
        o.set(k, v)
    o.step(6)
    L = orclib.lib()
    p = lc.par
    E = lc.nelem
    rng = np.random.default_rng(9)
    tri = lc.inpoel - 1
    U, T, vx, vy, G = o.get("U").reshape(-1, 4), o.get("T"), o.get("VEL_X"), o.get("VEL_Y"), o.get("GAMM")
    wx, wy = 5 * rng.standard_normal(lc.npoin), 5 * rng.standard_normal(lc.npoin)
    area, dNx, dNy = o.get("area"), o.get("dNx").reshape(-1, 3), o.get("dNy").reshape(-1, 3)
    # ---- deltat
    dt_o, dtmin_o = np.zeros(E), np.zeros(1)
    L.orc_deltat(E, lc.inpoel, area, T, vx, vy, wx, wy, p["FSAFE"], p["FR"], p["GAMA"], p["T_inf"], dt_o, dtmin_o)
    Tm = T[tri].mean(1)
    vu, vv = np.abs(vx - wx)[tri].max(1), np.abs(vy - wy)[tri].max(1)
    hh = np.sqrt(2 * area)
    vel = np.sqrt(vu**2 + vv**2)
    fmu = 0.017 * (Tm / p["T_inf"]) ** 1.5 * (p["T_inf"] + 110) / (Tm + 110)
    alpha = np.minimum(vel * hh / (2 * fmu) / 3, 1.0)
    dtu, dtc = 1 / (4 * fmu / hh**2 + alpha * vel / hh), 1 / (4 * fmu / hh**2)
    dt = p["FSAFE"] / (1 / dtc + 1 / dtu)
    assert abs(dt.min() - dtmin_o[0]) <= 1e-14 * dt.min()
    assert np.allclose(np.minimum(dt, 10 * dt.min()), dt_o, rtol=1e-13, atol=0)
    # ---- ESTAB
    out = [np.zeros(E) for _ in range(4)]
    L.orc_estab(E, lc.inpoel, o.get("U"), T, vx, vy, wx, wy, G, o.get("dNx"), o.get("dNy"), p["FR"], dtmin_o[0], p["RHO_inf"],
                p["T_inf"], *out)
    gm = G[tri].mean(1)
    rho_e = U[tri, 0].mean(1)
    VX, VY = vx[tri].mean(1) - wx[tri].mean(1), vy[tri].mean(1) - wy[tri].mean(1)
    vel2 = np.hypot(VX, VY)

    def grad(f):
        return (f[tri] * dNx).sum(1), (f[tri] * dNy).sum(1)

    drx, dry = grad(U[:, 0])
    dtx, dty = grad(T)
    dux, duy = vel2 * dNx[:, 0] + vel2 * dNx[:, 1] + vel2 * dNx[:, 2], vel2 * dNy[:, 0] + vel2 * dNy[:, 1] + vel2 * dNy[:, 2]
    dr2, dt2, du2 = np.hypot(drx, dry) + 1e-20, np.hypot(dtx, dty) + 1e-20, np.hypot(dux, duy) + 1e-20
    c = np.sqrt(gm * p["FR"] * Tm)
    fm = 0.017 * (Tm / p["T_inf"]) ** 1.5 * (p["T_inf"] + 110) / (Tm + 110)

    def proj(ax, ay):
        return np.abs(ax[:, None] * dNx + ay[:, None] * dNy).sum(1)

    with np.errstate(divide="ignore", invalid="ignore"):
        tau = 1 / (proj(VX, VY) + c * proj(drx / dr2, dry / dr2))
        h_rgne = 2 / proj(dtx / dt2, dty / dt2)
        h_rgn = 2 / proj(dux / du2, duy / du2)
        h_rgn = np.where(h_rgn > 10, 0.0, h_rgn)
        h_jgn = 2 / proj(drx / dr2, dry / dr2)
        h_jgn = np.where(h_jgn > 10, 0.0, h_jgn)
        tr1 = dr2 * h_jgn / rho_e
        shoc = (tr1 + tr1**2) * .5 * c**2 * h_jgn / (2 * c)
        res = 1 / tau**2 + (2 / dtmin_o[0]) ** 2
        t1 = res**-.5
        t2 = (res + 1 / (h_rgn**2 / (4 * fm / p["RHO_inf"])) ** 2) ** -.5
        t3 = (res + 1 / (h_rgne**2 / (4 * fm / p["RHO_inf"])) ** 2) ** -.5
    assert np.allclose(out[0], shoc, rtol=1e-9, atol=1e-300) and np.allclose(out[1], t1, rtol=1e-12)
    assert np.allclose(out[3], t3, rtol=1e-9)
    # T_SUGN2 rides on round-off noise (F9): compare only where both transcriptions see the same zero / non-zero DU
    same = (out[2] == 0) == (t2 == 0)
    assert same.mean() > 0.6 and np.allclose(out[2][same], t2[same], rtol=1e-9)
    assert set(np.unique(np.round(out[2][out[2] > 0] / out[1][out[2] > 0], 3))) <= {1.0}   # when alive, tau2 ~ tau1


def test_laplace_and_bicg_against_scipy():
    """Mlaplace::laplace values against an independent scipy assembly, and biCG's answer against a direct sparse solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    lc = deck.load(meshgen.ale_body(nt=40, nr=12))
    o = Oracle(lc)
    P = lc.npoin
    tri = lc.inpoel - 1
    X, Y = o.get("X"), o.get("Y")
    dNx, dNy = o.get("dNx").reshape(-1, 3), o.get("dNy").reshape(-1, 3)
    x, y = X[tri], Y[tri]
    a2 = x[:, 1] * y[:, 2] + x[:, 2] * y[:, 0] + x[:, 0] * y[:, 1] - (x[:, 1] * y[:, 0] + x[:, 2] * y[:, 1] + x[:, 0] * y[:, 2])
    l = sum((x[:, i] - x[:, j]) ** 2 + (y[:, i] - y[:, j]) ** 2 for i, j in ((2, 1), (0, 2), (1, 0)))
    q = 1.0 / (3.46410161513775 * a2 / l) ** 2
    rows, cols, vals = [], [], []
    for i in range(3):
        for j in range(3):
            rows.append(tri[:, i]); cols.append(tri[:, j]); vals.append((dNx[:, i] * dNx[:, j] + dNy[:, i] * dNy[:, j]) * q)
    K = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(P, P)).tocsr()
    spv, idx, rp, dg = o.get("lap_sparse"), o.get("lap_idx"), o.get("lap_rowptr"), o.get("lap_diag")
    A = sp.csr_matrix((spv, idx - 1, rp), shape=(P, P))
    assert abs(A - K).max() <= 1e-12 * abs(K).max()
    assert np.allclose(dg, K.diagonal(), rtol=1e-12)
    # Dirichlet problem: body nodes displaced, outer ring fixed; interior rows of K x = 0
    fix = lc.ilaux - 1
    xf = np.concatenate([1e-3 * np.cos(np.arange(lc.i_m.size)), np.zeros(lc.ifm.size)])
    free = np.setdiff1d(np.arange(P), fix)
    xs = np.zeros(P)
    xs[fix] = xf
    xs[free] = spla.spsolve(K[free][:, free].tocsc(), -K[free][:, fix] @ xf)
    xo = np.zeros(P)
    it = orclib.lib().orc_bicg(spv, idx, rp, dg, xo, np.zeros(P), xf, lc.ilaux, P, fix.size)
    # the reference stops at |r.z| <= 1e-10, an ABSOLUTE tolerance: ~1e-6 on displacements of 1e-3 (measured 1.4e-6)
    assert it > 3 and np.max(np.abs(xo - xs)) < 5e-3 * np.abs(xf).max()
