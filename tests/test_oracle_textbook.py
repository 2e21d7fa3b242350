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
