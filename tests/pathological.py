"""Call-site parity on operands outside the fast paths of the branch-free kernels (exact.cuh: Recip::div, sqrt_nb, div3_nb).

A handful of elements of a small channel mesh get subnormal, zero, infinite, negative or overflowing nodal values; those
elements leave the optimistic straight-line code and are recomputed by the plain form (calcrhs_one_plain, estab_plain,
deltat_plain).  The CUDA result must equal the oracle's bit for bit (NaN == NaN whatever the payload: x86 and the GPU
generate different default NaNs).  Used by tests/test_gpu_parity.py and by tests/opt_worker.py (so that it also runs with
CFDB_CALCRHS_NB forced on and off)."""
import numpy as np


def _same(a, b, name):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    both_nan = np.isnan(a) & np.isnan(b)
    ne = (a.view(np.uint64) != b.view(np.uint64)) & ~both_nan
    if ne.any():
        i = int(np.flatnonzero(ne)[0])
        raise AssertionError(f"{name}: {int(ne.sum())}/{a.size} entries differ; first at {i}: {a[i]!r} vs {b[i]!r}")


def check(lc, g, o):
    """lc: a loaded viscous channel case; g: NSComp2D; o: Oracle.  Returns the number of elements made pathological."""
    from cfd_b200.meshgen import density_bump
    from oracle import orclib

    for k, v in density_bump(lc).items():
        g.set(k, v)
        o.set(k, v)
    o.step(3)
    L = orclib.lib()
    P, E = lc.npoin, lc.nelem
    p = lc.par
    inp = np.asarray(lc.inpoel).reshape(E, 3) - 1          # (3,E) Fortran order, 1-based
    U0, T0 = o.get("U").copy(), o.get("T").copy()
    vx0, vy0, GAMM = o.get("VEL_X").copy(), o.get("VEL_Y").copy(), o.get("GAMM")
    dNx, dNy, area0 = o.get("dNx"), o.get("dNy"), o.get("area").copy()
    shoc, t1, t2, t3 = o.get("SHOC"), o.get("T_SUGN1"), o.get("T_SUGN2"), o.get("T_SUGN3")
    dtl = o.get("DTL").copy()

    # elements far enough apart not to share nodes
    picks, used = [], set()
    for e in range(0, E, 7):
        ns = set(int(n) for n in inp[e])
        if not (ns & used):
            picks.append(e)
            used |= ns
        if len(picks) == 16:
            break
    assert len(picks) == 16
    U, T, vx, vy, area = U0.copy().reshape(P, 4), T0.copy(), vx0.copy(), vy0.copy(), area0.copy()
    tiny, sub = 1e-200, 5e-320
    edits = [
        lambda n: U.__setitem__((n, 1), sub),                       # subnormal numerators
        lambda n: U.__setitem__((n, 2), -sub),
        lambda n: U.__setitem__((n, 0), 1e-310),                    # subnormal divisor
        lambda n: U.__setitem__((n, 0), 0.0),                       # x/0
        lambda n: U.__setitem__((n, 0), -1.0),                      # negative density
        lambda n: U.__setitem__((n, 3), np.inf),
        lambda n: U.__setitem__((n, slice(1, 3)), 1e200),           # squares overflow
        lambda n: U.__setitem__((n, slice(1, 3)), tiny),            # squares underflow
        lambda n: T.__setitem__(n, 0.0),                            # fmu == 0 branch of estab
        lambda n: T.__setitem__(n, -5.0),                           # sqrt of a negative
        lambda n: T.__setitem__(n, np.inf),
        lambda n: T.__setitem__(n, 1e-310),
        lambda n: (vx.__setitem__(n, tiny), vy.__setitem__(n, -tiny)),
        lambda n: (vx.__setitem__(n, 1e200), vy.__setitem__(n, 1e200)),
        lambda n: (vx.__setitem__(n, sub), vy.__setitem__(n, 0.0)),
        lambda n: (vx.__setitem__(n, np.nan)),
    ]
    for e, ed in zip(picks, edits):
        for n in inp[e]:
            ed(int(n))
    area[picks[0]] = 1e-300        # rt*area*dtl in the subnormal range
    area[picks[1]] = 0.0
    U = U.ravel()
    theta = np.zeros(4 * P)
    rng = np.random.default_rng(3)
    wx, wy = np.zeros(P), np.zeros(P)
    with np.errstate(all="ignore"):
        for mu_ref in (0.0, p["FMU"]):
            rhs0 = rng.standard_normal(4 * P)
            r_o = rhs0.copy()
            L.orc_calcrhs(r_o, U, theta, T, dNx, dNy, area, shoc, dtl, t1, t2, t3, lc.inpoel, E, P, p["FCv"], p["FK"], mu_ref,
                          p["GAMA"], p["T_inf"], p["CTE"])
            r_g = g.calcrhs(rhs0.copy(), U, theta, T, dNx, dNy, area, shoc, dtl, t1, t2, t3, p["FCv"], p["FK"], mu_ref,
                            p["GAMA"], p["T_inf"], p["CTE"])
            _same(r_g, r_o, f"calcrhs mu_ref={mu_ref}")
            assert np.isnan(r_o).any() and np.isfinite(r_o).sum() > 0.8 * r_o.size
        dt_o, dtmin_o = np.zeros(E), np.zeros(1)
        L.orc_deltat(E, lc.inpoel, area, T, vx, vy, wx, wy, p["FSAFE"], p["FR"], p["GAMA"], p["T_inf"], dt_o, dtmin_o)
        dtmin_g, dt_g = g.deltat(area, T, vx, vy, wx, wy, p["FSAFE"], p["FR"], p["GAMA"], p["T_inf"])
        _same(dt_g, dt_o, "DT")
        _same([dtmin_g], dtmin_o, "DTMIN")
        outs_o = [np.zeros(E) for _ in range(4)]
        dtmin = 1e-6
        L.orc_estab(E, lc.inpoel, U, T, vx, vy, wx, wy, GAMM, dNx, dNy, p["FR"], dtmin, p["RHO_inf"], p["T_inf"], *outs_o)
        outs_g = g.estab(U, T, vx, vy, wx, wy, GAMM, dNx, dNy, p["FR"], dtmin, p["RHO_inf"], p["T_inf"])
        for a, b, n in zip(outs_g, outs_o, ("SHOC", "T_SUGN1", "T_SUGN2", "T_SUGN3")):
            _same(a, b, n)
    return len(picks)
