"""Seeded synthetic triangulations for the BASELINE.json configs.

All meshes are jittered structured lattices mapped to the physical domain; every lattice quad
is split along the diagonal that satisfies the Delaunay (empty-circumcircle) criterion for its
four corners, which gives an irregular, valence-4..8 connectivity in O(N) time at 64 M
triangles.  `delaunay=True` re-triangulates the same point cloud with scipy/Qhull (true global
Delaunay; practical up to a few million points) and keeps only triangles whose centroid lies
in the domain.  Nodes are numbered lattice-row-major and elements quad-row-major, so contiguous
element ranges are compact strips (the multi-GPU partitioner relies on that, never on the
generator itself).  Triangles are counter-clockwise (positive `area` in deriv,
subrutinas.f90:106).

The reference ships no meshes or decks; these cases are ours.
"""
from __future__ import annotations

import math

import numpy as np

from .deck import I32, RawCase


def _lattice(nx, ny, jitter, seed, periodic_i=False):
    """Parameter-space coordinates (xi, eta) in [0,1]^2, interior nodes jittered."""
    rng = np.random.default_rng(seed)
    i = np.arange(nx, dtype=np.float64)
    j = np.arange(ny, dtype=np.float64)
    hx = 1.0 / (nx if periodic_i else nx - 1)
    hy = 1.0 / (ny - 1)
    XI, ETA = np.meshgrid(i * hx, j * hy)  # shape (ny, nx); node id = j*nx + i
    jx = rng.uniform(-jitter, jitter, size=(ny, nx)) * hx
    jy = rng.uniform(-jitter, jitter, size=(ny, nx)) * hy
    jy[0, :] = 0.0
    jy[-1, :] = 0.0
    jx[0, :] = 0.0
    jx[-1, :] = 0.0
    if not periodic_i:
        jx[:, 0] = 0.0
        jx[:, -1] = 0.0
        jy[:, 0] = 0.0
        jy[:, -1] = 0.0
    return XI + jx, ETA + jy


def _incircle(ax, ay, bx, by, cx, cy, dx, dy):
    """>0 when d lies strictly inside the circumcircle of the CCW triangle a,b,c."""
    adx, ady = ax - dx, ay - dy
    bdx, bdy = bx - dx, by - dy
    cdx, cdy = cx - dx, cy - dy
    ad, bd, cd = adx * adx + ady * ady, bdx * bdx + bdy * bdy, cdx * cdx + cdy * cdy
    return adx * (bdy * cd - bd * cdy) - ady * (bdx * cd - bd * cdx) + ad * (bdx * cdy - bdy * cdx)


def _triangulate(X, Y, nx, ny, periodic_i=False):
    """Split every lattice quad along its locally-Delaunay diagonal; returns (nelem,3) 1-based, CCW."""
    nqx = nx if periodic_i else nx - 1
    ii, jj = np.meshgrid(np.arange(nqx), np.arange(ny - 1))
    ii = ii.ravel()
    jj = jj.ravel()
    ip = (ii + 1) % nx
    a = jj * nx + ii
    b = jj * nx + ip
    c = (jj + 1) * nx + ip
    d = (jj + 1) * nx + ii
    Xf, Yf = X.ravel(), Y.ravel()
    orient = (Xf[b] - Xf[a]) * (Yf[d] - Yf[a]) - (Xf[d] - Xf[a]) * (Yf[b] - Yf[a])
    sgn = np.where(orient >= 0, 1.0, -1.0)  # mapped lattices may be clockwise
    inc = _incircle(Xf[a], Yf[a], Xf[b], Yf[b], Xf[c], Yf[c], Xf[d], Yf[d]) * sgn
    use_bd = inc > 0  # d inside circle(a,b,c): diagonal a-c is not Delaunay
    t1 = np.where(use_bd[:, None], np.stack([a, b, d], 1), np.stack([a, b, c], 1))
    t2 = np.where(use_bd[:, None], np.stack([b, c, d], 1), np.stack([a, c, d], 1))
    tri = np.empty((2 * a.size, 3), dtype=np.int64)
    tri[0::2] = t1
    tri[1::2] = t2
    return _ccw(Xf, Yf, tri)


def _ccw(Xf, Yf, tri):
    x1, x2, x3 = Xf[tri[:, 0]], Xf[tri[:, 1]], Xf[tri[:, 2]]
    y1, y2, y3 = Yf[tri[:, 0]], Yf[tri[:, 1]], Yf[tri[:, 2]]
    area2 = (x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)
    flip = area2 < 0
    tri[flip, 1], tri[flip, 2] = tri[flip, 2].copy(), tri[flip, 1].copy()
    if np.any(area2 == 0):
        raise ValueError("degenerate triangle in synthetic mesh")
    return (tri + 1).astype(I32)


def _scipy_delaunay(Xf, Yf, inside):
    from scipy.spatial import Delaunay

    tri = Delaunay(np.stack([Xf, Yf], 1)).simplices.astype(np.int64)
    cx, cy = Xf[tri].mean(1), Yf[tri].mean(1)
    keep = inside(cx, cy)
    tri, cx, cy = tri[keep], cx[keep], cy[keep]
    order = np.lexsort((cx, np.floor(cy * 64)))  # strip-sorted for locality
    return _ccw(Xf, Yf, tri[order])


def _boundary_edges_row(nx, j, reverse=False):
    n = j * nx + np.arange(nx) + 1
    e = np.stack([n[:-1], n[1:]], 1)
    return e[:, ::-1] if reverse else e


def channel(nx=201, ny=51, Lx=4.0, Ly=1.0, bump=0.05, mach=0.5, jitter=0.25, seed=12345, name="channel", **kw) -> RawCase:
    """Config 1: ~10k-node channel with a sin^2 bump on the lower wall; inflow x=0, slip walls."""
    XI, ETA = _lattice(nx, ny, jitter, seed)
    x = XI * Lx
    yb = np.where((x > 1.5) & (x < 2.5), bump * np.sin(np.pi * (x - 1.5)) ** 2, 0.0)
    y = yb + ETA * (Ly - yb)
    inpoel = _triangulate(x, y, nx, ny)
    left = (np.arange(ny) * nx + 1).astype(I32)
    wall = np.concatenate([_boundary_edges_row(nx, 0), _boundary_edges_row(nx, ny - 1, reverse=True)]).astype(I32)
    bnd = np.unique(np.concatenate([wall.ravel(), left, (np.arange(ny) * nx + nx).astype(I32)])).astype(I32)
    return RawCase(
        name=name, X=x.ravel().copy(), Y=y.ravel().copy(), inpoel=inpoel, MACH_inf=mach,
        fixrho=(left, np.ones(left.size)), fixvi=(left, np.ones(left.size), np.ones(left.size)),
        wall=wall, ifm=bnd, **kw,
    )


def wedge(nx=1001, ny=501, Lx=2.0, Ly=1.0, x0=0.5, angle_deg=15.0, mach=2.5, jitter=0.25, seed=12345, name="wedge", **kw) -> RawCase:
    """Config 2: supersonic compression ramp; supersonic inflow fixes rho, velocity and T."""
    XI, ETA = _lattice(nx, ny, jitter, seed)
    x = XI * Lx
    yb = np.maximum(0.0, (x - x0) * math.tan(math.radians(angle_deg)))
    y = yb + ETA * (Ly - yb)
    inpoel = _triangulate(x, y, nx, ny)
    left = (np.arange(ny) * nx + 1).astype(I32)
    wall = np.concatenate([_boundary_edges_row(nx, 0), _boundary_edges_row(nx, ny - 1, reverse=True)]).astype(I32)
    bnd = np.unique(np.concatenate([wall.ravel(), left, (np.arange(ny) * nx + nx).astype(I32)])).astype(I32)
    return RawCase(
        name=name, X=x.ravel().copy(), Y=y.ravel().copy(), inpoel=inpoel, MACH_inf=mach,
        fixrho=(left, np.ones(left.size)), fixvi=(left, np.ones(left.size), np.ones(left.size)),
        fixt=(left, np.ones(left.size)), wall=wall, ifm=bnd, **kw,
    )


def ale_body(nt=256, nr=64, a=0.5, b=0.06, R=8.0, mach=0.5, jitter=0.2, seed=12345, name="alebody", **kw) -> RawCase:
    """Config 3: O-mesh around an ellipse centred on (XREF1,YREF1)=(0,0) that pitches (meshMove.f90:70).

    Body nodes: I_M (moving) + slip-wall edges + ISET set 1; outer ring: IFM (fixed) + far-field
    inflow values.  MOVING=1 so NORMALES/DERIV/MASAS/laplace are recomputed every step.
    """
    XI, ETA = _lattice(nt, nr, jitter, seed, periodic_i=True)
    th = -2.0 * np.pi * XI  # clockwise in i so (i,j) maps counter-clockwise
    s = (np.expm1(3.0 * ETA)) / math.expm1(3.0)  # radial stretching
    xb, yb = a * np.cos(th), b * np.sin(th)
    xo, yo = R * np.cos(th), R * np.sin(th)
    x = xb + s * (xo - xb)
    y = yb + s * (yo - yb)
    inpoel = _triangulate(x, y, nt, nr, periodic_i=True)
    body = (np.arange(nt) + 1).astype(I32)
    outer = ((nr - 1) * nt + np.arange(nt) + 1).astype(I32)
    nxt = (np.arange(nt) + 1) % nt + 1
    wall = np.stack([body, nxt], 1).astype(I32)
    # element adjacent to each body edge: first triangle of quad (i, j=0)
    elem = (2 * np.arange(nt) + 1).astype(I32)
    sets = np.stack([elem, body, nxt.astype(I32), np.ones(nt, I32)], 1).astype(I32)
    kw.setdefault("MOVING", 1)
    return RawCase(
        name=name, X=x.ravel().copy(), Y=y.ravel().copy(), inpoel=inpoel, MACH_inf=mach,
        fixrho=(outer, np.ones(nt)), fixvi=(outer, np.ones(nt), np.ones(nt)), wall=wall, sets=sets,
        ifm=outer, i_m=body, XREF1=0.0, YREF1=0.0, **kw,
    )


def square(n=2829, mach=0.5, jitter=0.3, seed=12345, name="square", nx=None, ny=None, **kw) -> RawCase:
    """Configs 4/5: unit-square (or nx x ny strip) jittered-lattice domain; n=2829 -> 16.0 M triangles."""
    nx = nx or n
    ny = ny or n
    XI, ETA = _lattice(nx, ny, jitter, seed)
    Ly = (ny - 1) / (nx - 1)
    x, y = XI, ETA * Ly
    inpoel = _triangulate(x, y, nx, ny)
    left = (np.arange(ny) * nx + 1).astype(I32)
    right = (np.arange(ny) * nx + nx).astype(I32)
    wall = np.concatenate([_boundary_edges_row(nx, 0), _boundary_edges_row(nx, ny - 1, reverse=True)]).astype(I32)
    bnd = np.unique(np.concatenate([wall.ravel(), left, right])).astype(I32)
    return RawCase(
        name=name, X=x.ravel().copy(), Y=y.ravel().copy(), inpoel=inpoel, MACH_inf=mach,
        fixrho=(left, np.ones(left.size)), fixvi=(left, np.ones(left.size), np.ones(left.size)),
        wall=wall, ifm=bnd, **kw,
    )


def retriangulate_delaunay(raw: RawCase, inside=None) -> RawCase:
    """Replace the connectivity by the true Delaunay triangulation (scipy/Qhull) of the same nodes."""
    inside = inside or (lambda cx, cy: np.ones(cx.shape, bool))
    raw.inpoel = _scipy_delaunay(raw.X, raw.Y, inside)
    return raw


def density_bump(lc, amp=0.1, x0=None, y0=None, sigma=None, period_y=None):
    """Smooth initial density perturbation at free-stream pressure and velocity.

    Returns dict(U, T, VEL_X, VEL_Y) to be set on both the oracle and the CUDA solver after init.
    """
    p = lc.par
    X, Y = lc.X, lc.Y
    x0 = 0.5 * (X.min() + X.max()) if x0 is None else x0
    y0 = 0.5 * (Y.min() + Y.max()) if y0 is None else y0
    sigma = 0.15 * min(X.max() - X.min(), Y.max() - Y.min()) if sigma is None else sigma
    dy = Y - y0
    if period_y:  # one bump per strip of height period_y (multi-GPU weak scaling: same flow on every rank)
        dy = (Y - np.floor(Y / period_y) * period_y) - y0
    rho = p["RHO_inf"] * (1.0 + amp * np.exp(-((X - x0) ** 2 + dy**2) / sigma**2))
    pres = p["RHO_inf"] * p["FR"] * p["T_inf"]
    u, v = p["U_inf"], p["V_inf"]
    e = pres / ((p["GAMA"] - 1.0) * rho) + 0.5 * (u * u + v * v)
    U = np.stack([rho, rho * u, rho * v, rho * e], 1)
    return dict(U=np.ascontiguousarray(U), T=pres / (rho * p["FR"]), VEL_X=np.full(X.size, u), VEL_Y=np.full(X.size, v))


# ---- strip-decomposable square (multi-GPU weak scaling) -------------------------------------------------------
def _rows_lattice(nx, j0, j1, ny_total, jitter, seed):
    """Rows j0..j1 (inclusive) of an nx x ny_total jittered lattice with spacing h = 1/(nx-1).
    Every row has its own RNG stream, so any window of rows reproduces the global mesh exactly."""
    h = 1.0 / (nx - 1)
    rows = np.arange(j0, j1 + 1)
    x = np.tile(np.arange(nx, dtype=np.float64) * h, (rows.size, 1))
    y = np.repeat((rows.astype(np.float64) * h)[:, None], nx, axis=1)
    for k, j in enumerate(rows):
        rng = np.random.default_rng([seed, int(j)])
        jx = rng.uniform(-jitter, jitter, nx) * h
        jy = rng.uniform(-jitter, jitter, nx) * h
        jx[0] = jx[-1] = 0.0
        jy[0] = jy[-1] = 0.0
        if j == 0 or j == ny_total - 1:
            jx[:] = 0.0
            jy[:] = 0.0
        x[k] += jx
        y[k] += jy
    return x, y


def square_rows(n, nranks, rank, mach=0.5, jitter=0.3, seed=12345, name="square", all_rows=False, rows_per=None, jrange=None, **kw):
    """Window (own quad rows +-1) of the `nranks`-strip square mesh; returns (RawCase, first global row).
    Each strip has n nodes per row and `rows_per` quad rows (default n-1: weak scaling, one n x n lattice per rank;
    strong scaling passes rows_per = (n-1)/nranks so that the whole domain stays n x n)."""
    nx = n
    rows_per = (n - 1) if rows_per is None else rows_per
    ny_total = nranks * rows_per + 1
    j0 = 0 if all_rows else max(0, rank * rows_per - 1)
    j1 = ny_total - 1 if all_rows else min(ny_total - 1, (rank + 1) * rows_per + 1)
    if jrange is not None:   # node rows chosen by the caller (partition.square_window: chunk-aligned ownership)
        j0, j1 = jrange
    x, y = _rows_lattice(nx, j0, j1, ny_total, jitter, seed)
    nrows = j1 - j0 + 1
    inpoel = _triangulate(x, y, nx, nrows)
    left = (np.arange(nrows) * nx + 1).astype(I32)
    right = (np.arange(nrows) * nx + nx).astype(I32)
    walls = []
    if j0 == 0:
        walls.append(_boundary_edges_row(nx, 0))
    if j1 == ny_total - 1:
        walls.append(_boundary_edges_row(nx, nrows - 1, reverse=True))
    wall = np.concatenate(walls).astype(I32) if walls else np.zeros((0, 2), I32)
    bnd = np.unique(np.concatenate([wall.ravel(), left, right])).astype(I32)
    raw = RawCase(
        name=name, X=x.ravel().copy(), Y=y.ravel().copy(), inpoel=inpoel, MACH_inf=mach,
        fixrho=(left, np.ones(left.size)), fixvi=(left, np.ones(left.size), np.ones(left.size)),
        wall=wall, ifm=bnd, **kw,
    )
    return raw, j0


def square_global(n, nranks, rows_per=None, **kw):
    """The whole `nranks`-strip square mesh in one piece (tests; small sizes)."""
    return square_rows(n, nranks, 0, all_rows=True, rows_per=rows_per, **kw)[0]
