"""Sub-domain decomposition for multi-GPU runs (new in this build; the reference is single-process).

Definition, deterministic and bit-exact against tests' brute-force restatement:
  * the element list *as given* is cut into N equal contiguous ranges, rank(e) = ranges containing e;
  * a node is OWNED by the lowest rank among the elements touching it;
  * rank r computes every element that touches a node it owns (its own range plus a one-element-deep
    layer of higher-rank elements), so the element->node sums of owned nodes are complete and are
    accumulated in ascending global element order — bit-identical to the single-GPU run, with no
    partial-sum exchange at all;
  * the other nodes of those elements are GHOSTS: their state is received from the owner after every
    RK stage (one packed message per neighbour).
Local numbering: owned nodes first (ascending global id), then ghosts (ascending global id); local
elements in ascending global id.  Exchange lists are sorted by global node id on both sides.

Reduction alignment (round 2).  Inner products (biCG) and the residual norms use one fixed "canonical" order over
the GLOBAL node index: chunks of RED_CHUNK = 4096 consecutive nodes, then a recursive tree over the chunk sums
(oracle/orc_math.h canon_sum).  To make an N-rank run reproduce the single-GPU bits, every chunk must be summed by
ONE rank.  When the owned sets of the "lowest rank touching" rule are contiguous ranges of the global numbering (strip
partitions of meshes numbered along the element order -- every mesh of this repository), the range boundaries are
therefore moved to the NEAREST multiple of RED_CHUNK (`align_owner`): rank r owns [B_r, B_{r+1}), B_r % 4096 == 0.
Everything else follows from the ownership as before (a rank computes every element touching a node it owns, a few
thousand elements more or less than with the natural boundary).  The ranks then exchange chunk sums, never partial
sums of a chunk, and run the upper tree levels redundantly: bit-identical to one GPU.  Meshes whose owned sets are
not contiguous (or ranks with fewer than two chunks) keep the natural ownership; their multi-rank inner products
then agree with the single-GPU ones to round-off only (`LocalPart.red_aligned` says which).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np

from .deck import I32, LoadedCase


RED_CHUNK = 4096   # first-level chunk of the canonical reduction order (oracle/orc_math.h, kernels.cuh: dot_chunks)


def element_ranges(nelem: int, nranks: int) -> np.ndarray:
    """bounds[r] .. bounds[r+1] = contiguous element range of rank r (0-based, equal split)."""
    return np.array([(r * nelem) // nranks for r in range(nranks + 1)], dtype=np.int64)


@dataclass
class LocalPart:
    rank: int
    nranks: int
    lc: LoadedCase              # local mesh + BC lists in local 1-based numbering
    node_gid: np.ndarray        # local node -> global node (0-based)
    elem_gid: np.ndarray        # local element -> global element (0-based)
    n_owned: int                # local nodes [0,n_owned) are owned, the rest are ghosts
    elem_own: np.ndarray        # bool: element belongs to this rank's own range (counted once globally)
    neighbors: list = field(default_factory=list)
    send: dict = field(default_factory=dict)   # rank -> local ids of owned nodes that rank needs (by gid)
    recv: dict = field(default_factory=dict)   # rank -> local ids of ghost nodes that rank owns (by gid)
    moving: bool = False        # the GLOBAL case has body sets (fluidStructure moves the mesh on every rank, also on
                                # ranks whose local deck holds none of the set edges)
    red_aligned: bool = False   # owned nodes = the global range [gid0, gid0 + n_owned) with gid0 % RED_CHUNK == 0
    gid0: int = 0               # global id of the first owned node (meaningful when red_aligned)
    npoin_global: int = 0
    set_gidx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))  # local ISET row -> global ISET row

    def halo_arrays(self):
        """Flattened CSR form for cfdb_set_halo: ranks, send_ptr, send_idx, recv_ptr, recv_idx (0-based)."""
        ranks = np.array(self.neighbors, dtype=I32)
        sp, rp = [0], [0]
        si, ri = [], []
        for s in self.neighbors:
            si.append(self.send.get(s, np.zeros(0, I32)))
            ri.append(self.recv.get(s, np.zeros(0, I32)))
            sp.append(sp[-1] + si[-1].size)
            rp.append(rp[-1] + ri[-1].size)
        cat = lambda l: np.ascontiguousarray(np.concatenate(l).astype(I32)) if l else np.zeros(0, I32)  # noqa: E731
        return ranks, np.array(sp, I32), cat(si), np.array(rp, I32), cat(ri)


def node_owner(inpoel0: np.ndarray, elem_rank: np.ndarray, npoin: int) -> np.ndarray:
    """Lowest rank among the elements touching each node (npoin,) ; nodes touched by no element get -1."""
    owner = np.full(npoin, np.iinfo(np.int32).max, dtype=np.int64)
    np.minimum.at(owner, inpoel0.ravel(), np.repeat(elem_rank, 3))
    owner[owner == np.iinfo(np.int32).max] = -1
    return owner


def aligned_boundaries(first_owned: np.ndarray, npoin_global: int, chunk: int = RED_CHUNK):
    """first_owned[r] = global id of the first node rank r owns under the natural rule (first_owned[0] == 0), ranges
    contiguous and ascending.  Returns B (nranks+1,) with B[r] the nearest multiple of `chunk`, or None when a rank would be
    left with fewer than two chunks (tiny meshes keep the natural ownership)."""
    g = np.asarray(first_owned, np.int64)
    B = np.concatenate([((g + chunk // 2) // chunk) * chunk, [npoin_global]]).astype(np.int64)
    B[0] = 0
    if (np.diff(B) < 2 * chunk).any():
        return None
    return B


def align_owner(owner: np.ndarray, node_gid: np.ndarray, nranks: int, npoin_global: int):
    """Natural owner (global mesh) -> chunk-aligned owner, or (owner, None) when the natural owned sets are not contiguous
    ascending ranges of the global numbering."""
    order = np.argsort(node_gid, kind="stable")
    o = owner[order]
    if nranks < 2 or (o < 0).any() or (np.diff(o) < 0).any() or node_gid.size != npoin_global:
        return owner, None
    first = np.searchsorted(o, np.arange(nranks), side="left")          # position == global id (all nodes present)
    if (np.diff(np.concatenate([first, [npoin_global]])) <= 0).any():
        return owner, None
    B = aligned_boundaries(node_gid[order][first], npoin_global)
    if B is None:
        return owner, None
    return (np.searchsorted(B, node_gid, side="right") - 1).astype(np.int64), B


def _need_pairs(inpoel0, owner):
    """Unique (node, rank) pairs: `rank` needs the state of `node`, which it does not own."""
    on = owner[inpoel0]                                 # (E,3) owner of each local node
    mixed = (on.min(1) != on.max(1))
    el, own = inpoel0[mixed], on[mixed]
    pairs = []
    for i in range(3):
        for j in range(3):
            if i != j:
                m = own[:, i] != own[:, j]
                pairs.append(np.stack([el[m, i], own[m, j]], 1))
    if not pairs:
        return np.zeros((0, 2), np.int64)
    return np.unique(np.concatenate(pairs), axis=0)


def build_local(lc: LoadedCase, nranks: int, rank: int, elem_rank: np.ndarray | None = None,
                node_gid: np.ndarray | None = None, elem_gid: np.ndarray | None = None,
                bounds: np.ndarray | None = None, npoin_global: int | None = None, align: bool = True) -> LocalPart:
    """Local part of `rank`.  `lc` is the global mesh, or a window of it that contains every element
    touching a node of the rank's local elements (then pass the window's elem_rank / node_gid / elem_gid, and the
    chunk-aligned ownership boundaries `bounds` the window's generator knows analytically, or None for the natural rule)."""
    E, P = lc.nelem, lc.npoin
    inp0 = lc.inpoel.astype(np.int64) - 1
    if elem_rank is None:
        b = element_ranges(E, nranks)
        elem_rank = np.searchsorted(b, np.arange(E), side="right") - 1
    elem_rank = np.asarray(elem_rank, np.int64)
    node_gid = np.arange(P, dtype=np.int64) if node_gid is None else np.asarray(node_gid, np.int64)
    elem_gid = np.arange(E, dtype=np.int64) if elem_gid is None else np.asarray(elem_gid, np.int64)
    window = npoin_global is not None and int(npoin_global) != P
    npoin_global = P if npoin_global is None else int(npoin_global)
    owner = node_owner(inp0, elem_rank, P)
    B = None
    if align and nranks > 1:
        if window:
            if bounds is not None:
                B = np.asarray(bounds, np.int64)
                owner = (np.searchsorted(B, node_gid, side="right") - 1).astype(np.int64)
        else:
            owner, B = align_owner(owner, node_gid, nranks, npoin_global)
    emask = (owner[inp0] == rank).any(1)
    loc_el = np.flatnonzero(emask)                       # ascending window id == ascending global id
    nodes = np.unique(inp0[loc_el])
    own_nodes = nodes[owner[nodes] == rank]
    ghost_nodes = nodes[owner[nodes] != rank]
    own_nodes = own_nodes[np.argsort(node_gid[own_nodes], kind="stable")]
    ghost_nodes = ghost_nodes[np.argsort(node_gid[ghost_nodes], kind="stable")]
    order = np.concatenate([own_nodes, ghost_nodes])
    g2l = np.full(P, -1, dtype=np.int64)
    g2l[order] = np.arange(order.size)
    n_owned = own_nodes.size

    pairs = _need_pairs(inp0[loc_el], owner)             # all pairs involving this rank live in its local elements
    send, recv = {}, {}
    for s in np.unique(pairs[:, 1]) if pairs.size else []:
        if s == rank:
            continue
        m = (pairs[:, 1] == s) & (owner[pairs[:, 0]] == rank)
        if m.any():
            n = pairs[m, 0]
            send[int(s)] = g2l[n[np.argsort(node_gid[n], kind="stable")]].astype(I32)
    mine = pairs[pairs[:, 1] == rank] if pairs.size else pairs
    for s in np.unique(owner[mine[:, 0]]) if mine.size else []:
        n = mine[owner[mine[:, 0]] == s, 0]
        recv[int(s)] = g2l[n[np.argsort(node_gid[n], kind="stable")]].astype(I32)
    got = np.sort(np.concatenate([v for v in recv.values()])) if recv else np.zeros(0, np.int64)
    assert np.array_equal(got, np.arange(n_owned, order.size)), "ghost set and receive lists disagree"
    neighbors = sorted(set(send) | set(recv))

    # ---- local LoadedCase: lists keep their order (last-entry-wins semantics), nodes renumbered --------------
    def keep(nodes1):
        return g2l[np.asarray(nodes1, np.int64) - 1] >= 0

    def ren(nodes1):
        return (g2l[np.asarray(nodes1, np.int64) - 1] + 1).astype(I32)

    m_rho, m_v, m_t = keep(lc.ifixrho_node), keep(lc.ifixv_node), keep(lc.ifixt_node)
    m_w = keep(lc.wall[:, 0]) & keep(lc.wall[:, 1]) if lc.wall.size else np.zeros(0, bool)
    m_im, m_ifm = keep(lc.i_m), keep(lc.ifm)
    if lc.sets.size:
        m_s = keep(lc.sets[:, 1]) & keep(lc.sets[:, 2])
        s = lc.sets[m_s]
        e_l = np.searchsorted(loc_el, s[:, 0].astype(np.int64) - 1)
        e_l = np.where((e_l < loc_el.size) & (loc_el[np.minimum(e_l, loc_el.size - 1)] == s[:, 0] - 1), e_l + 1, 0)
        sets = np.stack([e_l.astype(I32), ren(s[:, 1]), ren(s[:, 2]), s[:, 3]], 1).astype(I32)
        set_gidx = np.flatnonzero(m_s).astype(np.int64)
    else:
        sets = np.zeros((0, 4), I32)
        set_gidx = np.zeros(0, np.int64)
    i_m, ifm = ren(lc.i_m[m_im]), ren(lc.ifm[m_ifm])
    fix = np.zeros(order.size, np.uint8)
    fix[i_m - 1] = 1
    fix[ifm - 1] = 1
    local = replace(
        lc, X=np.ascontiguousarray(lc.X[order]), Y=np.ascontiguousarray(lc.Y[order]),
        inpoel=np.ascontiguousarray((g2l[inp0[loc_el]] + 1).astype(I32)),
        ifixrho_node=ren(lc.ifixrho_node[m_rho]), rfixrho_value=np.ascontiguousarray(lc.rfixrho_value[m_rho]),
        ifixv_node=ren(lc.ifixv_node[m_v]), rfixv_valuex=np.ascontiguousarray(lc.rfixv_valuex[m_v]),
        rfixv_valuey=np.ascontiguousarray(lc.rfixv_valuey[m_v]),
        wall=np.ascontiguousarray(np.stack([ren(lc.wall[m_w, 0]), ren(lc.wall[m_w, 1])], 1).astype(I32)) if m_w.any() else np.zeros((0, 2), I32),
        ifixt_node=ren(lc.ifixt_node[m_t]), rfixt_value=np.ascontiguousarray(lc.rfixt_value[m_t]),
        sets=sets, ifm=ifm, i_m=i_m, ilaux=np.concatenate([i_m, ifm]).astype(I32), smooth_fix=fix,
    )
    return LocalPart(rank=rank, nranks=nranks, lc=local, node_gid=node_gid[order], elem_gid=elem_gid[loc_el],
                     n_owned=int(n_owned), elem_own=(elem_rank[loc_el] == rank), neighbors=neighbors, send=send, recv=recv,
                     moving=bool(lc.sets.size), red_aligned=B is not None, gid0=int(B[rank]) if B is not None else 0,
                     npoin_global=npoin_global, set_gidx=set_gidx)


def square_window(n: int, nranks: int, rank: int, rows_per: int | None = None, align: bool = True, **kw):
    """Weak-scaling bench mesh: rank's window of the global `nranks`-strip square mesh (each strip is the
    n x n lattice of meshgen.square, stacked in y), generated without building the global mesh.

    Returns (window LoadedCase, elem_rank, node_gid, elem_gid, bounds, npoin_global) ready for build_local().  The window
    holds every node row that carries a node this rank owns plus one row on either side, which is enough to know the owner
    of every node of the rank's local elements; the ownership boundaries are the chunk-aligned ones (module docstring),
    known analytically here: the natural first owned node of rank r >= 1 is the first node of row r*rows_per + 1.
    meshgen.square_global(n, nranks) builds the same mesh in one piece.
    """
    from . import deck, meshgen

    nx = n
    nqx = nx - 1
    rows_per = (n - 1) if rows_per is None else rows_per  # quad rows per rank
    ny_total = nranks * rows_per + 1
    npoin_global = nx * ny_total
    B = None
    if align and nranks > 1:
        first = np.array([0] + [(r * rows_per + 1) * nx for r in range(1, nranks)], np.int64)
        B = aligned_boundaries(first, npoin_global)
    jrange = None
    if B is not None:
        j0 = max(0, int(B[rank] // nx) - 1)
        j1 = min(ny_total - 1, int((B[rank + 1] - 1) // nx) + 1)
        # the natural window (own quad rows +-1) is kept inside: elem_rank of the window's elements stays meaningful
        j0 = min(j0, max(0, rank * rows_per - 1))
        j1 = max(j1, min(ny_total - 1, (rank + 1) * rows_per + 1))
        jrange = (j0, j1)
    raw, j0 = meshgen.square_rows(n, nranks, rank, rows_per=rows_per, jrange=jrange, **kw)
    lc = deck.load(raw)
    qrow = (np.arange(lc.nelem) // (2 * nqx)) + j0        # global quad row of each window element
    elem_rank = np.minimum(qrow // rows_per, nranks - 1)
    elem_gid = 2 * nqx * j0 + np.arange(lc.nelem, dtype=np.int64)
    node_gid = nx * j0 + np.arange(lc.npoin, dtype=np.int64)
    return lc, elem_rank, node_gid, elem_gid, B, npoin_global
