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
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np

from .deck import I32, LoadedCase


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
                node_gid: np.ndarray | None = None, elem_gid: np.ndarray | None = None) -> LocalPart:
    """Local part of `rank`.  `lc` is the global mesh, or a window of it that contains every element
    touching a node of the rank's local elements (then pass the window's elem_rank / node_gid / elem_gid)."""
    E, P = lc.nelem, lc.npoin
    inp0 = lc.inpoel.astype(np.int64) - 1
    if elem_rank is None:
        b = element_ranges(E, nranks)
        elem_rank = np.searchsorted(b, np.arange(E), side="right") - 1
    elem_rank = np.asarray(elem_rank, np.int64)
    node_gid = np.arange(P, dtype=np.int64) if node_gid is None else np.asarray(node_gid, np.int64)
    elem_gid = np.arange(E, dtype=np.int64) if elem_gid is None else np.asarray(elem_gid, np.int64)
    owner = node_owner(inp0, elem_rank, P)
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
    else:
        sets = np.zeros((0, 4), I32)
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
                     moving=bool(lc.sets.size))


def square_window(n: int, nranks: int, rank: int, rows_per: int | None = None, **kw):
    """Weak-scaling bench mesh: rank's window of the global `nranks`-strip square mesh (each strip is the
    n x n lattice of meshgen.square, stacked in y), generated without building the global mesh.

    Returns (window LoadedCase, elem_rank, node_gid, elem_gid) ready for build_local().  The window holds the
    rank's own quad rows plus one quad row on either side, which is enough to know the owner of every node
    of the rank's local elements.  meshgen.square_global(n, nranks) builds the same mesh in one piece.
    """
    from . import deck, meshgen

    raw, j0 = meshgen.square_rows(n, nranks, rank, rows_per=rows_per, **kw)
    lc = deck.load(raw)
    nx = n
    nqx = nx - 1
    rows_per = (n - 1) if rows_per is None else rows_per  # quad rows per rank
    qrow = (np.arange(lc.nelem) // (2 * nqx)) + j0        # global quad row of each window element
    elem_rank = np.minimum(qrow // rows_per, nranks - 1)
    elem_gid = 2 * nqx * j0 + np.arange(lc.nelem, dtype=np.int64)
    node_gid = nx * j0 + np.arange(lc.npoin, dtype=np.int64)
    return lc, elem_rank, node_gid, elem_gid
