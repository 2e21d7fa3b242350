"""Host-side multi-GPU logic on CPU: the partitioner and (world_size 2 and 3, gloo) the exchange protocol."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from cfd_b200 import deck, meshgen, partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def brute_owner(inpoel0, bounds, npoin):
    own = np.full(npoin, 10**9)
    for e, tri in enumerate(inpoel0):
        r = int(np.searchsorted(bounds, e, side="right") - 1)
        for n in tri:
            own[n] = min(own[n], r)
    return own


@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
def test_partition_invariants(nranks):
    glc = deck.load(meshgen.channel(nx=21, ny=9))
    inp0 = glc.inpoel.astype(np.int64) - 1
    bounds = partition.element_ranges(glc.nelem, nranks)
    assert bounds[0] == 0 and bounds[-1] == glc.nelem and (np.diff(bounds) >= glc.nelem // nranks).all()
    owner = brute_owner(inp0, bounds, glc.npoin)
    parts = [partition.build_local(glc, nranks, r) for r in range(nranks)]
    seen = np.zeros(glc.npoin, int)
    own_elems = 0
    for r, p in enumerate(parts):
        gid = p.node_gid
        assert (owner[gid[: p.n_owned]] == r).all() and (owner[gid[p.n_owned:]] != r).all()
        assert (np.diff(gid[: p.n_owned]) > 0).all() and (np.diff(gid[p.n_owned:]) > 0).all()
        assert (np.diff(p.elem_gid) > 0).all()
        seen[gid[: p.n_owned]] += 1
        own_elems += int(p.elem_own.sum())
        # every element touching an owned node is local, and local connectivity maps back to the global one
        need = np.flatnonzero((owner[inp0] == r).any(1))
        assert np.array_equal(need, p.elem_gid)
        assert np.array_equal(gid[p.lc.inpoel - 1], inp0[p.elem_gid])
        assert np.array_equal(p.lc.X.view(np.uint64), glc.X[gid].view(np.uint64))
        # BC lists: restricted, order kept, renumbered
        keep = np.isin(glc.ifixv_node - 1, gid)
        assert np.array_equal(gid[p.lc.ifixv_node - 1], glc.ifixv_node[keep] - 1)
        assert np.array_equal(p.lc.rfixv_valuex, glc.rfixv_valuex[keep])
        wk = np.isin(glc.wall[:, 0] - 1, gid) & np.isin(glc.wall[:, 1] - 1, gid)
        assert np.array_equal(gid[p.lc.wall - 1], glc.wall[wk] - 1)
        # all wall edges of an owned wall node are local
        for a, b in glc.wall - 1:
            if owner[a] == r or owner[b] == r:
                assert a in gid and b in gid
    # (an element all of whose nodes belong to lower ranks is computed only there, hence <=)
    assert (seen == 1).all() and own_elems <= glc.nelem
    for r, p in enumerate(parts):
        for s in p.neighbors:
            a = p.node_gid[p.send[s]] if s in p.send else np.zeros(0, int)
            b = parts[s].node_gid[parts[s].recv[r]] if r in parts[s].recv else np.zeros(0, int)
            assert np.array_equal(a, b), (r, s)
        ranks, sp, si, rp, ri = p.halo_arrays()
        assert sp[-1] == si.size and rp[-1] == ri.size and (si < p.n_owned).all() and (ri >= p.n_owned).all()
        assert np.array_equal(np.sort(ri), np.arange(p.n_owned, p.lc.npoin))


@pytest.mark.parametrize("nranks", [2, 3])
def test_strip_window_equals_global(nranks):
    n = 9
    glc = deck.load(meshgen.square_global(n, nranks))
    for r in range(nranks):
        a = partition.build_local(glc, nranks, r)
        w = partition.square_window(n, nranks, r)
        b = partition.build_local(w[0], nranks, r, *w[1:])
        assert np.array_equal(a.node_gid, b.node_gid) and np.array_equal(a.elem_gid, b.elem_gid)
        assert np.array_equal(a.lc.inpoel, b.lc.inpoel) and a.n_owned == b.n_owned and a.neighbors == b.neighbors
        assert np.array_equal(a.lc.X.view(np.uint64), b.lc.X.view(np.uint64))
        for s in a.neighbors:
            assert np.array_equal(a.send.get(s, []), b.send.get(s, [])) and np.array_equal(a.recv.get(s, []), b.recv.get(s, []))
        for f in ("ifixrho_node", "ifixv_node", "wall", "ifm", "ilaux", "rfixv_valuex"):
            assert np.array_equal(getattr(a.lc, f), getattr(b.lc, f)), f


@pytest.mark.parametrize("nranks", [2, 3])
def test_chunk_aligned_ownership(nranks):
    """Round 2: when every rank owns at least two reduction chunks, the ownership boundaries sit on multiples of 4096 in the
    global node numbering (partition module docstring), the window generator and the global partitioner agree, every other
    invariant of the partition still holds, and the rank-wise protocol of the canonical reduction -- chunk sums placed at
    their global positions, summed with zeros, upper tree levels on every rank -- reproduces the undivided canonical sum
    bit for bit."""
    from oracle.orclib import lib as orclib

    n = 129
    glc = deck.load(meshgen.square_global(n, nranks))
    inp0 = glc.inpoel.astype(np.int64) - 1
    parts = [partition.build_local(glc, nranks, r) for r in range(nranks)]
    B = [p.gid0 for p in parts] + [glc.npoin]
    assert all(p.red_aligned for p in parts) and B[0] == 0 and all(b % partition.RED_CHUNK == 0 for b in B[:-1])
    nat = partition.node_owner(inp0, np.searchsorted(partition.element_ranges(glc.nelem, nranks), np.arange(glc.nelem), side="right") - 1, glc.npoin)
    rng = np.random.default_rng(7)
    x, y = rng.standard_normal(glc.npoin), rng.standard_normal(glc.npoin)
    L = orclib()
    nchunk = (glc.npoin + 4095) // 4096
    G = np.zeros(nchunk)
    for r, p in enumerate(parts):
        gid = p.node_gid
        assert np.array_equal(gid[: p.n_owned], np.arange(B[r], B[r + 1]))          # contiguous, aligned
        first_nat = np.flatnonzero(nat == r)[0]
        assert abs(B[r] - first_nat) <= partition.RED_CHUNK // 2 or r == 0            # nearest multiple of the chunk
        owner = np.searchsorted(np.array(B), np.arange(glc.npoin), side="right") - 1
        need = np.flatnonzero((owner[inp0] == r).any(1))
        assert np.array_equal(need, p.elem_gid)                                       # every element touching an owned node
        assert np.array_equal(np.sort(np.concatenate([v for v in p.recv.values()])), np.arange(p.n_owned, p.lc.npoin))
        w = partition.square_window(n, nranks, r)
        b = partition.build_local(w[0], nranks, r, *w[1:])
        assert b.red_aligned and b.gid0 == p.gid0 and b.npoin_global == glc.npoin
        assert np.array_equal(p.node_gid, b.node_gid) and np.array_equal(p.elem_gid, b.elem_gid) and p.n_owned == b.n_owned
        assert np.array_equal(p.lc.inpoel, b.lc.inpoel) and np.array_equal(p.lc.X.view(np.uint64), b.lc.X.view(np.uint64))
        for s_ in p.neighbors:
            assert np.array_equal(p.send.get(s_, []), b.send.get(s_, [])) and np.array_equal(p.recv.get(s_, []), b.recv.get(s_, []))
        # this rank's first-level chunk sums over its owned prefix, at their global chunk positions
        xl, yl = x[gid[: p.n_owned]], y[gid[: p.n_owned]]
        for cidx in range((p.n_owned + 4095) // 4096):
            lo, hi = 4096 * cidx, min(4096 * (cidx + 1), p.n_owned)
            G[B[r] // 4096 + cidx] += L.orc_vecdot(hi - lo, np.ascontiguousarray(xl[lo:hi]), np.ascontiguousarray(yl[lo:hi]))
    want = L.orc_vecdot(glc.npoin, x, y)
    got = L.orc_canon_sum(nchunk, G)   # nchunk <= 4096: one more level of the same tree
    assert np.float64(got).view(np.uint64) == np.float64(want).view(np.uint64)


def test_small_meshes_keep_the_natural_ownership():
    glc = deck.load(meshgen.square_global(9, 2))
    assert not any(partition.build_local(glc, 2, r).red_aligned for r in range(2))
    assert partition.square_window(9, 2, 0)[4] is None


def test_every_rank_of_a_moving_mesh_case_is_flagged_moving():
    """The body sets of the ALE case end up on one rank only, but fluidStructure moves the mesh everywhere: LocalPart.moving
    (-> cfdb_set_option "ale": FUENTE and the mesh-velocity terms of ESTAB/deltat) must be set on every rank; fixed-mesh
    cases are not flagged."""
    from cfd_b200 import deck, meshgen, partition

    glc = deck.load(meshgen.ale_body(nt=48, nr=14))
    parts = [partition.build_local(glc, 2, r) for r in range(2)]
    assert any(p.lc.sets.size == 0 for p in parts) and any(p.lc.sets.size > 0 for p in parts)
    assert all(p.moving for p in parts)
    flc = deck.load(meshgen.channel(nx=21, ny=9))
    assert not any(partition.build_local(flc, 2, r).moving for r in range(2))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case,world", [("square_visc", 2), ("channel", 3), ("channel_last_stage_only", 2)])
def test_gloo_exchange_protocol_matches_undivided_oracle(case, world):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_oracle_worker.py"), case, "6"],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "DIST_ORACLE_OK" in outs[0]
