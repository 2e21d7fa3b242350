"""World-size-N CPU harness (gloo): every rank runs the ORACLE on its sub-domain and performs the
exchanges where the GPU path has them (global DTMIN min, ghost refresh after every RK stage); rank 0
checks the assembled result bit-for-bit against the oracle on the undivided mesh.  Launched by
tests/test_partition.py with RANK/WORLD_SIZE/MASTER_* in the environment."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cfd_b200 import deck, meshgen, partition  # noqa: E402
from oracle.orclib import Oracle  # noqa: E402


def main():
    case, steps = sys.argv[1], int(sys.argv[2])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if case == "square_visc":
        glc = deck.load(meshgen.square_global(9, world, FMU=1.8e-5, FK=0.0257))
    elif case in ("channel", "channel_last_stage_only"):
        glc = deck.load(meshgen.channel(nx=33, ny=11))
    else:
        raise SystemExit("unknown case")
    part = partition.build_local(glc, world, rank)
    o = Oracle(part.lc)
    st = meshgen.density_bump(glc)
    gid = part.node_gid
    o.set("U", st["U"][gid])
    for k in ("T", "VEL_X", "VEL_Y"):
        o.set(k, st[k][gid])

    def exchange():
        reqs, rbuf = [], {}
        # the ghost message of libcfdb200 (kernels.cuh: halo_pack_state): U1(4), T, VEL_X, VEL_Y, E, P, RMACH; RHO = U1(1)
        views = {"U1": o.view("U1").reshape(-1, 4), "T": o.view("T"), "VEL_X": o.view("VEL_X"), "VEL_Y": o.view("VEL_Y"),
                 "E": o.view("E"), "P": o.view("P"), "RMACH": o.view("RMACH"), "RHO": o.view("RHO")}
        for s in part.neighbors:
            if s in part.send:
                idx = part.send[s]
                pack = np.concatenate([views["U1"][idx]] + [views[k][idx, None] for k in ("T", "VEL_X", "VEL_Y", "E", "P", "RMACH")], 1)
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(pack)), dst=s))
            if s in part.recv:
                rbuf[s] = torch.empty((part.recv[s].size, 10), dtype=torch.float64)
                reqs.append(dist.irecv(rbuf[s], src=s))
        for r in reqs:
            r.wait()
        for s, b in rbuf.items():
            idx, b = part.recv[s], b.numpy()
            views["U1"][idx] = b[:, :4]
            views["T"][idx] = b[:, 4]
            views["VEL_X"][idx] = b[:, 5]
            views["VEL_Y"][idx] = b[:, 6]
            views["E"][idx], views["P"][idx], views["RMACH"][idx] = b[:, 7], b[:, 8], b[:, 9]
            views["RHO"][idx] = b[:, 0]

    for _ in range(steps):
        d = torch.tensor([o.step_part1()], dtype=torch.float64)
        dist.all_reduce(d, op=dist.ReduceOp.MIN)
        o.step_part2(float(d.item()))
        # Euler flow on a fixed mesh: every stage evaluates calcRHS at U and reads nothing a stage writes, so the
        # ghost refresh after stages 1-3 can be dropped (what cfdb.cu: run_rk does on the fused path)
        last_only = case == "channel_last_stage_only"
        for irk in (1, 2, 3, 4):
            o.rk_stage(irk)
            if irk == 4 or not last_only:
                exchange()
        o.step_part3()

    box = [None] * world
    own = slice(0, part.n_owned)
    # ghosts carry every nodal array of the owner after the exchange (body forces read P at both ends of an edge)
    gh = slice(part.n_owned, part.lc.npoin)
    dist.all_gather_object(box, (gid[own], o.get("U").reshape(-1, 4)[own], o.get("T")[own], o.scalar("DTMIN"), o.scalar("TIME"),
                                 gid[gh], o.get("P")[gh], o.get("RMACH")[gh]))
    if rank == 0:
        ref = Oracle(glc)
        ref.set("U", st["U"])
        for k in ("T", "VEL_X", "VEL_Y"):
            ref.set(k, st[k])
        ref.step(steps)
        U, T = np.zeros((glc.npoin, 4)), np.zeros(glc.npoin)
        seen = np.zeros(glc.npoin, int)
        for g, u, t, dtmin, time, gg, pg, mg_ in box:
            U[g], T[g] = u, t
            seen[g] += 1
            assert dtmin == ref.scalar("DTMIN") and time == ref.scalar("TIME")
            assert np.array_equal(pg.view(np.uint64), ref.get("P")[gg].view(np.uint64)), "ghost P differs from the owner's"
            assert np.array_equal(mg_.view(np.uint64), ref.get("RMACH")[gg].view(np.uint64)), "ghost RMACH differs"
        assert (seen == 1).all(), "every node must be owned exactly once"
        assert np.array_equal(U.view(np.uint64), ref.get("U").reshape(-1, 4).view(np.uint64)), "U differs"
        assert np.array_equal(T.view(np.uint64), ref.get("T").view(np.uint64)), "T differs"
        print("DIST_ORACLE_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
