"""Multi-GPU worker (one rank per GPU, NCCL inside libcfdb200.so).  Launched by tests/test_multigpu.py through
torch.distributed.run; rank 0 compares the assembled owned-node state with the oracle on the undivided mesh."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cfd_b200 import deck, meshgen  # noqa: E402
from cfd_b200.dist import gather_owned, make_rank_solver  # noqa: E402
from oracle.orclib import Oracle  # noqa: E402


def main():
    case, steps = sys.argv[1], int(sys.argv[2])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    if case == "square_visc":
        glc = deck.load(meshgen.square_global(17, world, FMU=1.8e-5, FK=0.0257))
    elif case == "channel_itlocal":
        glc = deck.load(meshgen.channel(nx=41, ny=13, ITLOCAL=50))
    elif case == "ale":
        glc = deck.load(meshgen.ale_body(nt=48, nr=14))            # too small for chunk-aligned ownership: round-off agreement
    elif case == "ale_aligned":
        glc = deck.load(meshgen.ale_body(nt=192, nr=96))           # 18 432 nodes: every rank owns >= 2 reduction chunks
    elif case == "ale_visc_aligned":
        glc = deck.load(meshgen.ale_body(nt=192, nr=96, FMU=1.8e-5, FK=0.0257, IPRINT=3))
    elif case == "square_aligned":
        glc = deck.load(meshgen.square_global(129, world, IPRINT=2))
    else:
        raise SystemExit("unknown case")
    g, part = make_rank_solver(glc, rank, world, local, dist)
    st = None
    if not case.startswith("ale"):
        st = meshgen.density_bump(glc)
        g.set("U", st["U"][part.node_gid])
        for k in ("T", "VEL_X", "VEL_Y"):
            g.set(k, st[k][part.node_gid])
    g.step(steps)
    g.sync()
    U = gather_owned(g, part, "U", 4, dist, glc.npoin)
    T = gather_owned(g, part, "T", 1, dist, glc.npoin)[:, 0]
    X = gather_owned(g, part, "X", 1, dist, glc.npoin)[:, 0]
    P_ = gather_owned(g, part, "P", 1, dist, glc.npoin)[:, 0]
    W_ = gather_owned(g, part, "W_X", 1, dist, glc.npoin)[:, 0]
    er, err = g.norms()
    ser, serr = g.step_norms()
    fx = np.concatenate([g.get("FX"), g.get("FY"), g.get("RM")])
    fv = np.concatenate(g.force_visc()[:2]) if case == "ale_visc_aligned" else None
    dtmin, time = g.scalar("DTMIN"), g.scalar("TIME")
    if case.endswith("aligned"):
        assert part.red_aligned, "this case is sized for chunk-aligned ownership"
    if rank == 0:
        ref = Oracle(glc)
        if st:
            ref.set("U", st["U"])
            for k in ("T", "VEL_X", "VEL_Y"):
                ref.set(k, st[k])
        ref.step(steps)
        Ur, Tr, Xr = ref.get("U").reshape(-1, 4), ref.get("T"), ref.get("X")
        if case.endswith("aligned"):
            # chunk-aligned ownership: the inner products of biCG and the residual norms are the global canonical sums, so the
            # moving-mesh run is bit-identical to the undivided one, mesh solve included
            print(case, "diag: bicg", g.scalar("bicg_x"), ref.scalar("bicg_x"), g.scalar("bicg_y"), ref.scalar("bicg_y"), flush=True)
            assert g.scalar("bicg_x") == ref.scalar("bicg_x") and g.scalar("bicg_y") == ref.scalar("bicg_y")
            assert dtmin == ref.scalar("DTMIN") and time == ref.scalar("TIME")
            for nm, a, b in (("U", U, Ur), ("T", T, Tr), ("X", X, Xr), ("P", P_, ref.get("P")), ("W_X", W_, ref.get("W_X"))):
                assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), f"{nm} differs, max {np.max(np.abs(a - b))}"
            ero, erro = ref.norms()
            assert np.array_equal(err.view(np.uint64), erro.view(np.uint64)) and np.array_equal(er.view(np.uint64), ero.view(np.uint64))
            if case == "square_aligned":   # norms of the last print step (before U = U1): the 8 canonical sums of run_norms
                sero, serro = ref.step_norms()
                assert np.array_equal(ser.view(np.uint64), sero.view(np.uint64)) and np.array_equal(serr.view(np.uint64), serro.view(np.uint64))
            # body forces are per-rank sequential sums added across ranks (edge order within a rank kept): round-off
            fxo = np.concatenate([ref.get("FX"), ref.get("FY"), ref.get("RM")])
            assert np.allclose(fx, fxo, rtol=1e-12, atol=1e-12 * np.abs(fxo).max()), (fx, fxo)
            if fv is not None:
                fvo = np.concatenate(ref.force_visc()[:2])
                assert np.allclose(fv, fvo, rtol=1e-12, atol=1e-12 * np.abs(fvo).max()), (fv, fvo)
        elif case == "ale":
            # inner products are reduced per rank then summed: round-off level differences in the mesh solve
            print("ale diag: bicg", g.scalar("bicg_x"), ref.scalar("bicg_x"), g.scalar("bicg_y"), ref.scalar("bicg_y"),
                  "dX", np.max(np.abs(X - Xr)), "dU", np.max(np.abs(U - Ur) / np.abs(Ur).max(0)), "dt", dtmin, ref.scalar("DTMIN"), flush=True)
            assert abs(g.scalar("bicg_y") - ref.scalar("bicg_y")) <= 1
            assert np.max(np.abs(X - Xr)) < 1e-13, np.max(np.abs(X - Xr))
            assert np.max(np.abs(U - Ur) / np.abs(Ur).max(0)) < 1e-10
        else:
            assert dtmin == ref.scalar("DTMIN") and time == ref.scalar("TIME")
            assert np.array_equal(U.view(np.uint64), Ur.view(np.uint64)), f"U differs, max {np.max(np.abs(U - Ur))}"
            assert np.array_equal(T.view(np.uint64), Tr.view(np.uint64)), "T differs"
            ero, erro = ref.norms()  # U==U1 after the step on both sides: ER=0, ERR = sum U^2
            assert np.allclose(err, erro, rtol=1e-13) and np.all(er == 0)
        print("MULTIGPU_OK", case, world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
