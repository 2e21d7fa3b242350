#!/usr/bin/env python
"""bench.py — triangle-element updates/s of the per-timestep hot path on B200 (BASELINE.json metric).

One "step" = one pass of the reference's time loop (ns2DComp.ALE.f90:138-282: deltat, dt logic,
estab, 4 x [calcRHS + FUENTE + nodal update + BCs], fluidStructure) over the whole mesh.
Workload (N=1 and per GPU for N>1, weak scaling): BASELINE.json configs[4] — a 16.0 M-triangle
jittered-lattice Delaunay-diagonal unit-square mesh per GPU, the mesh north_star's target is stated
on (configs[1], the 1 M wedge, fits in L2 and is a parity-test case, see tests/).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--lattice N] [--strong] [--impl reference]

Prints ONE JSON line (rank 0).  Keys: value (device-resident throughput, CUDA events on the
library's stream, max over ranks), e2e (same step through the C ABI with host buffers: pinned
H2D of the state every step + D2H of the result), roofline (dominant kernel, algorithmic bytes /
measured per-launch time vs MEASURED_PEAKS.json), cpu_baseline (the oracle's OpenMP build on the
host cores, bounded sample), clocks, gpu_launches.
`--impl reference` times the CPU restatement of the reference (oracle/, OpenMP build) — the
reference itself is Fortran and cannot be built in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "triangle-element updates/s"
UNIT = "elements/s"
# SURVEY.md §8(d) / BASELINE.md §3: algorithmic bytes per element (r = npoin/nelem = 0.5)
BYTES_STAGE_CALCRHS = 100 + 20 + 16      # element stream + gathers(U,T) + RHS write   (per element-stage)
BYTES_STAGE_UPDATE = 16 + 68             # RHS read + nodal update                     (per element-stage)
BYTES_STEP = 1136                        # whole fixed-mesh step


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_throughput(n_lattice, steps, warmup=1, omp=True):
    """Element updates/s of the oracle (C++ restatement of the reference) on the host: the OpenMP build on all cores, or the
    sequential build (the order the parity tests use) on one."""
    from cfd_b200 import deck, meshgen
    from oracle.orclib import Oracle

    lc = deck.load(meshgen.square(n=n_lattice, IPRINT=10**9, MAXITER=10**9))
    o = Oracle(lc, omp=omp)
    o.set_scalar("norms_every_step", 0)
    st = meshgen.density_bump(lc)
    for k, v in st.items():
        o.set(k, v)
    o.step(warmup)
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    return lc.nelem * steps / dt, (o.L.orc_omp_threads() if omp else 1), lc.nelem, dt


def run_reference(args, rank, out):
    """Reference arm: the CPU implementation of the path (oracle OpenMP build = C++ restatement of chanshing/cfd with
    the reference's own pragmas/atomics; the Fortran original cannot be built here) on all host cores, bounded sample."""
    if rank != 0:
        return
    from cfd_b200 import deck, meshgen
    from oracle.orclib import Oracle

    n = args.ref_n
    lc = deck.load(meshgen.square(n=n, IPRINT=10**9, MAXITER=10**9))
    o = Oracle(lc, omp=True)
    o.set_scalar("norms_every_step", 0)
    for k, v in meshgen.density_bump(lc).items():
        o.set(k, v)
    o.step(args.warmup)
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    val = lc.nelem * args.steps / dt
    cores = o.L.orc_omp_threads()
    sample = f"{lc.nelem}-triangle square mesh (lattice {n}), {args.steps} steps in {dt:.1f} s, oracle OpenMP build, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"bounded CPU sample of the GPU arm's workload (same generator, same flow): {sample}"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C++ restatement of chanshing/cfd (oracle/), not the Fortran binary: no Fortran compiler in the image",
    }), file=out, flush=True)


def main():
    # stdout carries exactly one JSON line: libraries that chat on fd 1 (NCCL prints its version there) go to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--lattice", "--n", dest="n", type=int, default=2829, help="lattice nodes per side per GPU (2829 -> 16.0 M triangles)")
    ap.add_argument("--ref-n", type=int, default=1415, help="lattice of the bounded CPU sample (1415 -> 4.0 M triangles)")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--strong", action="store_true", help="strong scaling (BASELINE configs[3]): one n x n domain cut into N strips")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, real_stdout)
        return

    import torch
    import torch.distributed as dist

    from cfd_b200 import deck, meshgen
    from cfd_b200.solver import NSComp2D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — cfd_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- workload: this rank's sub-domain (weak scaling: fixed work per GPU) --------------------
    # The global mesh is `world` 16 M-triangle strips stacked in y; each rank builds only its window of it
    # (cfd_b200/partition.py), computes its own range plus a one-element ghost layer, and refreshes ghost nodes
    # over NCCL after every RK stage.
    t_gen = time.perf_counter()
    from cfd_b200 import partition
    from cfd_b200.dist import make_rank_solver

    rows_per = None
    if args.strong:
        if (args.n - 1) % world:
            raise SystemExit("--strong needs (n-1) divisible by the number of GPUs")
        rows_per = (args.n - 1) // world
    win = partition.square_window(args.n, world, rank, rows_per=rows_per, IPRINT=10**9, MAXITER=10**9)
    g, part = make_rank_solver(win, rank, world, local, dist if world > 1 else None)
    lc = part.lc
    E = 2 * (args.n - 1) * (rows_per if args.strong else args.n - 1)   # elements of this rank's own range
    P = part.n_owned
    if args.strong:
        bump = meshgen.density_bump(lc, x0=0.5, y0=0.5, sigma=0.15)    # one bump in the middle of the fixed domain
    else:
        bump = meshgen.density_bump(lc, x0=0.5, y0=0.5, sigma=0.15, period_y=1.0)  # one bump per strip
    for k, v in bump.items():
        g.set(k, v)
    t_gen = time.perf_counter() - t_gen
    stream = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", local))

    def barrier():
        g.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing -------------------------------------------------------------------
    g.step(args.warmup)
    barrier()
    l0 = g.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    g.step(args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = g.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * E * args.steps / (ms * 1e-3)

    # ---- per-kernel launch durations (events on the launching stream; separate pass) ---------------
    g.profile(True)
    g.step(3)
    g.sync()
    prof = {}
    for kname in ("calcrhs_elem", "node_update", "estab", "deltat", "spmv", "dot", "vec"):
        t_ms, n = g.profile_get(kname)
        if n:
            prof[kname] = {"avg_ms": t_ms / n, "launches_per_step": n / 3.0}
    g.profile(False)
    peak, peak_src = peaks()
    stage_ms = prof["calcrhs_elem"]["avg_ms"] + prof["node_update"]["avg_ms"]
    dom = "calcrhs_elem" if prof["calcrhs_elem"]["avg_ms"] >= prof["node_update"]["avg_ms"] else "node_update"
    dom_bytes = BYTES_STAGE_CALCRHS if dom == "calcrhs_elem" else BYTES_STAGE_UPDATE
    achieved = dom_bytes * E / (prof[dom]["avg_ms"] * 1e-3) / 1e9
    stage_achieved = (BYTES_STAGE_CALCRHS + BYTES_STAGE_UPDATE) * E / (stage_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")   # dram__bytes_read+write per launch from the committed ncu capture
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(dom)
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_element": dom_bytes,
        "stage": {"kernels": "calcrhs_elem+node_update", "bytes_per_element_stage": BYTES_STAGE_CALCRHS + BYTES_STAGE_UPDATE,
                  "ms": stage_ms, "achieved": stage_achieved, "frac": stage_achieved / peak,
                  "element_stage_per_s": E / (stage_ms * 1e-3)},
        "step": {"bytes_per_element_step": BYTES_STEP, "achieved": BYTES_STEP * E * args.steps / (ms * 1e-3) / 1e9,
                 "frac": BYTES_STEP * E * args.steps / (ms * 1e-3) / 1e9 / peak},
        "kernels": prof,
    }

    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        names = ["U", "T", "VEL_X", "VEL_Y"]
        host = {nme: torch.empty(g.L.cfdb_field_size(g.h, nme.encode()), dtype=torch.float64, pin_memory=True) for nme in names}
        for nme in names:
            host[nme].numpy()[:] = g.get(nme)
        h2d = sum(8 * h.numel() for h in host.values())
        d2h = h2d + 64
        ksteps = max(3, min(args.steps, 10))

        hv = {nme: host[nme].numpy() for nme in names}

        def one():
            for nme in names:
                g.set_from(nme, hv[nme])   # pinned host -> HBM
            g.step(1)
            # HBM -> pinned host, straight into the host program's own arrays: U = U1 (ns2DComp.ALE.f90:277-281) and the
            # primitives the next DELTAT/ESTAB read, so the host copy of the state stays consistent step after step
            for nme in names:
                g.get_into(nme, hv[nme])
            g.norms()

        one()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            one()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * E * ksteps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": ksteps, "what": "state U,T,VEL_X,VEL_Y uploaded from pinned host memory, cfdb_step(1), "
               "the same four arrays and the residual norms downloaded into the host program's arrays, every step"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, ne, dt = cpu_oracle_throughput(args.ref_n, args.cpu_steps)
        v1, _, ne1, dt1 = cpu_oracle_throughput(709, 3, omp=False)     # threads = 1 next to all cores
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{ne}-triangle square mesh, {args.cpu_steps} steps ({dt:.1f} s), oracle OpenMP build "
                         "(C++ restatement; the Fortran reference cannot be built here)",
               "single_thread": {"value": v1, "cores": 1, "sample": f"{ne1}-triangle square mesh, 3 steps ({dt1:.1f} s), "
                                 "sequential build (the summation order of the parity tests)"}}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"square16M: {E}-triangle / {P}-node jittered-lattice Delaunay-diagonal strip per GPU "
                                   "(BASELINE configs[4]; the mesh north_star's target is stated on), Euler, fixed mesh, "
                                   "density-bump initial state", "l2": "inputs larger than L2 (no flush needed)",
                       "elements_per_gpu": E, "local_elements_incl_ghost_layer": lc.nelem,
                       "parallelism": f"{world} contiguous strips, owner-computes + NCCL ghost refresh per RK stage",
                       "element_stage_updates_per_s": 4 * value, "setup_s": t_gen},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": launches,
        }), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
