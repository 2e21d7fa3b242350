#!/usr/bin/env python
"""bench.py — triangle-element updates/s of the per-timestep hot path on B200 (BASELINE.json metric).

One "step" = one pass of the reference's time loop (ns2DComp.ALE.f90:138-282: deltat, dt logic,
estab, 4 x [calcRHS + FUENTE + nodal update + BCs], fluidStructure) over the whole mesh.
Workload (N=1 and per GPU for N>1, weak scaling): BASELINE.json configs[4] — a 16.0 M-triangle
jittered-lattice Delaunay-diagonal unit-square mesh per GPU, the mesh north_star's target is stated
on (configs[1], the 1 M wedge, fits in L2 and is a parity-test case, see tests/).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--lattice N] [--strong] [--impl reference]

Prints ONE JSON line (rank 0).  Keys:
  value         device-resident throughput (CUDA events on the library's stream, max over ranks)
  roofline      the RK stage (fused tile kernel + tile-boundary pass; north_star's unit): algorithmic bytes / measured time
                vs MEASURED_PEAKS.json, per-kernel launch durations measured live with event pairs
  e2e           the same step through the C ABI with HOST buffers (cfdb_step_streamed: pinned H2D of the state and D2H of the
                result every step, pipelined on copy streams); e2e.callsite = the literal per-subroutine drop-in
  config.secondary   short timings of the other BASELINE configs: viscous 16 M, ALE 4 M (+ biCG-iteration roofline) at N=1,
                strong-scaling 64 M at N>1
  multi_gpu_parity   N>1: a small moving-mesh case and a small viscous case run on the N ranks and, on rank 0, on one GPU
                with the same library; owned state compared byte for byte before anything is timed
  cpu_baseline  the oracle's OpenMP build on all host cores (bounded sample), clocks, gpu_launches.
`--impl reference` times the CPU restatement of the reference (oracle/, OpenMP build, all host cores) — the
reference itself is Fortran and cannot be built in this image.
"""
import argparse
import glob
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
BYTES_STAGE = BYTES_STAGE_CALCRHS + BYTES_STAGE_UPDATE
BYTES_STEP = 1136                        # whole fixed-mesh step
BYTES_BICG_ITER_PER_NODE = 192           # 12*nnz + 108*P, nnz ~ 7P


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    """threads the CPU arms use: every core this process may run on (not what a launcher exported as OMP_NUM_THREADS:
    torch.distributed.run sets that to 1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """SM clock, power and clock-event (throttle) reasons of one GPU sampled every 5 ms during the timed region, through NVML
    (nvidia_ml_py; an `nvidia-smi -lms` child takes longer to start than the timed region lasts on an 8-GPU box)."""

    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"),
               ("hw_power_brake", "nvmlClocksEventReasonHwPowerBrakeSlowdown"))

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.h = None
        self.stop_flag = threading.Event()
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            try:      # the device this rank computes on, whatever CUDA_VISIBLE_DEVICES says
                import torch
                pr = torch.cuda.get_device_properties(index)
                bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:   # no NVML: say so in the line instead of inventing numbers
            self.err = f"NVML unavailable: {e}"
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.h is None:
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [getattr(self, "err", "NVML unavailable")], "samples": 0}
        self.stop_flag.set()
        self.t.join(timeout=2)
        sm = [r[0] for r in self.rows]
        mask = 0
        for r in self.rows:
            mask |= r[2]
        reasons = sorted(n for n, const in self.REASONS if mask & int(getattr(self.nv, const, 0)))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": self.max_sm,
                "power_w_max": max((r[1] for r in self.rows), default=None), "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_numa_node(local):
    """first-touch the pinned host arrays on the NUMA node the GPU hangs off (e2e copies then stay on one socket)"""
    try:
        bus = subprocess.run(["nvidia-smi", f"--id={local}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus              # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def cpu_oracle_throughput(n_lattice, steps, warmup=1, omp=True):
    """Element updates/s of the oracle (C++ restatement of the reference) on the host: the OpenMP build on all cores, or the
    sequential build (the order the parity tests use) on one."""
    from cfd_b200 import deck, meshgen
    from oracle.orclib import Oracle

    lc = deck.load(meshgen.square(n=n_lattice, IPRINT=10**9, MAXITER=10**9))
    o = Oracle(lc, omp=omp)
    threads = o.L.orc_set_omp_threads(host_threads()) if omp else 1
    o.set_scalar("norms_every_step", 0)
    st = meshgen.density_bump(lc)
    for k, v in st.items():
        o.set(k, v)
    o.step(warmup)
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    return lc.nelem * steps / dt, threads, lc.nelem, dt


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
    cores = o.L.orc_set_omp_threads(host_threads())     # explicit: never the launcher's OMP_NUM_THREADS
    o.set_scalar("norms_every_step", 0)
    for k, v in meshgen.density_bump(lc).items():
        o.set(k, v)
    o.step(args.warmup)
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    val = lc.nelem * args.steps / dt
    sample = (f"{lc.nelem}-triangle square mesh (lattice {n}; the GPU arm's generator and flow on a smaller mesh -- a per-element "
              f"rate), {args.steps} steps in {dt:.1f} s, oracle OpenMP build, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"bounded CPU sample of the GPU arm's workload: {sample}"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C++ restatement of chanshing/cfd (oracle/), not the Fortran binary: no Fortran compiler in the image",
    }), file=out, flush=True)


PROF_KERNELS = ("stage_fused", "calcrhs_elem", "node_update", "estab", "deltat", "spmv", "dot", "vec", "scalar", "fixrows",
                "move_apply", "halo", "laplace", "deriv", "masas", "gcl", "fill", "layout")


def kernel_profile(g, steps=3):
    g.profile(True)
    g.step(steps)
    g.sync()
    prof = {}
    for kname in PROF_KERNELS:
        t_ms, n = g.profile_get(kname)
        if n:
            prof[kname] = {"avg_ms": t_ms / n, "launches_per_step": n / float(steps)}
    g.profile(False)
    return prof


def timed_steps(g, torch, stream, steps, barrier):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    g.step(steps)
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1)


def parity_selfcheck(torch, dist, rank, world, local):
    """Correctness carried by the scaling record: a small moving-mesh case (biCG, global canonical reductions) and a small
    viscous fixed-mesh case on the N ranks, against the same library on ONE GPU (rank 0), owned state compared byte for byte."""
    from cfd_b200 import deck, meshgen
    from cfd_b200.dist import make_rank_solver
    from cfd_b200.solver import NSComp2D

    cases = {"ale": (lambda: meshgen.ale_body(nt=512, nr=160), False, 4),
             "viscous": (lambda: meshgen.square_global(129, world, FMU=1.8e-5, FK=0.0257), True, 4)}
    verdict = []
    for name, (mk, bump, steps) in cases.items():
        glc = deck.load(mk())
        g, part = make_rank_solver(glc, rank, world, local, dist)
        st = meshgen.density_bump(glc) if bump else None
        if st:
            g.set("U", st["U"][part.node_gid])
            for k in ("T", "VEL_X", "VEL_Y"):
                g.set(k, st[k][part.node_gid])
        g.step(steps)
        g.sync()
        own = slice(0, part.n_owned)
        mine = {f: g.get(f).reshape(part.lc.npoin, -1)[own] for f in ("U", "T", "X", "P")}
        box = [None] * world
        dist.all_gather_object(box, (part.node_gid[own], mine, part.red_aligned))
        g.close()
        if rank == 0:
            ref = NSComp2D(glc, device=local)
            if st:
                for k, v in st.items():
                    ref.set(k, v)
            ref.step(steps)
            ok = all(b[2] for b in box) or name != "ale"
            for f in ("U", "T", "X", "P"):
                want = ref.get(f).reshape(glc.npoin, -1)
                for gid, vals, _ in box:
                    ok = ok and np.array_equal(vals[f].view(np.uint64), want[gid].view(np.uint64))
            ref.close()
            verdict.append((name, ok))
    if rank == 0:
        bad = [n for n, ok in verdict if not ok]
        return "bit-exact" if not bad else "FAILED: " + ",".join(bad)
    return None


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
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
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
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    parity = None
    if world > 1 and not args.no_parity:
        parity = parity_selfcheck(torch, dist, rank, world, local)

    # ---- workload: this rank's sub-domain (weak scaling: fixed work per GPU) --------------------
    # The global mesh is `world` 16 M-triangle strips stacked in y; each rank builds only its window of it
    # (cfd_b200/partition.py), computes its own range plus a one-element ghost layer, and refreshes ghost nodes
    # over NCCL after every RK stage.
    from cfd_b200 import partition
    from cfd_b200.dist import make_rank_solver

    def build(n, strong, **kw):
        t0 = time.perf_counter()
        rows_per = None
        if strong:
            if (n - 1) % world:
                raise SystemExit("--strong needs (n-1) divisible by the number of GPUs")
            rows_per = (n - 1) // world
        win = partition.square_window(n, world, rank, rows_per=rows_per, IPRINT=10**9, MAXITER=10**9, **kw)
        g, part = make_rank_solver(win, rank, world, local, dist if world > 1 else None)
        E = 2 * (n - 1) * (rows_per if strong else n - 1)   # elements of this rank's own range
        if strong:
            bump = meshgen.density_bump(part.lc, x0=0.5, y0=0.5, sigma=0.15)    # one bump in the middle of the fixed domain
        else:
            bump = meshgen.density_bump(part.lc, x0=0.5, y0=0.5, sigma=0.15, period_y=1.0)  # one bump per strip
        for k, v in bump.items():
            g.set(k, v)
        return g, part, E, time.perf_counter() - t0

    g, part, E, t_gen = build(args.n, args.strong)
    lc = part.lc
    P = part.n_owned
    stream = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", local))

    def barrier():
        g.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- device-resident timing -------------------------------------------------------------------
    g.step(args.warmup)
    barrier()
    l0 = g.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed_steps(g, torch, stream, args.steps, barrier)
    clocks = sampler.stop()
    launches = g.launch_count() - l0
    ms = max_over_ranks(ms)
    value = world * E * args.steps / (ms * 1e-3)

    # ---- per-kernel launch durations (event pairs on the launching stream; separate pass, stream launches) ---------------
    prof = kernel_profile(g)
    peak, peak_src = peaks()
    fused = "stage_fused" in prof
    if fused:   # the RK stage = fused tile kernel + boundary_update over the tile-boundary nodes (profile slot "node_update")
        stage_ms = prof["stage_fused"]["avg_ms"] + prof["node_update"]["avg_ms"]
        stage_kernels = "stage_fused+boundary_update(tile-boundary nodes)"
        dom = "stage_fused"
    else:
        stage_ms = prof["calcrhs_elem"]["avg_ms"] + prof["node_update"]["avg_ms"]
        stage_kernels = "calcrhs_elem+node_update"
        dom = "calcrhs_elem" if prof["calcrhs_elem"]["avg_ms"] >= prof["node_update"]["avg_ms"] else "node_update"
    El = lc.nelem    # elements this rank computes (own range + ghost layer); the stage's algorithmic bytes follow the work done
    stage_achieved = BYTES_STAGE * El / (stage_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if cands:   # dram__bytes_read+write per launch from the newest committed ncu capture (NOT measured in this run)
        with open(cands[-1]) as f:
            tj = json.load(f)
        traffic = sum(tj.get(k, 0) for k in (("stage_fused", "boundary_update") if fused else ("calcrhs_elem", "node_update")) if k in tj) or None
        traffic_src = f"profiles/{os.path.basename(cands[-1])} (committed ncu --set full capture, bytes per stage; not measured in this run)"
    roofline = {
        "bound": "hbm", "kernel": stage_kernels, "achieved": stage_achieved, "peak": peak, "unit": "GB/s", "frac": stage_achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_element": BYTES_STAGE, "launch_ms": stage_ms,
        "what": "the RK stage (calcRHS + ordered sum + nodal update + BCs), north_star's unit: 220 algorithmic B per element-stage "
                "x elements per launch / event-timed duration of the stage's launches",
        "dominant_kernel": {"name": dom, "avg_ms": prof[dom]["avg_ms"]},
        "stage": {"kernels": stage_kernels, "bytes_per_element_stage": BYTES_STAGE, "ms": stage_ms, "achieved": stage_achieved,
                  "frac": stage_achieved / peak, "element_stage_per_s": El / (stage_ms * 1e-3)},
        "step": {"bytes_per_element_step": BYTES_STEP, "achieved": BYTES_STEP * E * args.steps / (ms * 1e-3) / 1e9,
                 "frac": BYTES_STEP * E * args.steps / (ms * 1e-3) / 1e9 / peak},
        "kernels": prof,
    }

    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        names = ["U", "T", "VEL_X", "VEL_Y"]
        sizes = {nme: g.L.cfdb_field_size(g.h, nme.encode()) for nme in names}
        # two sets of pinned host arrays on each side (the host program double-buffers its state): call k reads set k%2 and
        # delivers into set k%2 of the outputs; uploads, steps and downloads of consecutive calls overlap (cfdb_step_streamed)
        hin = [{nme: torch.empty(sizes[nme], dtype=torch.float64, pin_memory=True) for nme in names} for _ in range(2)]
        hout = [{nme: torch.empty(sizes[nme], dtype=torch.float64, pin_memory=True) for nme in names} for _ in range(2)]
        hnorm = [torch.empty(8, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        for s in range(2):
            for nme in names:
                hin[s][nme].numpy()[:] = g.get(nme)
        h2d = sum(8 * sizes[nme] for nme in names)
        d2h = h2d + 64
        ksteps = max(3, min(args.steps, 50))   # the same K as the resident timing: the pipeline's fill (first upload) and drain (last download) are inside the timed region

        def one(kk):
            s = kk & 1
            g.step_streamed({nme: hin[s][nme].numpy() for nme in names}, {nme: hout[s][nme].numpy() for nme in names}, hnorm[s].numpy())

        one(0)
        one(1)
        g.streamed_wait()
        barrier()
        t0 = time.perf_counter()
        for kk in range(ksteps):
            one(kk)
        g.streamed_wait()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * E * ksteps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": ksteps, "ms_per_step": 1e3 * dt / ksteps, "numa_node": numa,
               "what": "cfdb_step_streamed: every step takes its state U,T,VEL_X,VEL_Y from pinned host arrays and delivers the new "
                       "state and the residual norms into pinned host arrays (double-buffered on the host side); upload of call "
                       "k+1, step k and download of call k-1 run concurrently on three streams (wall clock, max over ranks)"}
        del hin, hout
        # the literal per-subroutine drop-in (pageable host arrays in the Fortran layouts through the call-site entries): deltat,
        # RK and fluidStructure of one pass of the loop, every array of every call crossing PCIe both ways
        if world == 1:
            Pn, En = lc.npoin, lc.nelem
            arr = {k: g.get(k) for k in ("U", "T", "VEL_X", "VEL_Y", "W_X", "W_Y", "GAMM", "area", "X", "Y", "X1", "Y1", "P", "xpos", "ypos")}
            outs = {k: np.zeros(4 * Pn) for k in ("U1", "RHS", "RHS1", "RHS2", "RHS3")}
            outs.update({k: np.zeros(Pn) for k in ("RHO", "E", "RMACH")})
            outs.update({k: np.zeros(En) for k in ("SHOC", "T_SUGN1", "T_SUGN2", "T_SUGN3")})
            par = lc.par
            kcs = 2

            def callsite_step():
                dtmin, dt_e = g.deltat(arr["area"], arr["T"], arr["VEL_X"], arr["VEL_Y"], arr["W_X"], arr["W_Y"], par["FSAFE"], par["FR"],
                                       par["GAMA"], par["T_inf"])
                dtl = np.full(En, dtmin)
                g.rk_callsite(dtmin, 4, 5, arr["GAMM"], dtl, arr["U"], outs["U1"], outs["RHS"], outs["RHS1"], outs["RHS2"], outs["RHS3"],
                              arr["T"], arr["P"], outs["RHO"], outs["E"], outs["RMACH"], arr["VEL_X"], arr["VEL_Y"], arr["W_X"], arr["W_Y"],
                              outs["SHOC"], outs["T_SUGN1"], outs["T_SUGN2"], outs["T_SUGN3"])
                g.mesh_move(dtmin, 0.0, arr["X"], arr["Y"], arr["X1"], arr["Y1"], arr["W_X"], arr["W_Y"], arr["P"], arr["xpos"], arr["ypos"])
                arr["U"][:] = outs["U1"]

            callsite_step()
            t0 = time.perf_counter()
            for _ in range(kcs):
                callsite_step()
            dtc = time.perf_counter() - t0
            e2e["callsite"] = {"value": E * kcs / dtc, "unit": UNIT, "steps": kcs, "ms_per_step": 1e3 * dtc / kcs,
                               "what": "one pass of the loop as three call-site calls with pageable host arrays (cfdb_deltat, cfdb_rk, "
                                       "cfdb_mesh_move): ~1.6 GB up and ~2.4 GB down per step; the context is a moving-mesh one afterwards"}
            del arr, outs

    # ---- the other BASELINE configs, short (config.secondary) ----------------------------------------------------------
    secondary = {}
    if not args.no_secondary and not args.strong:
        g.close()
        del g
        if world == 1:
            # viscous 16 M (calcRHS.f90:119-137 path)
            gv, pv, Ev, _ = build(args.n, False, FMU=1.8e-5, FK=0.0257)
            sv = torch.cuda.ExternalStream(gv.stream, device=torch.device("cuda", local))
            gv.step(3)
            msv = timed_steps(gv, torch, sv, 10, lambda: (gv.sync(), torch.cuda.synchronize()))
            pk = kernel_profile(gv, 2)
            st_ms = pk.get("stage_fused", pk.get("calcrhs_elem"))["avg_ms"] + pk["node_update"]["avg_ms"]
            secondary["viscous16M"] = {"workload": f"{Ev}-triangle strip, FMU=1.8e-5, FK=0.0257 (viscous terms of calcRHS.f90:119-137), fixed mesh",
                                       "ms_per_step": msv / 10, "value": Ev * 10 / (msv * 1e-3), "unit": UNIT, "stage_ms": st_ms,
                                       "stage_frac": (BYTES_STAGE + 4) * pv.lc.nelem / (st_ms * 1e-3) / 1e9 / peak,
                                       "kernels": {k: round(v["avg_ms"], 4) for k, v in pk.items()}}
            gv.close()
            del gv
            # ALE 4 M (BASELINE configs[2]): pitching body, meshMove Laplace biCG + geometry refresh + GCL every step
            lca = deck.load(meshgen.ale_body(nt=2000, nr=1000, IPRINT=10**9, MAXITER=10**9))
            ga = NSComp2D(lca, device=local, use_gcl=1)
            sa = torch.cuda.ExternalStream(ga.stream, device=torch.device("cuda", local))
            ga.step(2)
            its = []
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ga.sync()
            ea.record(sa)
            for _ in range(8):
                ga.step(1)
                its.append((int(ga.scalar("bicg_x")), int(ga.scalar("bicg_y"))))
            eb.record(sa)
            ga.sync()
            torch.cuda.synchronize()
            msa = ea.elapsed_time(eb) / 8
            # biCG-iteration roofline: the reference's tolerance is absolute (1e-10 on r.z), so a physical step needs 0-1
            # iterations; perturb the warm start so that one solve runs many iterations and time its kernels
            rng = np.random.default_rng(1)
            ga.set("xpos", 1e-3 * rng.standard_normal(lca.npoin))
            ga.profile(True)
            ga.step(1)
            ga.sync()
            nit = max(1, int(ga.scalar("bicg_x")) + max(0, int(ga.scalar("bicg_y"))))
            tb = 0.0
            per = {}
            for kname in ("vec", "spmv", "dot", "scalar"):
                t_ms, nk = ga.profile_get(kname)
                per[kname] = t_ms
                tb += t_ms
            ga.profile(False)
            it_ms = tb / nit
            bicg = {"iterations": nit, "ms_per_iteration": it_ms, "bytes_per_node": BYTES_BICG_ITER_PER_NODE,
                    "achieved": BYTES_BICG_ITER_PER_NODE * lca.npoin / (it_ms * 1e-3) / 1e9,
                    "frac": BYTES_BICG_ITER_PER_NODE * lca.npoin / (it_ms * 1e-3) / 1e9 / peak,
                    "kernel_ms": {k: round(v, 4) for k, v in per.items()},
                    "what": "one biCG solve with a perturbed warm start (forces iterations; the physical steps above need 0-1): "
                            "all vec/spmv/dot/scalar kernel time of that step / iterations, against 192 B per node"}
            secondary["ale4M"] = {"workload": f"{lca.nelem}-triangle / {lca.npoin}-node O-mesh around a pitching ellipse, MOVING=1, gcl on",
                                  "ms_per_step": msa, "value": lca.nelem / (msa * 1e-3), "unit": UNIT,
                                  "bicg_iters_per_step": its, "bicg_iteration": bicg}
            ga.close()
            del ga
        else:
            n64 = 5657
            if (n64 - 1) % world == 0:
                gs, ps, Es, _ = build(n64, True)
                ss = torch.cuda.ExternalStream(gs.stream, device=torch.device("cuda", local))

                def bar2():
                    gs.sync()
                    torch.cuda.synchronize()
                    dist.barrier()

                gs.step(3)
                mss = max_over_ranks(timed_steps(gs, torch, ss, 10, bar2))
                secondary["strong64M"] = {"workload": f"one {world * Es}-triangle domain (lattice {n64}) cut into {world} strips (BASELINE configs[3])",
                                          "ms_per_step": mss / 10, "value": world * Es * 10 / (mss * 1e-3), "unit": UNIT, "scaling": "strong"}
                gs.close()
                del gs

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, ne, dt = cpu_oracle_throughput(args.ref_n, args.cpu_steps)
        v1, _, ne1, dt1 = cpu_oracle_throughput(709, 3, omp=False)     # threads = 1 next to all cores
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{ne}-triangle square mesh (same generator and flow as the GPU arm, smaller mesh: a per-element rate), "
                         f"{args.cpu_steps} steps ({dt:.1f} s), oracle OpenMP build (C++ restatement; the Fortran reference cannot be built here)",
               "single_thread": {"value": v1, "cores": 1, "sample": f"{ne1}-triangle square mesh, 3 steps ({dt1:.1f} s), "
                                 "sequential build (the summation order of the parity tests)"}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"square16M: {E}-triangle / {P}-node jittered-lattice Delaunay-diagonal strip per GPU "
                                   "(BASELINE configs[4]; the mesh north_star's target is stated on), Euler, fixed mesh, "
                                   "density-bump initial state", "l2": "inputs larger than L2 (no flush needed)",
                       "elements_per_gpu": E, "local_elements_incl_ghost_layer": lc.nelem,
                       "parallelism": f"{world} contiguous strips, owner-computes + NCCL ghost refresh per RK stage",
                       "stage": "fused tile kernel (stage_fused) + tile-boundary pass" if fused else "two kernels",
                       "element_stage_updates_per_s": 4 * value, "setup_s": t_gen, "secondary": secondary},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": launches,
        }
        if world > 1:
            line["multi_gpu_parity"] = parity
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
