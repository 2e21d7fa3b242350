"""Round-2 additions that need no GPU: list-directed REAL(8) layout, the greedy element colouring against a brute-force
restatement, the oracle's ADAMSB pinned to the reference's own routine (interpreted from /root/reference), the static
schedule tool on a built object."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import assert_bit_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.f90ref import refrun  # noqa: E402

needs_ref = pytest.mark.skipif(not refrun.available(), reason="neither the reference sources nor oracle/_ref/refprog.py are here")


def _fmt(kind, v, w=0, d=0):
    from cfd_b200 import capi

    buf = C.create_string_buffer(128)
    capi.check(capi.lib().cfdb_format_real(ord(kind), float(v), w, d, buf, 128))
    return buf.value.decode()


def test_list_directed_real8_layout():
    """gfortran's list-directed REAL(8): width 25, 17 significant digits, F editing with five trailing blanks inside
    0.1 <= |x| < 1e17 and for zero, ES editing with a three-digit exponent outside (fortran/README.md build line)"""
    want = {3.14: "   3.1400000000000001     ", 1e-3: "   1.0000000000000000E-003", 0.0: "   0.0000000000000000     ",
            123456789.0: "   123456789.00000000     ", 1e20: "   1.0000000000000000E+020", -2.5e-7: "  -2.4999999999999999E-007",
            0.1: "  0.10000000000000001     ", 1e16: "   10000000000000000.     ", 1.0: "   1.0000000000000000     "}
    for v, s in want.items():
        got = " " + _fmt("L", v)          # a record starts with one blank
        assert got == s, (v, got, s)
        assert float(got) == v            # 17 significant digits round-trip
    rng = np.random.default_rng(0)
    for v in np.concatenate([rng.normal(size=200) * 10.0 ** rng.integers(-30, 30, 200), [1e-310, -1e308]]):
        s = _fmt("L", v)
        assert len(s) == 25 and float(s) == v, (v, s)


def test_greedy_element_colouring():
    """first-fit in ascending element order with a forbidden mask per node"""
    from cfd_b200 import capi, deck, meshgen

    lc = deck.load(meshgen.channel(nx=41, ny=13))
    col, nc = np.zeros(lc.nelem, np.int32), C.c_int32()
    capi.check(capi.lib().cfdb_color_elements(lc.inpoel, lc.nelem, lc.npoin, col, C.byref(nc)))
    used = [set() for _ in range(lc.npoin)]
    for e, tri in enumerate(lc.inpoel):
        forb = used[tri[0] - 1] | used[tri[1] - 1] | used[tri[2] - 1]
        c = 0
        while c in forb:
            c += 1
        assert col[e] == c, e
        for n in tri:
            used[n - 1].add(c)
    assert nc.value == col.max() + 1 and 6 <= nc.value <= 12
    for c in range(nc.value):                      # a colour class touches every node at most once
        nodes = lc.inpoel[col == c].ravel()
        assert np.unique(nodes).size == nodes.size


@needs_ref
def test_oracle_adamsb_equals_the_references_routine():
    """ADAMSB (subrutinas.f90:851-1034) is never called by PROGRAM NSComp2D (the call at ns2DComp.ALE.f90:177 is commented
    out): call the reference's subroutine on a post-run state, with NESTAB = 2 so that its CUARTO_ORDEN + ESTAB branch runs,
    and compare every array it writes with the oracle's restatement."""
    import importlib.util

    from cfd_b200 import deck, meshgen
    from oracle import orclib
    from oracle.f90ref.refrun import Reference
    from oracle.orclib import Oracle

    spec = importlib.util.spec_from_file_location("make_golden_ref", os.path.join(ROOT, "tests", "golden", "make_golden_ref.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    name = "ref_channel_visc"
    raw = mg.raw_case(name)
    lc = deck.load(raw)
    st = meshgen.density_bump(lc)
    ref = Reference()
    ref.run_program(raw, maxiter=3, initial_state=st)
    md, g, var, vel, est = (ref.mod(m) for m in ("meshdata", "mvariabgen", "mvariables", "mvelocidades", "mestabilizacion"))
    npoin, nelem = int(md.npoin), int(md.nelem)
    orclib.lib().orc_smoothing(lc.X, lc.Y, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem)
    o = Oracle(lc)
    for k, v in st.items():
        o.set(k, v)
    o.step(3)
    assert_bit_equal(o.get("U"), g.u.T.ravel(), "state before ADAMSB")
    dtmin = o.scalar("DTMIN")
    rng = np.random.default_rng(11)
    hist = [1e-3 * rng.normal(size=(4, npoin)) for _ in range(3)]       # some RHS history
    for a, h in zip((g.rhs1, g.rhs2, g.rhs3), hist):
        a[...] = h
    for nm, h in zip(("RHS1", "RHS2", "RHS3"), hist):
        o.set(nm, h.T.ravel())
    g.u1[...] = g.u                                                        # ns2DComp.ALE.f90:168-172
    gamm = np.full(npoin, float(ref.mod("inputdata").gama))
    dtl = np.full(nelem, dtmin)
    with np.errstate(all="ignore"):
        ref.proc("adamsb")(np.float64(dtmin), 2, gamm, dtl)
    o.set("DTL", dtl)
    o.set("U1", o.get("U"))
    o.set_scalar("NESTAB", 2)
    o.L.orc_adamsb(o.h)
    for nm, a in (("U1", g.u1), ("RHS", g.rhs), ("RHS1", g.rhs1), ("RHS2", g.rhs2), ("RHS3", g.rhs3), ("UN", g.un)):
        assert_bit_equal(o.get(nm), a.T.ravel(), f"ADAMSB {nm}")
    for nm, a in (("T", var.t), ("P", var.p), ("RHO", var.rho), ("E", var.e), ("RMACH", var.rmach), ("VEL_X", vel.vel_x),
                  ("VEL_Y", vel.vel_y), ("SHOC", est.shoc), ("T_SUGN2", est.t_sugn2)):
        assert_bit_equal(o.get(nm), a, f"ADAMSB {nm}")
    assert np.max(np.abs(g.un)) > 0       # the CUARTO_ORDEN projection is kept here (no UN = 0.0 as in RK)


def test_static_schedule_tool_runs_on_the_built_object():
    import subprocess

    obj = os.path.join(ROOT, "cfd_b200", "csrc", "cfdb.o")
    if not os.path.exists(obj):
        pytest.skip("cfdb.o not built here")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_stalls.py"), obj, "stage_fusedILb0ELi12"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "sum of stall fields" in r.stdout, r.stdout + r.stderr


def test_tile_decomposition_of_the_fused_stage():
    """host_topology.h: rcb_split / build_tiling through the host-only entry cfdb_tile_elements.  The internal element order
    is a permutation; every tile but the last is full; recursive coordinate bisection leaves more nodes interior to one
    tile than runs of a Morton curve; the result does not depend on the number of OpenMP threads."""
    from cfd_b200 import capi, deck, meshgen

    lc = deck.load(meshgen.square(201))
    L = capi.lib()
    out = {}
    for order in (0, 1, 2):
        i2e, st = np.zeros(lc.nelem, np.int32), np.zeros(8)
        capi.check(L.cfdb_tile_elements(lc.inpoel, lc.nelem, lc.npoin, lc.X, lc.Y, 384, order, i2e, st))
        assert np.array_equal(np.sort(i2e), np.arange(lc.nelem)), "i2e must be a permutation"
        assert st[1] == -(-lc.nelem // 384)
        assert st[6] + round(st[0] * lc.npoin) == lc.npoin          # interior + tile-boundary (incl. orphans) = all nodes
        out[order] = (i2e, st)
    assert np.array_equal(out[0][0], np.arange(lc.nelem))            # order 0: the file's order
    assert out[2][1][0] > out[1][1][0] > 0.7                         # rcb beats Morton runs; both keep most nodes interior
    assert out[2][1][0] > 0.80
    # a tile of the rcb order is a compact patch: its bounding box holds about as many element centroids as the tile has elements
    i2e = out[2][0]
    cx = lc.X[lc.inpoel - 1].mean(1)
    cy = lc.Y[lc.inpoel - 1].mean(1)
    worst = 0.0
    for t in range(0, lc.nelem // 384, 7):
        e = i2e[384 * t:384 * (t + 1)]
        inside = ((cx >= cx[e].min()) & (cx <= cx[e].max()) & (cy >= cy[e].min()) & (cy <= cy[e].max())).sum()
        worst = max(worst, inside / 384.0)
    assert worst < 1.6, worst
    # elements inside a tile are in ascending original id
    for t in range(0, lc.nelem // 384, 11):
        e = i2e[384 * t:384 * (t + 1)]
        assert np.all(np.diff(e) > 0)
    old = os.environ.get("OMP_NUM_THREADS")
    again = np.zeros(lc.nelem, np.int32)
    capi.check(L.cfdb_tile_elements(lc.inpoel, lc.nelem, lc.npoin, lc.X, lc.Y, 384, 2, again, np.zeros(8)))
    assert np.array_equal(again, i2e)
    if old is not None:
        os.environ["OMP_NUM_THREADS"] = old
    with pytest.raises(RuntimeError):
        capi.check(L.cfdb_tile_elements(lc.inpoel, lc.nelem, lc.npoin, lc.X, lc.Y, 100, 2, again, np.zeros(8)))


@pytest.mark.parametrize("mesh", ["channel", "ale", "wedge"])
def test_tile_schedule_is_consistent_with_the_mesh(mesh):
    """cfdb_tile_elements runs host_topology.h: check_tiling on what build_tiling produced and fails on any violation: every
    element in exactly one tile position, lnode naming the element's own vertices, an interior node's slot list = its element
    list in ascending original element id, the boundary records a one-to-one map onto each tile-boundary node's run in the
    same order.  Unstructured-looking meshes (O-mesh around a body, wedge, channel), three element orders, three tile sizes."""
    from cfd_b200 import capi, deck, meshgen

    raw = {"channel": lambda: meshgen.channel(nx=61, ny=21), "ale": lambda: meshgen.ale_body(nt=96, nr=24),
           "wedge": lambda: meshgen.wedge(nx=61, ny=31)}[mesh]()
    lc = deck.load(raw)
    for TE in (64, 384, 512):
        for order in (0, 1, 2):
            st, i2e = np.zeros(8), np.zeros(lc.nelem, np.int32)
            capi.check(capi.lib().cfdb_tile_elements(lc.inpoel, lc.nelem, lc.npoin, lc.X, lc.Y, TE, order, i2e, st))
            assert st[1] == -(-lc.nelem // TE) and 0.0 <= st[0] <= 1.0


def test_committed_bench_lines_follow_the_contract():
    """The JSON lines committed under profiles/ (produced by bench.py on the GPU boxes) carry every key of the bench contract,
    with self-consistent numbers."""
    import json

    def line(name):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.loads(f.read().strip().split("\n")[-1])

    d = line("r2_bench_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    E = d["config"]["elements_per_gpu"]
    assert abs(d["value"] - E / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "traffic_source"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["achieved"] - 220 * E / (r["launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["traffic"] is not None and 0.9 < r["traffic"] / (220 * E) < 1.5          # measured DRAM bytes per stage vs algorithmic
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] > 0 and d["clocks"]["samples"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for name, n in (("r2_bench_n2.json", 2), ("r2_bench_n4.json", 4), ("r2_bench_n8.json", 8)):
        m = line(name)
        assert m["n_gpus"] == n and m["multi_gpu_parity"] == "bit-exact" and m["scaling"] == "weak"
        assert 0.9 < m["value"] / (n * d["value"]) <= 1.02                             # weak-scaling efficiency of the committed lines
    ref = line("r2_bench_ref_n1.json")
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["value"] == ref["value"]


def _qualities(lc):
    t = lc.inpoel - 1
    x, y = lc.X[t], lc.Y[t]
    a = x[:, 1] * y[:, 2] + x[:, 2] * y[:, 0] + x[:, 0] * y[:, 1] - (x[:, 1] * y[:, 0] + x[:, 2] * y[:, 1] + x[:, 0] * y[:, 2])
    l = sum((x[:, i] - x[:, j]) ** 2 + (y[:, i] - y[:, j]) ** 2 for i, j in ((2, 1), (0, 2), (1, 0)))
    return 3.46410161513775 * a / l          # mu of smoothing.f90:305-317 (1 = equilateral, <= 0 = inverted)


def test_colour_ordered_smoothing_is_thread_independent_and_improves_the_mesh(tmp_path):
    """cfdb_smoothing_colored runs the reference's per-node optimiser colour by colour on all host cores.  Its
    result is bit-identical for 1 and 4 OpenMP threads, leaves the fixed nodes where they are, produces no inverted element and
    improves the worst element about as much as the serial optimiser does (it is NOT the serial result: different node order)."""
    import subprocess

    from cfd_b200 import capi, deck, meshgen

    script = tmp_path / "w.py"
    script.write_text(
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from cfd_b200 import capi, deck, meshgen\n"
        "lc = deck.load(meshgen.channel(nx=41, ny=15, jitter=0.45, seed=4))\n"
        "n = capi.smoothing_colored(lc)\n"
        "np.savez(sys.argv[1], X=lc.X, Y=lc.Y, n=n)\n")
    outs = []
    for thr in ("1", "4"):
        f = str(tmp_path / f"o{thr}.npz")
        r = subprocess.run([sys.executable, str(script), f], env=dict(os.environ, OMP_NUM_THREADS=thr), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(np.load(f))
    assert_bit_equal(outs[0]["X"], outs[1]["X"], "X: 1 thread vs 4 threads")
    assert_bit_equal(outs[0]["Y"], outs[1]["Y"], "Y: 1 thread vs 4 threads")
    assert int(outs[0]["n"]) >= 1

    lc0 = deck.load(meshgen.channel(nx=41, ny=15, jitter=0.45, seed=4))
    q0 = _qualities(lc0)
    lcs = deck.load(meshgen.channel(nx=41, ny=15, jitter=0.45, seed=4))
    capi.smoothing(lcs)
    qs = _qualities(lcs)
    lcc = deck.load(meshgen.channel(nx=41, ny=15, jitter=0.45, seed=4))
    capi.smoothing_colored(lcc)
    qc = _qualities(lcc)
    assert_bit_equal(lcc.X, outs[0]["X"], "same result in this process")
    fixed = lc0.smooth_fix.astype(bool)
    assert np.array_equal(lcc.X[fixed], lc0.X[fixed]) and np.array_equal(lcc.Y[fixed], lc0.Y[fixed])
    assert qc.min() > 0 and qc.min() > q0.min()                       # no inverted element, the worst element got better
    assert qc.min() > 0.8 * qs.min() and np.mean(qc) > 0.98 * np.mean(qs)   # about as good as the serial optimiser
    assert not np.array_equal(lcc.X, lcs.X)                          # ... but not the same mesh
