"""host/ns2dcomp — the C++ mirror of PROGRAM NSComp2D: deck reading, smoothing (CPU), and the run itself (GPU)."""
import os
import subprocess

import numpy as np
import pytest

from cfd_b200 import deck, meshgen
from conftest import assert_bit_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "ns2dcomp")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")], stdout=subprocess.DEVNULL)


def test_deck_reader_and_smoothing_match_oracle(tmp_path):
    from oracle import orclib

    _build()
    raw = meshgen.channel(nx=25, ny=9, jitter=0.42, seed=5, mach=0.7, CTE=2.0)
    raw.fixv = np.array([30, 31], np.int32)
    deck.write_deck(raw, str(tmp_path))
    fx, fy = str(tmp_path / "x.bin"), str(tmp_path / "y.bin")
    r = subprocess.run([EXE, str(tmp_path), "--check-deck", "--dump", "X:" + fx, "--dump", "Y:" + fy], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lc = deck.load(raw)
    assert f"U_inf={lc.par['U_inf']!r}" in r.stdout.replace("U_inf=", "U_inf=") or repr(lc.par["U_inf"])[:15] in r.stdout
    assert f"nfixv={lc.ifixv_node.size} " in r.stdout and f"nfixt={lc.ifixt_node.size} " in r.stdout
    X, Y = lc.X.copy(), lc.Y.copy()
    sweeps = orclib.lib().orc_smoothing(X, Y, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem)
    assert f"sweeps:{sweeps}" in r.stdout and sweeps > 0
    assert_bit_equal(np.fromfile(fx), X, "smoothed X")
    assert_bit_equal(np.fromfile(fy), Y, "smoothed Y")


def test_binary_mesh_sidecar_reads_like_the_text_deck(tmp_path):
    """<name>.cfdbmesh (binary side-car for meshes whose text deck would be gigabytes): the C++ reader and the
    Python reader give exactly what they give for the text file -- same post-processed lists, same smoothing result."""
    _build()
    raw = meshgen.ale_body(nt=24, nr=6, FMU=1.8e-5, FK=0.0257)
    raw.fixv = np.array([30, 31], np.int32)
    raw.fixt = (np.array([40, 41], np.int32), np.array([1.1, 0.9]))
    a, b = tmp_path / "text", tmp_path / "bin"
    deck.write_deck(raw, str(a))
    deck.write_deck(raw, str(b), binary_mesh=True)
    assert not (b / (raw.name + ".dat")).exists() and (b / (raw.name + ".cfdbmesh")).exists()
    ra, rb = deck.read_deck(str(a)), deck.read_deck(str(b))
    la, lb = deck.load(ra), deck.load(rb)
    for f in ("X", "Y", "inpoel", "ifixrho_node", "rfixrho_value", "ifixv_node", "rfixv_valuex", "rfixv_valuey", "wall", "ifixt_node",
              "rfixt_value", "sets", "ifm", "i_m", "ilaux", "smooth_fix"):
        assert np.array_equal(getattr(la, f), getattr(lb, f)), f
    outs = []
    for d in (a, b):
        fx = str(d / "x.bin")
        r = subprocess.run([EXE, str(d), "--check-deck", "--dump", "X:" + fx], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append((r.stdout, np.fromfile(fx)))
    assert outs[0][0] == outs[1][0] and "sweeps:" in outs[0][0]
    assert_bit_equal(outs[0][1], outs[1][1], "smoothed X")
    (b / (raw.name + ".cfdbmesh")).write_bytes(b"garbage!" + bytes(100))
    r = subprocess.run([EXE, str(b), "--check-deck"], capture_output=True, text=True)
    assert r.returncode != 0 and "binary mesh" in r.stderr


def test_bad_deck_stops(tmp_path):
    _build()
    r = subprocess.run([EXE, str(tmp_path), "--check-deck"], capture_output=True, text=True)
    assert r.returncode != 0 and "EULER.DAT" in r.stderr


@pytest.mark.gpu
def test_driver_run_matches_oracle(tmp_path):
    from oracle import orclib
    from oracle.orclib import Oracle

    _build()
    raw = meshgen.channel(nx=41, ny=13, jitter=0.4, seed=9, MAXITER=12, IPRINT=5, FMU=1.8e-5, FK=0.0257)
    deck.write_deck(raw, str(tmp_path))
    fu = str(tmp_path / "u.bin")
    r = subprocess.run([EXE, str(tmp_path), "--dump", "U:" + fu], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lc = deck.load(deck.read_deck(str(tmp_path)))
    orclib.lib().orc_smoothing(lc.X, lc.Y, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem)
    o = Oracle(lc)
    o.set_scalar("norms_every_step", 0)
    rows = []
    for it in range(1, 13):
        o.step(1)
        if it % 5 == 0 or it == 12:
            rows.append((it, o.scalar("TIME")))
    assert_bit_equal(np.fromfile(fu), o.get("U"), "U after 12 steps")
    cnv = [l.split() for l in open(tmp_path / "channel.cnv").read().strip().split("\n")]
    assert [int(c[0]) for c in cnv] == [5, 10, 12]
    for c, (it, t) in zip(cnv, rows):
        assert abs(float(c[1]) - t) <= 5e-6 * t       # Fortran E14.6: six significant digits
        assert all(0 < float(v) < 1 for v in c[2:6])
    # PRINTFLAVIA's GiD file of the last print step (MOVIE = 0), all seven blocks on ('.si.' flags of the deck)
    fl = open(tmp_path / "channel.flavia.res").read().split("\n")
    assert fl[0].split() == ["VELOCITY", "2", "12", "2", "1", "1"] and sum(1 for l in fl if l[:1] != " " and l) == 10
