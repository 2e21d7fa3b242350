"""The dataLoader boundary: text decks round-trip and the loader's post-processing follows dataLoader.f90."""
import numpy as np

from cfd_b200 import deck, meshgen


def test_deck_round_trip(tmp_path):
    raw = meshgen.ale_body(nt=16, nr=6, FMU=1.8e-5, FK=0.0257, CTE=2.0, ITLOCAL=7)
    raw.fixv = np.array([3, 5], np.int32)
    raw.fixt = (np.array([7], np.int32), np.array([1.25]))
    deck.write_deck(raw, str(tmp_path))
    back = deck.read_deck(str(tmp_path))
    a, b = deck.load(raw), deck.load(back)
    for f in ("X", "Y", "inpoel", "ifixrho_node", "rfixrho_value", "ifixv_node", "rfixv_valuex", "rfixv_valuey", "wall",
              "ifixt_node", "rfixt_value", "sets", "ifm", "i_m", "ilaux", "smooth_fix"):
        x, y = getattr(a, f), getattr(b, f)
        assert x.dtype == y.dtype and np.array_equal(x, y), f
    assert a.par == b.par


def test_loader_post_processing():
    raw = meshgen.channel(nx=9, ny=5, mach=2.0, CTE=4.0)
    raw.fixv = np.array([2, 3], np.int32)
    raw.fixt = (np.array([9], np.int32), np.array([1.5]))
    raw.fixrho = (np.array([1, 10], np.int32), np.array([-1.0, 2.0]))
    lc = deck.load(raw)
    p = lc.par
    assert p["CTE"] == 0.25                                            # dataLoader.f90:59
    assert p["P_inf"] == p["RHO_inf"] * p["FR"] * p["T_inf"]           # :61
    assert p["U_inf"] == p["C_inf"] * 2.0 and p["V_inf"] == 0.0        # :63-64
    assert lc.rfixrho_value.tolist() == [1.225, 2.0 * p["RHO_inf"]]    # :124-128
    nvi = raw.fixvi[0].size
    assert lc.ifixv_node.size == nvi + 2 and lc.rfixv_valuex[-2:].tolist() == [0.0, 0.0]   # :149-160, :165
    twall = np.float32(p["T_inf"] * (1.0 + (p["GAMA"] - 1) / 2.0 * 4.0))
    assert lc.rfixt_value[:2].tolist() == [float(twall)] * 2           # TWALL is single precision (F11)
    assert float(twall) != p["T_inf"] * (1.0 + (p["GAMA"] - 1) / 2.0 * 4.0)
    assert lc.ifixt_node.tolist() == [2, 3, 9] and lc.rfixt_value[2] == 1.5 * p["T_inf"]
    assert np.array_equal(lc.ilaux, np.concatenate([lc.i_m, lc.ifm]))


def test_generated_meshes_are_valid():
    for raw in (meshgen.channel(nx=21, ny=9), meshgen.wedge(nx=21, ny=11), meshgen.ale_body(nt=24, nr=8), meshgen.square(n=12)):
        inp = raw.inpoel - 1
        x, y = raw.X[inp], raw.Y[inp]
        a2 = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
        assert (a2 > 0).all() and inp.min() == 0 and inp.max() == raw.npoin - 1
        assert np.unique(inp).size == raw.npoin
    a, b = meshgen.square(n=12, seed=3), meshgen.square(n=12, seed=3)
    assert np.array_equal(a.X, b.X) and np.array_equal(a.inpoel, b.inpoel)      # seeded


def test_restart_file_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    n = 37
    U, T, G = rng.standard_normal((n, 4)), rng.random(n) + 250, np.full(n, 1.4)
    p = str(tmp_path / "case.RST")
    deck.write_rst(p, 123, 0.5, U, T, G)
    raw = open(p, "rb").read()
    assert len(raw) == 20 + n * 56 and raw[:4] == (12).to_bytes(4, "little")      # Fortran record framing
    it, time, U2, T2, G2 = deck.read_rst(p, n)
    assert (it, time) == (123, 0.5) and np.array_equal(U, U2) and np.array_equal(T, T2) and np.array_equal(G, G2)
    import pytest

    with pytest.raises(ValueError):
        deck.read_rst(p, n + 1)
