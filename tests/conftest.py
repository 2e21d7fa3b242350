import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def bits(a):
    return np.ascontiguousarray(np.asarray(a, np.float64)).view(np.uint64)


def assert_bit_equal(a, b, name=""):
    """Bit-exact equality of two float64 arrays (NaNs with equal payload count as equal)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    ne = bits(a) != bits(b)
    if ne.any():
        i = int(np.flatnonzero(ne)[0])
        rel = np.abs(a[ne] - b[ne]) / np.maximum(np.abs(b[ne]), 1e-300)
        raise AssertionError(
            f"{name}: {int(ne.sum())}/{a.size} entries differ bitwise; first at {i}: {a[i]!r} vs {b[i]!r}; "
            f"max rel diff {np.nanmax(rel):.3e}")


@pytest.fixture(scope="session")
def cases():
    """Small instances of the BASELINE.json configs, loaded through the dataLoader mirror."""
    from cfd_b200 import deck, meshgen

    out = {}
    out["channel"] = deck.load(meshgen.channel(nx=41, ny=13))                       # Euler, slip walls, inflow
    out["channel_visc"] = deck.load(meshgen.channel(nx=41, ny=13, FMU=1.8e-5, FK=0.0257))
    out["channel_itlocal"] = deck.load(meshgen.channel(nx=31, ny=11, ITLOCAL=50))
    out["wedge"] = deck.load(meshgen.wedge(nx=49, ny=25, mach=2.5))
    out["ale"] = deck.load(meshgen.ale_body(nt=48, nr=14))
    out["square"] = deck.load(meshgen.square(n=33))
    # viscous channel with a no-slip lower wall (fixv -> also fixed wall temperature, single-precision TWALL),
    # a fixed-temperature patch, duplicated list entries (last entry wins) and a wall/no-slip overlap
    raw = meshgen.channel(nx=37, ny=11, FMU=1.8e-5, FK=0.0257, mach=0.6)
    nx = 37
    lower = np.arange(2, nx, dtype=np.int32)                       # interior nodes of the lower wall
    raw.wall = raw.wall[~np.isin(raw.wall, lower[5:]).any(1)]      # keep a few slip edges that overlap no-slip nodes
    raw.fixv = np.concatenate([lower, lower[:3]]).astype(np.int32)  # duplicates
    top = (10 * nx + np.arange(5, 15)).astype(np.int32)
    raw.fixt = (np.concatenate([top, top[:2]]).astype(np.int32), np.concatenate([np.full(10, 1.1), np.full(2, 0.9)]))
    raw.fixrho = (np.concatenate([raw.fixrho[0], raw.fixrho[0][:2]]).astype(np.int32),
                  np.concatenate([raw.fixrho[1], np.array([-1.0, 1.05])]))
    out["channel_noslip"] = deck.load(raw)
    return out
