"""Integer artefacts of the product's host code (cfd_b200/csrc/host_topology.h through the C ABI, no GPU needed)
are bit-exact against the oracle's restatement of pointNeighbor.f90."""
import numpy as np
import pytest

from cfd_b200 import capi
from oracle import orclib


@pytest.mark.parametrize("name", ["channel", "wedge", "ale", "square"])
def test_esup_psup_bit_exact(cases, name):
    lc = cases[name]
    L = orclib.lib()
    e1, e2 = np.zeros(3 * lc.nelem, np.int32), np.zeros(lc.npoin + 1, np.int32)
    L.orc_get_esup(lc.inpoel, lc.nelem, lc.npoin, e1, e2)
    g1, g2 = capi.get_esup(lc.inpoel, lc.npoin)
    assert np.array_equal(e1, g1) and np.array_equal(e2, g2)
    p1, p2 = np.zeros(8 * lc.nelem, np.int32), np.zeros(lc.npoin + 1, np.int32)
    cnt = L.orc_get_psup(lc.inpoel, lc.nelem, lc.npoin, p1, p1.size, p2)
    q1, q2 = capi.get_psup(lc.inpoel, lc.npoin)
    assert cnt == q1.size and np.array_equal(p1[:cnt], q1) and np.array_equal(p2, q2)


def test_psup_capacity_error(cases):
    import ctypes as C

    lc = cases["channel"]
    p1, p2, cnt = np.zeros(4, np.int32), np.zeros(lc.npoin + 1, np.int32), C.c_int32()
    rc = capi.lib().cfdb_get_psup(lc.inpoel, lc.nelem, lc.npoin, p1, 4, p2, C.byref(cnt))
    assert rc != 0 and b"capacity" in capi.lib().cfdb_last_error() and cnt.value > 4


def test_ragged_and_tiny_meshes():
    # a single triangle, and a fan with an isolated (unreferenced) node in the middle of the numbering
    for inpoel, npoin in ((np.array([[1, 2, 3]], np.int32), 3), (np.array([[1, 2, 5], [1, 5, 4], [2, 6, 5]], np.int32), 6)):
        L = orclib.lib()
        e1, e2 = np.zeros(3 * len(inpoel), np.int32), np.zeros(npoin + 1, np.int32)
        L.orc_get_esup(inpoel, len(inpoel), npoin, e1, e2)
        g1, g2 = capi.get_esup(inpoel, npoin)
        assert np.array_equal(e1, g1) and np.array_equal(e2, g2)
        q1, q2 = capi.get_psup(inpoel, npoin)
        p1, p2 = np.zeros(32, np.int32), np.zeros(npoin + 1, np.int32)
        cnt = L.orc_get_psup(inpoel, len(inpoel), npoin, p1, 32, p2)
        assert np.array_equal(p1[:cnt], q1) and np.array_equal(p2, q2)
    assert q2[3] == q2[2]  # node 3 has no neighbours


def test_product_smoothing_matches_oracle():
    """cfdb_smoothing (host code inside libcfdb200.so, no GPU needed) is bit-exact against the oracle's restatement."""
    from cfd_b200 import deck, meshgen

    for seed, jit in ((5, 0.42), (8, 0.45)):
        lc = deck.load(meshgen.channel(nx=31, ny=11, jitter=jit, seed=seed))
        X, Y = lc.X.copy(), lc.Y.copy()
        so = orclib.lib().orc_smoothing(X, Y, lc.inpoel, lc.smooth_fix, lc.npoin, lc.nelem)
        sp = capi.smoothing(lc)
        assert so == sp and so > 0
        assert np.array_equal(lc.X.view(np.uint64), X.view(np.uint64)) and np.array_equal(lc.Y.view(np.uint64), Y.view(np.uint64))
