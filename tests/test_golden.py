"""Frozen oracle outputs (tests/golden/, made by tests/golden/make_golden.py): the oracle must keep reproducing
them bit for bit (CPU), and the CUDA path must reproduce them without any oracle at run time (GPU)."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import assert_bit_equal

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def _check(solver, gold, lc):
    assert np.array_equal(lc.X, gold["mesh_X"]) and np.array_equal(lc.inpoel, gold["mesh_inpoel"]), "mesh generator drifted"
    for f in mg.FIELDS:
        assert_bit_equal(solver.get(f), gold[f], f)
    for f in ("esup1", "psup1", "lap_idx"):
        assert np.array_equal(solver.get(f), gold[f]), f
    sc = [solver.scalar(s) for s in ("DTMIN", "TIME", "ITER", "BANDERA", "bicg_x", "bicg_y")]
    assert sc == gold["scalars"].tolist()


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_reproduces_golden(name):
    from cfd_b200 import meshgen
    from oracle.orclib import Oracle

    lc, bump, gcl = mg.build(name)
    o = Oracle(lc, use_gcl=gcl)
    if bump:
        for k, v in meshgen.density_bump(lc).items():
            o.set(k, v)
    o.step(mg.STEPS)
    _check(o, np.load(os.path.join(HERE, "golden", name + ".npz")), lc)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.CASES))
def test_gpu_reproduces_golden(name):
    from cfd_b200 import meshgen
    from cfd_b200.solver import NSComp2D

    lc, bump, gcl = mg.build(name)
    g = NSComp2D(lc, use_gcl=gcl)
    if bump:
        for k, v in meshgen.density_bump(lc).items():
            g.set(k, v)
    g.step(mg.STEPS)
    _check(g, np.load(os.path.join(HERE, "golden", name + ".npz")), lc)
