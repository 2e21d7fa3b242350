"""cfd_b200 — B200 (sm_100a) implementation of the per-timestep hot path of chanshing/cfd.

The product is the C-ABI shared library `libcfdb200.so` (include/cfdb.h, cfd_b200/csrc/).
This package is the Python-side mirror of the reference's driver interface on top of it:

    deck     — the dataLoader input-deck boundary (EULER.DAT, <name>-1.dat, <name>.dat)
    meshgen  — seeded synthetic meshes for the BASELINE.json configs
    capi     — ctypes binding of include/cfdb.h (fails loudly if the CUDA library is missing)
    solver   — NSComp2D: the time loop of ns2DComp.ALE.f90 driven through the C ABI
    partition— contiguous sub-domain decomposition for multi-GPU runs

There is no CPU fallback anywhere in this package.
"""
from . import deck, meshgen  # noqa: F401
