"""Run the reference's own Fortran (translated statement by statement by translate.py) -- TEST INFRASTRUCTURE ONLY.

    ref = Reference()                      # parses /root/reference/*.f90 (raises ReferenceMissing on the GPU box)
    ref.run_program(raw_case, maxiter)     # PROGRAM NSComp2D itself: readInputData, loadMeshData, RESTART, smoothing,
                                           # normales/deriv/masas/laplace, then the time loop, on a deck written to a tmp dir
    ref.mod("mvariabgen").u                # module variables afterwards (Fortran shapes, e.g. U is (4, npoin))

Nothing is stubbed: formatted output (PRINTFLAVIA's GiD file, FORCES, ...) is kept as text records in
Reference.io.text, PRINTREST's unformatted restart dump as bytes in Reference.io.binary.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np

from . import runtime as rt
from . import translate

REF_DIR = os.environ.get("CFD_REFERENCE_DIR", "/root/reference")
FILES = ["commonModules", "dataLoader", "pointNeighbor", "calcRHS", "subrutinas", "biconjGrad", "mLaplace", "gcl",
         "smoothing", "meshMove", "ns2DComp.ALE"]


class ReferenceMissing(Exception):
    pass


class _ModView:
    def __init__(self, obj):
        object.__setattr__(self, "_o", obj)

    def __getattr__(self, name):
        return getattr(self._o, "v_" + name.lower())

    def __setattr__(self, name, value):
        setattr(self._o, "v_" + name.lower(), value)


class Reference:
    _cache = None

    def __init__(self, ref_dir=REF_DIR):
        paths = [os.path.join(ref_dir, f + ".f90") for f in FILES]
        if not all(os.path.exists(p) for p in paths):
            raise ReferenceMissing(ref_dir)
        if Reference._cache is None:
            src, ns = translate.build(paths)
            Reference._cache = (src, ns, {k: v for k, v in ns.items() if k.startswith("p_")})
        self.src, self.ns, self._orig = Reference._cache
        self.reset()

    def reset(self):
        self.ns.update(self._orig)      # drop the overrides of an earlier run
        self.ns["IO"].__init__()
        self.ns["_reset"]()

    def mod(self, name):
        return _ModView(self.ns["M_" + name.lower()])

    def proc(self, name, module=""):
        return self.ns[f"p_{module.lower()}__{name.lower()}"]

    @property
    def io(self):
        return self.ns["IO"]

    def run_program(self, raw, maxiter=None, hook=None, files=None):
        """write the deck (+ extra files: {name: bytes}), run PROGRAM NSComp2D; returns the records written to <name>.cnv"""
        from cfd_b200 import deck
        import copy

        raw = copy.copy(raw)
        if maxiter is not None:
            raw.MAXITER = int(maxiter)
        self.reset()
        if hook is not None:
            hook(self)
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as d, np.errstate(all="ignore"):
            deck.write_deck(raw, d)
            for fname, data in (files or {}).items():
                with open(os.path.join(d, fname), "wb") as f:
                    f.write(data)
            os.chdir(d)
            try:
                self.ns["p___nscomp2d"]()
            finally:
                os.chdir(cwd)
        return self.io.written.get(raw.name + ".cnv", [])


def f_array(a, dtype=None):
    """C-ordered (n, k) numpy array of the repo's layout -> Fortran (k, n) array sharing nothing"""
    a = np.asarray(a, dtype=dtype)
    return np.asfortranarray(a.T.copy(order="F")) if a.ndim == 2 else a.copy()
