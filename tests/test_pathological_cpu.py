"""CPU half of tests/pathological.py: the crafted operands (subnormal, zero, infinite, negative, overflowing) go through the
oracle's call-site entries on both sides of the comparison, so the helper that the GPU suite relies on for the fast-path
fallbacks stays runnable without a GPU, and its own sanity assertions (some NaNs, most entries finite) are checked here."""
import numpy as np

import pathological


class _OracleCallSites:
    """The subset of NSComp2D's call-site interface that pathological.check uses, answered by the oracle library."""

    def __init__(self, lc):
        from oracle import orclib

        self.lc, self.L = lc, orclib.lib()

    def set(self, name, value):
        pass

    def calcrhs(self, rhs, U, theta, T, dNx, dNy, area, shoc, dtl, t1, t2, t3, Cv, lam, mu, g0, T_inf, cte):
        lc = self.lc
        self.L.orc_calcrhs(rhs, U, theta, T, dNx, dNy, area, shoc, dtl, t1, t2, t3, lc.inpoel, lc.nelem, lc.npoin, Cv, lam, mu,
                           g0, T_inf, cte)
        return rhs

    def deltat(self, area, T, vx, vy, wx, wy, FSAFE, FR, GAMA, T_inf):
        lc = self.lc
        dt, dtmin = np.zeros(lc.nelem), np.zeros(1)
        self.L.orc_deltat(lc.nelem, lc.inpoel, area, T, vx, vy, wx, wy, FSAFE, FR, GAMA, T_inf, dt, dtmin)
        return dtmin[0], dt

    def estab(self, U, T, vx, vy, wx, wy, GAMM, dNx, dNy, FR, DTMIN, RHOINF, TINF):
        lc = self.lc
        out = [np.zeros(lc.nelem) for _ in range(4)]
        self.L.orc_estab(lc.nelem, lc.inpoel, U, T, vx, vy, wx, wy, GAMM, dNx, dNy, FR, DTMIN, RHOINF, TINF, *out)
        return out


def test_pathological_operands_are_well_formed_for_the_oracle():
    from cfd_b200 import deck, meshgen
    from oracle.orclib import Oracle

    lc = deck.load(meshgen.channel(nx=41, ny=13, FMU=1.8e-5, FK=0.0257))
    assert pathological.check(lc, _OracleCallSites(lc), Oracle(lc)) == 16
