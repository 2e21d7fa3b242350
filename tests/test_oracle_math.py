"""The oracle's arithmetic building blocks (oracle/orc_math.h), which the CUDA kernels repeat operation for
operation (cfd_b200/csrc/exact.cuh)."""
import math

import numpy as np

from oracle import orclib


def test_fixed_exponent_powers_are_the_correctly_rounded_libm_values():
    L = orclib.lib()
    rng = np.random.default_rng(1)
    xs = np.exp(rng.uniform(np.log(1e-6), np.log(1e6), 200000))
    for f, y in ((L.orc_pow15, 1.5), (L.orc_pow05, 0.5), (L.orc_powm05, -0.5)):
        got = np.array([f(float(x)) for x in xs[:60000]])
        ref = np.array([math.pow(float(x), y) for x in xs[:60000]])
        ulp = np.abs(got - ref) / np.spacing(np.abs(ref))
        # glibc's pow is accurate to ~0.52 ulp but not correctly rounded: the two may differ by one ulp, rarely
        assert ulp.max() <= 1.0
        assert (ulp == 0).mean() > 0.998, (y, (ulp == 0).mean())  # measured here: 0.9994 (1.5), 0.9990 (.5), 0.9993 (-.5)
    inf, nan = float("inf"), float("nan")
    assert L.orc_pow15(0.0) == 0.0 and L.orc_pow15(inf) == inf and math.isnan(L.orc_pow15(-1.0)) and math.isnan(L.orc_pow15(nan))
    assert L.orc_pow05(0.0) == 0.0 and math.copysign(1, L.orc_pow05(-0.0)) == 1.0 and L.orc_pow05(4.0) == 2.0
    assert L.orc_powm05(0.0) == inf and L.orc_powm05(inf) == 0.0 and math.isnan(L.orc_powm05(-2.0)) and L.orc_powm05(4.0) == 0.5
    # exact cases
    for x in (1.0, 4.0, 9.0, 0.25, 2.0**40):
        assert L.orc_pow15(x) == x * math.sqrt(x) and L.orc_powm05(x) == 1.0 / math.sqrt(x)


def test_division_by_three_sequence_is_exact():
    assert orclib.lib().orc_div3_mismatches(20_000_000, 12345) == 0


def _canon_py(v):
    def chunk(w):
        a = [0.0] * 256
        for l in range(256):
            acc = 0.0
            for i in range(l, len(w), 256):
                acc = acc + w[i]
            a[l] = acc
        s = 128
        while s >= 1:
            for l in range(s):
                a[l] = a[l] + a[l + s]
            s //= 2
        return a[0]

    v = list(v)
    if not v:
        return 0.0
    while True:
        p = [chunk(v[c:c + 4096]) for c in range(0, len(v), 4096)]
        if len(p) == 1:
            return p[0]
        v = p


def test_canonical_reduction_order():
    L = orclib.lib()
    rng = np.random.default_rng(3)
    for n in (0, 1, 255, 256, 257, 4096, 4097, 10000, 3 * 4096 * 2 + 17):
        v = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)
        got = L.orc_canon_sum(n, np.ascontiguousarray(v))
        assert got == _canon_py(v), n
        assert abs(got - math.fsum(v)) <= 1e-9 * (np.abs(v).sum() + 1e-300)
    x, y = rng.standard_normal(9000), rng.standard_normal(9000)
    assert L.orc_vecdot(9000, x, y) == _canon_py(x * y)


def test_power_of_two_scalings_fold_into_one_fma_exactly():
    """exact.cuh "exact scalings": for c in {0, +-1/2, +-2} the two IEEE operations of the source, c*x + t, give the
    correctly rounded value of the exact c*x + t (what one fused multiply-add returns), and (c*a)*b == c*(a*b).
    Exact rational arithmetic is the referee; float(Fraction) rounds to nearest-even."""
    from fractions import Fraction

    rng = np.random.default_rng(20261017)
    n = 20000
    mant = rng.random((3, n)) + 1.0
    # exponents near each other so that the additions round, signs mixed so that they cancel
    expo = rng.integers(-6, 7, size=(3, n))
    sign = rng.choice([-1.0, 1.0], size=(3, n))
    xs, ts, bs = (sign * np.ldexp(mant, expo)).tolist()
    for c in (0.0, 0.5, -0.5, 2.0, -2.0):
        fc = Fraction(c)
        for x, t, b in zip(xs, ts, bs):
            two_ops = c * x + t
            fused = float(fc * Fraction(x) + Fraction(t))
            assert two_ops == fused and math.copysign(1.0, two_ops) == math.copysign(1.0, fused)
            assert (c * x) * b == c * (x * b)
    # signed zeros: 0*x + t keeps IEEE's zero-sum rules on both sides (-0 only from (-0) + (-0))
    for x, t in ((3.0, -0.0), (-3.0, -0.0), (3.0, 0.0), (-3.0, 0.0)):
        r = 0.0 * x + t
        want_negative = (math.copysign(1.0, 0.0 * x) < 0) and (math.copysign(1.0, t) < 0)
        assert r == 0.0 and (math.copysign(1.0, r) < 0) == want_negative
