// ORACLE (test infrastructure, never shipped, never on the product path).
//
// Arithmetic rules of the CPU restatement of chanshing/cfd:
//   * every + - * / sqrt is a single IEEE-754 binary64 operation evaluated left to right exactly
//     as the Fortran source writes it; this file and oracle.cpp are compiled with
//     -ffp-contract=off so GCC never fuses a*b+c;
//   * x**2 and x**2.d0 are x*x (GCC folds pow(x,2) exactly);
//   * x**1.5d0, x**.5d0, x**(-.5d0) are calls into libm `pow` in a gfortran build
//     (subrutinas.f90:194,196,400,429,438,439; calcRHS.f90:50,51).  libm is a third-party
//     dependency that is not under /root/reference: glibc 2.39 here; its pow (Szabolcs Nagy's
//     table-driven algorithm, sysdeps/ieee754/dbl-64/e_pow.c) is accurate to 0.52 ULP, is NOT
//     correctly rounded, and is selected at run time between an FMA and a non-FMA build
//     (ifunc), so the reference binary's own bits differ between machines at this level.
//     Its lookup tables cannot be regenerated offline.  The oracle therefore pins these three
//     fixed-exponent powers to the value a correctly rounded pow returns, computed with
//     double-double residuals built only from IEEE fma/mul/add/div/sqrt, so that the CUDA path
//     can evaluate the *same* sequence of IEEE operations and agree bit for bit.
//     tests/test_oracle_math.py measures how often this equals the host glibc pow (>99.9 %).
//
// Why bit-exactness matters at all: ESTAB (subrutinas.f90:380-393) turns round-off noise in
// VEL2*dNx1+VEL2*dNx2+VEL2*dNx3 into an O(1) switch on T_SUGN2, so a one-ulp
// difference anywhere in the nodal state flips stabilisation terms on the next step.
#pragma once
#include <cmath>
#include <limits>

namespace orc {

// x**1.5d0  — correctly rounded x*sqrt(x) up to a 2^-104 relative slack.
static inline double pow15(double x) {
    if (!(x > 0.0)) {                       // 0, negatives, NaN
        if (x == 0.0) return 0.0;           // pow(+-0, 1.5) = +0
        return std::numeric_limits<double>::quiet_NaN();
    }
    if (x == std::numeric_limits<double>::infinity()) return x;
    double s = std::sqrt(x);                // RN(sqrt x)
    double e = __builtin_fma(-s, s, x);     // exact: x - s*s
    double d = e / (2.0 * s);               // sqrt(x) = s + d (+ O(2^-106))
    double ph = x * s;
    double pl = __builtin_fma(x, s, -ph);   // exact low part of x*s
    double t = pl + x * d;
    return ph + t;
}

// x**.5d0
static inline double pow05(double x) {
    if (x == 0.0) return 0.0;               // pow(-0, .5) = +0
    return std::sqrt(x);                    // NaN for x<0 like pow
}

// x**(-.5d0) — correctly rounded 1/sqrt(x) up to a 2^-104 relative slack.
static inline double powm05(double x) {
    if (!(x > 0.0)) {
        if (x == 0.0) return std::numeric_limits<double>::infinity();
        return std::numeric_limits<double>::quiet_NaN();
    }
    if (x == std::numeric_limits<double>::infinity()) return 0.0;
    double s = std::sqrt(x);
    double y = 1.0 / s;                     // first approximation of x^-1/2
    double t = x * y;
    double tl = __builtin_fma(x, y, -t);    // x*y = t + tl exactly
    double u = __builtin_fma(-t, y, 1.0);   // 1 - t*y
    double rho = __builtin_fma(-tl, y, u);  // 1 - x*y*y
    return __builtin_fma(y, 0.5 * rho, y);  // y*(1 + rho/2)
}

// Canonical reduction order for the sums whose order the reference leaves to OpenMP
// (`reduction(+:…)` at biconjGrad.f90:162 and ns2DComp.ALE.f90:192): fixed 4096-entry chunks;
// inside a chunk 256 lanes each add their stride-256 entries in ascending order, then a
// binary tree (stride 128,64,…,1) folds the lanes; chunk sums are reduced by the same rule
// recursively.  Any order is a legal realisation of the OpenMP reduction; this one is cheap
// on a GPU and is what both the oracle and the CUDA path implement.
template <class F>
static inline double canon_chunk(long lo, long hi, F term) {
    double a[256];
    for (int l = 0; l < 256; ++l) {
        double acc = 0.0;
        for (long i = lo + l; i < hi; i += 256) acc = acc + term(i);
        a[l] = acc;
    }
    for (int s = 128; s >= 1; s >>= 1)
        for (int l = 0; l < s; ++l) a[l] = a[l] + a[l + s];
    return a[0];
}

template <class F>
static inline double canon_sum(long n, F term) {
    if (n <= 0) return 0.0;
    long m = (n + 4095) / 4096;
    double* p = new double[m];
    for (long c = 0; c < m; ++c) {
        long lo = c * 4096, hi = lo + 4096 < n ? lo + 4096 : n;
        p[c] = canon_chunk(lo, hi, term);
    }
    while (m > 1) {
        long m2 = (m + 4095) / 4096;
        double* q = new double[m2];
        for (long c = 0; c < m2; ++c) {
            long lo = c * 4096, hi = lo + 4096 < m ? lo + 4096 : m;
            q[c] = canon_chunk(lo, hi, [&](long i) { return p[i]; });
        }
        delete[] p;
        p = q;
        m = m2;
    }
    double r = p[0];
    delete[] p;
    return r;
}

}  // namespace orc
