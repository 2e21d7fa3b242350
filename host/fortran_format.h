// Fortran formatted-output edit descriptors as gfortran lays them out (Ew.d, Fw.d, Iw, Aw), for the files the reference
// writes with FORMAT strings: <name>.cnv (ns2DComp.ALE.f90:199) and the GiD post file (PRINTFLAVIA, :701-817).
// tests/test_output_formats.py compares the bytes with the reference's own WRITE statements executed by oracle/f90ref.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace ffmt {

inline std::string rjust(const std::string& s, int w) {
    if ((int)s.size() > w) return std::string((size_t)w, '*');
    return std::string((size_t)(w - (int)s.size()), ' ') + s;
}

// Ew.d : [-]0.ddddE+ee, the leading zero dropped when the field is too narrow; three-digit exponents lose the 'E'
inline std::string E(double v, int w, int d) {
    std::string s;
    if (std::isnan(v)) s = "NaN";
    else if (std::isinf(v)) s = std::string(v < 0 ? "-" : "") + (w >= 8 + (v < 0 ? 1 : 0) ? "Infinity" : "Inf");
    else {
        const bool neg = std::signbit(v);
        std::string digits((size_t)d, '0');
        int ex = 0;
        if (v != 0.0) {
            char buf[64];
            std::snprintf(buf, sizeof buf, "%.*e", d - 1, std::fabs(v));  // d significant digits, correctly rounded
            std::string m(buf);
            size_t epos = m.find('e');
            ex = std::atoi(m.c_str() + epos + 1) + 1;
            digits.clear();
            for (size_t i = 0; i < epos; ++i)
                if (m[i] != '.') digits.push_back(m[i]);
        }
        char eb[16];
        if (std::abs(ex) <= 99) std::snprintf(eb, sizeof eb, "E%+03d", ex);
        else std::snprintf(eb, sizeof eb, "%+04d", ex);
        std::string body = "." + digits + eb;
        s = std::string(neg ? "-" : "") + "0" + body;
        if ((int)s.size() > w) s = std::string(neg ? "-" : "") + body;
    }
    return rjust(s, w);
}

inline std::string F(double v, int w, int d) {
    std::string s;
    if (std::isnan(v)) s = "NaN";
    else if (std::isinf(v)) s = std::string(v < 0 ? "-" : "") + (w >= 8 + (v < 0 ? 1 : 0) ? "Infinity" : "Inf");
    else {
        char buf[512];
        std::snprintf(buf, sizeof buf, "%.*f", d, v);
        s = buf;
        if ((int)s.size() > w) {
            if (s.rfind("0.", 0) == 0) s = s.substr(1);
            else if (s.rfind("-0.", 0) == 0) s = "-" + s.substr(2);
        }
    }
    return rjust(s, w);
}

inline std::string I(long v, int w) { return rjust(std::to_string(v), w); }

// Aw : right-justified when shorter than w, the leftmost w characters otherwise
inline std::string A(const std::string& s, int w) { return (int)s.size() >= w ? s.substr(0, (size_t)w) : rjust(s, w); }

// one record of <name>.cnv.  The reference's '(I7, 4E14.6)' has one slot too few for its six items (SURVEY.md F14:
// gfortran stops with a run-time error at the first print step); this is the evident intent, '(I7, 5E14.6)'.
inline std::string cnv_record(int iter, double time, const double r[4]) {
    std::string s = I(iter, 7) + E(time, 14, 6);
    for (int i = 0; i < 4; ++i) s += E(r[i], 14, 6);
    return s;
}

// One REAL(8) item of a list-directed WRITE as gfortran lays it out: the G25.17E3 editing with a scale factor of 1 that
// libgfortran applies to kind-8 reals -- 17 significant digits in a field of 25: F editing with five trailing blanks when
// 0.1 <= |x| < 1e17 (and for zero), otherwise d.ddddddddddddddddE+ddd.  List-directed output is processor-dependent
// (F2008 10.10.4); this is the layout of the build line the oracle assumes (fortran/README.md), e.g.
//   3.14d0 -> "   3.1400000000000001     "    1.d-3 -> "   1.0000000000000000E-003"    0.d0 -> "   0.0000000000000000     "
inline std::string list_r8(double v) {
    const int w = 25, d = 17;
    std::string s;
    if (std::isnan(v)) return rjust("NaN", w);
    if (std::isinf(v)) return rjust(v < 0 ? "-Infinity" : "Infinity", w);
    const double a = std::fabs(v);
    char buf[128];
    if (a == 0.0) {
        std::snprintf(buf, sizeof buf, "%.*f", d - 1, 0.0);
        s = std::string(std::signbit(v) ? "-" : "") + buf + "     ";
        return rjust(s, w);
    }
    // decimal exponent after rounding to d significant digits
    std::snprintf(buf, sizeof buf, "%.*e", d - 1, a);
    const int ex = std::atoi(std::strchr(buf, 'e') + 1);   // a = m x 10^ex, 1 <= m < 10
    if (ex >= -1 && ex < d) {
        // F editing: d significant digits in all (a leading "0." digit does not count)
        const int decimals = ex >= 0 ? d - 1 - ex : d;
        std::snprintf(buf, sizeof buf, "%.*f", decimals, a);
        s = std::string(v < 0 ? "-" : "") + buf + (decimals == 0 ? "." : "") + "     ";   // Fw.0 keeps its decimal point
    } else {
        std::string m(buf, std::strchr(buf, 'e') - buf);
        char eb[16];
        std::snprintf(eb, sizeof eb, "E%+04d", ex);
        s = std::string(v < 0 ? "-" : "") + m + eb;
    }
    return rjust(s, w);
}

}  // namespace ffmt
