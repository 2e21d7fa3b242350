"""fortran/cfdb_iface.f90 cannot be compiled here (no Fortran compiler): keep its bind(C) names and argument
counts in step with include/cfdb.h."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_protos():
    src = open(os.path.join(ROOT, "include", "cfdb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(cfdb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_bind_c_names_and_arity_match_header():
    protos = _c_protos()
    f = open(os.path.join(ROOT, "fortran", "cfdb_iface.f90")).read()
    f = re.sub(r"&\s*\n\s*", " ", f)
    found = 0
    for m in re.finditer(r"(?:function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)", f):
        fname, args, cname = m.groups()
        assert fname == cname and cname in protos, cname
        n = 0 if not args.strip() else args.count(",") + 1
        assert n == protos[cname], (cname, n, protos[cname])
        found += 1
    assert found >= 14
    shim = open(os.path.join(ROOT, "fortran", "calcRHS_gpu.f90")).read()
    for name in ("module calcRHS_mod", "module BiconjGrad", "module gcl_mod", "subroutine deriv(hmin)", "subroutine MASAS()",
                 "subroutine deltat(dtmin, dt)", "subroutine ESTAB(U, T, GAMA, FR, RMU, DTMIN, RHOINF, TINF, UINF, VINF, GAMM)",
                 "subroutine FUENTE(dtl)"):
        assert name in shim, name
