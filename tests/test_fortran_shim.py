"""The Fortran side of the boundary cannot be compiled here (no Fortran compiler).  What can be checked is:

  * include/cfdb.h, fortran/cfdb_iface.f90 and cfd_b200/_abi.py are exactly what the ONE ABI table (tools/gen_abi.py)
    generates -- they cannot drift apart;
  * every dummy argument of every bind(C) interface has the passing mode (value / reference) and the kind of the C
    parameter it stands for, parsed independently from the generated header;
  * the shim file keeps the reference's module / subroutine names and dummy lists for every call site SURVEY.md 8b names;
  * every actual argument list in the shims has the arity of the interface it calls;
  * both files go through the repository's Fortran front end (oracle/f90ref/translate.py: statement splitter, declaration
    and expression parser) without a syntax error.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _c_protos():
    """name -> [(ctype, name)] parsed from the header text (not from the table)"""
    src = open(os.path.join(ROOT, "include", "cfdb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = src[src.index("} cfdb_bc;"):]
    out = {}
    for m in re.finditer(r"([\w\*\s]+?)\b(cfdb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = " ".join(m.group(3).split())
        lst = []
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)(\[\d+\])?$", a)
                lst.append(((mm.group(1).strip() + ("*" if mm.group(3) else "")).replace(" *", "*"), mm.group(2)))
        out[m.group(2)] = (" ".join(m.group(1).split()), lst)
    return out


def _f_interfaces():
    """name -> (kind, [dummy names], {dummy: (type text, has_value, is_array)}) parsed from the Fortran text"""
    f = open(os.path.join(ROOT, "fortran", "cfdb_iface.f90")).read()
    f = re.sub(r"&\s*\n\s*", " ", f)
    body = f[f.index("  interface"):f.index("  end interface")]
    out = {}
    for m in re.finditer(r"(function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)(.*?)end \1", body, flags=re.S):
        kind, fname, args, cname, rest = m.groups()
        assert fname == cname
        names = [a.strip() for a in args.split(",") if a.strip()]
        decl = {}
        for line in rest.split("\n"):
            if "::" not in line:
                continue
            t, vs = line.split("::", 1)
            t = t.strip()
            for v in re.findall(r"(\w+)(\([^)]*\))?", vs):
                decl[v[0]] = (t.replace(", value", "").strip(), ", value" in t, bool(v[1]))
        out[cname] = (kind, names, decl)
    return out


C2F = {"double": "real(c_double)", "int32_t": "integer(c_int32_t)", "int64_t": "integer(c_int64_t)", "uint64_t": "integer(c_int64_t)",
       "int": "integer(c_int)", "unsigned char": "integer(c_int8_t)", "char": "character(kind=c_char)"}


def test_generated_files_match_the_table():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_abi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_every_dummy_matches_its_c_parameter():
    protos, ifs = _c_protos(), _f_interfaces()
    assert set(protos) == set(ifs), set(protos) ^ set(ifs)
    assert len(protos) >= 55
    for name, (ret, cargs) in protos.items():
        kind, fnames, decl = ifs[name]
        assert kind == ("subroutine" if ret == "void" else "function"), name
        assert fnames == [n for _, n in cargs], (name, fnames, cargs)
        for ctype, an in cargs:
            ftype, by_value, is_array = decl[an]
            base = ctype.replace("const ", "").strip()
            if base.endswith("**"):                       # cfdb_ctx** : a c_ptr by reference
                assert (ftype, by_value) == ("type(c_ptr)", False), (name, an)
            elif base in ("cfdb_ctx*", "void*"):          # opaque handles and raw buffers: c_ptr by value
                assert (ftype, by_value) == ("type(c_ptr)", True), (name, an)
            elif base in ("cfdb_params*", "cfdb_bc*"):    # structs by reference
                assert (ftype, by_value) == ("type(%s)" % base[:-1], False), (name, an)
            elif base.endswith("*"):                      # arrays and scalar results: by reference, same kind
                assert not by_value and ftype == C2F[base[:-1]], (name, an, ftype)
                if base[:-1] == "char":
                    assert is_array, (name, an)
            else:                                         # scalars: by value, same kind
                assert by_value and not is_array and ftype == C2F[base], (name, an, ftype)
        if ret != "void":
            assert decl["rc"][0] == {"int": "integer(c_int)", "int64_t": "integer(c_int64_t)", "const char*": "type(c_ptr)",
                                     "void*": "type(c_ptr)"}[ret], name


SHIMS = {   # reference routine (file:line) -> header the shim must carry, C entry it forwards to
    "calcRHS.f90:4": ("subroutine calcRHS(rhs, U, theta, dNx, dNy, area, shoc, dtl, t_sugn1, t_sugn2, t_sugn3, inpoel, nelem, npoin)", "cfdb_calcrhs"),
    "biconjGrad.f90:8": ("subroutine biCG(spMtx, spIdx, spRowptr, diagMtx, x, b, x_fix, x_fixIdx, npoin, nfix)", "cfdb_bicg"),
    "gcl.f90:8": ("subroutine main(M, W_x, W_y, dNx, dNy, area, inpoel, dt)", "cfdb_gcl_main"),
    "subrutinas.f90:7": ("subroutine normales", "cfdb_normales"),
    "subrutinas.f90:66": ("subroutine normalvel", "cfdb_normalvel"),
    "subrutinas.f90:88": ("subroutine deriv(hmin)", "cfdb_deriv"),
    "subrutinas.f90:128": ("subroutine MASAS()", "cfdb_masas"),
    "subrutinas.f90:155": ("subroutine deltat(dtmin, dt)", "cfdb_deltat"),
    "subrutinas.f90:332": ("subroutine ESTAB(U, T, GAMA, FR, RMU, DTMIN, RHOINF, TINF, UINF, VINF, GAMM)", "cfdb_estab"),
    "subrutinas.f90:601": ("subroutine fixvel", "cfdb_fixvel"),
    "subrutinas.f90:618": ("subroutine FIX(FR, GAMM)", "cfdb_fix"),
    "subrutinas.f90:645": ("subroutine RK(DTMIN, NRK, BANDERA, GAMM, dtl)", "cfdb_rk"),
    "subrutinas.f90:1036": ("subroutine FUENTE(dtl)", "cfdb_fuente"),
    "mLaplace.f90:7": ("subroutine laplace(inpoel, area, dNx, dNy, nelem, npoin)", "cfdb_laplace"),
    "meshMove.f90:28": ("subroutine fluidStructure(dtmin, time, SMOOTH_FIX, x1, y1)", "cfdb_mesh_move"),
    "smoothing.f90:21": ("subroutine smoothing(X, Y, inpoel, fixed, npoin0, nelem0)", "cfdb_smoothing"),
}


def test_shims_keep_the_reference_names_and_call_the_abi_with_the_right_arity():
    shim = open(os.path.join(ROOT, "fortran", "calcRHS_gpu.f90")).read()
    flat = re.sub(r"&\s*\n\s*", " ", shim)
    protos = _c_protos()
    for where, (header, centry) in SHIMS.items():
        assert header in shim, (where, header)
    for mod in ("module calcRHS_mod", "module BiconjGrad", "module gcl_mod", "module Mnormales", "module Mlaplace", "module MeshMove",
                "module smoothing_mod"):
        assert mod in shim, mod
    seen = set()
    for m in re.finditer(r"\b(cfdb_[a-z0-9_]+)\s*\(", flat):
        name = m.group(1)
        if name in ("cfdb_check",) or name not in protos:
            continue
        # argument list up to the matching parenthesis
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(flat[i], 0)
            i += 1
        args, d, n = flat[m.end():i - 1], 0, 1
        for ch in args:
            d += {"(": 1, ")": -1}.get(ch, 0)
            n += ch == "," and d == 0
        assert n == len(protos[name][1]), (name, n, len(protos[name][1]))
        seen.add(name)
    assert {c for _, c in SHIMS.values()} <= seen


def test_both_files_pass_the_fortran_front_end():
    """statement splitter (continuations, comments, strings), declaration parser and expression parser of the repository's
    Fortran-subset front end accept every statement of the two files"""
    from oracle.f90ref import translate

    n = 0
    for fn in ("cfdb_iface.f90", "calcRHS_gpu.f90"):
        text = open(os.path.join(ROOT, "fortran", fn)).read()
        stmts = translate.logical_statements(text)
        assert stmts, fn
        depth = 0
        for st in stmts:
            s = st[1]
            low = s.strip().lower()
            n += 1
            # block structure: every module / subroutine / function / interface / type / if-then / do closes
            if re.match(r"(module|subroutine|function|interface|type\b(?!\s*\())", low) and not low.startswith("module procedure"):
                depth += 1
            elif re.match(r"end\s*(module|subroutine|function|interface|type)\b", low):
                depth -= 1
            assert depth >= 0, (fn, s)
            # expressions on the right of assignments and inside call argument lists must parse
            m = re.match(r"call\s+\w+\s*\((.*)\)\s*$", s.strip(), flags=re.I | re.S)
            if m:
                for a in translate._split_top(m.group(1), ","):
                    translate.parse_expr(a.strip())
            elif "=" in s and "::" not in s and not low.startswith(("if", "use", "function", "subroutine")) and "=>" not in s:
                lhs, rhs = translate._split_top_assign(s)
                if rhs is not None:
                    translate.parse_expr(rhs.strip())
        assert depth == 0, fn
    assert n > 300
