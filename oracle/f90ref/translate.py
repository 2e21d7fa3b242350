"""Fortran-90 subset -> Python translator, used to EXECUTE the reference's own sources (read from /root/reference at
test/golden-generation time; never copied into the repo) so that the hand-written C++ oracle can be compared bit for bit
with what the reference's statements compute.  TEST INFRASTRUCTURE ONLY.

No Fortran compiler exists in the image, so this is the closest thing to "the reference itself, run here":
every assignment, loop, call and expression of the reference is parsed and evaluated as written -- operator precedence and
left-to-right association per the Fortran standard, implicit typing, array sections, vector subscripts, module variables,
SAVE'd locals, optional arguments, sequence association of explicit-shape dummies, `-fpp` function-like macros.
OpenMP directives are comments (one thread: the only well-defined order).  The arithmetic model is in
runtime.py.  Unsupported statements become `raise Unsupported` at the point of execution, so files can be loaded whole.
"""
from __future__ import annotations

import re

from . import runtime as rt

TYPE_KW = ("integer", "real", "doubleprecision", "double precision", "logical", "character", "complex")
PYKW = {"lambda", "in", "is", "or", "and", "not", "if", "else", "for", "while", "def", "class", "pass", "del", "from",
        "import", "as", "with", "try", "global", "return", "yield", "assert", "raise", "print", "exec", "None", "True", "False"}


# --------------------------------------------------------------------------------------------------------------------
# source -> logical statements
# --------------------------------------------------------------------------------------------------------------------
def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out)


def _lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        else:
            out.append(ch.lower())
    return "".join(out)


def _split_top(s, sep):
    """split on sep at parenthesis depth 0, outside strings"""
    parts, depth, q, cur = [], 0, None, []
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == "(":
            depth += 1
            cur.append(ch)
        elif ch == ")":
            depth -= 1
            cur.append(ch)
        elif ch == sep and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
        i += 1
    parts.append("".join(cur))
    return parts


def _match_paren(s, i):
    """s[i] == '(' -> index of the matching ')'"""
    depth, q = 0, None
    for j in range(i, len(s)):
        ch = s[j]
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return j
    raise SyntaxError("unbalanced parentheses: " + s)


def _expand_macros(line, macros):
    for name, (params, body) in macros.items():
        pos = 0
        while True:
            m = re.search(r"\b" + re.escape(name) + r"\s*\(", line[pos:])
            if not m:
                break
            start = pos + m.start()
            lp = pos + m.end() - 1
            rp = _match_paren(line, lp)
            args = [a.strip() for a in _split_top(line[lp + 1:rp], ",")]
            rep = body
            for p, a in zip(params, args):
                rep = re.sub(r"\b" + re.escape(p) + r"\b", a.replace("\\", "\\\\"), rep)
            line = line[:start] + rep + line[rp + 1:]
            pos = start + len(rep)
    return line


def logical_statements(text):
    """-> list of (label or None, lowered statement text, first line number)"""
    macros, stmts = {}, []
    cur, cur_line = "", 0
    for ln, raw in enumerate(text.split("\n"), 1):
        if raw.startswith("#"):
            m = re.match(r"#define\s+(\w+)\(([^)]*)\)\s+(.*)", raw)
            if m:
                macros[m.group(1)] = ([p.strip() for p in m.group(2).split(",")], m.group(3))
            continue
        line = _strip_comment(raw.replace("\t", " ")).rstrip()
        if macros:
            line = _expand_macros(line, macros)
        if not line.strip():
            continue
        body = line.strip()
        if cur:
            if body.startswith("&"):
                body = body[1:]
            cur += " " + body
        else:
            cur, cur_line = body, ln
        if cur.endswith("&"):
            cur = cur[:-1]
            continue
        for piece in _split_top(cur, ";"):
            piece = piece.strip()
            if piece:
                low = _lower_outside_strings(piece)
                m = re.match(r"^(\d+)\s+(.*)$", low)
                stmts.append((int(m.group(1)), m.group(2), cur_line) if m else (None, low, cur_line))
        cur = ""
    return stmts


# --------------------------------------------------------------------------------------------------------------------
# expressions
# --------------------------------------------------------------------------------------------------------------------
_TOK = re.compile(r"""\s*(?:
    (?P<dotop>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.) |
    (?P<num>(?:\d+\.(?![a-z]+\.)\d*|\.\d+|\d+)(?:[de][+-]?\d+)?(?:_\w+)?) |
    (?P<name>[a-z_]\w*) |
    (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*") |
    (?P<op>\*\*|//|==|/=|<=|>=|=>|\(/|/\)|[-+*/(),:<>=%])
)""", re.X)


def tokenize(s):
    toks, pos = [], 0
    s = s.rstrip()
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m or m.end() == pos:
            raise SyntaxError(f"cannot tokenize {s[pos:]!r} in {s!r}")
        pos = m.end()
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
    # "(/" directly after a name or ")" is "(" "/" -- not an array constructor; in practice never occurs
    return toks


_REL = {".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=",
        "==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">="}


class ExprParser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val:
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise SyntaxError(f"expected {val!r}, got {self.peek()} in {self.t}")

    def done(self):
        return self.i >= len(self.t)

    # precedence climbing, lowest first
    def expr(self):
        a = self.or_()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.next()[1]
            a = ("bin", op, a, self.or_())
        return a

    def or_(self):
        a = self.and_()
        while self.accept(".or."):
            a = ("bin", ".or.", a, self.and_())
        return a

    def and_(self):
        a = self.not_()
        while self.accept(".and."):
            a = ("bin", ".and.", a, self.not_())
        return a

    def not_(self):
        if self.accept(".not."):
            return ("un", ".not.", self.not_())
        return self.rel()

    def rel(self):
        a = self.concat()
        if self.peek()[1] in _REL and self.peek()[0] in ("op", "dotop"):
            op = self.next()[1]
            a = ("bin", _REL[op], a, self.concat())
        return a

    def concat(self):
        a = self.add()
        while self.accept("//"):
            a = ("bin", "//", a, self.add())
        return a

    def add(self):
        if self.peek()[1] in ("+", "-") and self.peek()[0] == "op":
            op = self.next()[1]
            a = self.mul()
            a = ("un", op, a) if op == "-" else a
        else:
            a = self.mul()
        while self.peek()[1] in ("+", "-") and self.peek()[0] == "op":
            op = self.next()[1]
            a = ("bin", op, a, self.mul())
        return a

    def mul(self):
        a = self.pow()
        while self.peek()[1] in ("*", "/") and self.peek()[0] == "op":
            op = self.next()[1]
            a = ("bin", op, a, self.pow())
        return a

    def pow(self):
        a = self.primary()
        if self.accept("**"):
            # right associative; the exponent may carry a sign only inside parentheses (standard), but accept -x too
            if self.peek()[1] in ("+", "-"):
                op = self.next()[1]
                b = self.pow()
                b = ("un", "-", b) if op == "-" else b
            else:
                b = self.pow()
            a = ("bin", "**", a, b)
        return a

    def primary(self):
        kind, val = self.next()
        if kind == "num":
            return ("num", val)
        if kind == "str":
            q = val[0]
            return ("str", val[1:-1].replace(q + q, q))
        if kind == "dotop" and val in (".true.", ".false."):
            return ("log", val == ".true.")
        if val == "(":
            e = self.expr()
            self.expect(")")
            return ("par", e)
        if val == "(/":
            items = []
            if not self.accept("/)"):
                while True:
                    items.append(self.expr())
                    if self.accept("/)"):
                        break
                    self.expect(",")
            return ("ac", items)
        if kind == "name":
            if self.accept("("):
                args = []
                if not self.accept(")"):
                    while True:
                        args.append(self.arg())
                        if self.accept(")"):
                            break
                        self.expect(",")
                return ("ref", val, args)
            return ("name", val)
        raise SyntaxError(f"unexpected token {val!r} in {self.t}")

    def arg(self):
        # keyword argument?
        if self.peek()[0] == "name" and self.i + 1 < len(self.t) and self.t[self.i + 1][1] == "=":
            name = self.next()[1]
            self.next()
            return ("kw", name, self.expr())
        lo = hi = st = None
        if self.peek()[1] != ":":
            lo = self.expr()
            if self.peek()[1] != ":":
                return lo
        self.expect(":")
        if self.peek()[1] not in (",", ")", ":"):
            hi = self.expr()
        if self.accept(":"):
            st = self.expr()
        return ("slice", lo, hi, st)


def parse_expr(s):
    p = ExprParser(tokenize(s))
    e = p.expr()
    if not p.done():
        raise SyntaxError(f"trailing tokens in expression {s!r}")
    return e


# --------------------------------------------------------------------------------------------------------------------
# program structure
# --------------------------------------------------------------------------------------------------------------------
class Var:
    def __init__(self, name, ftype, dims=None, **attrs):
        self.name, self.ftype, self.dims = name, ftype, dims   # ftype: 'i','r4','r8','l','c'; dims: list of (lb, ub) ASTs / None
        self.allocatable = attrs.get("allocatable", False)
        self.parameter = attrs.get("parameter", False)
        self.save = attrs.get("save", False)
        self.optional = attrs.get("optional", False)
        self.init = attrs.get("init")
        self.dummy = False

    @property
    def rank(self):
        return len(self.dims) if self.dims else 0


class Unit:
    """module or procedure"""

    def __init__(self, kind, name, args=None, parent=None, prefix_type=None, result=None):
        self.kind, self.name, self.args, self.parent = kind, name, args or [], parent
        self.prefix_type, self.result = prefix_type, result or name
        self.stmts = []          # (label, text, line)
        self.vars = {}
        self.implicit = {c: ("i" if "i" <= c <= "n" else "r4") for c in "abcdefghijklmnopqrstuvwxyz"}
        self.uses = []           # (module, only: bool, {local: remote})
        self.procs = {}
        self.file = None
        self.default_private = False
        self.access = {}

    def is_public(self, name):
        return self.access.get(name, "private" if self.default_private else "public") == "public"

    @property
    def pyname(self):
        return f"p_{self.parent.name if self.parent else ''}__{self.name}"


_DTYPE = {"i": "np.int32", "r4": "np.float32", "r8": "np.float64", "l": "np.bool_", "c": "object"}
_CONV = {"i": "rt.i4", "r4": "rt.f4", "r8": "rt.f8", "l": "rt.lg", "c": "str"}
_ZERO = {"i": "0", "r4": "rt.f4(0)", "r8": "rt.f8(0)", "l": "False", "c": "''"}

INTRINSICS = {
    "abs": "rt.abs_", "dabs": "rt.abs_", "sqrt": "rt.sqrt_", "dsqrt": "rt.sqrt_", "min": "rt.min_", "max": "rt.max_",
    "dmin1": "rt.min_", "dmax1": "rt.max_", "sum": "rt.sum_", "minval": "rt.minval", "maxval": "rt.maxval",
    "sin": "rt.sin_", "dsin": "rt.sin_", "cos": "rt.cos_", "dcos": "rt.cos_", "exp": "rt.exp_", "dexp": "rt.exp_",
    "acos": "rt.acos_", "dacos": "rt.acos_", "atan": "rt.atan_", "datan": "rt.atan_", "log": "rt.log_", "dlog": "rt.log_",
    "size": "rt.size_", "real": "rt.real_", "float": "rt.real_", "dble": "rt.dble_", "int": "rt.int_", "mod": "rt.mod_",
    "any": "rt.any_", "all": "rt.all_", "count": "rt.count_", "dfloat": "rt.dble_", "dlog10": "rt.log10_", "log10": "rt.log10_",
    "sign": "rt.sign_", "dsign": "rt.sign_", "tiny": "rt.tiny_", "huge": "rt.huge_", "trim": "rt.trim_", "reshape": "rt.reshape",
}


def _parse_type_prefix(s):
    """statement starting with a type keyword -> (ftype, rest) or None"""
    m = re.match(r"^(double\s*precision|integer|real|logical|character|complex)\b\s*", s)
    if not m:
        return None
    kw, rest = m.group(1), s[m.end():]
    kind = None
    if rest.startswith("*"):
        m2 = re.match(r"\*\s*(\d+|\(\s*\*\s*\))\s*", rest)
        kind, rest = m2.group(1), rest[m2.end():]
    elif rest.startswith("("):
        j = _match_paren(rest, 0)
        inner = rest[1:j].replace(" ", "")
        # "real(x)" used as an expression statement cannot start a statement, so this is a kind/len selector
        kind = inner.split("=")[-1]
        rest = rest[j + 1:].lstrip()
    if kw.startswith("double"):
        ft = "r8"
    elif kw == "integer":
        ft = "i"
    elif kw == "real":
        ft = "r8" if kind == "8" else "r4"
    elif kw == "logical":
        ft = "l"
    elif kw == "character":
        ft = "c"
    else:
        raise rt.Unsupported("complex")
    return ft, rest


class Program:
    def __init__(self):
        self.modules = {}
        self.externals = {}
        self.consts = {}
        self.lines = []

    # ---------------------------------------------------------------- loading
    def load(self, path):
        with open(path) as f:
            text = f.read()
        stmts = logical_statements(text)
        i, cur_mod, cur = 0, None, None
        stack = []
        for label, s, ln in stmts:
            head = s.split("(")[0].strip()
            m_end = re.match(r"^end\s*(module|subroutine|function|program)?\b", s)
            if m_end and (m_end.group(1) or s.strip() == "end") and (cur is not None or cur_mod is not None):
                if cur is not None:
                    cur = None
                else:
                    cur_mod = None
                continue
            if cur is None:
                m = re.match(r"^module\s+(\w+)$", s)
                if m and not s.startswith("module procedure"):
                    cur_mod = Unit("module", m.group(1))
                    cur_mod.file = path
                    self.modules[cur_mod.name] = cur_mod
                    continue
                m = re.match(r"^program\s+(\w+)$", s)
                if m:
                    cur = Unit("program", m.group(1), parent=cur_mod)
                    cur.file = path
                    self.externals[cur.name] = cur
                    continue
                m = re.match(r"^(?:(?:pure|elemental|recursive)\s+)*(.*?)\b(subroutine|function)\s+(\w+)\s*(\(.*?\))?\s*(?:result\s*\((\w+)\))?$", s)
                if m and (m.group(1).strip() == "" or _parse_type_prefix(m.group(1).strip())):
                    pre = m.group(1).strip()
                    args = [a.strip() for a in m.group(4)[1:-1].split(",")] if m.group(4) and m.group(4)[1:-1].strip() else []
                    cur = Unit(m.group(2), m.group(3), args, parent=cur_mod,
                               prefix_type=_parse_type_prefix(pre)[0] if pre else None, result=m.group(5))
                    cur.file = path
                    (cur_mod.procs if cur_mod else self.externals)[cur.name] = cur
                    continue
                if s == "contains":
                    continue
                if cur_mod is not None:
                    cur_mod.stmts.append((label, s, ln))
                    continue
                raise SyntaxError(f"{path}:{ln}: statement outside a program unit: {s}")
            cur.stmts.append((label, s, ln))

    # ---------------------------------------------------------------- declarations
    def _declare(self, unit, s):
        """process a specification statement; returns True if it was one"""
        if s == "implicit none":
            unit.implicit = None
            return True
        m = re.match(r"^implicit\s+(.*)$", s)
        if m:
            ft, rest = _parse_type_prefix(m.group(1))
            for rng in rest.strip()[1:-1].split(","):
                a, _, b = rng.strip().partition("-")
                for c in range(ord(a.strip()), ord((b or a).strip()) + 1):
                    unit.implicit[chr(c)] = ft
            return True
        m = re.match(r"^use\s+(\w+)\s*(?:,\s*(only\s*:)?\s*(.*))?$", s)
        if m:
            ren = {}
            if m.group(3):
                for item in m.group(3).split(","):
                    item = item.strip()
                    if "=>" in item:
                        loc, rem = [x.strip() for x in item.split("=>")]
                    else:
                        loc = rem = item
                    if loc:
                        ren[loc] = rem
            unit.uses.append((m.group(1), bool(m.group(2)), ren))
            return True
        m = re.match(r"^(private|public)\b\s*(?:::)?\s*(.*)$", s)
        if m:
            names = [n.strip() for n in m.group(2).split(",") if n.strip()]
            if not names:
                unit.default_private = m.group(1) == "private"
            for n in names:
                unit.access[n] = m.group(1)
            return True
        if re.match(r"^(save|intrinsic|external)\b", s):
            return True
        tp = _parse_type_prefix(s)
        if tp is None:
            return False
        ft, rest = tp
        attrs, dims_attr = {}, None
        if "::" in rest:
            astr, ents = rest.split("::", 1)
            for a in _split_top(astr, ","):
                a = a.strip()
                if not a:
                    continue
                if a.startswith("dimension"):
                    dims_attr = a[a.index("(") + 1:_match_paren(a, a.index("("))]
                elif a in ("allocatable", "parameter", "save", "optional"):
                    attrs[a] = True
                elif a in ("private", "public"):
                    attrs["access"] = a
        else:
            ents = rest
        for ent in _split_top(ents, ","):
            ent = ent.strip()
            if not ent:
                continue
            init = None
            parts = _split_top_assign(ent)
            if parts:
                ent, init = parts[0].strip(), parse_expr(parts[1])
            m = re.match(r"^(\w+)\s*(\(.*\))?\s*(\*\s*\d+)?$", ent)
            if not m:
                raise SyntaxError(f"cannot parse entity {ent!r} in {s!r}")
            name = m.group(1)
            dstr = m.group(2)[1:-1] if m.group(2) else dims_attr
            dims = None
            if dstr is not None:
                dims = []
                for d in _split_top(dstr, ","):
                    d = d.strip()
                    if ":" in d:
                        lo, hi = [x.strip() for x in d.split(":", 1)]
                        dims.append((parse_expr(lo) if lo else None, parse_expr(hi) if hi else None))
                    else:
                        dims.append((None, parse_expr(d)))
            v = Var(name, ft, dims, init=init, **attrs)
            if "access" in attrs:
                unit.access[name] = attrs["access"]
            if init is not None and not v.parameter:
                v.save = True
            old = unit.vars.get(name)
            if old is not None and old.dummy:
                v.dummy = True
            unit.vars[name] = v
        return True

    # ---------------------------------------------------------------- name resolution
    def lookup(self, unit, name, declare_implicit=True):
        """-> ('local', Var) | ('module', modname, Var) | ('proc', Unit) | None"""
        if name in unit.vars:
            return ("local", unit.vars[name])
        if unit.kind in ("function",) and name == unit.result:
            return ("local", self._implicit_var(unit, name))
        for modname, only, ren in unit.uses:
            mod = self.modules.get(modname)
            if mod is None:
                continue
            if name in ren:
                r = self._module_entity(mod, ren[name])
                if r:
                    return r
            elif not only:
                r = self._module_entity(mod, name)
                if r:
                    return r
        if unit.parent is not None:
            r = self._module_entity(unit.parent, name, own=True)
            if r:
                return r
        if name in self.externals:
            return ("proc", self.externals[name])
        return None

    def _module_entity(self, mod, name, own=False):
        self._prepare_module(mod)
        if not own and (name in mod.vars or name in mod.procs) and not mod.is_public(name):
            return None
        if name in mod.vars:
            return ("module", mod.name, mod.vars[name])
        if name in mod.procs:
            return ("proc", mod.procs[name])
        # entities a module re-exports through its own module-level USE
        for modname, only, ren in mod.uses:
            m2 = self.modules.get(modname)
            if m2 is None or m2 is mod:
                continue
            if name in ren:
                return self._module_entity(m2, ren[name])
            if not only:
                r = self._module_entity(m2, name)
                if r:
                    return r
        return None

    def _prepare_module(self, mod):
        if getattr(mod, "_prepared", False):
            return
        mod._prepared = True
        for label, s, ln in mod.stmts:
            if not self._declare(mod, s):
                raise SyntaxError(f"{mod.file}:{ln}: executable statement in module specification part: {s}")

    def _implicit_var(self, unit, name):
        if name in unit.vars:
            return unit.vars[name]
        if unit.kind == "function" and name == unit.result and unit.prefix_type:
            ft = unit.prefix_type
        elif unit.implicit is None:
            raise NameError(f"{unit.file}: {unit.name}: '{name}' has no type (implicit none)")
        else:
            ft = unit.implicit[name[0]]
        v = Var(name, ft)
        unit.vars[name] = v
        return v

    # ---------------------------------------------------------------- code generation: expressions
    def const(self, text):
        t = text.lower()
        if re.fullmatch(r"\d+", t):
            return t
        t = t.split("_")[0]
        if "d" in t:
            code = f"rt.f8('{t.replace('d', 'e')}')"
        else:
            code = f"rt.f4('{t}')"
        if code not in self.consts:
            self.consts[code] = f"K{len(self.consts)}"
        return self.consts[code]

    def varcode(self, unit, name):
        r = self.lookup(unit, name)
        if r is None:
            r = ("local", self._implicit_var(unit, name))
        if r[0] == "local":
            v = r[1]
            if v.save and not v.dummy:
                return f"SV.{unit.pyname}__{name}", v
            return f"v_{name}", v
        if r[0] == "module":
            return f"M_{r[1]}.v_{r[2].name}", r[2]
        raise NameError(f"{name} is a procedure, not a variable")

    def _lb(self, v, k):
        """static lower bound of dimension k (default 1)"""
        lo = v.dims[k][0]
        if lo is None:
            return 1
        if lo[0] == "num" and re.fullmatch(r"\d+", lo[1]):
            return int(lo[1])
        if lo[0] == "un" and lo[1] == "-" and lo[2][0] == "num":
            return -int(lo[2][1])
        raise rt.Unsupported(f"non-constant lower bound of {v.name}")

    def subscripts(self, unit, v, args):
        if len(args) != v.rank:
            raise SyntaxError(f"{v.name}: {len(args)} subscripts for rank {v.rank}")
        out = []
        for k, a in enumerate(args):
            lb = self._lb(v, k)
            if a[0] == "slice":
                lo = "" if a[1] is None else self._minus(self.ex(unit, a[1]), lb)
                hi = "" if a[2] is None else self._minus(self.ex(unit, a[2]), lb - 1)
                if a[3] is not None:
                    out.append(f"{lo}:{hi}:{self.ex(unit, a[3])}")
                else:
                    out.append(f"{lo}:{hi}")
            else:
                out.append(self._minus(self.ex(unit, a), lb))
        return ", ".join(out)

    @staticmethod
    def _minus(code, k):
        if k == 0:
            return code
        if re.fullmatch(r"\d+", code):
            return str(int(code) - k)
        return f"{code}-{k}" if k > 0 else f"{code}+{-k}"

    def ex(self, unit, e):
        k = e[0]
        if k == "num":
            return self.const(e[1])
        if k == "str":
            return repr(e[1])
        if k == "log":
            return "True" if e[1] else "False"
        if k == "par":
            return f"({self.ex(unit, e[1])})"
        if k == "name":
            r = self.lookup(unit, e[1])
            if r and r[0] == "proc":
                return f"{r[1].pyname}()"
            return self.varcode(unit, e[1])[0]
        if k == "un":
            return f"(-{self.ex(unit, e[2])})" if e[1] == "-" else f"(not {self.ex(unit, e[2])})"
        if k == "ac":
            return "rt.ac([" + ", ".join(self.ex(unit, x) for x in e[1]) + "])"
        if k == "bin":
            op, a, b = e[1], self.ex(unit, e[2]), self.ex(unit, e[3])
            if op in ("+", "-", "*"):
                return f"({a} {op} {b})"
            if op == "/":
                if _is_real_literal(e[2]) or _is_real_literal(e[3]):
                    return f"({a} / {b})"
                return f"rt.div({a}, {b})"
            if op == "**":
                return f"rt.pow_({a}, {b})"
            if op in ("==", "!=", "<", "<=", ">", ">="):
                return f"({a} {op} {b})"
            if op == ".and.":
                return f"({a} and {b})"
            if op == ".or.":
                return f"({a} or {b})"
            if op == ".eqv.":
                return f"(bool({a}) == bool({b}))"
            if op == ".neqv.":
                return f"(bool({a}) != bool({b}))"
            if op == "//":
                return f"({a} + {b})"
        if k == "ref":
            name, args = e[1], e[2]
            r = self.lookup(unit, name)
            if r is None and name not in INTRINSICS and name not in ("allocated", "present"):
                # undeclared + subscripted: an external function
                return f"p___{name}(" + ", ".join(self.actual(unit, a) for a in args) + ")"
            if r is not None and r[0] != "proc" and r[1 if r[0] == "local" else 2].rank > 0:
                code, v = self.varcode(unit, name)
                return f"{code}[{self.subscripts(unit, v, args)}]"
            if r is not None and r[0] == "local" and r[1].rank == 0 and name in self.externals:
                r = ("proc", self.externals[name])   # "real(8) f" declaring the type of an external function
            if r is not None and r[0] == "proc":
                return f"{r[1].pyname}(" + ", ".join(self.actual(unit, a) for a in args) + ")"
            if name == "allocated":
                return f"({self.ex(unit, args[0])} is not None)"
            if name == "present":
                return f"({self.ex(unit, args[0])} is not None)"
            if name in INTRINSICS:
                return f"{INTRINSICS[name]}(" + ", ".join(self.actual(unit, a) for a in args) + ")"
            if r is not None and r[0] in ("local", "module") and (r[1] if r[0] == "local" else r[2]).ftype == "c":
                raise rt.Unsupported("character substring")
            raise NameError(f"{unit.name}: cannot resolve {name}(...)")
        raise SyntaxError(f"bad expression node {e}")

    def actual(self, unit, a):
        if a[0] == "kw":
            return f"v_{a[1]}={self.ex(unit, a[2])}" if a[1] not in ("dim", "kind") else f"{a[1]}={self.ex(unit, a[2])}"
        return self.ex(unit, a)

    # ---------------------------------------------------------------- code generation: statements
    def lhs(self, unit, e):
        """-> (python target, Var, is_whole_array)"""
        if e[0] == "name":
            code, v = self.varcode(unit, e[1])
            return code, v, v.rank > 0
        if e[0] == "ref":
            code, v = self.varcode(unit, e[1])
            if v.rank == 0:
                raise SyntaxError(f"{e[1]} is not an array")
            return f"{code}[{self.subscripts(unit, v, e[2])}]", v, None
        raise SyntaxError(f"bad assignment target {e}")

    def assign(self, unit, target_ast, value_code):
        code, v, whole = self.lhs(unit, target_ast)
        if whole is True:
            return f"{code}[...] = {value_code}"
        if whole is None:
            return f"{code} = {value_code}"
        return f"{code} = {_CONV[v.ftype]}({value_code})"

    def gen_proc(self, unit):
        for a in unit.args:
            v = Var(a, None)
            v.dummy = True
            unit.vars[a] = v
        body_stmts = []
        for label, s, ln in unit.stmts:
            try:
                if label is None and self._declare(unit, s):
                    continue
            except rt.Unsupported as ex:
                raise rt.Unsupported(f"{unit.file}:{ln}: {ex}")
            body_stmts.append((label, s, ln))
        for a in unit.args:   # dummies never given a type get the implicit one
            v = unit.vars[a]
            v.dummy = True
            if v.ftype is None:
                if unit.implicit is None:
                    raise NameError(f"{unit.name}: dummy {a} has no type")
                v.ftype = unit.implicit[a[0]]
        self._body, self._pos, self._unit = body_stmts, 0, unit
        self._tmp = 0
        body = self.block(unit, 1, terminators=())
        # prologue
        pro = []
        for a in unit.args:
            v = unit.vars[a]
            if v.rank and all(d[1] is not None for d in v.dims):
                pro.append(f"    v_{a} = rt.shape_(v_{a}, ({self._extents(unit, v)}))")
        statics = []
        for name, v in unit.vars.items():
            if v.dummy:
                continue
            if v.save:
                tgt = f"SV.{unit.pyname}__{name}"
                if v.rank and v.allocatable:
                    statics.append(f"{tgt} = None")
                elif v.rank:
                    raise rt.Unsupported("saved fixed-size array")
                else:
                    statics.append(f"{tgt} = {_CONV[v.ftype]}({self.ex(unit, v.init)})" if v.init is not None else f"{tgt} = {_ZERO[v.ftype]}")
                continue
            if v.parameter:
                if v.rank:
                    pro.append(f"    v_{name} = rt.alloc({_DTYPE[v.ftype]}, ({self._extents(unit, v)}))")
                    pro.append(f"    v_{name}[...] = {self.ex(unit, v.init)}")
                else:
                    pro.append(f"    v_{name} = {_CONV[v.ftype]}({self.ex(unit, v.init)})")
            elif v.rank and v.allocatable:
                pro.append(f"    v_{name} = None")
            elif v.rank:
                pro.append(f"    v_{name} = rt.alloc({_DTYPE[v.ftype]}, ({self._extents(unit, v)}))")
            else:
                pro.append(f"    v_{name} = {_ZERO[v.ftype]}")
        ret = self._ret(unit)
        sig = ", ".join(f"v_{a}=None" for a in unit.args)
        out = [f"def {unit.pyname}({sig}{', ' if sig else ''}*_extra):"] + pro + body + [f"    {ret}", ""]
        return statics, out

    def _extents(self, unit, v):
        ext = []
        for lo, hi in v.dims:
            if lo is None:
                ext.append(self.ex(unit, hi))
            else:
                ext.append(f"({self.ex(unit, hi)}) - ({self.ex(unit, lo)}) + 1")
        return ", ".join(ext) + ","

    def _ret(self, unit):
        if unit.kind == "function":
            return f"return v_{unit.result}"
        return "return (" + "".join(f"v_{a}, " for a in unit.args) + ")"

    def block(self, unit, ind, terminators, until_label=None):
        """translate statements until one whose leading keyword is in terminators (left unconsumed)"""
        out = []
        pad = "    " * ind
        while self._pos < len(self._body):
            label, s, ln = self._body[self._pos]
            if until_label is not None and label == until_label:
                break
            if _leading(s) in terminators:
                break
            self._pos += 1
            try:
                out.extend(self.stmt(unit, ind, label, s, ln))
            except rt.Unsupported as ex:
                out.append(f"{pad}raise rt.Unsupported({f'{unit.file}:{ln}: {ex}: {s}'!r})")
            except (SyntaxError, NameError) as ex:
                raise type(ex)(f"{unit.file}:{ln}: {s}\n   {ex}")
        if not out:
            out.append(pad + "pass")
        return out

    def stmt(self, unit, ind, label, s, ln):
        pad = "    " * ind
        lead = _leading(s)
        if lead == "if":
            j = _match_paren(s, s.index("("))
            cond = self.ex(unit, parse_expr(s[s.index("(") + 1:j]))
            rest = s[j + 1:].strip()
            if rest == "then":
                out = [f"{pad}if {cond}:"] + self.block(unit, ind + 1, ("elseif", "else", "endif"))
                while True:
                    _, s2, ln2 = self._body[self._pos]
                    self._pos += 1
                    l2 = _leading(s2)
                    if l2 == "endif":
                        break
                    if l2 == "elseif":
                        j2 = _match_paren(s2, s2.index("("))
                        c2 = self.ex(unit, parse_expr(s2[s2.index("(") + 1:j2]))
                        out += [f"{pad}elif {c2}:"] + self.block(unit, ind + 1, ("elseif", "else", "endif"))
                    else:
                        out += [f"{pad}else:"] + self.block(unit, ind + 1, ("endif",))
                return out
            m = re.match(r"^go\s*to\s+(\d+)$", rest)
            if m:   # forward jump over the following statements of this block
                body = self.block(unit, ind + 1, (), until_label=int(m.group(1)))
                return [f"{pad}if not ({cond}):"] + body
            return [f"{pad}if {cond}:"] + self.stmt(unit, ind + 1, None, rest, ln)
        if lead == "do":
            m = re.match(r"^do\s+while\s*\((.*)\)$", s)
            if m:
                out = [f"{pad}while {self.ex(unit, parse_expr(m.group(1)))}:"] + self.block(unit, ind + 1, ("enddo",))
                self._pos += 1
                return out
            if s == "do":
                out = [f"{pad}while True:"] + self.block(unit, ind + 1, ("enddo",))
                self._pos += 1
                return out
            m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", s)
            if not m:
                raise rt.Unsupported("labelled do")
            parts = [self.ex(unit, parse_expr(p)) for p in _split_top(m.group(2), ",")]
            var, v = self.varcode(unit, m.group(1))
            out = [f"{pad}for {var} in rt.dorange({', '.join(parts)}):"] + self.block(unit, ind + 1, ("enddo",))
            out += [f"{pad}else:", f"{pad}    {var} = rt.doend({', '.join(parts)})"]
            self._pos += 1
            return out
        if lead == "forall":
            m = re.match(r"^forall\s*\(\s*(\w+)\s*=\s*(.*)\)$", s)
            lo, hi = _split_top(m.group(2), ":")
            var, v = self.varcode(unit, m.group(1))
            out = [f"{pad}for {var} in rt.dorange({self.ex(unit, parse_expr(lo))}, {self.ex(unit, parse_expr(hi))}):"]
            out += self.block(unit, ind + 1, ("endforall",))
            self._pos += 1
            return out
        if lead == "call":
            return self.call(unit, ind, s[4:].strip())
        if lead == "return":
            return [pad + self._ret(unit)]
        if lead == "stop":
            return [f"{pad}raise rt.FortranStop({s!r})"]
        if lead == "exit":
            return [pad + "break"]
        if lead == "cycle":
            return [pad + "continue"]
        if lead == "continue":
            return [pad + "pass"]
        if lead in ("print", "format"):
            return [pad + "pass"]
        if lead == "goto":
            raise rt.Unsupported("unconditional goto")
        if lead == "allocate":
            out = []
            for item in _split_top(s[s.index("(") + 1:_match_paren(s, s.index("("))], ","):
                e = parse_expr(item.strip())
                if e[0] == "kw":
                    continue
                code, v = self.varcode(unit, e[1])
                if any(a[0] == "slice" for a in e[2]):
                    raise rt.Unsupported("allocate with explicit lower bound")
                out.append(f"{pad}{code} = rt.alloc({_DTYPE[v.ftype]}, ({', '.join(self.ex(unit, a) for a in e[2])},))")
            return out
        if lead == "deallocate":
            out = []
            for item in _split_top(s[s.index("(") + 1:_match_paren(s, s.index("("))], ","):
                out.append(f"{pad}{self.varcode(unit, item.strip())[0]} = None")
            return out
        if lead in ("open", "close"):
            j = _match_paren(s, s.index("("))
            args = [parse_expr(a.strip()) if "=" not in a else ("kw", a.split("=")[0].strip(), parse_expr(a.split("=", 1)[1].strip()))
                    for a in _split_top(s[s.index("(") + 1:j], ",")]
            code = ", ".join(f"{a[1]}={self.ex(unit, a[2])}" if a[0] == "kw" else self.ex(unit, a) for a in args)
            return [f"{pad}IO.{lead}({code})"]
        if lead == "write":
            j = _match_paren(s, s.index("("))
            ctl = _split_top(s[s.index("(") + 1:j], ",")
            unit_code = "'*'" if ctl[0].strip() == "*" else self.ex(unit, parse_expr(ctl[0].strip()))
            if len(ctl) == 1:   # unformatted: one record of raw bytes
                return [f"{pad}IO.write_unf({unit_code}, [{', '.join(self._unf_write_items(unit, self._io_items(s[j + 1:])))}])"]
            try:
                items = [self.ex(unit, parse_expr(a.strip())) for a in _split_top(s[j + 1:], ",") if a.strip()]
            except SyntaxError:
                raise rt.Unsupported("write with implied-do")
            fmt_code = "None"
            if len(ctl) > 1 and ctl[1].strip()[:1] in ("'", '"'):
                fmt_code = self.ex(unit, parse_expr(ctl[1].strip()))
            return [f"{pad}IO.write({unit_code}, {fmt_code}, [{', '.join(items)}])"]
        if lead == "read":
            return self.read(unit, ind, s)
        # assignment
        parts = _split_top_assign(s)
        if parts:
            tgt = parse_expr(parts[0].strip())
            return [pad + self.assign(unit, tgt, self.ex(unit, parse_expr(parts[1].strip())))]
        raise rt.Unsupported("statement")


    # ---------------------------------------------------------------- unformatted (binary) I/O lists
    def _io_items(self, text):
        """split an I/O list at top-level commas; '(a(j), j=1,4)' stays one item"""
        return [a.strip() for a in _split_top(text, ",") if a.strip()]

    def _implied_do(self, item):
        """'(u(j, inod), j=1, 4)' -> (['u(j, inod)'], 'j', ['1', '4']) or None"""
        if not (item.startswith("(") and _match_paren(item, 0) == len(item) - 1):
            return None
        parts = _split_top(item[1:-1], ",")
        for k, part in enumerate(parts):
            m = re.match(r"^\s*(\w+)\s*=\s*(.*)$", part)
            if m and _split_top_assign(part):
                return [x.strip() for x in parts[:k]], m.group(1), [m.group(2).strip()] + [x.strip() for x in parts[k + 1:]]
        return None

    def _unf_write_items(self, unit, items):
        out = []
        for it in items:
            d = self._implied_do(it)
            if d:
                inner, var, rng = d
                v, _ = self.varcode(unit, var)
                body = ", ".join(self._unf_write_items(unit, inner))
                out.append(f"*[x for {v} in rt.dorange({', '.join(self.ex(unit, parse_expr(r)) for r in rng)}) for x in ({body},)]")
            else:
                out.append(self.ex(unit, parse_expr(it)))
        return out

    def _unf_read_stmts(self, unit, pad, items):
        out = []
        for it in items:
            d = self._implied_do(it)
            if d:
                inner, var, rng = d
                v, _ = self.varcode(unit, var)
                out.append(f"{pad}for {v} in rt.dorange({', '.join(self.ex(unit, parse_expr(r)) for r in rng)}):")
                out.extend(self._unf_read_stmts(unit, pad + "    ", inner))
            else:
                e = parse_expr(it)
                name = e[1]
                _, var = self.varcode(unit, name)
                out.append(pad + self.assign(unit, e, f"_rec.take({var.ftype!r})"))
        return out

    def read(self, unit, ind, s):
        pad = "    " * ind
        j = _match_paren(s, s.index("("))
        ctl = [c.strip() for c in _split_top(s[s.index("(") + 1:j], ",")]
        u = self.ex(unit, parse_expr(ctl[0]))
        items = [a.strip() for a in _split_top(s[j + 1:], ",") if a.strip()]
        if len(ctl) == 1:   # unformatted: one record
            return [f"{pad}_rec = IO.read_unf({u})"] + self._unf_read_stmts(unit, pad, self._io_items(s[j + 1:]))
        if len(ctl) > 1 and ctl[1] != "*":
            if ctl[1].replace(" ", "").lower() in ("'(a)'", '"(a)"') and len(items) == 1:
                return [pad + self.assign(unit, parse_expr(items[0]), f"IO.read_fmt_a({u})")]
            raise rt.Unsupported("formatted read")
        if not items:
            return [f"{pad}IO.read_list({u}, 0)"]
        out = [f"{pad}_tok = IO.read_list({u}, {len(items)})"]
        for k, it in enumerate(items):
            try:
                e = parse_expr(it)
            except SyntaxError:
                raise rt.Unsupported("implied-do in read")
            if e[0] == "par" or (e[0] == "ref" and any(a[0] == "kw" for a in e[2])):
                raise rt.Unsupported("implied-do in read")
            cur = self.ex(unit, e)
            out.append(pad + self.assign(unit, e, f"rt.conv_token(_tok[{k}], {cur})"))
        return out

    def call(self, unit, ind, s):
        pad = "    " * ind
        e = parse_expr(s)
        name, args = (e[1], e[2]) if e[0] == "ref" else (e[1], [])
        if name == "system_clock":
            out = []
            for k, a in enumerate(args[:2]):
                out.append(pad + self.assign(unit, a, "0" if k == 0 else "1000"))
            return out or [pad + "pass"]
        if name == "cpu_time":
            return [pad + self.assign(unit, args[0], "0")]
        r = self.lookup(unit, name)
        fn = r[1].pyname if r and r[0] == "proc" else f"p___{name}"
        self._tmp += 1
        t = f"_t{self._tmp}"
        out = [f"{pad}{t} = {fn}(" + ", ".join(self.actual(unit, a) for a in args) + ")"]
        back = []
        for k, a in enumerate(args):
            if a[0] == "name":
                rr = self.lookup(unit, a[1])
                if rr is None:
                    rr = ("local", self._implicit_var(unit, a[1]))
                if rr[0] == "proc":
                    continue
                v = rr[1] if rr[0] == "local" else rr[2]
                if v.rank == 0 and not v.parameter:
                    back.append(f"{pad}    if {t}[{k}] is not None: " + self.assign(unit, a, f"{t}[{k}]"))
            elif a[0] == "ref":
                rr = self.lookup(unit, a[1])
                if rr and rr[0] != "proc":
                    v = rr[1] if rr[0] == "local" else rr[2]
                    if v.rank and not any(x[0] == "slice" for x in a[2]):
                        back.append(f"{pad}    if {t}[{k}] is not None and type({t}[{k}]) is not np.ndarray: "
                                    + self.assign(unit, a, f"{t}[{k}]"))
        if back:
            out.append(f"{pad}if {t} is not None and len({t}) > {len(args) - 1}:")
            out.extend(back)
        return out

    # ---------------------------------------------------------------- whole-program generation
    def generate(self):
        mods, statics, procs = [], [], []
        for mod in self.modules.values():
            self._prepare_module(mod)
        units = [p for m in self.modules.values() for p in m.procs.values()] + list(self.externals.values())
        self.failed = {}
        for u in units:
            try:
                st, code = self.gen_proc(u)
            except (SyntaxError, NameError, rt.Unsupported) as ex:   # dead code in the reference the subset does not cover
                self.failed[u.pyname] = str(ex)
                st, code = [], [f"def {u.pyname}(*a, **k):", f"    raise rt.Unsupported({'not translated: ' + str(ex)!r})", ""]
            statics += st
            procs += code
        for mod in self.modules.values():
            mods.append(f"class _Mod_{mod.name}: pass")
            mods.append(f"M_{mod.name} = _Mod_{mod.name}()")
        modinit = ["def _init_modules():"]
        for mod in self.modules.values():
            for name, v in mod.vars.items():
                tgt = f"M_{mod.name}.v_{name}"
                if v.rank and (v.allocatable or any(d[1] is None for d in v.dims)):
                    modinit.append(f"    {tgt} = None")
                elif v.rank:
                    modinit.append(f"    {tgt} = rt.alloc({_DTYPE[v.ftype]}, ({self._extents(mod, v)}))")
                    if v.init is not None:
                        modinit.append(f"    {tgt}[...] = {self.ex(mod, v.init)}")
                elif v.init is not None:
                    modinit.append(f"    {tgt} = {_CONV[v.ftype]}({self.ex(mod, v.init)})")
                else:
                    modinit.append(f"    {tgt} = {_ZERO[v.ftype]}")
        modinit.append("    pass")
        head = ["import numpy as np", "from oracle.f90ref import runtime as rt", "IO = rt.IO()", "class _SV: pass", "SV = _SV()"]
        consts = [f"{name} = {code}" for code, name in self.consts.items()]
        reset = ["def _reset():", "    _init_modules()"] + [f"    {s}" for s in statics] + ["    pass", "_reset()"]
        return "\n".join(head + consts + mods + modinit + procs + reset) + "\n"


def _leading(s):
    m = re.match(r"^(end\s*if|end\s*do|end\s*forall|else\s*if|go\s*to|[a-z]+)", s)
    if not m:
        return ""
    w = m.group(1).replace(" ", "")
    if w in ("if", "elseif") and not re.match(r"^(else\s*)?if\s*\(", s):
        return ""
    if w in ("do", "call", "print", "read", "write", "open", "close", "allocate", "deallocate", "forall", "stop", "return",
             "exit", "cycle", "continue", "else", "format"):
        # "do = 3" style assignments to variables named like keywords are not used by the reference
        if _split_top_assign(s) and w not in ("do", "forall", "open", "close", "read", "write", "if", "allocate"):
            return ""
    return w


def _split_top_assign(s):
    """'a(i) = expr' -> [lhs, rhs] when there is an assignment '=' at depth 0 (not ==, /=, <=, >=, =>)"""
    depth, q = 0, None
    for i, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            if s[i + 1:i + 2] in ("=", ">") or s[i - 1:i] in ("=", "/", "<", ">"):
                continue
            return [s[:i], s[i + 1:]]
    return None


def _is_real_literal(e):
    return e[0] == "num" and not re.fullmatch(r"\d+", e[1])


def build(paths):
    """translate the given Fortran files -> (python source, namespace with the code executed)"""
    prog = Program()
    for p in paths:
        prog.load(p)
    src = prog.generate()
    ns = {"_failed": prog.failed}
    exec(compile(src, "<f90ref>", "exec"), ns)
    return src, ns
