"""Run-time support for the Python code oracle/f90ref/translate.py generates from the reference's Fortran sources.

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/): only tests/ and tests/golden/ generators import it.

Arithmetic model = what a `gfortran -O2` x86-64 build (no -ffast-math, no FMA contraction) evaluates:
  * real(8) -> numpy.float64 scalars, real(4) -> numpy.float32, integer -> Python int / numpy int32 arrays;
    every + - * / and sqrt is one IEEE operation, evaluated in the source's left-to-right order by the generated code;
  * integer / integer truncates toward zero;
  * x**2 and x**2.d0 are x*x (GCC folds pow(x, 2.0) exactly);
  * x**1.5d0, x**.5d0, x**(-.5d0) are libm pow() in a gfortran build.  glibc's pow is not correctly rounded (<1 ulp),
    so -- like oracle/orc_math.h and cfd_b200/csrc/exact.cuh -- they are pinned to the CORRECTLY ROUNDED value, computed
    here independently with exact integer arithmetic (cr_pow); `pow_stats` counts how often glibc's pow (math.pow) differs;
  * sin, cos, exp, acos go to glibc through the `math` module (the same libm a gfortran binary links);
  * SUM() adds sequentially from zero in array-element order; OpenMP directives are comments (one thread).
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

f8 = np.float64
f4 = np.float32
_INT = (int, np.int32, np.int64, np.intc)
pow_stats = {"calls": 0, "glibc_differs": 0}


class FortranStop(Exception):
    pass


class Unsupported(Exception):
    pass


def i4(x):
    """assignment to an integer variable: truncation toward zero"""
    return int(x)


def lg(x):
    return bool(x)


def div(a, b):
    ta, tb = type(a), type(b)
    if ta in _INT and tb in _INT:
        q = abs(int(a)) // abs(int(b))
        return q if (a < 0) == (b < 0) else -q
    if ta is np.ndarray or tb is np.ndarray:
        ka = a.dtype.kind if ta is np.ndarray else ("i" if ta in _INT else "f")
        kb = b.dtype.kind if tb is np.ndarray else ("i" if tb in _INT else "f")
        if ka == "i" and kb == "i":
            return np.trunc(np.true_divide(a, b)).astype(np.int32)
    return a / b


def _round_sqrt_fraction(v: Fraction) -> float:
    """correctly rounded (nearest-even) double of sqrt(v), v a positive rational in the normal range"""
    num, den = v.numerator, v.denominator
    # scale by 4**k so that the integer square root has exactly 54 bits
    k = (108 - (num.bit_length() - den.bit_length())) // 2
    while True:
        scaled_num = num << (2 * k) if k >= 0 else num
        scaled_den = den if k >= 0 else den << (-2 * k)
        fl = scaled_num // scaled_den
        t = math.isqrt(fl)
        if t.bit_length() > 54:
            k -= 1
        elif t.bit_length() < 54:
            k += 1
        else:
            break
    exact = (t * t == fl) and (fl * scaled_den == scaled_num)
    q, half = t >> 1, t & 1
    if half and (not exact or (q & 1)):
        q += 1
    try:
        return math.ldexp(q, 1 - k)
    except OverflowError:
        return math.inf


def cr_pow(x: float, p: float) -> float:
    """correctly rounded x**p for p in {1.5, 0.5, -0.5}; IEEE special cases as C pow()"""
    x = float(x)
    if x != x:
        return x
    if x == 0.0:
        return 0.0 if p > 0 else math.inf
    if x < 0:
        return math.nan if x != -math.inf else (math.inf if p > 0 else 0.0)
    if x == math.inf:
        return math.inf if p > 0 else 0.0
    fx = Fraction(x)
    if p == 1.5:
        v = fx * fx * fx
    elif p == 0.5:
        v = fx
    elif p == -0.5:
        v = 1 / fx
    else:
        raise ValueError(p)
    r = _round_sqrt_fraction(v)
    pow_stats["calls"] += 1
    if math.pow(x, p) != r:
        pow_stats["glibc_differs"] += 1
    return r


def _pow_real(a, b):
    b = float(b)
    if b == 2.0:
        return a * a
    if b in (1.5, 0.5, -0.5):
        if type(a) is np.ndarray:
            return np.array([cr_pow(v, b) for v in a.ravel(order="F")], np.float64).reshape(a.shape, order="F")
        return f8(cr_pow(a, b))
    if type(a) is np.ndarray:
        return np.array([math.pow(v, b) for v in a.ravel(order="F")], np.float64).reshape(a.shape, order="F")
    return f8(math.pow(float(a), b))


def pow_(a, b):
    if type(b) in _INT:
        n = int(b)
        if n == 2:
            return a * a
        if n == 1:
            return a
        if n == 0:
            return a * 0 + 1
        if n == 3:
            return a * a * a
        if type(a) in _INT:
            return int(a) ** n
        raise Unsupported(f"integer power {n} of a real: expansion order is compiler-specific")
    return _pow_real(a, b)


def dorange(a, b, c=1):
    a, b, c = int(a), int(b), int(c)
    return range(a, b + 1, c) if c > 0 else range(a, b - 1, c)


def doend(a, b, c=1):
    a, b, c = int(a), int(b), int(c)
    n = max(0, (b - a + c) // c)
    return a + n * c


def alloc(dtype, shape):
    return np.zeros(tuple(int(s) for s in shape), dtype, order="F")


def shape_(x, dims):
    """explicit-shape dummy argument: sequence association with the actual argument (must stay a view)"""
    if x is None:
        return None
    dims = tuple(int(d) for d in dims)
    if type(x) is not np.ndarray:
        raise Unsupported("scalar actual argument for an array dummy")
    if x.shape == dims:
        return x
    n = 1
    for d in dims:
        n *= d
    flat = x.reshape(-1, order="F")
    if flat.size < n:
        raise ValueError(f"actual argument has {flat.size} elements, dummy needs {dims}")
    r = flat[:n].reshape(dims, order="F")
    if n and not np.shares_memory(r, x):
        raise Unsupported("dummy argument would be a copy (non-contiguous actual)")
    return r


def ac(items):
    """array constructor (/ ... /)"""
    flat = []
    for it in items:
        if type(it) is np.ndarray:
            flat.extend(it.ravel(order="F").tolist() if it.dtype.kind == "i" else list(it.ravel(order="F")))
        else:
            flat.append(it)
    if all(type(v) in _INT for v in flat):
        return np.array([int(v) for v in flat], np.int32)
    if any(type(v) is f8 or type(v) is float for v in flat):
        return np.array(flat, np.float64)
    return np.array(flat, np.float32)


def reshape(src, shp, *a):
    return np.reshape(np.asarray(src), tuple(int(s) for s in np.asarray(shp).ravel()), order="F").copy(order="F")


def sum_(x, *a):
    if a:
        raise Unsupported("sum with dim/mask")
    if type(x) is not np.ndarray:
        return x
    s = x.dtype.type(0)
    for v in x.ravel(order="F"):
        s = s + v
    return s


def minval(x):
    return x.min()


def maxval(x):
    return x.max()


def min_(*a):
    if any(type(v) is np.ndarray for v in a):
        r = a[0]
        for v in a[1:]:
            r = np.where(v < r, v, r)
        return r
    r = a[0]
    for v in a[1:]:
        if v < r:
            r = v
    return r


def max_(*a):
    if any(type(v) is np.ndarray for v in a):
        r = a[0]
        for v in a[1:]:
            r = np.where(v > r, v, r)
        return r
    r = a[0]
    for v in a[1:]:
        if v > r:
            r = v
    return r


def abs_(x):
    return abs(x)


def sqrt_(x):
    if type(x) in _INT:
        raise Unsupported("sqrt of an integer")
    return np.sqrt(x)


def _libm(fn):
    def g(x):
        if type(x) is np.ndarray:
            return np.array([fn(float(v)) for v in x.ravel(order="F")], x.dtype).reshape(x.shape, order="F")
        if type(x) is f4:
            return f4(fn(float(x)))  # gfortran calls sinf/cosf for real(4); not on the hot path
        return f8(fn(float(x)))
    return g


def _exp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


sin_ = _libm(math.sin)
cos_ = _libm(math.cos)
exp_ = _libm(_exp)
acos_ = _libm(math.acos)
atan_ = _libm(math.atan)
log_ = _libm(math.log)


log10_ = _libm(math.log10)


def any_(x):
    return bool(np.any(x))


def all_(x):
    return bool(np.all(x))


def count_(x):
    return int(np.count_nonzero(x))


def size_(x, dim=None):
    return int(x.size) if dim is None else int(x.shape[int(dim) - 1])


def real_(x, kind=None):
    if kind is not None and int(kind) == 8:
        return np.float64(x)
    return np.float32(x)


def dble_(x):
    return np.float64(x)


def int_(x):
    if type(x) is np.ndarray:
        return np.trunc(x).astype(np.int32)
    return int(x)


def mod_(a, b):
    if type(a) in _INT and type(b) in _INT:
        return int(math.fmod(int(a), int(b)))
    return type(a)(math.fmod(a, b))


def sign_(a, b):
    return abs(a) if b >= 0 else -abs(a)


def tiny_(x):
    return np.finfo(np.asarray(x).dtype).tiny


def huge_(x):
    if type(x) in _INT:
        return 2147483647
    return np.finfo(np.asarray(x).dtype).max


def trim_(s):
    return s.rstrip()


def idx(e, lb):
    """subscript -> 0-based (scalar or vector subscript)"""
    return e - lb


class _Record:
    """one unformatted record being consumed by a READ list"""

    def __init__(self, data):
        self.data, self.pos = data, 0

    def take(self, ftype):
        import struct
        fmt, n = {"i": ("<i", 4), "r4": ("<f", 4), "r8": ("<d", 8)}[ftype]
        if self.pos + n > len(self.data):
            raise IOError("READ list longer than the record")
        (v,) = struct.unpack_from(fmt, self.data, self.pos)
        self.pos += n
        return v


class IO:
    """list-directed READ from text files, WRITE/PRINT recorded (never formatted): enough for dataLoader and the .cnv values"""

    def __init__(self):
        self.units = {}
        self.written = {}
        self.text = {}
        self.binary = {}

    def open(self, unit, file=None, status=None, form=None, **kw):
        unit = int(unit)
        file = str(file).strip()
        if form is not None and str(form).strip().lower() == "unformatted":
            # sequential unformatted file: 4-byte little-endian record markers around every record (gfortran's default)
            data = b""
            import os as _os
            if _os.path.exists(file):
                with open(file, "rb") as f:
                    data = f.read()
            self.units[unit] = {"lines": None, "name": file, "unf": True, "data": data, "pos": 0, "out": bytearray()}
            return
        if status is not None and str(status).strip().lower() == "old":
            with open(file) as f:
                self.units[unit] = {"lines": f.read().split("\n"), "pos": 0, "name": file}
        else:
            self.units[unit] = {"lines": None, "name": file}
            self.written.setdefault(file, [])
            self.text[file] = []          # a re-opened file is rewritten from its first record

    def close(self, unit, **kw):
        u = self.units.pop(int(unit), None)
        if u is not None and u.get("unf") and u["out"]:
            self.binary[u["name"]] = bytes(u["out"])
            with open(u["name"], "wb") as f:
                f.write(u["out"])

    def write_unf(self, unit, items):
        import struct
        payload = bytearray()
        for v in items:
            if type(v) in _INT:
                payload += struct.pack("<i", int(v))
            elif type(v) is f4:
                payload += struct.pack("<f", float(v))
            elif type(v) is np.ndarray:
                payload += np.ascontiguousarray(v.ravel(order="F")).tobytes()
            else:
                payload += struct.pack("<d", float(v))
        u = self.units[int(unit)]
        u["out"] += struct.pack("<i", len(payload)) + payload + struct.pack("<i", len(payload))

    def read_unf(self, unit):
        import struct
        u = self.units[int(unit)]
        (n,) = struct.unpack_from("<i", u["data"], u["pos"])
        rec = _Record(u["data"][u["pos"] + 4:u["pos"] + 4 + n])
        (n2,) = struct.unpack_from("<i", u["data"], u["pos"] + 4 + n)
        if n2 != n:
            raise IOError("corrupt unformatted record")
        u["pos"] += n + 8
        return rec

    def write(self, unit, fmt, items):
        if unit == "*":
            return
        u = self.units.get(int(unit))
        if u is not None and u["lines"] is None:
            self.written[u["name"]].append(items)
            if fmt is not None:      # formatted: also keep the text records (list-directed layout is compiler-specific)
                try:
                    self.text.setdefault(u["name"], []).extend(format_records(fmt, items))
                except FormatError as ex:
                    self.text.setdefault(u["name"], []).append(f"<run-time error: {ex}>")

    def read_list(self, unit, n):
        """next record(s) -> n list-directed tokens (n == 0 skips one record)"""
        u = self.units[int(unit)]
        if n == 0:
            u["pos"] += 1
            return []
        toks = []
        while len(toks) < n:
            if u["pos"] >= len(u["lines"]):
                raise EOFError(u["name"])
            line = u["lines"][u["pos"]]
            u["pos"] += 1
            toks.extend(line.replace(",", " ").split())
        return toks[:n]

    def read_fmt_a(self, unit):
        u = self.units[int(unit)]
        line = u["lines"][u["pos"]]
        u["pos"] += 1
        return line


def conv_token(tok, like):
    """list-directed input conversion to the type of the receiving variable"""
    if isinstance(like, str):
        return tok.strip("'\"")
    t = tok.lower().replace("d", "e")
    if type(like) in _INT or (type(like) is np.ndarray and like.dtype.kind == "i"):
        return int(t)
    if type(like) is f4 or (type(like) is np.ndarray and like.dtype == np.float32):
        return f4(t)
    return f8(t)


# ----------------------------------------------------------------------------------------------------------------
# formatted WRITE: the edit descriptors the reference uses (A, Aw, Iw, Ew.d, Fw.d, nX, /, repeat counts, groups,
# format reversion), laid out the way gfortran does (leading "0." when the field has room, two-digit exponent)
# ----------------------------------------------------------------------------------------------------------------
class FormatError(Exception):
    pass


def _fmt_E(v, w, d):
    v = float(v)
    if v != v:
        s = "NaN"
    elif v in (math.inf, -math.inf):
        s = ("-" if v < 0 else "") + ("Infinity" if w >= 8 + (v < 0) else "Inf")
    else:
        neg = math.copysign(1.0, v) < 0
        if v == 0.0:
            digits, ex = "0" * d, 0
        else:
            m, e = ("%.*e" % (d - 1, abs(v))).split("e")      # d significant digits, correctly rounded
            digits, ex = m.replace(".", ""), int(e) + 1       # 0.ddddE+ex
        es = ("E%+03d" % ex) if abs(ex) <= 99 else ("%+04d" % ex)
        body = "." + digits + es
        s = ("-" if neg else "") + "0" + body
        if len(s) > w:
            s = ("-" if neg else "") + body
    return s.rjust(w) if len(s) <= w else "*" * w


def _fmt_F(v, w, d):
    v = float(v)
    if v != v:
        s = "NaN"
    elif v in (math.inf, -math.inf):
        s = ("-" if v < 0 else "") + ("Infinity" if w >= 8 + (v < 0) else "Inf")
    else:
        s = "%.*f" % (d, v)
        if len(s) > w and (s.startswith("0.") or s.startswith("-0.")):
            s = s.replace("0.", ".", 1)
    return s.rjust(w) if len(s) <= w else "*" * w


def _parse_format(f):
    """'(I7, 4E14.6)' -> nested list of items: ('I',w) ('E',w,d) ('F',w,d) ('A',w|None) ('X',n) ('/',) ('S',text) (n,[group])"""
    f = f.strip()
    if not (f.startswith("(") and f.endswith(")")):
        raise FormatError(f)
    pos = 1

    def group():
        nonlocal pos
        items = []
        while True:
            while pos < len(f) and f[pos] in " ,":
                pos += 1
            ch = f[pos]
            if ch == ")":
                pos += 1
                return items
            if ch == "/":
                items.append(("/",))
                pos += 1
                continue
            if ch in "'\"":
                q, end = ch, f.index(ch, pos + 1)
                items.append(("S", f[pos + 1:end]))
                pos = end + 1
                continue
            n = 0
            has_n = False
            while f[pos].isdigit():
                n, has_n, pos = n * 10 + int(f[pos]), True, pos + 1
            rep = n if has_n else 1
            ch = f[pos].upper()
            if ch == "(":
                pos += 1
                items.append((rep, group()))
                continue
            pos += 1
            if ch == "X":
                items.append(("X", rep))
                continue
            num = ""
            while pos < len(f) and (f[pos].isdigit() or f[pos] == "."):
                num += f[pos]
                pos += 1
            if ch == "A":
                it = ("A", int(num) if num else None)
            elif ch == "I":
                it = ("I", int(num.split(".")[0]))
            elif ch in "EFDG":
                w, d = num.split(".")
                it = ("F" if ch == "F" else "E", int(w), int(d))
            else:
                raise FormatError(f"edit descriptor {ch} in {f}")
            items.extend([it] * rep)

    return group()


def format_records(fmt, values):
    """-> list of text records produced by WRITE(u, fmt) values"""
    items = _parse_format(fmt)
    vals = list(values)
    recs, cur = [], []
    vi = 0

    class Done(Exception):
        pass

    def emit(it):
        nonlocal vi, cur
        k = it[0]
        if k == "/":
            recs.append("".join(cur))
            cur = []
        elif k == "X":
            cur.append(" " * it[1])
        elif k == "S":
            cur.append(it[1])
        else:
            if vi >= len(vals):
                raise Done
            v = vals[vi]
            vi += 1
            if k == "A":
                if not isinstance(v, str):
                    raise FormatError("A edit descriptor with a non-character item")
                cur.append(v if it[1] is None else (v[:it[1]] if len(v) >= it[1] else v.rjust(it[1])))
            elif k == "I":
                if type(v) not in _INT:
                    raise FormatError("I edit descriptor with a non-integer item (gfortran: run-time error)")
                s = str(int(v))
                cur.append(s.rjust(it[1]) if len(s) <= it[1] else "*" * it[1])
            else:
                if type(v) in _INT or isinstance(v, str):
                    raise FormatError(f"{k} edit descriptor with a non-real item")
                cur.append(_fmt_E(v, it[1], it[2]) if k == "E" else _fmt_F(v, it[1], it[2]))

    def run(lst):
        for it in lst:
            if isinstance(it[0], int):
                for _ in range(it[0]):
                    run(it[1])
            else:
                emit(it)

    try:
        run(items)
        while vi < len(vals):   # format reversion: new record, restart at the last top-level group (or the whole format)
            recs.append("".join(cur))
            cur = []
            last = [it for it in items if isinstance(it[0], int)]
            run([last[-1]] if last else items)
    except Done:
        pass
    recs.append("".join(cur))
    return recs
