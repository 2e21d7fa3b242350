"""The Fortran-subset translator (oracle/f90ref) is the root of trust of the reference pin, so its own semantics are tested on
small programs written for this purpose (none of this is reference code): operator precedence and association, integer
division, mixed-kind arithmetic, implicit typing, array sections / vector subscripts / lower bounds, DO-loop semantics,
argument association (copy-out of scalars, sequence association of arrays, OPTIONAL), SAVE, module PRIVATE/ONLY/renames,
the -fpp function-like macro, forward GO TO, list-directed READ, FORMAT edit descriptors, unformatted records."""
import struct

import numpy as np
import pytest

from oracle.f90ref import runtime as rt
from oracle.f90ref import translate


def run(src, tmp_path, name="t.f90"):
    p = tmp_path / name
    p.write_text(src)
    _, ns = translate.build([str(p)])
    assert not ns["_failed"], ns["_failed"]
    return ns


def test_expression_semantics(tmp_path):
    ns = run("""
subroutine ex(a, b, c, r)
  implicit none
  real(8) a, b, c, r(12)
  integer i, j
  i = 7
  j = 2
  r(1) = a - b - c            ! left to right
  r(2) = a/b*c                ! (a/b)*c
  r(3) = -a**2                ! -(a**2)
  r(4) = a**2.d0 + 2*b**2     ! integer literal promoted, x**2 = x*x
  r(5) = i/j                  ! integer division, then conversion
  r(6) = (-i)/j               ! truncation toward zero
  r(7) = 1/3 + 1.d0/3         ! 0 + 0.333...
  r(8) = 0.1                  ! default-real literal is single precision
  r(9) = 0.1d0
  r(10) = a**(-.5d0)
  r(11) = (a + b)**.5d0
  r(12) = a**1.5d0
end subroutine
""", tmp_path)
    r = np.zeros(12)
    a, b, c = np.float64(2.5), np.float64(0.3), np.float64(1e-3)
    ns["p___ex"](a, b, c, r)
    assert r[0] == (a - b) - c and r[1] == (a / b) * c and r[2] == -(a * a) and r[3] == a * a + 2 * (b * b)
    assert r[4] == 3.0 and r[5] == -3.0 and r[6] == 1.0 / 3 and r[7] == float(np.float32(0.1)) and r[8] == 0.1
    assert r[9] == rt.cr_pow(2.5, -0.5) and r[10] == rt.cr_pow(2.8, 0.5) and r[11] == rt.cr_pow(2.5, 1.5)
    assert abs(r[11] - 2.5 ** 1.5) <= 2 ** -50 * r[11]


def test_implicit_typing_and_real4(tmp_path):
    ns = run("""
module m
  real(8) x
contains
  subroutine s(out)
    real(8) out(4)
    twall = 1.1d0             ! implicit REAL(4): rounded to single on assignment
    n = 2.9d0                 ! implicit INTEGER: truncated
    out(1) = twall
    out(2) = n
    out(3) = twall*2          ! single-precision product
    out(4) = twall*2.d0       ! promoted to double
  end subroutine
  subroutine t(out)
    implicit real(8) (a-h,o-z)
    real(8) out(2)
    v = 1.1d0
    k = 7
    out(1) = v
    out(2) = k/2
  end subroutine
end module
""", tmp_path)
    o = np.zeros(4)
    ns["p_m__s"](o)
    f = np.float32(1.1)
    assert o[0] == float(f) and o[1] == 2.0 and o[2] == float(f * np.float32(2)) and o[3] == float(f) * 2.0
    o = np.zeros(2)
    ns["p_m__t"](o)
    assert o[0] == 1.1 and o[1] == 3.0


def test_arrays_sections_and_loops(tmp_path):
    ns = run("""
subroutine arr(u, idx, n, out, last)
  implicit none
  integer n, idx(3), i, last, k(2:4)
  real(8) u(4, n), out(4), w(3), acc(2:4)
  out(:) = u(:, idx(1))*2.d0 + u(:, idx(2))
  w = u(1, idx)                         ! vector subscript
  out(1) = sum(w)
  acc(2:4) = (/ 1.d0, 2.d0, 3.d0 /)     ! lower bound 2
  k = (/ 10, 20, 30 /)
  out(2) = acc(3) + k(4)
  out(3) = 0.d0
  do i = n, 1, -2
     out(3) = out(3) + i
  end do
  last = i                               ! the DO variable after a complete loop
  do i = 1, 10
     if (i == 4) exit
     if (mod(i, 2) == 0) cycle
     out(4) = out(4) + 1.d0
  end do
  u(2:3, 1) = 0.d0
end subroutine
""", tmp_path)
    u = np.asfortranarray(np.arange(1, 21, dtype=np.float64).reshape(5, 4).T)      # u(4,5)
    out = np.zeros(4)
    ret = ns["p___arr"](u, np.array([2, 4, 5], np.int32), 5, out, 0)
    assert out[0] == u[0, 1] + u[0, 3] + u[0, 4] and out[1] == 2.0 + 30 and out[2] == 5 + 3 + 1
    assert ret[4] == -1 and out[3] == (2 * 8.0 + 16.0) + 2.0     # i = 5,3,1 then -1; out(4) from the section statement + odd i below 4
    assert u[1, 0] == 0 and u[2, 0] == 0 and u[0, 0] == 1 and u[3, 0] == 4


def test_argument_association_save_optional_private(tmp_path):
    ns = run("""
module a
  integer, private :: m
  real(8) shared
  integer, parameter :: three = 3
contains
  subroutine seta(v)
    integer v
    m = v
  end subroutine
  integer function geta()
    geta = m
  end function
end module
module b
  real(8) m(2)                 ! public m, while a's m is private
end module
subroutine bump(x, n, flat, opt)
  implicit none
  real(8) x, flat(6)
  integer n
  integer, optional :: opt
  integer, save :: calls = 0
  calls = calls + 1
  x = x + calls
  n = n*2
  flat(6) = 66.d0
  if (present(opt)) n = n + opt
end subroutine
subroutine driver(res)
  use a
  use b
  use a, only: sh => shared
  implicit none
  real(8) res(6), y, grid(2,3)
  integer k
  m = (/ 5.d0, 6.d0 /)         ! resolves to b's array
  call seta(41)
  sh = 9.d0
  y = 1.d0
  k = 3
  grid = 0.d0
  call bump(y, k, grid)        ! scalar copy-out, (2,3) actual for a flat(6) dummy
  call bump(y, k, grid, three)
  res(1) = y
  res(2) = k
  res(3) = grid(2,3)
  res(4) = geta() + 1
  res(5) = m(2)
  res(6) = shared
end subroutine
""", tmp_path)
    res = np.zeros(6)
    ns["p___driver"](res)
    assert res.tolist() == [1 + 1 + 2, (3 * 2) * 2 + 3, 66.0, 42.0, 6.0, 9.0]


def test_macro_goto_and_format(tmp_path):
    ns = run("""#define twice(call_it, acc) call call_it; call call_it; acc = acc + 1;
module c
  integer hits, n
contains
  subroutine hit(k)
    integer k
    hits = hits + k
  end subroutine
end module
subroutine go(flag, r)
  use c
  implicit none
  integer flag
  real(8) r
  hits = 0
  n = 0
  twice(hit(5), n)
  r = 1.d0
  if (flag.eq.1) go to 10
  r = 2.d0
10 continue
  r = r + hits + n
end subroutine
""", tmp_path)
    assert ns["p___go"](1, np.float64(0))[1] == 1.0 + 10 + 1
    assert ns["p___go"](0, np.float64(0))[1] == 2.0 + 10 + 1
    assert rt.format_records("(A15, 5(I8, 2X))", ["DENSITY", 2, 3, 1, 1, 1]) == ["        DENSITY       2         3         1         1         1  "]
    assert rt.format_records("(I8, 3E13.4)", [7, np.float64(-1234.5), np.float64(0.0)]) == ["       7  -0.1234E+04   0.0000E+00"]
    assert rt.format_records("(2F8.3/)", [np.float64(1.0005), np.float64(-2.5)]) == ["   1.000  -2.500", ""]   # 1.0005 is below the tie in binary
    assert rt.format_records("(I3)", [1, 2, 3]) == ["  1", "  2", "  3"]                      # format reversion: one record each
    with pytest.raises(rt.FormatError):
        rt.format_records("(I7, 2E14.6)", [1, np.float64(1), np.float64(2), np.float64(3)])   # a real meets I7 after reversion


def test_list_directed_read_and_unformatted_records(tmp_path):
    (tmp_path / "in.dat").write_text("header line\n 3  2.5d0, 7\n'abc' 1.e-3\n")
    ns = run(f"""
subroutine io(n, x, k, s, y)
  implicit none
  integer n, k, j
  real(8) x, y, v(3)
  character(4) s
  open(1, FILE='{tmp_path}/in.dat', STATUS='OLD')
  read(1, *)
  read(1, *) n, x, k
  read(1, *) s, y
  close(1)
  v = (/ 1.d0, 2.d0, 3.d0 /)
  open(2, FILE='{tmp_path}/out.bin', FORM='UNFORMATTED', STATUS='UNKNOWN')
  write(2) n, x
  write(2) (v(j), j=1, 3), k
  close(2)
  open(2, FILE='{tmp_path}/out.bin', FORM='UNFORMATTED', STATUS='UNKNOWN')
  read(2) n, x
  read(2) (v(j), j=3, 1, -1), k
  close(2)
  y = v(1)
end subroutine
""", tmp_path)
    n, x, k, s, y = ns["p___io"](0, np.float64(0), 0, "", np.float64(0))
    assert (n, x, k, s) == (3, 2.5, 7, "abc") and y == 3.0          # read back in reverse: v(3)=1, v(2)=2, v(1)=3
    raw = (tmp_path / "out.bin").read_bytes()
    assert raw == struct.pack("<iidi", 12, 3, 2.5, 12) + struct.pack("<idddii", 28, 1.0, 2.0, 3.0, 7, 28)
