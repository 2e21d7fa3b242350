! Drop-in replacements with the reference's module / subroutine names and argument lists; each forwards to
! the C ABI.  Link these instead of calcRHS.f90, biconjGrad.f90 and gcl.f90.  Source only (see cfdb_iface.f90).
module calcRHS_mod          ! replaces calcRHS.f90:1-156
  implicit none
contains
  subroutine calcRHS(rhs, U, theta, dNx, dNy, area, shoc, dtl, t_sugn1, t_sugn2, t_sugn3, inpoel, nelem, npoin)
    use cfdb_iface
    use InputData, only: Cv => FCV, lambda_ref => FK, mu_ref => FMU, gamma0 => gama, T_inf, cte
    use Mvariables, only: T
    integer, intent(in) :: npoin, nelem, inpoel(3,nelem)
    real*8, intent(inout) :: rhs(4,npoin)
    real*8, intent(in) :: U(4,npoin), theta(4,npoin), dNx(3,nelem), dNy(3,nelem)
    real*8, intent(in), dimension(nelem) :: area, dtl, shoc, t_sugn1, t_sugn2, t_sugn3
    call cfdb_check(cfdb_calcrhs(cfdb_ctx, rhs, U, theta, T, dNx, dNy, area, shoc, dtl, t_sugn1, t_sugn2, t_sugn3, &
         inpoel, nelem, npoin, Cv, lambda_ref, mu_ref, gamma0, T_inf, cte), 'calcRHS')
  end subroutine calcRHS
end module calcRHS_mod

module BiconjGrad           ! replaces biconjGrad.f90:1-191
  implicit none
  private
  public biCG
contains
  subroutine biCG(spMtx, spIdx, spRowptr, diagMtx, x, b, x_fix, x_fixIdx, npoin, nfix)
    use cfdb_iface
    integer npoin, nfix, iters
    integer spRowptr(npoin + 1), spIdx(spRowptr(npoin + 1)), x_fixIdx(nfix)
    real(8) spMtx(spRowptr(npoin + 1)), x(npoin), b(npoin), diagMtx(npoin), x_fix(nfix)
    call cfdb_check(cfdb_bicg(cfdb_ctx, spMtx, spIdx, spRowptr, diagMtx, x, b, x_fix, x_fixIdx, npoin, nfix, iters), 'biCG')
  end subroutine biCG
end module BiconjGrad

module gcl_mod              ! replaces gcl.f90:1-63 (putW/putArea keep the old fields on the Fortran side)
  implicit none
  private
  real*8, dimension(:), allocatable :: area_old, W_x_old, W_y_old
  public :: main, putW, putArea
contains
  subroutine main(M, W_x, W_y, dNx, dNy, area, inpoel, dt)
    use cfdb_iface
    real*8, intent(inout), dimension(:) :: M
    real*8, intent(in), dimension(:,:) :: dNx, dNy
    real*8, intent(in), dimension(:) :: W_x, W_y, area
    integer, intent(in), dimension(:,:) :: inpoel
    real*8, intent(in) :: dt
    if (.not.allocated(area_old) .or. .not.allocated(W_x_old) .or. .not.allocated(W_y_old)) stop 'Faltan valores (GCL)'
    call cfdb_check(cfdb_gcl_main(cfdb_ctx, M, W_x, W_y, W_x_old, W_y_old, area_old, dNx, dNy, area, inpoel, &
         size(inpoel,2), size(M), dt), 'gcl main')
  end subroutine
  subroutine putW(W_x, W_y)
    real*8, dimension(:), intent(in) :: W_x, W_y
    if (.not.allocated(W_x_old)) allocate(W_x_old(size(W_x)))
    if (.not.allocated(W_y_old)) allocate(W_y_old(size(W_y)))
    W_x_old = W_x; W_y_old = W_y
  end subroutine
  subroutine putArea(area)
    real*8, dimension(:), intent(in) :: area
    if (.not.allocated(area_old)) allocate(area_old(size(area)))
    area_old = area
  end subroutine
end module gcl_mod

! external subroutines of subrutinas.f90, same names and dummy lists
subroutine deriv(hmin)      ! subrutinas.f90:88
  use cfdb_iface
  use MeshData, only: X, Y, inpoel, area, HH, HHX, HHY, dNx, dNy, nelem, npoin
  implicit none
  real(8) hmin
  call cfdb_check(cfdb_deriv(cfdb_ctx, X, Y, inpoel, nelem, npoin, area, HH, HHX, HHY, dNx, dNy, hmin), 'deriv')
end subroutine deriv

subroutine MASAS()          ! subrutinas.f90:128
  use cfdb_iface
  use MeshData, only: M, inpoel, area, nelem, npoin
  implicit none
  call cfdb_check(cfdb_masas(cfdb_ctx, area, inpoel, nelem, npoin, M), 'MASAS')
end subroutine MASAS

subroutine deltat(dtmin, dt) ! subrutinas.f90:155
  use cfdb_iface
  use MeshData, only: inpoel, nelem, npoin, area
  use InputData, only: FSAFE, fr, gama, t_inf
  use MVELOCIDADES
  use MVARIABLES
  implicit none
  real(8) DT(nelem), dtmin
  call cfdb_check(cfdb_deltat(cfdb_ctx, dtmin, DT, inpoel, area, T, VEL_X, VEL_Y, W_X, W_Y, nelem, npoin, FSAFE, fr, gama, &
       t_inf), 'deltat')
end subroutine deltat

subroutine ESTAB(U, T, GAMA, FR, RMU, DTMIN, RHOINF, TINF, UINF, VINF, GAMM)   ! subrutinas.f90:332
  use cfdb_iface
  use MeshData
  use MVELOCIDADES
  use MESTABILIZACION
  implicit real(8) (A-H,O-Z)
  real(8) U(4,npoin), T(npoin), GAMM(npoin)
  call cfdb_check(cfdb_estab(cfdb_ctx, U, T, VEL_X, VEL_Y, W_X, W_Y, GAMM, dNx, dNy, inpoel, nelem, npoin, FR, DTMIN, &
       RHOINF, TINF, SHOC, T_SUGN1, T_SUGN2, T_SUGN3), 'ESTAB')
end subroutine ESTAB

subroutine FUENTE(dtl)      ! subrutinas.f90:1036
  use cfdb_iface
  use MeshData
  use MVELOCIDADES
  use MVARIABGEN
  implicit none
  real(8) dtl(nelem)
  call cfdb_check(cfdb_fuente(cfdb_ctx, RHS, U, W_X, W_Y, dNx, dNy, area, dtl, inpoel, nelem, npoin), 'FUENTE')
end subroutine FUENTE
