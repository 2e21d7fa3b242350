! Drop-in replacements with the reference's module / subroutine names and argument lists; each forwards to
! the C ABI.  Link these instead of calcRHS.f90, biconjGrad.f90, gcl.f90, mLaplace.f90, meshMove.f90 and smoothing.f90, and
! with the bodies of Mnormales, deriv, MASAS, deltat, ESTAB, FUENTE, fixvel, FIX and RK removed from subrutinas.f90.
! Source only (see cfdb_iface.f90, which is generated from the ABI table by tools/gen_abi.py).
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

module Mnormales            ! replaces subrutinas.f90:2-86 (the list stays module-private, as in the reference)
  implicit none
  integer, private :: m
  integer, dimension(:), allocatable, private :: n_ipoin
  real(8), dimension(:), allocatable, private :: n_x, n_y
contains
  subroutine normales
    use cfdb_iface
    use meshdata, only: wall, nwall, x, y, npoin
    if (.not.allocated(n_ipoin)) allocate(n_ipoin(npoin))
    if (.not.allocated(n_x)) allocate(n_x(npoin))
    if (.not.allocated(n_y)) allocate(n_y(npoin))
    call cfdb_check(cfdb_normales(cfdb_ctx, wall, nwall, x, y, npoin, m, n_ipoin, n_x, n_y), 'normales')
  end subroutine normales
  subroutine normalvel
    use cfdb_iface
    use mvelocidades, only: vel_x, vel_y, w_x, w_y
    use meshdata, only: npoin
    call cfdb_check(cfdb_normalvel(cfdb_ctx, m, n_ipoin, n_x, n_y, vel_x, vel_y, w_x, w_y, npoin), 'normalvel')
  end subroutine normalvel
end module Mnormales

module Mlaplace             ! replaces mLaplace.f90:1-109
  implicit none
  integer, dimension(:), allocatable :: lap_idx, lap_rowptr
  real(8), dimension(:), allocatable :: lap_sparse, lap_diag
contains
  subroutine laplace(inpoel, area, dNx, dNy, nelem, npoin)
    use cfdb_iface
    use MeshData, only: X, Y
    integer, intent(in) :: nelem, npoin
    integer, intent(in) :: inpoel(3,nelem)
    real*8, intent(in) :: dNx(3,nelem), dNy(3,nelem), area(nelem)
    integer(c_int64_t) :: nnz
    logical, save :: isFirstCall = .true.
    if (isFirstCall) then      ! Mlaplace::initialize (mLaplace.f90:60-94): the CSR pattern comes from the library
       nnz = cfdb_field_size(cfdb_ctx, 'lap_idx' // c_null_char)
       allocate(lap_idx(nnz), lap_rowptr(npoin + 1), lap_sparse(nnz), lap_diag(npoin))
       call cfdb_check(cfdb_get(cfdb_ctx, 'lap_idx' // c_null_char, c_loc(lap_idx), nnz), 'laplace: lap_idx')
       call cfdb_check(cfdb_get(cfdb_ctx, 'lap_rowptr' // c_null_char, c_loc(lap_rowptr), int(npoin + 1, c_int64_t)), 'laplace: lap_rowptr')
       isFirstCall = .false.
    end if
    call cfdb_check(cfdb_laplace(cfdb_ctx, inpoel, area, dNx, dNy, X, Y, nelem, npoin, lap_sparse, lap_diag), 'laplace')
  end subroutine laplace
end module Mlaplace

module MeshMove             ! replaces meshMove.f90 (fluidStructure and the public module variables it leaves behind)
  implicit none
  real(8) fx(10), fy(10), rm(10), f_vx(10), f_vy(10)
  real(8), dimension(:), allocatable :: xpos, ypos
  private
  public :: fluidStructure
  public :: fx, fy, rm, f_vx, f_vy, xpos, ypos
contains
  subroutine fluidStructure(dtmin, time, SMOOTH_FIX, x1, y1)
    use cfdb_iface
    use MeshData, only: X, Y, npoin
    use mvelocidades, only: W_X, W_Y
    use mvariables, only: P
    real*8 :: dtmin, time
    real*8 :: x1(npoin), y1(npoin)
    logical :: SMOOTH_FIX(npoin)
    if (.not. allocated(xpos)) then
       allocate(xpos(npoin), ypos(npoin))
       xpos = 0.d0; ypos = 0.d0      ! SURVEY.md F12: the reference relies on zero-initialised ALLOCATE memory here
    end if
    call cfdb_check(cfdb_mesh_move(cfdb_ctx, dtmin, time, X, Y, x1, y1, W_X, W_Y, P, xpos, ypos, fx, fy, rm, npoin), 'fluidStructure')
  end subroutine fluidStructure
end module MeshMove

module smoothing_mod        ! replaces smoothing.f90 (the optional dX, dY of the reference's dummy list are never passed, ns2DComp.ALE.f90:76)
  implicit none
  private
  public :: smoothing
contains
  subroutine smoothing(X, Y, inpoel, fixed, npoin0, nelem0)
    use cfdb_iface
    integer, intent(in) :: npoin0, nelem0, inpoel(3,nelem0)
    real*8, intent(inout) :: X(npoin0), Y(npoin0)
    logical, intent(in) :: fixed(npoin0)
    integer(c_int8_t) :: fixed8(npoin0)
    integer :: sweeps
    fixed8 = merge(1_c_int8_t, 0_c_int8_t, fixed)
    call cfdb_check(cfdb_smoothing(X, Y, inpoel, fixed8, npoin0, nelem0, sweeps), 'smoothing')
  end subroutine smoothing
end module smoothing_mod

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

subroutine fixvel           ! subrutinas.f90:601
  use cfdb_iface
  use mvelocidades, only: vel_x, vel_y
  use meshdata, only: nfixv, ifixv_node, rfixv_valuex, rfixv_valuey, npoin
  implicit none
  call cfdb_check(cfdb_fixvel(cfdb_ctx, nfixv, ifixv_node, rfixv_valuex, rfixv_valuey, vel_x, vel_y, npoin), 'fixvel')
end subroutine fixvel

subroutine FIX(FR, GAMM)    ! subrutinas.f90:618
  use cfdb_iface
  use MVELOCIDADES, only: VEL_X, VEL_Y
  use MVARIABLES, only: rho, T, E
  use MeshData, only: npoin, nfixrho, ifixrho_node, rfixrho_value, NFIXT, IFIXT_NODE, RFIXT_VALUE
  implicit none
  real(8) FR, GAMM(npoin)
  call cfdb_check(cfdb_fix(cfdb_ctx, FR, GAMM, nfixrho, ifixrho_node, rfixrho_value, NFIXT, IFIXT_NODE, RFIXT_VALUE, VEL_X, VEL_Y, &
       rho, T, E, npoin), 'FIX')
end subroutine FIX

subroutine RK(DTMIN, NRK, BANDERA, GAMM, dtl)   ! subrutinas.f90:645
  use cfdb_iface
  use MVELOCIDADES, only: VEL_X, VEL_Y, W_X, W_Y
  use MVARIABGEN, only: U, U1, RHS, RHS1, RHS2, RHS3
  use MeshData, only: nelem, npoin
  use MVARIABLES, only: T, P, RHO, E, RMACH
  use MESTABILIZACION, only: SHOC, T_SUGN1, T_SUGN2, T_SUGN3
  implicit none
  real(8) DTMIN, GAMM(npoin), dtl(nelem)
  integer NRK, BANDERA
  call cfdb_check(cfdb_rk(cfdb_ctx, DTMIN, NRK, BANDERA, GAMM, dtl, U, U1, RHS, RHS1, RHS2, RHS3, T, P, RHO, E, RMACH, VEL_X, VEL_Y, &
       W_X, W_Y, SHOC, T_SUGN1, T_SUGN2, T_SUGN3, nelem, npoin), 'RK')
end subroutine RK
