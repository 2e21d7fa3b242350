! ISO_C_BINDING interfaces to libcfdb200.so (include/cfdb.h).  Source only: no Fortran compiler exists in
! the build image (SURVEY.md F1), so this file has not been compiled here; it is kept in step with cfdb.h by
! tests/test_fortran_shim.py, which checks every bind(C) name and argument count against the header.
module cfdb_iface
  use iso_c_binding
  implicit none
  type(c_ptr), save :: cfdb_ctx = c_null_ptr   ! one context per process, bound to the mesh (like the SAVEd state of Mlaplace)

  type, bind(C) :: cfdb_params
     real(c_double) :: FSAFE, U_inf, V_inf, MACH_inf, T_inf, RHO_inf, P_inf, C_inf
     real(c_double) :: FMU, FGX, FGY, QH, FK, FR, FCv, GAMA, CTE
     real(c_double) :: XREF(10), YREF(10)
     integer(c_int32_t) :: IRESTART, MAXITER, IPRINT, MOVIE, ITLOCAL, MOVING, NGAS, use_gcl
  end type

  type, bind(C) :: cfdb_bc
     integer(c_int32_t) :: nfixrho;   type(c_ptr) :: ifixrho_node, rfixrho_value
     integer(c_int32_t) :: nfixv;     type(c_ptr) :: ifixv_node, rfixv_valuex, rfixv_valuey
     integer(c_int32_t) :: nwall;     type(c_ptr) :: wall
     integer(c_int32_t) :: nfixt;     type(c_ptr) :: ifixt_node, rfixt_value
     integer(c_int32_t) :: nsets;     type(c_ptr) :: iset_n1, iset_n2, iset_elem, iset_id
     integer(c_int32_t) :: nmove;     type(c_ptr) :: i_m
     integer(c_int32_t) :: nfix_move; type(c_ptr) :: ifm
  end type

  interface
     function cfdb_last_error() bind(C, name="cfdb_last_error") result(msg)
       import; type(c_ptr) :: msg
     end function
     function cfdb_create(ctx, par, npoin, nelem, X, Y, inpoel, bc, device) bind(C, name="cfdb_create") result(rc)
       import; type(c_ptr) :: ctx; type(cfdb_params) :: par; integer(c_int32_t), value :: npoin, nelem
       real(c_double) :: X(*), Y(*); integer(c_int32_t) :: inpoel(3,*); type(cfdb_bc) :: bc
       integer(c_int), value :: device; integer(c_int) :: rc
     end function
     subroutine cfdb_destroy(ctx) bind(C, name="cfdb_destroy")
       import; type(c_ptr), value :: ctx
     end subroutine
     function cfdb_init(ctx) bind(C, name="cfdb_init") result(rc)
       import; type(c_ptr), value :: ctx; integer(c_int) :: rc
     end function
     function cfdb_step(ctx, nsteps) bind(C, name="cfdb_step") result(rc)
       import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: nsteps; integer(c_int) :: rc
     end function
     function cfdb_force_visc(ctx) bind(C, name="cfdb_force_visc") result(rc)   ! FORCE_VISC, ns2DComp.ALE.f90:819-893
       import; type(c_ptr), value :: ctx; integer(c_int) :: rc
     end function
     function cfdb_get(ctx, name, host, count) bind(C, name="cfdb_get") result(rc)
       import; type(c_ptr), value :: ctx, host; character(kind=c_char) :: name(*)
       integer(c_int64_t), value :: count; integer(c_int) :: rc
     end function
     function cfdb_set(ctx, name, host, count) bind(C, name="cfdb_set") result(rc)
       import; type(c_ptr), value :: ctx, host; character(kind=c_char) :: name(*)
       integer(c_int64_t), value :: count; integer(c_int) :: rc
     end function
     function cfdb_calcrhs(ctx, rhs, U, theta, T, dNx, dNy, area, shoc, dtl, t_sugn1, t_sugn2, t_sugn3, inpoel, &
          nelem, npoin, Cv, lambda_ref, mu_ref, gamma0, T_inf, cte) bind(C, name="cfdb_calcrhs") result(rc)
       import; type(c_ptr), value :: ctx
       real(c_double) :: rhs(4,*), U(4,*), theta(4,*), T(*), dNx(3,*), dNy(3,*)
       real(c_double) :: area(*), shoc(*), dtl(*), t_sugn1(*), t_sugn2(*), t_sugn3(*)
       integer(c_int32_t) :: inpoel(3,*); integer(c_int32_t), value :: nelem, npoin
       real(c_double), value :: Cv, lambda_ref, mu_ref, gamma0, T_inf, cte; integer(c_int) :: rc
     end function
     function cfdb_fuente(ctx, rhs, U, w_x, w_y, dNx, dNy, area, dtl, inpoel, nelem, npoin) &
          bind(C, name="cfdb_fuente") result(rc)
       import; type(c_ptr), value :: ctx
       real(c_double) :: rhs(4,*), U(4,*), w_x(*), w_y(*), dNx(3,*), dNy(3,*), area(*), dtl(*)
       integer(c_int32_t) :: inpoel(3,*); integer(c_int32_t), value :: nelem, npoin; integer(c_int) :: rc
     end function
     function cfdb_deltat(ctx, dtmin, dt, inpoel, area, T, vel_x, vel_y, w_x, w_y, nelem, npoin, FSAFE, FR, GAMA, &
          T_inf) bind(C, name="cfdb_deltat") result(rc)
       import; type(c_ptr), value :: ctx; real(c_double) :: dtmin, dt(*), area(*), T(*), vel_x(*), vel_y(*), w_x(*), w_y(*)
       integer(c_int32_t) :: inpoel(3,*); integer(c_int32_t), value :: nelem, npoin
       real(c_double), value :: FSAFE, FR, GAMA, T_inf; integer(c_int) :: rc
     end function
     function cfdb_estab(ctx, U, T, vel_x, vel_y, w_x, w_y, GAMM, dNx, dNy, inpoel, nelem, npoin, FR, DTMIN, &
          RHOINF, TINF, shoc, t_sugn1, t_sugn2, t_sugn3) bind(C, name="cfdb_estab") result(rc)
       import; type(c_ptr), value :: ctx
       real(c_double) :: U(4,*), T(*), vel_x(*), vel_y(*), w_x(*), w_y(*), GAMM(*), dNx(3,*), dNy(3,*)
       integer(c_int32_t) :: inpoel(3,*); integer(c_int32_t), value :: nelem, npoin
       real(c_double), value :: FR, DTMIN, RHOINF, TINF
       real(c_double) :: shoc(*), t_sugn1(*), t_sugn2(*), t_sugn3(*); integer(c_int) :: rc
     end function
     function cfdb_deriv(ctx, X, Y, inpoel, nelem, npoin, area, HH, HHX, HHY, dNx, dNy, hmin) &
          bind(C, name="cfdb_deriv") result(rc)
       import; type(c_ptr), value :: ctx; real(c_double) :: X(*), Y(*); integer(c_int32_t) :: inpoel(3,*)
       integer(c_int32_t), value :: nelem, npoin
       real(c_double) :: area(*), HH(*), HHX(*), HHY(*), dNx(3,*), dNy(3,*), hmin; integer(c_int) :: rc
     end function
     function cfdb_masas(ctx, area, inpoel, nelem, npoin, M) bind(C, name="cfdb_masas") result(rc)
       import; type(c_ptr), value :: ctx; real(c_double) :: area(*), M(*); integer(c_int32_t) :: inpoel(3,*)
       integer(c_int32_t), value :: nelem, npoin; integer(c_int) :: rc
     end function
     function cfdb_laplace(ctx, inpoel, area, dNx, dNy, X, Y, nelem, npoin, lap_sparse, lap_diag) &
          bind(C, name="cfdb_laplace") result(rc)
       import; type(c_ptr), value :: ctx; integer(c_int32_t) :: inpoel(3,*)
       real(c_double) :: area(*), dNx(3,*), dNy(3,*), X(*), Y(*), lap_sparse(*), lap_diag(*)
       integer(c_int32_t), value :: nelem, npoin; integer(c_int) :: rc
     end function
     function cfdb_bicg(ctx, spMtx, spIdx, spRowptr, diagMtx, x, b, x_fix, x_fixIdx, npoin, nfix, iters) &
          bind(C, name="cfdb_bicg") result(rc)
       import; type(c_ptr), value :: ctx; real(c_double) :: spMtx(*), diagMtx(*), x(*), b(*), x_fix(*)
       integer(c_int32_t) :: spIdx(*), spRowptr(*), x_fixIdx(*), iters
       integer(c_int32_t), value :: npoin, nfix; integer(c_int) :: rc
     end function
     function cfdb_gcl_main(ctx, M, W_x, W_y, W_x_old, W_y_old, area_old, dNx, dNy, area, inpoel, nelem, npoin, dt) &
          bind(C, name="cfdb_gcl_main") result(rc)
       import; type(c_ptr), value :: ctx
       real(c_double) :: M(*), W_x(*), W_y(*), W_x_old(*), W_y_old(*), area_old(*), dNx(3,*), dNy(3,*), area(*)
       integer(c_int32_t) :: inpoel(3,*); integer(c_int32_t), value :: nelem, npoin
       real(c_double), value :: dt; integer(c_int) :: rc
     end function
  end interface
contains
  subroutine cfdb_check(rc, who)      ! the reference's error convention is STOP (dataLoader.f90:225,284; gcl.f90:25)
    integer(c_int), intent(in) :: rc
    character(*), intent(in) :: who
    if (rc /= 0) then
       write(*,*) 'libcfdb200 error in ', who
       stop 1
    end if
  end subroutine
end module cfdb_iface
