! Drop-in for utilities/operators.F90:17-224 (gradient, curl, divergence, filter3D of y-pencil fields).
! Same subroutine names and argument lists; `der` is accepted for source compatibility and the work is done by
! pdo_operators_* of libpadeops_b200.so, which differentiates along z
! WITHOUT transposing when the z-slabs allow it (distributed compact solve over NVLink; include/padeops_b200.h).
! One operators handle is cached per decomp: creating it is collective, like decomp_info_init.
module operators
    use iso_c_binding
    use padeops_b200_c
    use kind_parameters, only: rkind
    use decomp_2d, only: decomp_info
    use DerivativesMod, only: derivatives
    use FiltersMod, only: filters          ! the shim's filters type carries the library handle as fil%h (pattern of cd10.F90)
    use exits, only: GracefulExit
    implicit none
    private
    public :: gradient, curl, divergence, filter3D, operators_configure
    type(c_ptr), save :: hops = c_null_ptr, hops_decomp = c_null_ptr
    real(rkind), save :: cfg_dx = 0, cfg_dy = 0, cfg_dz = 0
    character(len=4), save :: cfg_method = "cd10"

contains

    ! The reference's `derivatives` type keeps its spacings and methods private, so the host program states them once,
    ! right after der%init (the one line a maintainer adds):  call operators_configure(dx, dy, dz, "cd10")
    subroutine operators_configure(dx, dy, dz, method)
        real(rkind), intent(in) :: dx, dy, dz
        character(len=*), intent(in) :: method
        integer(c_int) :: ierr
        cfg_dx = dx; cfg_dy = dy; cfg_dz = dz; cfg_method = method
        if (c_associated(hops)) ierr = pdo_operators_destroy(hops)
        hops = c_null_ptr
    end subroutine

    subroutine ensure_handle(decomp)
        type(decomp_info), intent(in) :: decomp
        integer(c_int) :: ierr
        if (c_associated(hops) .and. c_associated(hops_decomp, decomp%h)) return
        if (c_associated(hops)) ierr = pdo_operators_destroy(hops)
        ierr = pdo_operators_init(hops, decomp%h, cfg_dx, cfg_dy, cfg_dz, trim(cfg_method)//c_null_char, 1_c_int)
        if (ierr /= 0) call GracefulExit("padeops_b200: operators init failed (call operators_configure first)", ierr)
        hops_decomp = decomp%h
    end subroutine

    subroutine gradient(decomp, der, f, dfdx, dfdy, dfdz, x_bc_, y_bc_, z_bc_)          ! operators.F90:17
        type(decomp_info), intent(in) :: decomp
        type(derivatives), intent(in) :: der
        real(rkind), dimension(decomp%ysz(1), decomp%ysz(2), decomp%ysz(3)), intent(in),  target :: f
        real(rkind), dimension(size(f,1), size(f,2), size(f,3)),             intent(out), target :: dfdx, dfdy, dfdz
        integer, dimension(2), optional, intent(in) :: x_bc_, y_bc_, z_bc_
        integer(c_int) :: ierr
        call ensure_handle(decomp)
        ierr = pdo_operators_gradient(hops, c_loc(f), c_loc(dfdx), c_loc(dfdy), c_loc(dfdz), c_null_ptr)
        if (ierr /= 0) call GracefulExit("padeops_b200: gradient failed", ierr)
    end subroutine

    subroutine curl(decomp, der, u, v, w, curlu, x_bc_, y_bc_, z_bc_)                   ! operators.F90:55
        type(decomp_info), intent(in) :: decomp
        type(derivatives), intent(in) :: der
        real(rkind), dimension(decomp%ysz(1), decomp%ysz(2), decomp%ysz(3)), intent(in),  target :: u, v, w
        real(rkind), dimension(size(u,1), size(u,2), size(u,3), 3),          intent(out), target :: curlu
        integer, dimension(2), optional, intent(in) :: x_bc_, y_bc_, z_bc_
        integer(c_int) :: ierr
        call ensure_handle(decomp)
        ierr = pdo_operators_curl(hops, c_loc(u), c_loc(v), c_loc(w), c_loc(curlu), c_null_ptr)
        if (ierr /= 0) call GracefulExit("padeops_b200: curl failed", ierr)
    end subroutine

    subroutine divergence(decomp, der, u, v, w, div, x_bc_, y_bc_, z_bc_)               ! operators.F90:118
        type(decomp_info), intent(in) :: decomp
        type(derivatives) :: der
        real(rkind), dimension(decomp%ysz(1), decomp%ysz(2), decomp%ysz(3)), intent(in),  target :: u, v, w
        real(rkind), dimension(size(u,1), size(u,2), size(u,3)),             intent(out), target :: div
        integer, dimension(2), optional, intent(in) :: x_bc_, y_bc_, z_bc_
        integer(c_int) :: ierr
        call ensure_handle(decomp)
        ierr = pdo_operators_divergence(hops, c_loc(u), c_loc(v), c_loc(w), c_loc(div), c_null_ptr)
        if (ierr /= 0) call GracefulExit("padeops_b200: divergence failed", ierr)
    end subroutine

    subroutine filter3D(decomp, fil, arr, numtimes, x_bc_, y_bc_, z_bc_)                ! operators.F90:158
        type(decomp_info), intent(in) :: decomp
        type(filters),     intent(in) :: fil
        real(rkind), dimension(decomp%ysz(1), decomp%ysz(2), decomp%ysz(3)), intent(inout), target :: arr
        integer, optional, intent(in) :: numtimes
        integer, dimension(2), optional, intent(in) :: x_bc_, y_bc_, z_bc_
        integer(c_int), dimension(2), target :: x_bc, y_bc, z_bc
        integer(c_int) :: ierr, times2fil
        times2fil = 1; if (present(numtimes)) times2fil = numtimes
        x_bc = 0; if (present(x_bc_)) x_bc = x_bc_
        y_bc = 0; if (present(y_bc_)) y_bc = y_bc_
        z_bc = 0; if (present(z_bc_)) z_bc = z_bc_
        call ensure_handle(decomp)
        ierr = pdo_operators_filter3d(hops, fil%h, c_loc(arr), times2fil, c_loc(x_bc), c_loc(y_bc), c_loc(z_bc), c_null_ptr)
        if (ierr /= 0) call GracefulExit("padeops_b200: filter3D failed", ierr)
    end subroutine

end module
