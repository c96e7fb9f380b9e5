! Drop-in for src/utilities/PoissonPeriodic.F90: module PoissonPeriodicMod, type PoissonPeriodic with
! init / poisson_solve (generic: in-place and out-of-place) / destroy (PoissonPeriodic.F90:27-34).
module PoissonPeriodicMod
    use kind_parameters, only: rkind
    use decomp_2d,       only: decomp_info
    use exits,           only: GracefulExit
    use iso_c_binding
    use padeops_b200_c
    implicit none
    private
    public :: PoissonPeriodic

    type :: PoissonPeriodic
        private
        type(c_ptr) :: h = c_null_ptr
        integer :: nx_in, ny_in, nz_in
    contains
        procedure :: init
        procedure, private :: poisson_solve_inplace
        procedure, private :: poisson_solve_outofplace
        generic :: poisson_solve => poisson_solve_inplace, poisson_solve_outofplace
        procedure :: destroy
    end type

contains

    subroutine init(this, dx, dy, dz, gp, dir_id, useExhaustiveFFT, Get_ModKx, Get_ModKy, Get_ModKz)
        class(PoissonPeriodic), intent(inout) :: this
        real(rkind), intent(in) :: dx, dy, dz
        type(decomp_info), intent(in) :: gp
        integer, intent(in) :: dir_id
        logical, intent(in), optional :: useExhaustiveFFT
        external :: Get_ModKx, Get_ModKy, Get_ModKz
        optional :: Get_ModKx, Get_ModKy, Get_ModKz
        real(rkind), allocatable, target :: kx(:), ky(:), kz(:)
        type(c_ptr) :: pkx, pky, pkz
        integer :: nx, ny, nz, ierr
        nx = gp%xsz(1); ny = gp%ysz(2); nz = gp%zsz(3)
        pkx = c_null_ptr; pky = c_null_ptr; pkz = c_null_ptr
        ! the reference hands k*d to the callback and divides by d afterwards (PoissonPeriodic.F90:181-204)
        if (present(Get_ModKx)) then
            allocate(kx(nx)); kx = GetWaveNums(nx, dx)*dx; call Get_ModKx(kx); kx = kx/dx; pkx = c_loc(kx)
        end if
        if (present(Get_ModKy)) then
            allocate(ky(ny)); ky = GetWaveNums(ny, dy)*dy; call Get_ModKy(ky); ky = ky/dy; pky = c_loc(ky)
        end if
        if (present(Get_ModKz)) then
            allocate(kz(nz)); kz = GetWaveNums(nz, dz)*dz; call Get_ModKz(kz); kz = kz/dz; pkz = c_loc(kz)
        end if
        ierr = pdo_poisson_init(this%h, int(nx, c_int), int(ny, c_int), int(nz, c_int), real(dx, c_double), real(dy, c_double), &
                                real(dz, c_double), 0_c_int, 0_c_int, int(dir_id, c_int), pkx, pky, pkz)
        if (ierr /= 0) call GracefulExit("Couldn't initialize 3d FFT inside SPECTRAL derived type", 123)
        select case (dir_id)
        case (1); this%nx_in = gp%xsz(1); this%ny_in = gp%xsz(2); this%nz_in = gp%xsz(3)
        case (2); this%nx_in = gp%ysz(1); this%ny_in = gp%ysz(2); this%nz_in = gp%ysz(3)
        end select
    end subroutine

    subroutine poisson_solve_inplace(this, rhs)
        class(PoissonPeriodic), intent(inout) :: this
        real(rkind), dimension(this%nx_in, this%ny_in, this%nz_in), intent(inout), target :: rhs
        integer :: ierr
        ierr = pdo_poisson_solve(this%h, c_loc(rhs), c_loc(rhs), c_null_ptr)
        if (ierr /= 0) call GracefulExit("padeops_b200: poisson_solve failed", ierr)
    end subroutine

    subroutine poisson_solve_outofplace(this, rhs, f)
        class(PoissonPeriodic), intent(inout) :: this
        real(rkind), dimension(this%nx_in, this%ny_in, this%nz_in), intent(in), target  :: rhs
        real(rkind), dimension(this%nx_in, this%ny_in, this%nz_in), intent(out), target :: f
        integer :: ierr
        ierr = pdo_poisson_solve(this%h, c_loc(rhs), c_loc(f), c_null_ptr)
        if (ierr /= 0) call GracefulExit("padeops_b200: poisson_solve failed", ierr)
    end subroutine

    subroutine destroy(this)
        class(PoissonPeriodic), intent(inout) :: this
        integer :: ierr
        ierr = pdo_poisson_destroy(this%h)
        this%h = c_null_ptr
    end subroutine

    pure function GetWaveNums(nx, dx) result(k)
        integer, intent(in) :: nx
        real(rkind), intent(in) :: dx
        real(rkind), dimension(nx) :: k
        real(rkind), parameter :: pi = 3.141592653589793238462643383279502884197_rkind
        integer :: i, even, h
        even = nx - mod(nx, 2)
        h = merge(nx/2, (nx + 1)/2 - 1, mod(nx, 2) == 0)
        do i = 1, nx
            k(i) = (-pi + real(mod(i - 1 + h, nx), rkind)*2._rkind*pi/real(even, rkind))/dx
        end do
    end function

end module PoissonPeriodicMod
