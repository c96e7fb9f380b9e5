! Drop-in replacement for src/derivatives/cd10.F90: same module name, type name and type-bound procedure
! names / argument lists (cd10.F90:108-183, 195, 2029-2447); the body forwards to libpadeops_b200.so.
! Host arrays are passed with c_loc (the library stages them); a caller that keeps fields on the device
! passes the device address instead via the *_dev specifics (type(c_ptr) arguments).
module cd10stuff
    use kind_parameters, only: rkind
    use exits,           only: GracefulExit
    use iso_c_binding
    use padeops_b200_c
    implicit none
    private
    public :: cd10

    type cd10
        private
        integer     :: n = 0
        type(c_ptr) :: h = c_null_ptr
        type(c_ptr), public :: stream = c_null_ptr   ! cudaStream_t to enqueue on (default stream if null)
    contains
        procedure :: init
        procedure :: destroy
        procedure :: GetSize
        procedure :: dd1
        procedure :: dd2
        procedure :: dd3
        procedure :: d2d1
        procedure :: d2d2
        procedure :: d2d3
        procedure :: dd1_dev      ! device-resident fields: f, df are device addresses
        procedure :: dd2_dev
        procedure :: dd3_dev
    end type

contains

    function init(this, n_, dx_, periodic_, bc1_, bcn_) result(ierr)
        class(cd10), intent(inout) :: this
        integer, intent(in) :: n_, bc1_, bcn_
        real(rkind), intent(in) :: dx_
        logical, intent(in) :: periodic_
        integer :: ierr
        this%n = n_
        ierr = pdo_cd10_init(this%h, int(n_, c_int), real(dx_, c_double), merge(1_c_int, 0_c_int, periodic_), &
                             int(bc1_, c_int), int(bcn_, c_int))
    end function

    subroutine destroy(this)
        class(cd10), intent(inout) :: this
        integer :: ierr
        ierr = pdo_cd10_destroy(this%h)
        this%h = c_null_ptr
    end subroutine

    pure function GetSize(this) result(val)
        class(cd10), intent(in) :: this
        integer :: val
        val = this%n
    end function

#define CD10_HOST_FN(NAME, CNAME, D1, D2, D3) \
    subroutine NAME(this, f, df, na, nb, bc1_, bcn_); \
        class(cd10), intent(in) :: this; \
        integer, intent(in) :: na, nb; \
        integer, optional, intent(in) :: bc1_, bcn_; \
        real(rkind), dimension(D1, D2, D3), intent(in), target  :: f; \
        real(rkind), dimension(D1, D2, D3), intent(out), target :: df; \
        integer :: bc1, bcn, ierr; \
        bc1 = 0; bcn = 0; \
        if (present(bc1_)) bc1 = bc1_; \
        if (present(bcn_)) bcn = bcn_; \
        ierr = CNAME(this%h, c_loc(f), c_loc(df), int(na, c_int), int(nb, c_int), int(bc1, c_int), int(bcn, c_int), this%stream); \
        if (ierr /= 0) call GracefulExit("padeops_b200: cd10 call failed", ierr); \
    end subroutine
    CD10_HOST_FN(dd1,  pdo_cd10_dd1,  this%n, na, nb)
    CD10_HOST_FN(dd2,  pdo_cd10_dd2,  na, this%n, nb)
    CD10_HOST_FN(dd3,  pdo_cd10_dd3,  na, nb, this%n)
    CD10_HOST_FN(d2d1, pdo_cd10_d2d1, this%n, na, nb)
    CD10_HOST_FN(d2d2, pdo_cd10_d2d2, na, this%n, nb)
    CD10_HOST_FN(d2d3, pdo_cd10_d2d3, na, nb, this%n)

#define CD10_DEV_FN(NAME, CNAME) \
    subroutine NAME(this, f, df, na, nb); \
        class(cd10), intent(in) :: this; \
        type(c_ptr), intent(in) :: f, df; \
        integer, intent(in) :: na, nb; \
        integer :: ierr; \
        ierr = CNAME(this%h, f, df, int(na, c_int), int(nb, c_int), 0_c_int, 0_c_int, this%stream); \
        if (ierr /= 0) call GracefulExit("padeops_b200: cd10 call failed", ierr); \
    end subroutine
    CD10_DEV_FN(dd1_dev, pdo_cd10_dd1)
    CD10_DEV_FN(dd2_dev, pdo_cd10_dd2)
    CD10_DEV_FN(dd3_dev, pdo_cd10_dd3)

end module cd10stuff
