! Drop-in for the two decomp_2d_io entry points the incompressible solver uses for restart and field files
! (2D» io.f90 generics decomp_2d_write_one / decomp_2d_read_one; io_write_one.f90:23-80, io_read_one.f90): one distributed
! array <-> the flat global (nx, ny, nz) array in Fortran order, native doubles, no header.  Files written by either
! implementation are read by the other.  Arrays may live on the host or on the device (INTEGRATION.md 4).
module decomp_2d_io
    use iso_c_binding
    use padeops_b200_c
    use decomp_2d, only: DECOMP_INFO, decomp_main, mytype
    implicit none
    private
    public :: decomp_2d_write_one, decomp_2d_read_one

    interface decomp_2d_write_one
        module procedure write_one_real, write_one_complex
    end interface
    interface decomp_2d_read_one
        module procedure read_one_real, read_one_complex
    end interface

contains

    subroutine write_one_real(ipencil, var, filename, opt_decomp)
        integer, intent(in) :: ipencil
        real(mytype), dimension(:,:,:), intent(in), target, contiguous :: var
        character(len=*), intent(in) :: filename
        type(DECOMP_INFO), intent(in), optional :: opt_decomp
        integer(c_int) :: ierr
        type(c_ptr) :: h
        h = decomp_main%h; if (present(opt_decomp)) h = opt_decomp%h
        ierr = pdo_decomp_write_one(h, int(ipencil, c_int), c_loc(var), 1_c_int, trim(filename)//c_null_char)
        if (ierr /= 0) stop "padeops_b200: decomp_2d_write_one failed"
    end subroutine

    subroutine write_one_complex(ipencil, var, filename, opt_decomp)
        integer, intent(in) :: ipencil
        complex(mytype), dimension(:,:,:), intent(in), target, contiguous :: var
        character(len=*), intent(in) :: filename
        type(DECOMP_INFO), intent(in), optional :: opt_decomp
        integer(c_int) :: ierr
        type(c_ptr) :: h
        h = decomp_main%h; if (present(opt_decomp)) h = opt_decomp%h
        ierr = pdo_decomp_write_one(h, int(ipencil, c_int), c_loc(var), 2_c_int, trim(filename)//c_null_char)
        if (ierr /= 0) stop "padeops_b200: decomp_2d_write_one failed"
    end subroutine

    subroutine read_one_real(ipencil, var, filename, opt_decomp)
        integer, intent(in) :: ipencil
        real(mytype), dimension(:,:,:), intent(inout), target, contiguous :: var
        character(len=*), intent(in) :: filename
        type(DECOMP_INFO), intent(in), optional :: opt_decomp
        integer(c_int) :: ierr
        type(c_ptr) :: h
        h = decomp_main%h; if (present(opt_decomp)) h = opt_decomp%h
        ierr = pdo_decomp_read_one(h, int(ipencil, c_int), c_loc(var), 1_c_int, trim(filename)//c_null_char)
        if (ierr /= 0) stop "padeops_b200: decomp_2d_read_one failed"
    end subroutine

    subroutine read_one_complex(ipencil, var, filename, opt_decomp)
        integer, intent(in) :: ipencil
        complex(mytype), dimension(:,:,:), intent(inout), target, contiguous :: var
        character(len=*), intent(in) :: filename
        type(DECOMP_INFO), intent(in), optional :: opt_decomp
        integer(c_int) :: ierr
        type(c_ptr) :: h
        h = decomp_main%h; if (present(opt_decomp)) h = opt_decomp%h
        ierr = pdo_decomp_read_one(h, int(ipencil, c_int), c_loc(var), 2_c_int, trim(filename)//c_null_char)
        if (ierr /= 0) stop "padeops_b200: decomp_2d_read_one failed"
    end subroutine

end module decomp_2d_io
