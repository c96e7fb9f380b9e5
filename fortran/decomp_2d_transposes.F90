! Drop-in for the 2DECOMP&FFT entry points the hot path uses (2D» decomp_2d.f90:150-205, transpose_*.f90):
! decomp_2d_init, decomp_info_init, get_decomp_info, transpose_{x_to_y,y_to_x,y_to_z,z_to_y} (real/complex
! generics), nrank, nproc, xstart/xend/xsize ...  MPI is still what launches the ranks (one per GPU); it is used
! once, to broadcast the 128-byte NCCL id.  All data movement goes through libpadeops_b200.so.
module decomp_2d
    use iso_c_binding
    use padeops_b200_c
    use mpi
    implicit none
    private
    integer, parameter, public :: mytype = kind(0.d0)
    integer, save, public :: nrank, nproc
    integer, save, public, dimension(3) :: xstart, xend, xsize, ystart, yend, ysize, zstart, zend, zsize

    type, public :: DECOMP_INFO
        integer, dimension(3) :: xst, xen, xsz, yst, yen, ysz, zst, zen, zsz
        type(c_ptr) :: h = c_null_ptr
    end type
    type(DECOMP_INFO), save, public :: decomp_main
    integer, save :: p_row_ = 0, p_col_ = 0

    public :: decomp_2d_init, decomp_2d_finalize, decomp_info_init, decomp_info_finalize, get_decomp_info
    public :: transpose_x_to_y, transpose_y_to_x, transpose_y_to_z, transpose_z_to_y

    interface transpose_x_to_y
        module procedure transpose_x_to_y_real, transpose_x_to_y_complex
    end interface
    interface transpose_y_to_x
        module procedure transpose_y_to_x_real, transpose_y_to_x_complex
    end interface
    interface transpose_y_to_z
        module procedure transpose_y_to_z_real, transpose_y_to_z_complex
    end interface
    interface transpose_z_to_y
        module procedure transpose_z_to_y_real, transpose_z_to_y_complex
    end interface

contains

    subroutine decomp_2d_init(nx, ny, nz, p_row, p_col, periodic_bc)
        integer, intent(in) :: nx, ny, nz, p_row, p_col
        logical, dimension(3), intent(in), optional :: periodic_bc
        character(kind=c_char) :: id(128)
        integer :: ierr
        call MPI_COMM_RANK(MPI_COMM_WORLD, nrank, ierr)
        call MPI_COMM_SIZE(MPI_COMM_WORLD, nproc, ierr)
        if (nrank == 0) ierr = pdo_comm_unique_id(id)
        call MPI_BCAST(id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierr)
        ierr = pdo_comm_init(int(nrank, c_int), int(nproc, c_int), id)
        p_row_ = p_row; p_col_ = p_col      ! 0,0 → the library picks 1 x nproc (results never depend on the grid)
        call decomp_info_init(nx, ny, nz, decomp_main)
        xstart = decomp_main%xst; xend = decomp_main%xen; xsize = decomp_main%xsz
        ystart = decomp_main%yst; yend = decomp_main%yen; ysize = decomp_main%ysz
        zstart = decomp_main%zst; zend = decomp_main%zen; zsize = decomp_main%zsz
    end subroutine

    subroutine decomp_2d_finalize
        integer :: ierr
        call decomp_info_finalize(decomp_main)
        ierr = pdo_comm_finalize()
    end subroutine

    subroutine decomp_info_init(nx, ny, nz, decomp)
        integer, intent(in) :: nx, ny, nz
        type(DECOMP_INFO), intent(inout) :: decomp
        integer(c_int) :: info(27), ierr
        ierr = pdo_decomp_init(decomp%h, int(nx, c_int), int(ny, c_int), int(nz, c_int), int(p_row_, c_int), int(p_col_, c_int))
        if (ierr /= 0) call MPI_ABORT(MPI_COMM_WORLD, ierr, ierr)
        ierr = pdo_decomp_get_info(decomp%h, info)
        decomp%xst = info(1:3);   decomp%xen = info(4:6);   decomp%xsz = info(7:9)
        decomp%yst = info(10:12); decomp%yen = info(13:15); decomp%ysz = info(16:18)
        decomp%zst = info(19:21); decomp%zen = info(22:24); decomp%zsz = info(25:27)
    end subroutine

    subroutine decomp_info_finalize(decomp)
        type(DECOMP_INFO), intent(inout) :: decomp
        integer :: ierr
        ierr = pdo_decomp_destroy(decomp%h)
        decomp%h = c_null_ptr
    end subroutine

    subroutine get_decomp_info(decomp)
        type(DECOMP_INFO), intent(out) :: decomp
        decomp = decomp_main
    end subroutine

#define TRANSPOSE_PAIR(BASE, CNAME) \
    subroutine BASE##_real(src, dst, opt_decomp); \
        real(mytype), dimension(:,:,:), intent(in), target, contiguous  :: src; \
        real(mytype), dimension(:,:,:), intent(out), target, contiguous :: dst; \
        type(DECOMP_INFO), intent(in), optional :: opt_decomp; \
        integer :: ierr; \
        if (present(opt_decomp)) then; \
            ierr = CNAME(opt_decomp%h, c_loc(src), c_loc(dst), 1_c_int, c_null_ptr); \
        else; \
            ierr = CNAME(decomp_main%h, c_loc(src), c_loc(dst), 1_c_int, c_null_ptr); \
        end if; \
        if (ierr /= 0) call MPI_ABORT(MPI_COMM_WORLD, ierr, ierr); \
    end subroutine; \
    subroutine BASE##_complex(src, dst, opt_decomp); \
        complex(mytype), dimension(:,:,:), intent(in), target, contiguous  :: src; \
        complex(mytype), dimension(:,:,:), intent(out), target, contiguous :: dst; \
        type(DECOMP_INFO), intent(in), optional :: opt_decomp; \
        integer :: ierr; \
        if (present(opt_decomp)) then; \
            ierr = CNAME(opt_decomp%h, c_loc(src), c_loc(dst), 2_c_int, c_null_ptr); \
        else; \
            ierr = CNAME(decomp_main%h, c_loc(src), c_loc(dst), 2_c_int, c_null_ptr); \
        end if; \
        if (ierr /= 0) call MPI_ABORT(MPI_COMM_WORLD, ierr, ierr); \
    end subroutine
    TRANSPOSE_PAIR(transpose_x_to_y, pdo_transpose_x_to_y)
    TRANSPOSE_PAIR(transpose_y_to_x, pdo_transpose_y_to_x)
    TRANSPOSE_PAIR(transpose_y_to_z, pdo_transpose_y_to_z)
    TRANSPOSE_PAIR(transpose_z_to_y, pdo_transpose_z_to_y)

end module decomp_2d
