! padeops_b200_c.F90 — ISO_C_BINDING interface block for libpadeops_b200.so (include/padeops_b200.h).
! Not compiled in the build image (no Fortran compiler there); kept in lock-step with the header by
! tests/test_abi.py::test_fortran_shims_bind_declared_symbols.
module padeops_b200_c
    use iso_c_binding
    implicit none
    interface
        function pdo_malloc(dptr, bytes) bind(C, name="pdo_malloc") result(ierr)
            import :: c_ptr, c_size_t, c_int
            type(c_ptr), intent(out) :: dptr
            integer(c_size_t), value :: bytes
            integer(c_int) :: ierr
        end function
        function pdo_free(dptr) bind(C, name="pdo_free") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: dptr
            integer(c_int) :: ierr
        end function
        function pdo_h2d(dst, src, bytes, stream) bind(C, name="pdo_h2d") result(ierr)
            import :: c_ptr, c_size_t, c_int
            type(c_ptr), value :: dst, src, stream
            integer(c_size_t), value :: bytes
            integer(c_int) :: ierr
        end function
        function pdo_d2h(dst, src, bytes, stream) bind(C, name="pdo_d2h") result(ierr)
            import :: c_ptr, c_size_t, c_int
            type(c_ptr), value :: dst, src, stream
            integer(c_size_t), value :: bytes
            integer(c_int) :: ierr
        end function
        function pdo_stream_sync(stream) bind(C, name="pdo_stream_sync") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: stream
            integer(c_int) :: ierr
        end function

        ! ---- cd10 ----
        function pdo_cd10_init(h, n, dx, periodic, bc1, bcn) bind(C, name="pdo_cd10_init") result(ierr)
            import :: c_ptr, c_int, c_double
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: n, periodic, bc1, bcn
            real(c_double), value :: dx
            integer(c_int) :: ierr
        end function
        function pdo_cd10_destroy(h) bind(C, name="pdo_cd10_destroy") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int) :: ierr
        end function
#define PDO_LINE_FN(NAME) \
        function NAME(h, f, df, na, nb, bc1, bcn, stream) bind(C, name=#NAME) result(ierr); \
            import :: c_ptr, c_int; \
            type(c_ptr), value :: h, f, df, stream; \
            integer(c_int), value :: na, nb, bc1, bcn; \
            integer(c_int) :: ierr; \
        end function
        PDO_LINE_FN(pdo_cd10_dd1)
        PDO_LINE_FN(pdo_cd10_dd2)
        PDO_LINE_FN(pdo_cd10_dd3)
        PDO_LINE_FN(pdo_cd10_d2d1)
        PDO_LINE_FN(pdo_cd10_d2d2)
        PDO_LINE_FN(pdo_cd10_d2d3)
        ! ---- cd06 ----
        function pdo_cd06_init(h, n, dx, periodic, bc1, bcn) bind(C, name="pdo_cd06_init") result(ierr)
            import :: c_ptr, c_int, c_double
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: n, periodic, bc1, bcn
            real(c_double), value :: dx
            integer(c_int) :: ierr
        end function
        function pdo_cd06_destroy(h) bind(C, name="pdo_cd06_destroy") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int) :: ierr
        end function
        PDO_LINE_FN(pdo_cd06_dd1)
        PDO_LINE_FN(pdo_cd06_dd2)
        PDO_LINE_FN(pdo_cd06_dd3)
        ! ---- cf90 / gaussian ----
        function pdo_cf90_init(h, n, periodic) bind(C, name="pdo_cf90_init") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: n, periodic
            integer(c_int) :: ierr
        end function
        function pdo_cf90_destroy(h) bind(C, name="pdo_cf90_destroy") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int) :: ierr
        end function
        PDO_LINE_FN(pdo_cf90_filter1)
        PDO_LINE_FN(pdo_cf90_filter2)
        PDO_LINE_FN(pdo_cf90_filter3)
        function pdo_gaussian_init(h, n, periodic) bind(C, name="pdo_gaussian_init") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: n, periodic
            integer(c_int) :: ierr
        end function
        function pdo_gaussian_destroy(h) bind(C, name="pdo_gaussian_destroy") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int) :: ierr
        end function
        PDO_LINE_FN(pdo_gaussian_filter1)
        PDO_LINE_FN(pdo_gaussian_filter2)
        PDO_LINE_FN(pdo_gaussian_filter3)
        ! ---- cd06stagg ----
        function pdo_cd06stagg_init_periodic(h, n, dx) bind(C, name="pdo_cd06stagg_init_periodic") result(ierr)
            import :: c_ptr, c_int, c_double
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: n
            real(c_double), value :: dx
            integer(c_int) :: ierr
        end function
        function pdo_cd06stagg_destroy(h) bind(C, name="pdo_cd06stagg_destroy") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int) :: ierr
        end function
#define PDO_STAGG_FN(NAME) \
        function NAME(h, fin, fout, n1, n2, is_complex, stream) bind(C, name=#NAME) result(ierr); \
            import :: c_ptr, c_int; \
            type(c_ptr), value :: h, fin, fout, stream; \
            integer(c_int), value :: n1, n2, is_complex; \
            integer(c_int) :: ierr; \
        end function
        PDO_STAGG_FN(pdo_cd06stagg_ddz_E2C)
        PDO_STAGG_FN(pdo_cd06stagg_ddz_C2E)
        PDO_STAGG_FN(pdo_cd06stagg_interpz_E2C)
        PDO_STAGG_FN(pdo_cd06stagg_interpz_C2E)
        PDO_STAGG_FN(pdo_cd06stagg_d2dz2_C2C)
        PDO_STAGG_FN(pdo_cd06stagg_d2dz2_E2E)
        ! ---- decomp_2d ----
        function pdo_comm_unique_id(id) bind(C, name="pdo_comm_unique_id") result(ierr)
            import :: c_char, c_int
            character(kind=c_char), intent(out) :: id(128)
            integer(c_int) :: ierr
        end function
        function pdo_comm_init(rank, nproc, id) bind(C, name="pdo_comm_init") result(ierr)
            import :: c_char, c_int
            integer(c_int), value :: rank, nproc
            character(kind=c_char), intent(in) :: id(128)
            integer(c_int) :: ierr
        end function
        function pdo_comm_finalize() bind(C, name="pdo_comm_finalize") result(ierr)
            import :: c_int
            integer(c_int) :: ierr
        end function
        function pdo_decomp_init(h, nx, ny, nz, p_row, p_col) bind(C, name="pdo_decomp_init") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: nx, ny, nz, p_row, p_col
            integer(c_int) :: ierr
        end function
        function pdo_decomp_destroy(h) bind(C, name="pdo_decomp_destroy") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int) :: ierr
        end function
        function pdo_decomp_get_info(h, info) bind(C, name="pdo_decomp_get_info") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int), intent(out) :: info(27)   ! xst,xen,xsz,yst,yen,ysz,zst,zen,zsz (3 each)
            integer(c_int) :: ierr
        end function
#define PDO_TRANSPOSE_FN(NAME) \
        function NAME(h, src, dst, elem_doubles, stream) bind(C, name=#NAME) result(ierr); \
            import :: c_ptr, c_int; \
            type(c_ptr), value :: h, src, dst, stream; \
            integer(c_int), value :: elem_doubles; \
            integer(c_int) :: ierr; \
        end function
        PDO_TRANSPOSE_FN(pdo_transpose_x_to_y)
        PDO_TRANSPOSE_FN(pdo_transpose_y_to_x)
        PDO_TRANSPOSE_FN(pdo_transpose_y_to_z)
        PDO_TRANSPOSE_FN(pdo_transpose_z_to_y)
        function pdo_decomp_write_one(h, ipencil, var, elem_doubles, filename) bind(C, name="pdo_decomp_write_one") result(ierr)
            import :: c_ptr, c_int, c_char
            type(c_ptr), value :: h, var
            integer(c_int), value :: ipencil, elem_doubles
            character(kind=c_char), dimension(*), intent(in) :: filename
            integer(c_int) :: ierr
        end function
        function pdo_decomp_read_one(h, ipencil, var, elem_doubles, filename) bind(C, name="pdo_decomp_read_one") result(ierr)
            import :: c_ptr, c_int, c_char
            type(c_ptr), value :: h, var
            integer(c_int), value :: ipencil, elem_doubles
            character(kind=c_char), dimension(*), intent(in) :: filename
            integer(c_int) :: ierr
        end function
        function pdo_p_maxval(xloc, xglob) bind(C, name="pdo_p_maxval") result(ierr)
            import :: c_double, c_int
            real(c_double), value :: xloc
            real(c_double), intent(out) :: xglob
            integer(c_int) :: ierr
        end function
        function pdo_p_sum(xloc, xglob) bind(C, name="pdo_p_sum") result(ierr)
            import :: c_double, c_int
            real(c_double), value :: xloc
            real(c_double), intent(out) :: xglob
            integer(c_int) :: ierr
        end function
        ! ---- fft_3d / PoissonPeriodic ----
        function pdo_fft3d_init(h, nx, ny, nz, dx, dy, dz, p_row, p_col) bind(C, name="pdo_fft3d_init") result(ierr)
            import :: c_ptr, c_int, c_double
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: nx, ny, nz, p_row, p_col
            real(c_double), value :: dx, dy, dz
            integer(c_int) :: ierr
        end function
        function pdo_fft3d_destroy(h) bind(C, name="pdo_fft3d_destroy") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int) :: ierr
        end function
#define PDO_FFT_FN(NAME) \
        function NAME(h, fin, fout, stream) bind(C, name=#NAME) result(ierr); \
            import :: c_ptr, c_int; \
            type(c_ptr), value :: h, fin, fout, stream; \
            integer(c_int) :: ierr; \
        end function
        PDO_FFT_FN(pdo_fft3d_fft3_x2z)
        PDO_FFT_FN(pdo_fft3d_ifft3_z2x)
        PDO_FFT_FN(pdo_fft3d_fft2_x2y)
        function pdo_fft3d_ifft2_y2x(h, fin, fout, set_oddball, stream) bind(C, name="pdo_fft3d_ifft2_y2x") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h, fin, fout, stream
            integer(c_int), value :: set_oddball
            integer(c_int) :: ierr
        end function
        function pdo_poisson_init(h, nx, ny, nz, dx, dy, dz, p_row, p_col, dir_id, modkx, modky, modkz) &
                 bind(C, name="pdo_poisson_init") result(ierr)
            import :: c_ptr, c_int, c_double
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: nx, ny, nz, p_row, p_col, dir_id
            real(c_double), value :: dx, dy, dz
            type(c_ptr), value :: modkx, modky, modkz
            integer(c_int) :: ierr
        end function
        function pdo_poisson_destroy(h) bind(C, name="pdo_poisson_destroy") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int) :: ierr
        end function
        PDO_FFT_FN(pdo_poisson_solve)
        function pdo_operators_init(h, gp, dx, dy, dz, method, allow_zslab) bind(C, name="pdo_operators_init") result(ierr)
            import :: c_ptr, c_int, c_double, c_char
            type(c_ptr), intent(out) :: h
            type(c_ptr), value :: gp
            real(c_double), value :: dx, dy, dz
            character(kind=c_char), dimension(*), intent(in) :: method
            integer(c_int), value :: allow_zslab
            integer(c_int) :: ierr
        end function
        function pdo_operators_destroy(h) bind(C, name="pdo_operators_destroy") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int) :: ierr
        end function
        function pdo_operators_gradient(h, f, dfdx, dfdy, dfdz, stream) bind(C, name="pdo_operators_gradient") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h, f, dfdx, dfdy, dfdz, stream
            integer(c_int) :: ierr
        end function
        function pdo_operators_curl(h, u, v, w, curlu, stream) bind(C, name="pdo_operators_curl") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h, u, v, w, curlu, stream
            integer(c_int) :: ierr
        end function
        function pdo_operators_divergence(h, u, v, w, div, stream) bind(C, name="pdo_operators_divergence") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h, u, v, w, div, stream
            integer(c_int) :: ierr
        end function
        function pdo_operators_filter3d(h, fil, arr, numtimes, x_bc, y_bc, z_bc, stream) bind(C, name="pdo_operators_filter3d") result(ierr)
            import :: c_ptr, c_int
            type(c_ptr), value :: h, fil, arr, stream
            integer(c_int), value :: numtimes
            integer(c_int), dimension(2), intent(in) :: x_bc, y_bc, z_bc
            integer(c_int) :: ierr
        end function
    end interface
end module padeops_b200_c
