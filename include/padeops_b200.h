/*
 * padeops_b200.h — C ABI of the B200-native PadeOps operator hot path (libpadeops_b200.so).
 *
 * The reference has no FFI layer: its boundary is the set of Fortran module types with type-bound
 * procedures.  Each entry point below replaces one of those procedures and cites it
 * (paths relative to the reference's src/; "2D»" = dependencies/2decomp_fft-1.5.847.tar.gz » src).
 * fortran/ *.F90 holds the ISO_C_BINDING shim modules that keep the reference's module / type /
 * procedure names on top of this ABI; INTEGRATION.md shows how a maintainer links them.
 *
 * Conventions
 *   - arrays are column-major (first index fastest), contiguous, double precision; complex = (re,im) pairs;
 *   - field pointers may be DEVICE pointers (fields stay resident between calls: the fast path) or HOST
 *     pointers (the library stages H2D / D2H through an internal pool: the drop-in path);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are stream-ordered for
 *     device pointers and synchronous for host pointers;
 *   - every function returns int: 0 = ok; the reference's own error codes are preserved where it has
 *     them (cd10 init → 2, cd06 → 3, cf90 → 7, bad bc → 324, cd06stagg n<=4 → 21, bad 2D grid → 6);
 *     PDO_E_* for everything else; pdo_last_error() returns a message for the calling thread;
 *   - input and output must not alias (the reference's intent(in)/intent(out) contract).
 */
#ifndef PADEOPS_B200_H
#define PADEOPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDO_OK 0
#define PDO_E_BADARG 1001        /* null handle, bad axis, bad size */
#define PDO_E_UNSUPPORTED 1002   /* a branch SURVEY.md §8 marks out of scope (CD06 / staggered non-periodic closures, ...) */
#define PDO_E_CUDA 1003          /* a CUDA / cuFFT / NCCL call failed; see pdo_last_error() */
#define PDO_E_NODEVICE 1004      /* no CUDA device: this library has no CPU fallback */

const char* pdo_last_error(void);
int pdo_version(void);
/* number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
int64_t pdo_launch_count(void);

/* ---- device memory helpers so Fortran can hold type(c_ptr) device fields -------------------- */
int pdo_malloc(void** dptr, size_t bytes);
int pdo_free(void* dptr);
int pdo_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream);
int pdo_d2h(void* dst_host, const void* src_dev, size_t bytes, void* stream);
int pdo_stream_sync(void* stream);

/* ---- cd10stuff::cd10  (derivatives/cd10.F90) ------------------------------------------------- */
typedef struct pdo_cd10_s* pdo_cd10_t;
/* cd10%init(n, dx, periodic, bc1, bcn)                                     cd10.F90:195-315
 * periodic = 1: cyclic pentadiagonal operators (the fast chunked kernels).  periodic = 0: the non-periodic closures
 * (cd10.F90:29-96, 429-707), all nine tables built like the reference; dd* / d2d* then select by the bc1, bcn passed at the
 * call (0 one-sided, 1 symmetric, -1 antisymmetric; anything else -> 324).  Correctness path (one thread per line). */
int pdo_cd10_init(pdo_cd10_t* h, int n, double dx, int periodic, int bc1, int bcn);
int pdo_cd10_destroy(pdo_cd10_t h);                                      /* cd10.F90:317-341 */
int pdo_cd10_getsize(pdo_cd10_t h);                                      /* GetSize */
/* dd1/dd2/dd3(f, df, na, nb, bc1, bcn): f(n,na,nb) / f(na,n,nb) / f(na,nb,n)   cd10.F90:2029-2237 */
int pdo_cd10_dd1(pdo_cd10_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream);
int pdo_cd10_dd2(pdo_cd10_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream);
int pdo_cd10_dd3(pdo_cd10_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream);
/* d2d1/d2d2/d2d3                                                          cd10.F90:2239-2447 */
int pdo_cd10_d2d1(pdo_cd10_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream);
int pdo_cd10_d2d2(pdo_cd10_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream);
int pdo_cd10_d2d3(pdo_cd10_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream);

/* ---- cd06stuff::cd06  (derivatives/cd06.F90) ------------------------------------------------- */
typedef struct pdo_cd06_s* pdo_cd06_t;
int pdo_cd06_init(pdo_cd06_t* h, int n, double dx, int periodic, int bc1, int bcn);   /* cd06.F90:129-219; periodic = 0: the
                                                                                          one-sided closure (bc1 = bcn = 0 only) */
int pdo_cd06_destroy(pdo_cd06_t h);
int pdo_cd06_getsize(pdo_cd06_t h);
int pdo_cd06_dd1(pdo_cd06_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream); /* :775 */
int pdo_cd06_dd2(pdo_cd06_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream); /* :797 */
int pdo_cd06_dd3(pdo_cd06_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream); /* :819 */

/* ---- cf90stuff::cf90  (filters/cf90.F90) ----------------------------------------------------- */
typedef struct pdo_cf90_s* pdo_cf90_t;
int pdo_cf90_init(pdo_cf90_t* h, int n, int periodic);                                 /* cf90.F90:107-196; periodic = 0: the
                                                                                          closures of cf90.F90:24-47, 276-418 */
int pdo_cf90_destroy(pdo_cf90_t h);
int pdo_cf90_filter1(pdo_cf90_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :1020 */
int pdo_cf90_filter2(pdo_cf90_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :1090 */
int pdo_cf90_filter3(pdo_cf90_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :1160 */

/* ---- gaussianstuff::gaussian  (filters/gaussian.F90) ----------------------------------------- */
typedef struct pdo_gaussian_s* pdo_gaussian_t;
/* periodic = 0: the explicit boundary rows b1..b4 (bc 0) or the interior stencil on the even / odd reflection (bc +1 / -1),
   gaussian.F90:22-46, 215-330; needs n >= 8 */
int pdo_gaussian_init(pdo_gaussian_t* h, int n, int periodic);                          /* gaussian.F90:74-102 */
int pdo_gaussian_destroy(pdo_gaussian_t h);
int pdo_gaussian_filter1(pdo_gaussian_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :104 */
int pdo_gaussian_filter2(pdo_gaussian_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :336 */
int pdo_gaussian_filter3(pdo_gaussian_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :564 */

/* ---- kernel planning (no counterpart in the reference's operators; the analogue of FFTW's planner, which the reference runs
 *      inside fft_3d%init with FFTW_MEASURE / FFTW_EXHAUSTIVE, fft_3d.F90:150-158) ------------------------------------------
 * Without a plan every operator call picks its kernel from a fixed table: same kernel on every box, nothing timed, nothing
 * synchronised inside a call (stream order and graph capture are safe from the first call on).  pdo_*_plan times the kernel
 * candidates for ONE shape — axis 0 / 1 / 2 = x / y / z, (na, nb) as the dd* / filter* calls take them — on scratch arrays,
 * stores the winner in the handle and reports its code in *variant (may be NULL).  It synchronises the device; call it at
 * set-up time.  All candidates agree to rounding, so a plan changes the speed, not the result beyond 1e-15. */
int pdo_cd10_plan(pdo_cd10_t h, int axis, int na, int nb, int* variant_d1, int* variant_d2);
int pdo_cd06_plan(pdo_cd06_t h, int axis, int na, int nb, int* variant);
int pdo_cf90_plan(pdo_cf90_t h, int axis, int na, int nb, int* variant);
int pdo_gaussian_plan(pdo_gaussian_t h, int axis, int na, int nb, int* variant);
/* enable != 0: the first large call of an operator on a shape plans by itself (same as the environment variable PDO_TUNE=1) */
int pdo_plan_on_first_call(int enable);

/* ---- cd06staggstuff::cd06stagg, periodic  (derivatives/cd06stagg.F90) ------------------------ */
typedef struct pdo_cd06stagg_s* pdo_cd06stagg_t;
int pdo_cd06stagg_init_periodic(pdo_cd06stagg_t* h, int n, double dx);                  /* cd06stagg.F90:170-195 */
/* init(nx, dx, isTopEven, isBotEven, isTopSided, isBotSided): walls in z, cd06stagg.F90:197-231 + the STAGG_CD06_files includes.
   The FIELD is even / odd about a wall (is*Even), or that wall takes the one-sided closure (is*Sided).  Correctness path:
   one thread per z-line, right-hand side formed inside the Thomas sweep (csrc/stagg_np.cu). */
int pdo_cd06stagg_init_nonperiodic(pdo_cd06stagg_t* h, int nx, double dx, int is_top_even, int is_bot_even, int is_top_sided,
                                   int is_bot_sided);
int pdo_cd06stagg_destroy(pdo_cd06stagg_t h);
/* (in, out, n1, n2); cells have n planes, edges n+1.  `is_complex` = the CMPLX specific of the generic.
   ddz_E2C :820-848  ddz_C2E :850-881  InterpZ_E2C :928-956  InterpZ_C2E :958-993
   d2dz2_C2C :995-1023  d2dz2_E2E :1025-1059 */
int pdo_cd06stagg_ddz_E2C(pdo_cd06stagg_t h, const double* fE, double* dfC, int n1, int n2, int is_complex, void* stream);
int pdo_cd06stagg_ddz_C2E(pdo_cd06stagg_t h, const double* fC, double* dfE, int n1, int n2, int is_complex, void* stream);
int pdo_cd06stagg_interpz_E2C(pdo_cd06stagg_t h, const double* fE, double* fC, int n1, int n2, int is_complex, void* stream);
int pdo_cd06stagg_interpz_C2E(pdo_cd06stagg_t h, const double* fC, double* fE, int n1, int n2, int is_complex, void* stream);
int pdo_cd06stagg_d2dz2_C2C(pdo_cd06stagg_t h, const double* fC, double* d2fC, int n1, int n2, int is_complex, void* stream);
int pdo_cd06stagg_d2dz2_E2E(pdo_cd06stagg_t h, const double* fE, double* d2fE, int n1, int n2, int is_complex, void* stream);
/* collocated first derivatives on the staggered grids (cd06stagg.F90:883-925); non-periodic handles only, as in the reference */
int pdo_cd06stagg_ddz_C2C(pdo_cd06stagg_t h, const double* fC, double* dfC, int n1, int n2, int is_complex, void* stream);
int pdo_cd06stagg_ddz_E2E(pdo_cd06stagg_t h, const double* fE, double* dfE, int n1, int n2, int is_complex, void* stream);

/* ---- DerivativesMod::derivatives  (derivatives/derivatives.F90) ------------------------------ */
typedef struct pdo_derivatives_s* pdo_derivatives_t;
/* derivatives%init(gp, dx,dy,dz, periodicx,y,z, methodx,y,z)              derivatives.F90:189-236
   xsz/ysz/zsz = gp%xsz, gp%ysz, gp%zsz (local pencil sizes); method strings "cd10" | "cd06". */
int pdo_derivatives_init(pdo_derivatives_t* h, const int xsz[3], const int ysz[3], const int zsz[3], double dx, double dy,
                         double dz, int periodicx, int periodicy, int periodicz, const char* methodx, const char* methody,
                         const char* methodz);
int pdo_derivatives_destroy(pdo_derivatives_t h);
int pdo_derivatives_ddx(pdo_derivatives_t h, const double* f, double* dfdx, int bc1, int bcn, void* stream);   /* :447-468 */
int pdo_derivatives_ddy(pdo_derivatives_t h, const double* f, double* dfdy, int bc1, int bcn, void* stream);   /* :470-491 */
int pdo_derivatives_ddz(pdo_derivatives_t h, const double* f, double* dfdz, int bc1, int bcn, void* stream);   /* :493-513 */
int pdo_derivatives_d2dx2(pdo_derivatives_t h, const double* f, double* d2f, int bc1, int bcn, void* stream);  /* :515-532 */
int pdo_derivatives_d2dy2(pdo_derivatives_t h, const double* f, double* d2f, int bc1, int bcn, void* stream);  /* :534-550 */
int pdo_derivatives_d2dz2(pdo_derivatives_t h, const double* f, double* d2f, int bc1, int bcn, void* stream);  /* :552-569 */

/* ---- lstsqstuff::lstsq  (filters/lstsq.F90): explicit 9-point least-squares filter --------------- */
typedef struct pdo_lstsq_s* pdo_lstsq_t;
int pdo_lstsq_init(pdo_lstsq_t* h, int n, int periodic);   /* periodic = 0: one-sided rows b1..b4 at both ends (:169-212), n >= 8 */
int pdo_lstsq_destroy(pdo_lstsq_t h);
int pdo_lstsq_filter1(pdo_lstsq_t h, const double* f, double* fil, int na, int nb, void* stream);
int pdo_lstsq_filter2(pdo_lstsq_t h, const double* f, double* fil, int na, int nb, void* stream);
int pdo_lstsq_filter3(pdo_lstsq_t h, const double* f, double* fil, int na, int nb, void* stream);

/* ---- FiltersMod::filters  (filters/filters.F90) ---------------------------------------------- */
typedef struct pdo_filters_s* pdo_filters_t;
/* filters%init(gp, periodicx,y,z, methodx,y,z)   filters.F90:274-297; methods "cf90" | "gaussian" | "lstsq" (else 52) */
int pdo_filters_init(pdo_filters_t* h, const int xsz[3], const int ysz[3], const int zsz[3], int periodicx, int periodicy,
                     int periodicz, const char* methodx, const char* methody, const char* methodz);
int pdo_filters_destroy(pdo_filters_t h);
int pdo_filters_filterx(pdo_filters_t h, const double* f, double* fil, int bc1, int bcn, void* stream);        /* :220-235 */
int pdo_filters_filtery(pdo_filters_t h, const double* f, double* fil, int bc1, int bcn, void* stream);        /* :237-252 */
int pdo_filters_filterz(pdo_filters_t h, const double* f, double* fil, int bc1, int bcn, void* stream);        /* :254-269 */

/* ---- decomp_2d  (2D» decomp_2d.f90, transpose_*.f90) ----------------------------------------- */
typedef struct pdo_decomp_s* pdo_decomp_t;
typedef struct {
    int xst[3], xen[3], xsz[3];
    int yst[3], yen[3], ysz[3];
    int zst[3], zen[3], zsz[3];
} pdo_decomp_info;
/* Communicator bootstrap replacing MPI_Init + MPI_CART_CREATE (2D» decomp_2d.f90:297-369).  One process
   per GPU.  `unique_id` = 128 bytes from pdo_comm_unique_id() on rank 0, broadcast by the host program
   (MPI_Bcast in Fortran, torch.distributed in the Python harness).  nproc == 1 needs no id. */
int pdo_comm_unique_id(char id[128]);
int pdo_comm_init(int rank, int nproc, const char unique_id[128]);
int pdo_comm_finalize(void);
/* Collective over all ranks, same order everywhere: registers a device buffer (CUDA IPC) so that transposes whose
   DESTINATION lies inside it run as one fused pack + NVLink store + unpack kernel instead of pack / NCCL / unpack.
   The destination pointer must be the registered base pointer itself.  Unregistered
   destinations (and PDO_P2P=0) use the NCCL path; results are bit-identical either way. */
int pdo_comm_register_buffer(void* dev_ptr, size_t bytes);
/* Local: forget a registered buffer.  MUST be called before the buffer is freed (a later allocation at the same address
   would otherwise be mistaken for it on this rank only, and the ranks would disagree on the path). */
int pdo_comm_deregister_buffer(void* dev_ptr);
int pdo_comm_p2p_enabled(void);
int pdo_comm_rank(void);   /* nrank */
int pdo_comm_size(void);   /* nproc */
/* decomp_info_init(nx,ny,nz,decomp) on a p_row x p_col grid (decomp_2d_init's grid; p_row*p_col must equal
   nproc; 0,0 picks 1 x nproc — the reference's auto-tune is timing-based and results never depend on it,
   SURVEY A.7 #10)                                                          2D» decomp_2d.f90:498-580 */
int pdo_decomp_init(pdo_decomp_t* h, int nx, int ny, int nz, int p_row, int p_col);
int pdo_decomp_destroy(pdo_decomp_t h);
int pdo_decomp_get_info(pdo_decomp_t h, pdo_decomp_info* info);          /* get_decomp_info :477-486 */
/* pure arithmetic, no communicator needed: what rank `rank` of a p_row x p_col grid would own */
int pdo_decomp_info_for(int nx, int ny, int nz, int p_row, int p_col, int rank, pdo_decomp_info* info);
/* transpose_a_to_b(src, dst, decomp); elem_doubles = 1 (real specific) or 2 (complex specific)
   2D» transpose_x_to_y.f90:14-91,172-249  transpose_y_to_x.f90  transpose_y_to_z.f90:14-100  transpose_z_to_y.f90 */
int pdo_transpose_x_to_y(pdo_decomp_t h, const double* src, double* dst, int elem_doubles, void* stream);
int pdo_transpose_y_to_x(pdo_decomp_t h, const double* src, double* dst, int elem_doubles, void* stream);
int pdo_transpose_y_to_z(pdo_decomp_t h, const double* src, double* dst, int elem_doubles, void* stream);
int pdo_transpose_z_to_y(pdo_decomp_t h, const double* src, double* dst, int elem_doubles, void* stream);
/* decomp_2d_io: decomp_2d_write_one(ipencil, var, filename, opt_decomp) / decomp_2d_read_one       2D» io_write_one.f90:23-80, io_read_one.f90
   ONE distributed array <-> the flat GLOBAL (nx, ny, nz) array in Fortran order, native doubles, no header (the format of the
   reference's restart and field dumps).  ipencil 1 / 2 / 3 = var is this rank's x- / y- / z-pencil; host or device pointer;
   elem_doubles 1 (real) or 2 (complex).  Collective; the file is truncated first; a missing file on read -> 321. */
int pdo_decomp_write_one(pdo_decomp_t h, int ipencil, const double* var, int elem_doubles, const char* filename);
int pdo_decomp_read_one(pdo_decomp_t h, int ipencil, double* var, int elem_doubles, const char* filename);
/* the per-rank piece (pure host code, no communicator): sub-box (sub, st 0-based) of the global (sizes) array; create != 0
   truncates / creates the file first (rank 0's job in the collective call) */
int pdo_io_write_block(const char* filename, const int sizes[3], const int sub[3], const int st[3], int elem_doubles, const double* data,
                       int create);
int pdo_io_read_block(const char* filename, const int sizes[3], const int sub[3], const int st[3], int elem_doubles, double* data);
/* Fortran's G15.5 edit descriptor (the restart info file's format, igrid.F90:2797); out holds 15 characters + NUL */
int pdo_io_format_g15_5(double x, char out[16]);
/* reductions::p_maxval / p_sum of one double over all ranks              utilities/reductions.F90:29-225 */
int pdo_p_maxval(double local, double* global);
int pdo_p_sum(double local, double* global);

/* ---- operators  (utilities/operators.F90:17-151): vector calculus on y-pencil fields ----------
 * gradient(decomp, der, f, dfdx, dfdy, dfdz), curl(decomp, der, u, v, w, curlu), divergence(decomp, der, u, v, w, div):
 * every argument is a y-pencil array of `gp`; periodic, method "cd10" | "cd06" on all three axes.  Where the reference
 * transposes to the pencil of the axis and back (operators.F90:43-51), this library does the same on its own
 * transposes EXCEPT along z when the z-slabs are even multiples of 32 planes and the GPUs have peer access: then the
 * compact solve itself is distributed (neighbours exchange a few halo planes and the edge pieces of the reduced system
 * over NVLink; no transpose).  allow_zslab = 0 forces the reference's choreography.  Collective over all ranks. */
typedef struct pdo_operators_s* pdo_operators_t;
int pdo_operators_init(pdo_operators_t* h, pdo_decomp_t gp, double dx, double dy, double dz, const char* method, int allow_zslab);
int pdo_operators_destroy(pdo_operators_t h);
int pdo_operators_zmode(pdo_operators_t h);  /* 0: z local on this grid, 1: distributed z-slab solve, 2: transposes */
int pdo_operators_ddx(pdo_operators_t h, const double* f, double* dfdx, void* stream);
int pdo_operators_ddy(pdo_operators_t h, const double* f, double* dfdy, void* stream);
int pdo_operators_ddz(pdo_operators_t h, const double* f, double* dfdz, void* stream);
int pdo_operators_gradient(pdo_operators_t h, const double* f, double* dfdx, double* dfdy, double* dfdz, void* stream);      /* :17-53 */
int pdo_operators_curl(pdo_operators_t h, const double* u, const double* v, const double* w, double* curlu, void* stream);   /* :55-116 */
int pdo_operators_divergence(pdo_operators_t h, const double* u, const double* v, const double* w, double* div, void* stream); /* :118-151 */
/* filter3D(decomp, fil, arr, numtimes, x_bc_, y_bc_, z_bc_)                                            operators.F90:158-224
   `numtimes` passes of fil%filtery, then (through y->x / x->y) of fil%filterx, then (through y->z / z->y) of fil%filterz, in
   place on the y-pencil field `arr` (device pointer).  `fil` must have been built from the same decomposition (else code 234).
   x_bc / y_bc / z_bc: int[2] boundary codes handed to the filters, or NULL for (0, 0). */
int pdo_operators_filter3d(pdo_operators_t h, pdo_filters_t fil, double* arr, int numtimes, const int* x_bc, const int* y_bc,
                           const int* z_bc, void* stream);

/* ---- fft_3d_stuff::fft_3d, "x" base pencil  (utilities/fft_3d.F90) ---------------------------- */
typedef struct pdo_fft3d_s* pdo_fft3d_t;
/* fft_3d%init(nx,ny,nz,"x",dx,dy,dz,...) on the communicator's grid       fft_3d.F90:109-468 */
int pdo_fft3d_init(pdo_fft3d_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col);
int pdo_fft3d_destroy(pdo_fft3d_t h);
int pdo_fft3d_get_complex_output_size(pdo_fft3d_t h, int sz[3]);         /* z-pencil of the spectral decomp */
int pdo_fft3d_get_physical_info(pdo_fft3d_t h, pdo_decomp_info* info);   /* decomp of (nx,ny,nz) */
int pdo_fft3d_get_spectral_info(pdo_fft3d_t h, pdo_decomp_info* info);   /* link_spectral_gp */
int pdo_fft3d_fft3_x2z(pdo_fft3d_t h, const double* in_real_x, double* out_cplx_z, void* stream);   /* :588-613 */
int pdo_fft3d_ifft3_z2x(pdo_fft3d_t h, const double* in_cplx_z, double* out_real_x, void* stream);  /* :670-696 */
int pdo_fft3d_fft2_x2y(pdo_fft3d_t h, const double* in_real_x, double* out_cplx_y, void* stream);   /* :645-663 */
/* set_oddball != 0 zeroes the x-Nyquist mode before the c2r pass, as the reference's setOddBall does */
int pdo_fft3d_ifft2_y2x(pdo_fft3d_t h, const double* in_cplx_y, double* out_real_x, int set_oddball, void* stream);  /* :616-643 */

/* ---- PoissonPeriodicMod::PoissonPeriodic  (utilities/PoissonPeriodic.F90) --------------------- */
typedef struct pdo_poisson_s* pdo_poisson_t;
/* init(dx,dy,dz,gp,dir_id[,...,Get_ModKx,Get_ModKy,Get_ModKz]); dir_id 1 (x-pencil in/out), 2 (y-pencil) or 3 (z-pencil:
   brought to the x-pencil and back instead of the reference's z-base transforms; same solution); else -> 31243.
   modk{x,y,z}: optional full-length (nx, ny, nz) arrays of ALREADY MODIFIED wavenumbers replacing
   GetWaveNums output (what the Get_ModK* callbacks produce, :181-204); NULL = spectral wavenumbers. */
int pdo_poisson_init(pdo_poisson_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col,
                     int dir_id, const double* modkx, const double* modky, const double* modkz);
int pdo_poisson_destroy(pdo_poisson_t h);
/* poisson_solve(rhs, f): out-of-place (:62-87); f == rhs → in-place specific (:37-60) */
int pdo_poisson_solve(pdo_poisson_t h, const double* rhs, double* f, void* stream);


/* ---- spectralMod::spectral, pencil "x", dimTransform = 2  (incompressible/spectral.F90) --------
   The igrid flavour: 2-D FFTs in (x, y) per z-plane; real fields live in x-pencils of (nx,ny,nz), transforms in
   y-pencils of the spectral decomposition (nx/2+1, ny, nz).  init_periodicInZ adds the z-periodic procedures
   (3-D dealiasing through c2c-z, edge-field dealiasing, z FFTs).  For an edge-grid type pass nz+1 as nz and
   init_periodicInZ = 0, as igrid does for spectE (igrid.F90:487-495). */
typedef struct pdo_spectral_s* pdo_spectral_t;
/* spectral%init("x", nx,ny,nz, dx,dy,dz, "four", filt, 2, fixOddball, ..., init_periodicInZ, dealiasF)   spectral.F90:867-930 */
int pdo_spectral_init(pdo_spectral_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col,
                      int fix_oddball, int init_periodic_in_z, double dealias_fact);
int pdo_spectral_destroy(pdo_spectral_t h);
int pdo_spectral_get_physical_info(pdo_spectral_t h, pdo_decomp_info* info);   /* physdecomp  */
int pdo_spectral_get_spectral_info(pdo_spectral_t h, pdo_decomp_info* info);   /* spectdecomp */
int pdo_spectral_fft(pdo_spectral_t h, const double* in_real_x, double* out_cplx_y, void* stream);                   /* :1413-1429 */
int pdo_spectral_ifft(pdo_spectral_t h, const double* in_cplx_y, double* out_real_x, int set_oddball, void* stream);  /* :1431-1453 */
int pdo_spectral_mtimes_ik1_oop(pdo_spectral_t h, const double* fin, double* fout, void* stream);   /* :235-253 */
int pdo_spectral_mtimes_ik2_oop(pdo_spectral_t h, const double* fin, double* fout, void* stream);   /* :255-273 */
int pdo_spectral_mtimes_ik1_ip(pdo_spectral_t h, double* f, void* stream);                          /* :276-293 */
int pdo_spectral_mtimes_ik2_ip(pdo_spectral_t h, double* f, void* stream);                          /* :295-312 */
int pdo_spectral_dealias(pdo_spectral_t h, double* fhat_cplx_y, void* stream);                      /* :314-341 */
/* z-pencil edge field with nz+1 planes, dealiased with this (cell, periodic) type's tables          :343-363 */
int pdo_spectral_dealias_edgefield(pdo_spectral_t h, double* fhatE_cplx_z, void* stream);
int pdo_spectral_take_fft1d_z2z_ip(pdo_spectral_t h, double* a_cplx_z, void* stream);               /* :1483-1487 */
int pdo_spectral_take_ifft1d_z2z_ip(pdo_spectral_t h, double* a_cplx_z, void* stream);              /* :1489-1494 */
/* z-Fourier operators of the periodic-in-z type.  Complex arrays: z-pencil of the spectral decomposition; the real one: z-pencil
   of the physical decomposition, transformed r2c / c2r in the reference with "the oddball ignored": modes 0 .. nz/2-1 are
   multiplied, the Nyquist mode goes back as it came (here: pairs of real columns through one c2c, same result). */
int pdo_spectral_ddz_c2c_real_ip(pdo_spectral_t h, double* a_real_z, void* stream);                 /* :507-526 */
int pdo_spectral_ddz_c2c_complex_ip(pdo_spectral_t h, double* a_cplx_z, void* stream);              /* :528-547 */
int pdo_spectral_shiftz_e2c(pdo_spectral_t h, double* ahat_cplx_z, void* stream);                   /* :409-422, array already z-transformed */
int pdo_spectral_shiftz_c2e(pdo_spectral_t h, double* ahat_cplx_z, void* stream);                   /* :424-437 */
/* the 1-D tables behind k1 / k2 / kabs_sq / Gdealias: full global length (nx/2+1, ny, nz); NULL entries are skipped */
int pdo_spectral_get_tables(pdo_spectral_t h, double* k1, double* k2, double* gdealias_x, double* gdealias_y, double* gdealias_z);

/* ---- PadeDerOps::Pade6stagg  (incompressible/PadeDerOps.F90) --------------------------------- */
typedef struct pdo_pade6stagg_s* pdo_pade6stagg_t;
#define PDO_SCHEME_FD02 0
#define PDO_SCHEME_CD06 1
#define PDO_SCHEME_FOURIER 2
/* Pade6stagg%init(gpC, sp_gpC, gpE, sp_gpE, dz, scheme, isPeriodic, spectC)                        PadeDerOps.F90:57-88
   gp_zsz / sp_zsz: z-pencil sizes of the physical / spectral CELL decompositions.  scheme: cd06 here (fourierColl needs
   the spectC of pdo_pade6stagg_init2 and returns code 43 through this entry; fd02 returns PDO_E_UNSUPPORTED). */
int pdo_pade6stagg_init(pdo_pade6stagg_t* h, const int gp_zsz[3], const int sp_zsz[3], double dz, int scheme, int is_periodic);
/* the reference's init takes the optional spectC (PadeDerOps.F90:57-78): required for scheme = fourierColl (else code 43), where
   every operator becomes c2c-z forward, x table(k3), c2c-z backward, x 1/nz (spectral.F90:387-680, 843-856; complex arrays only) */
int pdo_pade6stagg_init2(pdo_pade6stagg_t* h, const int gp_zsz[3], const int sp_zsz[3], double dz, int scheme, int is_periodic,
                         struct pdo_spectral_s* spectC);
int pdo_pade6stagg_destroy(pdo_pade6stagg_t h);
/* generic over real (is_complex = 0, sizes from gp) / complex (is_complex = 1, sizes from sp_gp).  isPeriodic = .true.: the
   bot / top integers are accepted and ignored, as in the reference (:146-160, 404-418, 572-585, 689-702, 879-892).
   isPeriodic = .false. (scheme cd06; any other scheme -> 323): init builds the nine wall operators derOO .. derSS (:92-110) and
   bot / top select one per call: -1 the field is odd about that wall, +1 even, 0 one-sided closure (first-order operators
   only); any other combination gives output = 0, as the reference's select-case does (:185-205, 449-482). */
int pdo_pade6stagg_ddz_C2E(pdo_pade6stagg_t h, const double* in, double* out, int is_complex, int bot, int top, void* stream);
int pdo_pade6stagg_ddz_E2C(pdo_pade6stagg_t h, const double* in, double* out, int is_complex, int bot, int top, void* stream);
int pdo_pade6stagg_interpz_C2E(pdo_pade6stagg_t h, const double* in, double* out, int is_complex, int bot, int top, void* stream);
int pdo_pade6stagg_interpz_E2C(pdo_pade6stagg_t h, const double* in, double* out, int is_complex, int bot, int top, void* stream);
int pdo_pade6stagg_d2dz2_C2C(pdo_pade6stagg_t h, const double* in, double* out, int is_complex, int bot, int top, void* stream);
int pdo_pade6stagg_d2dz2_E2E(pdo_pade6stagg_t h, const double* in, double* out, int is_complex, int bot, int top, void* stream);
int pdo_pade6stagg_get_modified_wavenumbers(pdo_pade6stagg_t h, const double* k, double* kp, int n);   /* :997-1053 */

/* ---- PadePoissonMod::padepoisson  (incompressible/PadePoisson.F90) ----------------------------- */
typedef struct pdo_padepoisson_s* pdo_padepoisson_t;
/* padepoisson%init(dx,dy,dz, sp, spE, computeStokesPressure=F, Lz, storePressure, gpC, derivZ, PeriodicInZ=T)   :130-180, 76-128 */
int pdo_padepoisson_init(pdo_padepoisson_t* h, double dx, double dy, double dz, pdo_spectral_t sp, pdo_spectral_t spE,
                         pdo_pade6stagg_t derivZ);
/* the same with the reference's PeriodicInZ argument.  PeriodicInZ = .false., computeStokesPressure = .false. (:180-230, 459-623):
   walls at both ends of z — PressureProjection extends the horizontal divergence evenly and w oddly to 2 nz planes, solves in
   z-Fourier space with the z scheme's modified wavenumber and the half-cell shifts, and leaves w = 0 on both walls;
   DivergenceCheck uses derivZ%ddz_E2C(-1, -1).  derivZ must have been initialised with the same periodicity; computeStokesPressure: init3.  The pressure
   getters run with walls too (:762-896, 963-1160): getPressure projects copies of its intent(in) arguments and returns
   phat (+ phat_z1 + phat_z2 with computeStokesPressure, the pieces GetStokesPressure :641-714 keeps); getPressureAndUpdateRHS
   projects in place and, as in the reference, adds the Stokes pieces of the LAST getPressure call (:1146-1156; zero before any). */
int pdo_padepoisson_init2(pdo_padepoisson_t* h, double dx, double dy, double dz, pdo_spectral_t sp, pdo_spectral_t spE,
                          pdo_pade6stagg_t derivZ, int periodic_in_z);
/* ... and with computeStokesPressure and Lz (:232-296, 320-384, 444-458, 597-609; walls only): before the projection the harmonic
   pressure chat cosh(lambda (Lz - z)) / chat cosh(lambda z), lambda = |(k1, k2)|, cancels w on the bottom and then the top wall,
   so w* need not vanish there on input. */
int pdo_padepoisson_init3(pdo_padepoisson_t* h, double dx, double dy, double dz, pdo_spectral_t sp, pdo_spectral_t spE,
                          pdo_pade6stagg_t derivZ, int periodic_in_z, int compute_stokes_pressure, double Lz);
int pdo_padepoisson_destroy(pdo_padepoisson_t h);
/* uhat, vhat: complex y-pencils of sp; what: complex y-pencil of spE (nz+1 planes); all updated in place   :386-432 */
int pdo_padepoisson_pressure_projection(pdo_padepoisson_t h, double* uhat, double* vhat, double* what, void* stream);
int pdo_padepoisson_get_pressure(pdo_padepoisson_t h, const double* uhat, const double* vhat, const double* what,
                                 double* pressure_real_x, void* stream);                                     /* :716-750 */
int pdo_padepoisson_get_pressure_and_update_rhs(pdo_padepoisson_t h, double* uhat, double* vhat, double* what,
                                                double* pressure_real_x, void* stream);                      /* :900-949 */
/* divergence: real x-pencil out; fix_div != 0 re-projects as the reference does when max(div) > 1e-13 / 1e-10;
   max_div (optional) receives p_maxval(maxval(divergence)) of the last evaluation                          :1165-1244 */
int pdo_padepoisson_divergence_check(pdo_padepoisson_t h, double* uhat, double* vhat, double* what, double* divergence_real_x,
                                     int fix_div, double* max_div, void* stream);

/* ---- IncompressibleGrid::igrid, the RK substep  (incompressible/igrid.F90) ----------------------
   Scope: PeriodicInZ (the hot path) or walls in z with slip / no-slip stencils (get_boundary_conditions_stencil :5148-5204, wall
   closures of the staggered operators, even / odd Poisson solver with the Stokes-pressure step); NumericalSchemeVert = 1 (CD06)
   or 2 (Fourier collocation, periodic only); AdvectionTerm = 1 (skew-symmetric) or 0 (rotational); TimeSteppingScheme 1
   (TVD-RK3) or 2 (SSP-RK45); viscous or inviscid; optional SGS model and HIT forcing (periodic only, pdo_igrid_enable_*);
   no Coriolis / stratification / turbines / fringe.
   The namelist file of igrid%init is replaced by this struct. */
typedef struct pdo_igrid_s* pdo_igrid_t;
typedef struct pdo_hit_forcing_s* pdo_hit_forcing_t;   /* the forcing type, declared with its entry points further down */
typedef struct {
    int nx, ny, nz;
    double Lx, Ly, Lz;
    double Re;
    int is_inviscid;
    double dealias_fact;          /* dealiasFact, reference default 2/3 */
    int t_divergence_check;       /* t_DivergenceCheck, reference default 10 */
    int time_stepping_scheme;     /* 1 TVD-RK3, 2 SSP-RK45 */
    int p_row, p_col;             /* 0,0 = 1 x nproc */
    int use_d2dz2_c2c;            /* 1: d2dz2_C2C for the viscous z term (slip/periodic BC codes), 0: ddz_E2C(ddz_C2E) (igrid.F90:2642-2660) */
    int compute_all_gradients;    /* 1: all 18 duidxjC/E fields like the reference; 0: only the 9 the substep reads */
    int rotational_advection;     /* 0: AdvectionTerm = 1, skew-symmetric (igrid.F90:1572-1679); 1: AdvectionTerm = 0, u x omega (:1527-1555) */
    int fourier_collocation_z;    /* 0: NumericalSchemeVert = 1, cd06 staggered operators; 1: NumericalSchemeVert = 2, Fourier collocation in z */
    int wall_bounded;             /* 0: PeriodicInZ = .true.; 1: PeriodicInZ = .false. — walls at z = 0 and z = Lz (&BCs namelist) */
    int top_wall, bot_wall;       /* topWall / botWall when wall_bounded: 1 no-slip, 2 slip (3, the wall model, is out of scope) */
    int no_stokes_pressure;       /* wall_bounded only.  0: ComputeStokesPressure = .true. (the reference's default): the projection first
                                     removes the wall-normal velocity with the harmonic Stokes pressure; 1: .false. */
} pdo_igrid_params;
/* igrid%init: u, v on the cell grid, w on the edge grid (nz+1 planes, plane nz+1 == plane 1), x-pencil local blocks,
   host or device pointers (initfields_wallM is the caller's job).  Runs the fft / dealias / projection / gradient
   sequence of igrid.F90:625-655. */
int pdo_igrid_init(pdo_igrid_t* h, const pdo_igrid_params* p, const double* u, const double* v, const double* w);
int pdo_igrid_destroy(pdo_igrid_t h);
int pdo_igrid_time_advance(pdo_igrid_t h, double dt, void* stream);        /* timeAdvance(dtforced)  :1057-1299 */
/* physical fields (x-pencils): 0 u, 1 v, 2 w (edge), 3 wC, 4 uE, 5 vE, 6 divergence;  spectral (y-pencils, complex):
   10 uhat, 11 vhat, 12 what (edge).  Copies into `out` (host or device). */
int pdo_igrid_get_field(pdo_igrid_t h, int which, double* out, void* stream);
int pdo_igrid_get_decomp_info(pdo_igrid_t h, int which /*0 gpC, 1 gpE, 2 sp_gpC, 3 sp_gpE*/, pdo_decomp_info* info);
int pdo_igrid_get_state(pdo_igrid_t h, int* step, double* tsim);
/* Restart and field files exactly as the reference writes them (x-pencils of gpC / gpE through decomp_2d_write_one):
   dumpRestartFile :2763-2803  ->  <dir>/RESTART_Run<rid>_{u,v,w}.<step, 6 digits> + RESTART_Run<rid>_info.<step> (tsim, g15.5)
   readRestartFile :2719-2761 + init's :589-591, 625-655  ->  step = tid, tsim from the info file, fields projected, state rebuilt
   dumpFullField   :2806-2823  ->  <dir>/Run<rid>_<label>_t<step>.out for field ids 0 u, 1 v, 2 w, 3 wC, 4 uE, 5 vE, 6 divergence */
/* useSGS = .true. (igrid.F90:1866-1871; sgsmod_igrid.F90:156-268, sgs_models/{smagorinsky, sigma, AMD, eddyViscosity}.F90): the
   eddy-viscosity models with a global constant on the periodic box — SGSModelID 0 Smagorinsky, 1 sigma, 2 AMD; Csgs;
   explicitCalcEdgeEddyViscosity (0: nu is interpolated cells -> edges and clipped at zero).  Wall damping, the dynamic
   procedure and wall models are out of scope.  The handle must have been initialised with compute_all_gradients = 1.
   Bad model id -> 213.  The term enters every right-hand side after the viscous term. */
int pdo_igrid_enable_sgs(pdo_igrid_t h, int sgs_model_id, double csgs, int explicit_calc_edge_eddy_viscosity);
/* useHITForcing = .true. (igrid.F90:940-944, 1907-1910): call once after init; the forcing is added to every right-hand side
   after the viscous term, with a new draw at the first stage of every time step (tidStart = the current step) */
int pdo_igrid_enable_hit_forcing(pdo_igrid_t h, double kmin, double kmax, int nwaves, double eps_amplitude, int rand_seed_to_add);
pdo_hit_forcing_t pdo_igrid_hit_forcing(pdo_igrid_t h);   /* borrowed; NULL without forcing.  set_wavenumbers on it before a time step injects that step's draw */
int pdo_igrid_dump_restart(pdo_igrid_t h, const char* outputdir, int run_id);
int pdo_igrid_read_restart(pdo_igrid_t h, const char* inputdir, int run_id, int tid);
int pdo_igrid_dump_full_field(pdo_igrid_t h, int which, const char* label4, const char* outputdir, int run_id);
/* compute_deltaT with useCFL (igrid.F90:1372-1396) */
int pdo_igrid_compute_delta_t(pdo_igrid_t h, double cfl, double* dt, void* stream);
int pdo_igrid_max_divergence(pdo_igrid_t h, double* max_div, void* stream);   /* printDivergence + p_maxval(|div|) */

/* ---- forcingmod::HIT_shell_forcing  (incompressible/forcingIsotropic.F90:45-314) ---------------
   Every time step Nwaves integer wavenumber triplets are drawn on the shell kmin <= |k| <= kmax; each forced mode gets
   f_hat += normfact EpsAmplitude / (|u_hat|^2 + |v_hat|^2 + |w_hat|^2 + 1e-14) / Nwaves * conjg(u_hat) in the fully transformed
   space, w shifted edges -> cells and back.  Evaluated as a direct DFT of the Nwaves forced columns and plane-wave updates of
   the right-hand sides instead of the reference's six whole-field z transforms.  The &HIT_Forcing namelist enters as arguments.
   The random draw: Fortran's random_number is compiler-specific, the library uses SplitMix64 on the reference's seed
   arithmetic (:122-128); set_wavenumbers injects the reference's own draw for an A/B run (kept through the next new time step).  Arrays: DEVICE pointers, complex
   y-pencils of the cell (u, v) and edge (w) spectral decompositions. */
int pdo_hit_forcing_init(pdo_hit_forcing_t* h, pdo_spectral_t spectC, pdo_spectral_t spectE, double kmin, double kmax, int nwaves,
                         double eps_amplitude, int tid_start, int rand_seed_to_add);
int pdo_hit_forcing_destroy(pdo_hit_forcing_t h);
int pdo_hit_forcing_set_wavenumbers(pdo_hit_forcing_t h, const int* wave_x, const int* wave_y, const int* wave_z);
int pdo_hit_forcing_get_wavenumbers(pdo_hit_forcing_t h, int* wave_x, int* wave_y, int* wave_z);
int pdo_hit_forcing_get_rhs(pdo_hit_forcing_t h, double* urhs_xy, double* vrhs_xy, double* wrhs_xy, const double* uhat_xy,
                            const double* vhat_xy, const double* what_xy, int new_timestep, void* stream);   /* :254-311 */

/* ---- igrid_Operators_Periodic::Ops_Periodic  (incompressible/igrid_operators_periodic.F90:13-161) ----------------
   Fourier operators on x-pencil fields of a triply periodic box (the reference's post-processing programs use it):
   its own spectral type (pencil "x", 2-D transforms, init_periodicInZ, 2/3 dealiasing, fixOddball = .false.) and a
   PoissonPeriodic with spectral wavenumbers.  Real arguments are x-pencil arrays of the physical decomposition, the complex
   one a y-pencil array of the spectral decomposition; host or device pointers.  gp enters as its process grid. */
typedef struct pdo_ops_periodic_s* pdo_ops_periodic_t;
int pdo_ops_periodic_init(pdo_ops_periodic_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col);  /* :86-109 */
int pdo_ops_periodic_destroy(pdo_ops_periodic_t h);
pdo_spectral_t pdo_ops_periodic_spect(pdo_ops_periodic_t h);                                          /* link_spect :46-52 */
int pdo_ops_periodic_ddx(pdo_ops_periodic_t h, const double* f, double* dfdx, void* stream);          /* :117-125 */
int pdo_ops_periodic_ddy(pdo_ops_periodic_t h, const double* f, double* dfdy, void* stream);          /* :127-135 */
int pdo_ops_periodic_ddz(pdo_ops_periodic_t h, const double* f, double* dfdz, void* stream);          /* :149-160 */
int pdo_ops_periodic_ddz_cmplx2cmplx(pdo_ops_periodic_t h, double* fhat_cplx_y, void* stream);        /* :137-145 */
int pdo_ops_periodic_solve_poisson(pdo_ops_periodic_t h, const double* rhs, double* p, void* stream); /* _oop :70-76; p == rhs: _ip :78-84 */
int pdo_ops_periodic_dealias_field(pdo_ops_periodic_t h, double* f, void* stream);                    /* :56-62 */
/* WriteField3D :189-205 / ReadField3D :162-187: "<dir>/Run<runID, 2 digits>_<label, 4 chars>_t<tidx, 6 digits>.out" in the
   decomp_2d_io format; a missing file on read -> 321 */
int pdo_ops_periodic_write_field3d(pdo_ops_periodic_t h, const double* field, const char* label4, int tidx, int run_id, const char* outputdir);
int pdo_ops_periodic_read_field3d(pdo_ops_periodic_t h, double* field, const char* label4, int tidx, int run_id, const char* inputdir);

#ifdef __cplusplus
}
#endif
#endif /* PADEOPS_B200_H */
