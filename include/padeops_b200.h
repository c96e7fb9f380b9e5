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
#define PDO_E_UNSUPPORTED 1002   /* a branch SURVEY.md §8 marks out of scope (non-periodic closures, ...) */
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
/* cd10%init(n, dx, periodic, bc1, bcn)                                     cd10.F90:195-315 */
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
int pdo_cd06_init(pdo_cd06_t* h, int n, double dx, int periodic, int bc1, int bcn);   /* cd06.F90:129-219 */
int pdo_cd06_destroy(pdo_cd06_t h);
int pdo_cd06_getsize(pdo_cd06_t h);
int pdo_cd06_dd1(pdo_cd06_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream); /* :775 */
int pdo_cd06_dd2(pdo_cd06_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream); /* :797 */
int pdo_cd06_dd3(pdo_cd06_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream); /* :819 */

/* ---- cf90stuff::cf90  (filters/cf90.F90) ----------------------------------------------------- */
typedef struct pdo_cf90_s* pdo_cf90_t;
int pdo_cf90_init(pdo_cf90_t* h, int n, int periodic);                                 /* cf90.F90:107-196 */
int pdo_cf90_destroy(pdo_cf90_t h);
int pdo_cf90_filter1(pdo_cf90_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :1020 */
int pdo_cf90_filter2(pdo_cf90_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :1090 */
int pdo_cf90_filter3(pdo_cf90_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :1160 */

/* ---- gaussianstuff::gaussian  (filters/gaussian.F90) ----------------------------------------- */
typedef struct pdo_gaussian_s* pdo_gaussian_t;
int pdo_gaussian_init(pdo_gaussian_t* h, int n, int periodic);                          /* gaussian.F90:74-102 */
int pdo_gaussian_destroy(pdo_gaussian_t h);
int pdo_gaussian_filter1(pdo_gaussian_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :104 */
int pdo_gaussian_filter2(pdo_gaussian_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :336 */
int pdo_gaussian_filter3(pdo_gaussian_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream); /* :564 */

/* ---- cd06staggstuff::cd06stagg, periodic  (derivatives/cd06stagg.F90) ------------------------ */
typedef struct pdo_cd06stagg_s* pdo_cd06stagg_t;
int pdo_cd06stagg_init_periodic(pdo_cd06stagg_t* h, int n, double dx);                  /* cd06stagg.F90:170-195 */
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

/* ---- FiltersMod::filters  (filters/filters.F90) ---------------------------------------------- */
typedef struct pdo_filters_s* pdo_filters_t;
/* filters%init(gp, periodicx,y,z, methodx,y,z)   filters.F90:274-297; methods "cf90" | "gaussian" */
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
/* reductions::p_maxval / p_sum of one double over all ranks              utilities/reductions.F90:29-225 */
int pdo_p_maxval(double local, double* global);
int pdo_p_sum(double local, double* global);

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
/* init(dx,dy,dz,gp,dir_id[,...,Get_ModKx,Get_ModKy,Get_ModKz]); dir_id 1 (x-pencil in/out) or 2 (y-pencil).
   modk{x,y,z}: optional full-length (nx, ny, nz) arrays of ALREADY MODIFIED wavenumbers replacing
   GetWaveNums output (what the Get_ModK* callbacks produce, :181-204); NULL = spectral wavenumbers. */
int pdo_poisson_init(pdo_poisson_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col,
                     int dir_id, const double* modkx, const double* modky, const double* modkz);
int pdo_poisson_destroy(pdo_poisson_t h);
/* poisson_solve(rhs, f): out-of-place (:62-87); f == rhs → in-place specific (:37-60) */
int pdo_poisson_solve(pdo_poisson_t h, const double* rhs, double* f, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PADEOPS_B200_H */
