// spectral_internal.cuh — device-pointer entry points of decomp.cu / spectral.cu used by the other translation units.
#pragma once
#include <cuda_runtime.h>

#include "../../include/padeops_b200.h"
#include "fft2d.cuh"

namespace pdo {
// dir: 0 x->y, 1 y->x, 2 y->z, 3 z->y; w = doubles per element (1 real, 2 complex)
int decomp_transpose_device(pdo_decomp_t h, int dir, const double* src, double* dst, int w, cudaStream_t st);
// Collective (all ranks, same order): makes a device buffer writable by peer GPUs so that transposes INTO it take the
// fused NVLink path.  No-op on one rank or when peer access is unavailable.
void comm_register_buffer_quiet(void* p, size_t bytes);
void comm_deregister_buffer(void* p);  // local; call before freeing a registered buffer
// Collective symmetric allocation (zeroed; peers[world rank] = that rank's copy mapped here); -1 when peer access is unavailable
int comm_sym_alloc(size_t bytes, void** local, void** peers);
void comm_sym_free(void* p);
// peer-writable library-owned memory: pooled symmetric allocation with > 1 rank and peer access (collective, same order on every
// rank; never cudaFree'd while a peer may still map it), cudaMalloc / cudaFree otherwise
int comm_shared_malloc(void** p, size_t bytes);
void comm_shared_free(void* p);
void comm_info(int* rank, int* nproc, int* p2p);
int comm_allreduce_sum(double* dev, int count, cudaStream_t st);   // in place, stream-ordered
void decomp_grid(pdo_decomp_t h, int* p_row, int* p_col, int* c1, int* c2);
int fft3d_forward_xy(pdo_fft3d_t f, const double* in_real_x, double2* out_cplx_y, cudaStream_t st);
int fft3d_backward_yx(pdo_fft3d_t f, const double2* in_cplx_y, double* out_real_x, bool set_oddball, cudaStream_t st);
int fft3d_backward_yx_scratch(pdo_fft3d_t f, double2* prescaled_scratch_cplx_y, double* out_real_x, cudaStream_t st);
int fft3d_z_inplace(pdo_fft3d_t f, double2* a_cplx_z, int dir, cudaStream_t st);
// fused forms (hand-written passes when the shape allows, pointwise pass + cuFFT otherwise; see spectral.cu)
int fft3d_backward_yx_mul(pdo_fft3d_t f, const double2* in_cplx_y, int which, const double* ktab, double scale, bool set_oddball,
                          double* out_real_x, cudaStream_t st);
int fft3d_forward_xy_pro(pdo_fft3d_t f, const RealPro& pro, double2* out_cplx_y, cudaStream_t st);
int fft3d_z_pro(pdo_fft3d_t f, const double2* in, double2* out, int dir, const double* gx, const double* gy, const double* gz, double scale,
                cudaStream_t st);
int fft3d_z_fused(pdo_fft3d_t f, const double2* in, double2* out, int dir, FftPro pro, cudaStream_t st);
bool fft3d_own_z(pdo_fft3d_t f);
bool fft3d_own_xy(pdo_fft3d_t f);
// c2c along the slowest index of an array (nz, cols) with ANY column count (the real z-Fourier procedures transform pairs of
// real columns as one complex column); the plan is cached in `p` and rebuilt when the shape changes.
struct ZColsPlan { int plan = -1; long long cols = 0; int nz = 0; };
int zcols_exec(ZColsPlan* p, int nz, long long cols, double2* a, int dir, cudaStream_t st);
void zcols_destroy(ZColsPlan* p);
pdo_decomp_t fft3d_phys_decomp(pdo_fft3d_t f);
pdo_decomp_t fft3d_spec_decomp(pdo_fft3d_t f);
}  // namespace pdo
