// fft2d.cuh — hand-written batched FFT passes for the spectral / igrid path (fft2d.cu), with the pointwise work of the callers
// fused into the first load of a pass.  Declarations only; device pointers everywhere.
//
// Replaces, for power-of-two extents on slab grids (p_row == 1), the FFTW plans of utilities/fft_3d.F90:256-306 (r2c / c2r in x,
// strided c2c in y and z) that spectral.cu otherwise hands to cuFFT, and absorbs the passes the reference runs around them:
// the products of igrid.F90:1527-1679 (AddNonLinearTerm_*), mTimes_ik1_oop / mTimes_ik2_oop (spectral.F90:235-341), the 1/(nx ny)
// of ifft2_y2x (fft_3d.F90:633-641) with its oddball zeroing, and the Gdealias multiply of spectral.F90:343-363.
#pragma once
#include <cuda_runtime.h>

namespace pdo {

// First load of a strided c2c pass, with c0 = col % n1, c1 = col / n1:
//   U, V:      v += i (KU[c0] U + KV[c1] V)        (U, V: arrays of the input's shape; the divergence of PadePoisson.F90:392-401)
//   poisson:   v *= (s <= 1e-14 ? 0 : -scale / s), s = A[c0] + B[c1] + C[row]   (kradsq_inv and mfact, PadePoisson.F90:103-108, 413-415)
//   otherwise: v *= scale * A[c0] * B[c1] * C[row]  (null table = 1)
// then * i when times_i, then the column with c0 == nyq is zeroed (nyq < 0: none).
struct FftPro {
    const double* A = nullptr;
    const double* B = nullptr;
    const double* C = nullptr;
    const double2* U = nullptr;
    const double2* V = nullptr;
    const double* KU = nullptr;
    const double* KV = nullptr;
    double scale = 1.0;
    int n1 = 1;
    int times_i = 0;
    int nyq = -1;
    int poisson = 0;
    int active = 0;   // 0: plain load
};

// First load of the r2c pass: the real line is built from up to six arrays of the same shape.
//   0: a      1: a*b      2: a*b + c*d      3: (a - b)*c      4: (a - b)*c + (d - e)*f
struct RealPro {
    const double* p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int mode = 0;
};

// extents the kernels cover: powers of two, 16 <= nx <= 2048 (x pass), 16 <= n <= 1024 (strided pass)
bool fft2d_x_ok(int nx);
bool fft2d_cols_ok(int n);
bool fft2d_enabled();
// build the twiddle tables of a transform length ahead of the first pass (plan time; the passes then only launch kernels)
int fft2d_prepare_x(int nx);
int fft2d_prepare_cols(int n);   // PDO_FFT=cufft turns the hand-written passes off (A/B measurements)

// c2c along a strided axis: element (col, row, plane) at plane*plane_stride + row*row_stride + col, col < ncols, row < n.
// dir = -1 forward (e^{-i}), +1 backward (unnormalised).  in == out is allowed (a tile of whole columns is read before it is written).
int fft2d_cols(int n, long long ncols, long long nplanes, long long row_stride, long long plane_stride, const double2* in, double2* out,
               int dir, const FftPro& pro, cudaStream_t st);
// real lines (nx, nlines) -> complex lines (nx/2+1, nlines)
int fft2d_r2c_lines(int nx, long long nlines, const RealPro& pro, double2* out, cudaStream_t st);
// complex lines (nx/2+1, nlines) -> real lines (nx, nlines); the imaginary parts of modes 0 and nx/2 are ignored, as FFTW's c2r does
int fft2d_c2r_lines(int nx, long long nlines, const double2* in, double* out, cudaStream_t st);

}  // namespace pdo
