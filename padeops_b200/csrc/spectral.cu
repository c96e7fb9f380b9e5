// spectral.cu — pencil-decomposed 3-D / 2-D FFTs on cuFFT and the periodic Poisson solve.
//
// Replaces:
//   fft_3d%init ("x" base), fft3_x2z, ifft3_z2x, fft2_x2y, ifft2_y2x     utilities/fft_3d.F90:109-468, 588-696
//   PoissonPeriodic%init / poisson_solve / poisson3D_multiply / GetWaveNums  utilities/PoissonPeriodic.F90:37-261
// The reference composes FFTW 1-D batched plans with 2DECOMP transposes: r2c in x, c2c in y run ONCE PER
// z-PLANE (a strided plan has a single batch dimension), c2c in z.  Here:
//   * when the row communicator has one rank (slab grids 1 x P — x- and y-pencils coincide) the x and y
//     passes collapse into ONE batched 2-D D2Z / Z2D cuFFT plan over all local planes;
//   * otherwise the reference's pass structure is kept (x pass, x<->y transpose, per-plane y pass);
//   * when the column communicator has one rank the y<->z "transpose" is the out-of-place z pass itself;
//   * the 1/(nx ny nz) normalisation of the inverse and the Poisson multiply share one pointwise kernel.
// cuFFT is a library call by design; the pointwise kernels are hand-written.
#include <cufft.h>

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "fft2d.cuh"

using namespace pdo;

#include "spectral_internal.cuh"

namespace {

#define PDO_CUFFT(expr)                                                                                  \
    do {                                                                                                 \
        cufftResult _r = (expr);                                                                         \
        if (_r != CUFFT_SUCCESS) return fail(PDO_E_CUDA, "%s:%d %s -> cufft error %d", __FILE__, __LINE__, #expr, (int)_r); \
    } while (0)

// ---- pointwise kernels ----
__global__ void scale_kernel(double* __restrict__ a, long long n, double s) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] *= s;
}

// out = in * s (* i k when ktab: k along index 1 (which == 1) or 2 of a complex array with extents n1, n2, ...), optionally zeroing
// the x-Nyquist column i == inyq
__global__ void copy_scale_oddball_kernel(const double2* __restrict__ in, double2* __restrict__ out, long long n, int n1, int n2,
                                          int inyq, double s, int which, const double* __restrict__ ktab) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double2 v = in[i];
        const int i1 = (int)(i % n1);
        if (inyq >= 0 && i1 == inyq) { v.x = 0.0; v.y = 0.0; }
        else if (which) {
            const double kv = (which == 1 ? ktab[i1] : ktab[(int)((i / n1) % n2)]) * s;
            v = make_double2(-kv * v.y, kv * v.x);
        } else { v.x *= s; v.y *= s; }
        out[i] = v;
    }
}

// the products of fft2d.cuh's RealPro as a stand-alone pass (cuFFT path)
__global__ void real_pro_kernel(RealPro pro, double* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double a = pro.p[0][i];
        double r = a;
        switch (pro.mode) {
            case 1: r = a * pro.p[1][i]; break;
            case 2: r = a * pro.p[1][i] + pro.p[2][i] * pro.p[3][i]; break;
            case 3: r = (a - pro.p[1][i]) * pro.p[2][i]; break;
            case 4: { const double t1 = (a - pro.p[1][i]) * pro.p[2][i]; const double t2 = (pro.p[3][i] - pro.p[4][i]) * pro.p[5][i]; r = t1 + t2; break; }
            default: break;
        }
        out[i] = r;
    }
}

// poisson3D_multiply (PoissonPeriodic.F90:89-111) on the complex z-pencil a(n1,n2,n3), times `scale`
// (the inverse transform's normfactor, folded in so the real field needs no extra pass).
__global__ void poisson_multiply_kernel(double2* __restrict__ a, int n1, int n2, int n3, const double* __restrict__ kx,
                                        const double* __restrict__ ky, const double* __restrict__ kz, double scale,
                                        int have_zero) {
    const long long tot = (long long)n1 * n2 * n3;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n1);
        const long long t = idx / n1;
        const int j = (int)(t % n2);
        const int k = (int)(t / n2);
        const double kxv = kx[i], kyv = ky[j], kzv = kz[k];
        const double ky_sq = kyv * kyv, kz_sq = kzv * kzv;
        double m = -1.0 / (kxv * kxv + ky_sq + kz_sq + 1.e-20);
        double2 v = a[idx];
        v.x = v.x * m * scale;
        v.y = v.y * m * scale;
        if (have_zero && idx == 0) { v.x = 0.0; v.y = 0.0; }
        a[idx] = v;
    }
}

inline int grid_for(long long n, int thr) {
    long long b = (n + thr - 1) / thr;
    const long long cap = 148LL * 32;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

// GetWaveNums + ifftshift (PoissonPeriodic.F90:226-261, fft_3d.F90:899-934)
std::vector<double> wavenums(int n, double d) {
    const double pi = 3.141592653589793238462643383279502884197;
    const int even = n - (n % 2);
    std::vector<double> raw(n), k(n);
    for (int i = 0; i < n; ++i) raw[i] = (-pi + (double)i * 2.0 * pi / (double)even) / d;
    const int h = (n % 2 == 0) ? n / 2 : (n + 1) / 2 - 1;  // first index of the non-negative half
    for (int i = 0; i < n; ++i) k[i] = raw[(i + h) % n];
    return k;
}

}  // namespace

struct pdo_fft3d_s {
    int nx, ny, nz, nxh, p_row, p_col;
    pdo_decomp_t phys = nullptr, spec = nullptr;
    pdo_decomp_info pi, si;
    bool slab2d = false;  // p_row == 1 → batched 2-D plans
    cufftHandle plan2d_f = 0, plan2d_b = 0, planx_f = 0, planx_b = 0, plany = 0, planz = 0;
    bool has2d = false, hasx = false, hasy = false, hasz = false;
    double2 *bufX = nullptr, *bufY = nullptr, *bufZ = nullptr;
    double* rtmp = nullptr;   // real x-pencil scratch of the cuFFT path's product pass (allocated on first use)
    double normfactor, normfactor2d;
    // hand-written passes (fft2d.cu): slab grids with power-of-two extents; cuFFT otherwise
    bool own_xy = false, own_z = false;
};

namespace {

long long cvol(const int* s) { return (long long)s[0] * s[1] * s[2]; }

int y_pass(pdo_fft3d_s* f, double2* a, int dir, cudaStream_t st) {  // per-plane strided c2c (fft_3d.F90:602-604)
    PDO_CUFFT(cufftSetStream(f->plany, st));
    const long long plane = (long long)f->si.ysz[0] * f->si.ysz[1];
    for (int k = 0; k < f->si.ysz[2]; ++k) {
        PDO_CUFFT(cufftExecZ2Z(f->plany, (cufftDoubleComplex*)(a + k * plane), (cufftDoubleComplex*)(a + k * plane), dir));
        g_launches += 1;
    }
    return 0;
}

// real x-pencil → complex y-pencil (spectral decomp), written to `outY`
int own_y_pass(pdo_fft3d_s* f, const double2* in, double2* out, int dir, const FftPro& pro, cudaStream_t st) {
    const long long nxh = f->si.ysz[0], ny = f->si.ysz[1];
    return fft2d_cols((int)ny, nxh, f->si.ysz[2], nxh, nxh * ny, in, out, dir, pro, st);
}

int forward_xy(pdo_fft3d_s* f, const double* in, double2* outY, cudaStream_t st) {
    if (f->own_xy) {
        RealPro rp;
        rp.p[0] = in;
        if (int rc = fft2d_r2c_lines(f->nx, (long long)f->pi.xsz[1] * f->pi.xsz[2], rp, outY, st)) return rc;
        return own_y_pass(f, outY, outY, -1, FftPro(), st);
    }
    if (f->slab2d) {
        PDO_CUFFT(cufftSetStream(f->plan2d_f, st));
        PDO_CUFFT(cufftExecD2Z(f->plan2d_f, (cufftDoubleReal*)in, (cufftDoubleComplex*)outY));
        g_launches += 1;
        return 0;
    }
    PDO_CUFFT(cufftSetStream(f->planx_f, st));
    PDO_CUFFT(cufftExecD2Z(f->planx_f, (cufftDoubleReal*)in, (cufftDoubleComplex*)f->bufX));
    g_launches += 1;
    if (int rc = decomp_transpose_device(f->spec, 0, (const double*)f->bufX, (double*)outY, 2, st)) return rc;
    return y_pass(f, outY, CUFFT_FORWARD, st);
}

// complex y-pencil in `Y` (destroyed) → real x-pencil
int backward_yx(pdo_fft3d_s* f, double2* Y, double* out, cudaStream_t st) {
    if (f->own_xy) {
        if (int rc = own_y_pass(f, Y, Y, +1, FftPro(), st)) return rc;
        return fft2d_c2r_lines(f->nx, (long long)f->pi.xsz[1] * f->pi.xsz[2], Y, out, st);
    }
    if (f->slab2d) {
        PDO_CUFFT(cufftSetStream(f->plan2d_b, st));
        PDO_CUFFT(cufftExecZ2D(f->plan2d_b, (cufftDoubleComplex*)Y, (cufftDoubleReal*)out));
        g_launches += 1;
        return 0;
    }
    if (int rc = y_pass(f, Y, CUFFT_INVERSE, st)) return rc;
    if (int rc = decomp_transpose_device(f->spec, 1, (const double*)Y, (double*)f->bufX, 2, st)) return rc;
    PDO_CUFFT(cufftSetStream(f->planx_b, st));
    PDO_CUFFT(cufftExecZ2D(f->planx_b, (cufftDoubleComplex*)f->bufX, (cufftDoubleReal*)out));
    g_launches += 1;
    return 0;
}

int own_z_pass(pdo_fft3d_s* f, const double2* in, double2* out, int dir, const FftPro& pro, cudaStream_t st) {
    const long long cols = (long long)f->si.zsz[0] * f->si.zsz[1];
    return fft2d_cols(f->nz, cols, 1, cols, 0, in, out, dir, pro, st);
}

int fft3_x2z_dev(pdo_fft3d_s* f, const double* in, double2* out, cudaStream_t st) {
    if (int rc = forward_xy(f, in, f->bufY, st)) return rc;
    if (f->own_z) {
        if (f->p_col == 1) return own_z_pass(f, f->bufY, out, -1, FftPro(), st);
        if (int rc = decomp_transpose_device(f->spec, 2, (const double*)f->bufY, (double*)out, 2, st)) return rc;
        return own_z_pass(f, out, out, -1, FftPro(), st);
    }
    PDO_CUFFT(cufftSetStream(f->planz, st));
    if (f->p_col == 1) {  // y- and z-pencils coincide: the out-of-place z pass is the "transpose"
        PDO_CUFFT(cufftExecZ2Z(f->planz, (cufftDoubleComplex*)f->bufY, (cufftDoubleComplex*)out, CUFFT_FORWARD));
    } else {
        if (int rc = decomp_transpose_device(f->spec, 2, (const double*)f->bufY, (double*)out, 2, st)) return rc;
        PDO_CUFFT(cufftExecZ2Z(f->planz, (cufftDoubleComplex*)out, (cufftDoubleComplex*)out, CUFFT_FORWARD));
    }
    g_launches += 1;
    return 0;
}

// scale == 0 → caller already folded the normalisation in
int ifft3_z2x_dev(pdo_fft3d_s* f, const double2* in, double* out, bool do_scale, cudaStream_t st) {
    if (f->own_z) {
        FftPro pro;
        if (do_scale) { pro.active = 1; pro.scale = f->normfactor; }   // linear: the normalisation rides on the first load
        if (f->p_col == 1) {
            if (int rc = own_z_pass(f, in, f->bufY, +1, pro, st)) return rc;
        } else {
            if (int rc = own_z_pass(f, in, f->bufZ, +1, pro, st)) return rc;
            if (int rc = decomp_transpose_device(f->spec, 3, (const double*)f->bufZ, (double*)f->bufY, 2, st)) return rc;
        }
        return backward_yx(f, f->bufY, out, st);
    }
    PDO_CUFFT(cufftSetStream(f->planz, st));
    if (f->p_col == 1) {
        PDO_CUFFT(cufftExecZ2Z(f->planz, (cufftDoubleComplex*)in, (cufftDoubleComplex*)f->bufY, CUFFT_INVERSE));
    } else {
        PDO_CUFFT(cufftExecZ2Z(f->planz, (cufftDoubleComplex*)in, (cufftDoubleComplex*)f->bufZ, CUFFT_INVERSE));
        if (int rc = decomp_transpose_device(f->spec, 3, (const double*)f->bufZ, (double*)f->bufY, 2, st)) return rc;
    }
    g_launches += 1;
    if (int rc = backward_yx(f, f->bufY, out, st)) return rc;
    if (do_scale) {
        const long long n = cvol(f->pi.xsz);
        scale_kernel<<<grid_for(n, 256), 256, 0, st>>>(out, n, f->normfactor);
        PDO_CUDA(cudaGetLastError());
        g_launches += 1;
    }
    return 0;
}

}  // namespace

namespace pdo {
// Device-pointer entry points shared with igrid.cu (declared in spectral_internal.cuh).
int fft3d_forward_xy(pdo_fft3d_t f, const double* in_real_x, double2* out_cplx_y, cudaStream_t st) {
    return forward_xy(f, in_real_x, out_cplx_y, st);
}
// ifft2_y2x: the input is intent(in) — it is staged into bufY, folding in 1/(nx ny) and the oddball zeroing
// (fft_3d.F90:633-641; zeroing the x-Nyquist column commutes with the y pass).
int fft3d_backward_yx(pdo_fft3d_t f, const double2* in_cplx_y, double* out_real_x, bool set_oddball, cudaStream_t st) {
    return fft3d_backward_yx_mul(f, in_cplx_y, 0, nullptr, f->normfactor2d, set_oddball, out_real_x, st);
}
// out = ifft2(in * scale * (i k)): which = 0 no wavenumber, 1 k along index 1 (ktab = the LOCAL slice of k1), 2 along index 2.
// `in` is left untouched: the hand-written y pass reads it with the factor on its first load and writes the scratch pencil; on
// the cuFFT path one pointwise pass stages it.
int fft3d_backward_yx_mul(pdo_fft3d_t f, const double2* in_cplx_y, int which, const double* ktab, double scale, bool set_oddball,
                          double* out_real_x, cudaStream_t st) {
    const long long n = cvol(f->si.ysz);
    int inyq = -1;
    if (set_oddball) {
        const int g = f->nx / 2;  // 0-based global index of mode nx/2+1
        if (g >= f->si.yst[0] - 1 && g <= f->si.yen[0] - 1) inyq = g - (f->si.yst[0] - 1);
    }
    if (f->own_xy) {
        FftPro pro;
        pro.active = 1;
        pro.scale = scale;
        pro.n1 = f->si.ysz[0];
        pro.nyq = inyq;
        if (which == 1) { pro.A = ktab; pro.times_i = 1; }
        if (which == 2) { pro.C = ktab; pro.times_i = 1; }
        if (int rc = own_y_pass(f, in_cplx_y, f->bufY, +1, pro, st)) return rc;
        return fft2d_c2r_lines(f->nx, (long long)f->pi.xsz[1] * f->pi.xsz[2], f->bufY, out_real_x, st);
    }
    copy_scale_oddball_kernel<<<grid_for(n, 256), 256, 0, st>>>(in_cplx_y, f->bufY, n, f->si.ysz[0], f->si.ysz[1], inyq, scale, which, ktab);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return backward_yx(f, f->bufY, out_real_x, st);
}
// fft2 of a pointwise combination of real x-pencil arrays (RealPro): the hand-written x pass forms it on its first load
int fft3d_forward_xy_pro(pdo_fft3d_t f, const RealPro& pro, double2* out_cplx_y, cudaStream_t st) {
    if (f->own_xy) {
        if (int rc = fft2d_r2c_lines(f->nx, (long long)f->pi.xsz[1] * f->pi.xsz[2], pro, out_cplx_y, st)) return rc;
        return own_y_pass(f, out_cplx_y, out_cplx_y, -1, FftPro(), st);
    }
    if (pro.mode == 0) return forward_xy(f, pro.p[0], out_cplx_y, st);
    const long long n = cvol(f->pi.xsz);
    if (!f->rtmp) PDO_CUDA(cudaMalloc(&f->rtmp, sizeof(double) * (size_t)n));
    real_pro_kernel<<<grid_for(n, 256), 256, 0, st>>>(pro, f->rtmp, n);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return forward_xy(f, f->rtmp, out_cplx_y, st);
}
// c2c along z on a z-pencil array with a factor table on the first load: v *= scale * gx(i) gy(j) gz(k) (null = 1)
int fft3d_z_pro(pdo_fft3d_t f, const double2* in, double2* out, int dir, const double* gx, const double* gy, const double* gz, double scale,
                cudaStream_t st) {
    if (!f->own_z) return fail(PDO_E_UNSUPPORTED, "fft3d_z_pro needs the hand-written z pass");
    FftPro pro;
    pro.active = 1; pro.scale = scale; pro.n1 = f->si.zsz[0]; pro.A = gx; pro.B = gy; pro.C = gz;
    return own_z_pass(f, in, out, dir, pro, st);
}
// the same with a caller-built first-load recipe (n1 and active are filled in here)
int fft3d_z_fused(pdo_fft3d_t f, const double2* in, double2* out, int dir, FftPro pro, cudaStream_t st) {
    if (!f->own_z) return fail(PDO_E_UNSUPPORTED, "fft3d_z_fused needs the hand-written z pass");
    pro.active = 1;
    pro.n1 = f->si.zsz[0];
    return own_z_pass(f, in, out, dir, pro, st);
}
bool fft3d_own_z(pdo_fft3d_t f) { return f->own_z; }
bool fft3d_own_xy(pdo_fft3d_t f) { return f->own_xy; }
// Same inverse for a caller-owned scratch array that already carries the 1/(nx ny) factor: it is consumed in place
// (no staging copy).
int fft3d_backward_yx_scratch(pdo_fft3d_t f, double2* scratch_cplx_y, double* out_real_x, cudaStream_t st) {
    return backward_yx(f, scratch_cplx_y, out_real_x, st);
}
// c2c along z, in place, on a z-pencil array of the spectral decomposition (or on the first nz planes of an edge
// field, which has the same zsz(1:2)); dir = -1 forward, +1 backward, unnormalised like FFTW.
int fft3d_z_inplace(pdo_fft3d_t f, double2* a_cplx_z, int dir, cudaStream_t st) {
    if (f->own_z) return own_z_pass(f, a_cplx_z, a_cplx_z, dir, FftPro(), st);
    PDO_CUFFT(cufftSetStream(f->planz, st));
    PDO_CUFFT(cufftExecZ2Z(f->planz, (cufftDoubleComplex*)a_cplx_z, (cufftDoubleComplex*)a_cplx_z, dir < 0 ? CUFFT_FORWARD : CUFFT_INVERSE));
    g_launches += 1;
    return 0;
}
int zcols_exec(ZColsPlan* p, int nz, long long cols, double2* a, int dir, cudaStream_t st) {
    if (!p || nz < 1 || cols < 1 || cols > 0x7fffffffLL) return fail(PDO_E_BADARG, "zcols: bad shape");
    if (p->plan < 0 || p->nz != nz || p->cols != cols) {
        zcols_destroy(p);
        cufftHandle h;
        int n1[1] = {nz}, emb[1] = {nz};
        const int c = (int)cols;
        PDO_CUFFT(cufftPlanMany(&h, 1, n1, emb, c, 1, emb, c, 1, CUFFT_Z2Z, c));   // same shape of plan as planz (fft_3d.F90:295-306)
        p->plan = (int)h; p->nz = nz; p->cols = cols;
    }
    PDO_CUFFT(cufftSetStream((cufftHandle)p->plan, st));
    PDO_CUFFT(cufftExecZ2Z((cufftHandle)p->plan, (cufftDoubleComplex*)a, (cufftDoubleComplex*)a, dir < 0 ? CUFFT_FORWARD : CUFFT_INVERSE));
    g_launches += 1;
    return 0;
}
void zcols_destroy(ZColsPlan* p) {
    if (p && p->plan >= 0) cufftDestroy((cufftHandle)p->plan);
    if (p) { p->plan = -1; p->cols = 0; p->nz = 0; }
}
pdo_decomp_t fft3d_phys_decomp(pdo_fft3d_t f) { return f->phys; }
pdo_decomp_t fft3d_spec_decomp(pdo_fft3d_t f) { return f->spec; }
}  // namespace pdo

struct pdo_poisson_s {
    pdo_fft3d_t ft = nullptr;
    int dir_id = 1;
    double2* hat = nullptr;
    double* rbuf = nullptr;  // x-pencil real scratch for dir_id == 2, 3
    double* ybuf = nullptr;  // y-pencil real scratch for dir_id == 3
    double *kx = nullptr, *ky = nullptr, *kz = nullptr;
    int have_zero = 0;
};

extern "C" {

int pdo_fft3d_init(pdo_fft3d_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col) {
    (void)dx; (void)dy; (void)dz;
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    if (nx < 2 || ny < 1 || nz < 1) return fail(PDO_E_BADARG, "bad sizes");
    if (p_row == 0 && p_col == 0) { p_row = 1; p_col = pdo_comm_size(); }
    pdo_fft3d_s* f = new (std::nothrow) pdo_fft3d_s();
    if (!f) return fail(PDO_E_BADARG, "out of memory");
    f->nx = nx; f->ny = ny; f->nz = nz; f->nxh = nx / 2 + 1; f->p_row = p_row; f->p_col = p_col;
    int rc = pdo_decomp_init(&f->phys, nx, ny, nz, p_row, p_col);
    if (!rc) rc = pdo_decomp_init(&f->spec, f->nxh, ny, nz, p_row, p_col);  // fft_3d.F90:248
    if (rc) { pdo_fft3d_destroy(f); return rc; }
    pdo_decomp_get_info(f->phys, &f->pi);
    pdo_decomp_get_info(f->spec, &f->si);
    f->normfactor = 1.0 / ((double)nx * (double)ny * (double)nz);  // in floating point: 2048^3 overflows int
    f->normfactor2d = 1.0 / ((double)nx * (double)ny);
    f->slab2d = (p_row == 1);
    f->own_xy = f->slab2d && fft2d_enabled() && fft2d_x_ok(nx) && fft2d_cols_ok(ny);
    f->own_z = fft2d_enabled() && fft2d_cols_ok(nz);
    if (f->own_xy) { rc = fft2d_prepare_x(nx); if (!rc) rc = fft2d_prepare_cols(ny); }
    if (!rc && f->own_z) rc = fft2d_prepare_cols(nz);
    if (rc) { pdo_fft3d_destroy(f); return rc; }
    const long long ny_c = cvol(f->si.ysz), nz_c = cvol(f->si.zsz), nx_c = cvol(f->si.xsz);
    // scratch pencils are transpose destinations: peer-writable for the fused NVLink path (collective, same order everywhere)
    rc = pdo::comm_shared_malloc((void**)&f->bufY, sizeof(double2) * (size_t)ny_c);
    if (!rc && p_col > 1) rc = pdo::comm_shared_malloc((void**)&f->bufZ, sizeof(double2) * (size_t)nz_c);
    if (!rc && !f->slab2d) rc = pdo::comm_shared_malloc((void**)&f->bufX, sizeof(double2) * (size_t)nx_c);
    if (rc) { pdo_fft3d_destroy(f); return rc; }
    cufftResult r = CUFFT_SUCCESS;
    if (f->own_xy) {
        // no cuFFT plans (and none of their work areas) for passes that never run
    } else if (f->slab2d) {
        int n2[2] = {ny, nx};
        int inembed[2] = {ny, nx}, onembed[2] = {ny, f->nxh};
        const int batch = f->pi.xsz[2];
        r = cufftPlanMany(&f->plan2d_f, 2, n2, inembed, 1, nx * ny, onembed, 1, f->nxh * ny, CUFFT_D2Z, batch);
        if (r == CUFFT_SUCCESS) r = cufftPlanMany(&f->plan2d_b, 2, n2, onembed, 1, f->nxh * ny, inembed, 1, nx * ny, CUFFT_Z2D, batch);
        f->has2d = (r == CUFFT_SUCCESS);
    } else {
        int n1[1] = {nx};
        const int batch = f->pi.xsz[1] * f->pi.xsz[2];
        int ie[1] = {nx}, oe[1] = {f->nxh};
        r = cufftPlanMany(&f->planx_f, 1, n1, ie, 1, nx, oe, 1, f->nxh, CUFFT_D2Z, batch);  // fft_3d.F90:256-265
        if (r == CUFFT_SUCCESS) r = cufftPlanMany(&f->planx_b, 1, n1, oe, 1, f->nxh, ie, 1, nx, CUFFT_Z2D, batch);
        f->hasx = (r == CUFFT_SUCCESS);
        if (r == CUFFT_SUCCESS) {
            int ny1[1] = {ny};
            int emb[1] = {ny};
            const int ys1 = f->si.ysz[0];
            r = cufftPlanMany(&f->plany, 1, ny1, emb, ys1, 1, emb, ys1, 1, CUFFT_Z2Z, ys1);  // fft_3d.F90:274-289
            f->hasy = (r == CUFFT_SUCCESS);
        }
    }
    if (r == CUFFT_SUCCESS && !f->own_z) {
        int nz1[1] = {nz};
        int emb[1] = {nz};
        const int zs = f->si.zsz[0] * f->si.zsz[1];
        r = cufftPlanMany(&f->planz, 1, nz1, emb, zs, 1, emb, zs, 1, CUFFT_Z2Z, zs);  // fft_3d.F90:295-306
        f->hasz = (r == CUFFT_SUCCESS);
    }
    if (r != CUFFT_SUCCESS) { pdo_fft3d_destroy(f); return fail(PDO_E_CUDA, "cufftPlanMany failed: %d", (int)r); }
    *h = f;
    return 0;
}

int pdo_fft3d_destroy(pdo_fft3d_t f) {
    if (!f) return 0;
    if (f->has2d) { cufftDestroy(f->plan2d_f); cufftDestroy(f->plan2d_b); }
    else { if (f->plan2d_f) cufftDestroy(f->plan2d_f); }
    if (f->hasx) { cufftDestroy(f->planx_f); cufftDestroy(f->planx_b); }
    if (f->hasy) cufftDestroy(f->plany);
    if (f->hasz) cufftDestroy(f->planz);
    pdo::comm_shared_free(f->bufY);     // same order as the allocation on every rank
    pdo::comm_shared_free(f->bufZ);
    pdo::comm_shared_free(f->bufX);
    if (f->rtmp) cudaFree(f->rtmp);
    pdo_decomp_destroy(f->phys);
    pdo_decomp_destroy(f->spec);
    delete f;
    return 0;
}

int pdo_fft3d_get_complex_output_size(pdo_fft3d_t f, int sz[3]) {
    if (!f || !sz) return fail(PDO_E_BADARG, "null argument");
    for (int i = 0; i < 3; ++i) sz[i] = f->si.zsz[i];
    return 0;
}
int pdo_fft3d_get_spectral_info(pdo_fft3d_t f, pdo_decomp_info* info) {
    if (!f || !info) return fail(PDO_E_BADARG, "null argument");
    *info = f->si;
    return 0;
}
int pdo_fft3d_get_physical_info(pdo_fft3d_t f, pdo_decomp_info* info) {
    if (!f || !info) return fail(PDO_E_BADARG, "null argument");
    *info = f->pi;
    return 0;
}

int pdo_fft3d_fft3_x2z(pdo_fft3d_t f, const double* in, double* out, void* stream) {
    if (!f || !in || !out) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    return with_device_views(in, sizeof(double) * cvol(f->pi.xsz), out, sizeof(double2) * cvol(f->si.zsz), st,
                             [&](const void* di, void* d_o) { return fft3_x2z_dev(f, (const double*)di, (double2*)d_o, st); });
}
int pdo_fft3d_ifft3_z2x(pdo_fft3d_t f, const double* in, double* out, void* stream) {
    if (!f || !in || !out) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    return with_device_views(in, sizeof(double2) * cvol(f->si.zsz), out, sizeof(double) * cvol(f->pi.xsz), st,
                             [&](const void* di, void* d_o) { return ifft3_z2x_dev(f, (const double2*)di, (double*)d_o, true, st); });
}
int pdo_fft3d_fft2_x2y(pdo_fft3d_t f, const double* in, double* out, void* stream) {
    if (!f || !in || !out) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    return with_device_views(in, sizeof(double) * cvol(f->pi.xsz), out, sizeof(double2) * cvol(f->si.ysz), st,
                             [&](const void* di, void* d_o) { return forward_xy(f, (const double*)di, (double2*)d_o, st); });
}
int pdo_fft3d_ifft2_y2x(pdo_fft3d_t f, const double* in, double* out, int set_oddball, void* stream) {
    if (!f || !in || !out) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    return with_device_views(in, sizeof(double2) * cvol(f->si.ysz), out, sizeof(double) * cvol(f->pi.xsz), st,
                             [&](const void* di, void* d_o) -> int {
                                 return pdo::fft3d_backward_yx(f, (const double2*)di, (double*)d_o, set_oddball != 0, st);
                             });
}

// ---------------- PoissonPeriodic ----------------
int pdo_poisson_init(pdo_poisson_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col,
                     int dir_id, const double* modkx, const double* modky, const double* modkz) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    // dir_id = 3 (z-pencil in / out): the reference switches fft_3d to its "z" base (PoissonPeriodic.F90:151-154); here the
    // field is brought to the x-pencil (z->y->x), solved on the x-base transforms and taken back — the same solution, the
    // transform order only changes the rounding.
    if (dir_id != 1 && dir_id != 2 && dir_id != 3) return fail(31243, "Incorrect option for DIR_ID");  // PoissonPeriodic.F90:156
    pdo_poisson_s* p = new (std::nothrow) pdo_poisson_s();
    if (!p) return fail(PDO_E_BADARG, "out of memory");
    p->dir_id = dir_id;
    int rc = pdo_fft3d_init(&p->ft, nx, ny, nz, dx, dy, dz, p_row, p_col);
    if (rc) { delete p; return rc; }
    pdo_fft3d_s* f = p->ft;
    const int* zs = f->si.zsz;
    std::vector<double> kx = modkx ? std::vector<double>(modkx, modkx + nx) : wavenums(nx, dx);
    std::vector<double> ky = modky ? std::vector<double>(modky, modky + ny) : wavenums(ny, dy);
    std::vector<double> kz = modkz ? std::vector<double>(modkz, modkz + nz) : wavenums(nz, dz);
    cudaError_t e = cudaMalloc(&p->hat, sizeof(double2) * (size_t)cvol(zs));
    if (e == cudaSuccess) e = cudaMalloc(&p->kx, sizeof(double) * zs[0]);
    if (e == cudaSuccess) e = cudaMalloc(&p->ky, sizeof(double) * zs[1]);
    if (e == cudaSuccess) e = cudaMalloc(&p->kz, sizeof(double) * zs[2]);
    if (e == cudaSuccess && dir_id >= 2) e = cudaMalloc(&p->rbuf, sizeof(double) * (size_t)cvol(f->pi.xsz));
    if (e == cudaSuccess && dir_id == 3) e = cudaMalloc(&p->ybuf, sizeof(double) * (size_t)cvol(f->pi.ysz));
    // local slices kx(xst:xen) etc. of the spectral z-pencil (PoissonPeriodic.F90:169-179)
    if (e == cudaSuccess) e = cudaMemcpy(p->kx, kx.data() + (f->si.zst[0] - 1), sizeof(double) * zs[0], cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->ky, ky.data() + (f->si.zst[1] - 1), sizeof(double) * zs[1], cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->kz, kz.data() + (f->si.zst[2] - 1), sizeof(double) * zs[2], cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { pdo_poisson_destroy(p); return fail(PDO_E_CUDA, "poisson init: %s", cudaGetErrorString(e)); }
    p->have_zero = (f->si.zst[0] == 1 && f->si.zst[1] == 1 && f->si.zst[2] == 1) ? 1 : 0;  // :206-210
    *h = p;
    return 0;
}

int pdo_poisson_destroy(pdo_poisson_t p) {
    if (!p) return 0;
    if (p->hat) cudaFree(p->hat);
    if (p->rbuf) cudaFree(p->rbuf);
    if (p->ybuf) cudaFree(p->ybuf);
    if (p->kx) cudaFree(p->kx);
    if (p->ky) cudaFree(p->ky);
    if (p->kz) cudaFree(p->kz);
    pdo_fft3d_destroy(p->ft);
    delete p;
    return 0;
}

int pdo_poisson_solve(pdo_poisson_t p, const double* rhs, double* fout, void* stream) {
    if (!p || !rhs || !fout) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    pdo_fft3d_s* f = p->ft;
    const int* insz = (p->dir_id == 1) ? f->pi.xsz : (p->dir_id == 2 ? f->pi.ysz : f->pi.zsz);
    const size_t bytes = sizeof(double) * (size_t)cvol(insz);
    return with_device_views(rhs, bytes, fout, bytes, st, [&](const void* di, void* d_o) -> int {
        const double* x_in = (const double*)di;
        double* x_out = (double*)d_o;
        if (p->dir_id == 2) {  // PoissonPeriodic.F90:75-80
            if (int rc = decomp_transpose_device(f->phys, 1, (const double*)di, p->rbuf, 1, st)) return rc;
            x_in = p->rbuf;
            x_out = p->rbuf;
        } else if (p->dir_id == 3) {
            if (int rc = decomp_transpose_device(f->phys, 3, (const double*)di, p->ybuf, 1, st)) return rc;   // z -> y
            if (int rc = decomp_transpose_device(f->phys, 1, p->ybuf, p->rbuf, 1, st)) return rc;             // y -> x
            x_in = p->rbuf;
            x_out = p->rbuf;
        }
        if (int rc = fft3_x2z_dev(f, x_in, p->hat, st)) return rc;
        const int* zs = f->si.zsz;
        const long long n = cvol(zs);
        poisson_multiply_kernel<<<grid_for(n, 256), 256, 0, st>>>(p->hat, zs[0], zs[1], zs[2], p->kx, p->ky, p->kz, f->normfactor,
                                                                   p->have_zero);
        PDO_CUDA(cudaGetLastError());
        g_launches += 1;
        if (int rc = ifft3_z2x_dev(f, p->hat, x_out, false, st)) return rc;
        if (p->dir_id == 2) return decomp_transpose_device(f->phys, 0, p->rbuf, (double*)d_o, 1, st);
        if (p->dir_id == 3) {
            if (int rc = decomp_transpose_device(f->phys, 0, p->rbuf, p->ybuf, 1, st)) return rc;             // x -> y
            return decomp_transpose_device(f->phys, 2, p->ybuf, (double*)d_o, 1, st);                           // y -> z
        }
        return 0;
    });
}

}  // extern "C"
