// igrid.cu — the igrid periodic substep and the spectral / projection types it is made of, device-resident.
//
// Replaces (paths relative to the reference's src/incompressible):
//   spectralMod::spectral ("x" pencil, dimTransform = 2)   spectral.F90:235-363, 755-865, 867-1200, 1413-1509
//   PadeDerOps::Pade6stagg (periodic, scheme cd06)         PadeDerOps.F90:57-88, 146-160, 404-418, 572-585, 689-702, 879-892, 997-1053
//   PadePoissonMod::padepoisson (PeriodicInZ)              PadePoisson.F90:76-180, 386-432, 716-750, 900-949, 1165-1244
//   IncompressibleGrid::igrid (substep)                    igrid.F90:625-655, 1020-1037, 1105-1299, 1372-1396, 1423-1447,
//                                                          1572-1679, 1793-1941, 1961-1990, 2553-2683
// Differences in form, not in results:
//   * the reference stores k1, k2, kabs_sq, Gdealias and kradsq_inv as full 3-D arrays and streams them through
//     every pointwise pass; here they are 1-D tables (a few KB, L1/L2 resident) combined on the fly, so a pointwise
//     pass moves only the field;
//   * scalings that the reference applies as separate passes (normfactz after the inverse z FFT, mfact in the
//     projection) are folded into the preceding pointwise kernel (linear, commutes with the FFT);
//   * only the nine velocity-gradient fields the skew-symmetric substep reads are formed unless the caller asks for
//     all eighteen (compute_all_gradients);
//   * when the column communicator has one rank the y- and z-pencils of a spectral array are the same memory
//     layout, and the z-periodic dealiasing runs in place without the two transposes.
// All fields stay in HBM between calls; host pointers are accepted only at the API boundary (init / get_field).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "sgs_kernels.cuh"
#include "spectral_internal.cuh"

using namespace pdo;

namespace {

constexpr double kPi = 3.141592653589793238462643383279502884197;

template <class F>
__global__ void __launch_bounds__(256) ew_kernel(long long n, F f) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) f(i);
}
template <class F>
int launch_ew(long long n, cudaStream_t st, F f) {
    if (n <= 0) return 0;
    long long b = (n + 255) / 256;
    const long long cap = 148LL * 16;
    if (b > cap) b = cap;
    ew_kernel<<<(unsigned)b, 256, 0, st>>>(n, f);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return 0;
}

inline long long vol(const int* s) { return (long long)s[0] * s[1] * s[2]; }

// GetWaveNums + ifftshift (utilities/fft_3d.F90:899-934)
std::vector<double> wavenums(int n, double d) {
    const int even = n - (n % 2);
    std::vector<double> raw(n), k(n);
    for (int i = 0; i < n; ++i) raw[i] = (-kPi + (double)i * 2.0 * kPi / (double)even) / d;
    const int h = (n % 2 == 0) ? n / 2 : (n + 1) / 2 - 1;
    for (int i = 0; i < n; ++i) k[i] = raw[(i + h) % n];
    return k;
}

int upload(double** d, const std::vector<double>& h, size_t off, size_t cnt) {
    PDO_CUDA(cudaMalloc(d, sizeof(double) * (cnt ? cnt : 1)));
    if (cnt) PDO_CUDA(cudaMemcpy(*d, h.data() + off, sizeof(double) * cnt, cudaMemcpyHostToDevice));
    return 0;
}

// max over a real array; result on the host (synchronises the stream)
__global__ void __launch_bounds__(256) max_kernel(const double* __restrict__ a, long long n, int use_abs, double* __restrict__ partial) {
    __shared__ double sm[256];
    double m = -1.0e300;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = use_abs ? fabs(a[i]) : a[i];
        m = v > m ? v : m;
    }
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = sm[threadIdx.x] > sm[threadIdx.x + s] ? sm[threadIdx.x] : sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

}  // namespace

// ================================================================================================
// spectral
// ================================================================================================
struct pdo_spectral_s {
    int nx, ny, nz, nxh, p_row, p_col;
    double dx, dy, dz;
    pdo_fft3d_t ft = nullptr;
    pdo_decomp_info pi, si;
    bool periodicInZ = false;
    double normfactz = 1.0;
    std::vector<double> h_k1, h_k2, h_gx, h_gy, h_gz;  // global 1-D tables (nxh, ny, nxh, ny, nz)
    double *k1y = nullptr, *k2 = nullptr;               // local slice of k1 (ysz0 == zsz0 entries), full k2
    double *gx = nullptr, *gy = nullptr, *gyz = nullptr, *gz = nullptr;  // dealias masks: x slice, y full, y slice of the z-pencil, z
    double2* ctmpz = nullptr;
    double* partial = nullptr;  // reduction scratch
    // z-Fourier tables of init_periodic_inZ_procedures (spectral.F90:843-856), built on first use: 6 tables of nz complex numbers
    // (k3_E2Cshift, k3_C2Eshift, E2Cshift, C2Eshift, mk3sq, k3_C2Cder), then the same six with the oddball entry set to 1 for the
    // REAL procedures, which leave that mode untouched
    double2* ztab = nullptr;
    ZColsPlan rz_plan;          // c2c-z over pairs of real columns
    double2* rz_work = nullptr;
    size_t rz_cap = 0;
};
enum { ZT_K3_E2C = 0, ZT_K3_C2E = 1, ZT_E2C = 2, ZT_C2E = 3, ZT_MK3SQ = 4, ZT_K3_C2C = 5, ZT_COUNT = 6 };

namespace {

// fout = i k fin (* scale: lets a caller fold the inverse transform's 1/(nx ny) into this pass)
int spectral_mtimes(pdo_spectral_s* s, int which, const double2* fin, double2* fout, cudaStream_t st, double scale = 1.0) {
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const long long n = vol(s->si.ysz);
    const double* k = which == 1 ? s->k1y : s->k2;
    if (which == 1)
        return launch_ew(n, st, [=] __device__(long long i) {
            const double kv = k[(int)(i % n1)] * scale;
            const double2 v = fin[i];
            fout[i] = make_double2(-kv * v.y, kv * v.x);
        });
    return launch_ew(n, st, [=] __device__(long long i) {
        const double kv = k[(int)((i / n1) % n2)] * scale;
        const double2 v = fin[i];
        fout[i] = make_double2(-kv * v.y, kv * v.x);
    });
}

// ---- z-Fourier operators: c2c-z forward, x table(k), c2c-z backward, x 1/nz (spectral.F90:365-702) ----
// host side of the tables: [2][ZT_COUNT][nz], complex procedures first, then the REAL procedures' twins (oddball entry = 1)
std::vector<double2> build_ztables_host(int nz, double dz) {
    std::vector<double> k3 = wavenums(nz, dz);   // GetWaveNums(nz, dz), no sign flip of the oddball (spectral.F90:845)
    std::vector<double2> t(2 * ZT_COUNT * (size_t)nz);
    for (int k = 0; k < nz; ++k) {
        const double kk = k3[k], ph = kk * dz / 2.0, c = std::cos(ph), sn = std::sin(ph);
        t[(size_t)ZT_K3_E2C * nz + k] = make_double2(-kk * sn, kk * c);    // i k e^{+i k dz/2}
        t[(size_t)ZT_K3_C2E * nz + k] = make_double2(kk * sn, kk * c);     // i k e^{-i k dz/2}
        t[(size_t)ZT_E2C * nz + k] = make_double2(c, sn);                  // e^{+i k dz/2}
        t[(size_t)ZT_C2E * nz + k] = make_double2(c, -sn);                 // e^{-i k dz/2}
        t[(size_t)ZT_MK3SQ * nz + k] = make_double2(-(kk * kk), 0.0);      // -k^2
        t[(size_t)ZT_K3_C2C * nz + k] = make_double2(0.0, kk);             // i k
    }
    for (int i = 0; i < ZT_COUNT; ++i)
        for (int k = 0; k < nz; ++k)
            t[(size_t)(ZT_COUNT + i) * nz + k] = k == nz / 2 ? make_double2(1.0, 0.0) : t[(size_t)i * nz + k];
    return t;
}
int spectral_ztables(pdo_spectral_s* s) {
    if (s->ztab) return 0;
    if (!s->periodicInZ) return fail(PDO_E_BADARG, "spectral type was not initialised with init_periodicInZ");
    std::vector<double2> t = build_ztables_host(s->nz, s->dz);
    PDO_CUDA(cudaMalloc(&s->ztab, sizeof(double2) * t.size()));
    PDO_CUDA(cudaMemcpy(s->ztab, t.data(), sizeof(double2) * t.size(), cudaMemcpyHostToDevice));
    return 0;
}
inline const double2* ztable(const pdo_spectral_s* s, int which, bool real_variant) {
    return s->ztab + (size_t)((real_variant ? ZT_COUNT : 0) + which) * s->nz;
}
// w(cols, nz) *= tab(k) * scale
int ztable_multiply(double2* w, long long cols, int nz, const double2* tab, double scale, cudaStream_t st) {
    return launch_ew(cols * nz, st, [=] __device__(long long i) {
        double2 t = tab[(int)(i / cols)];
        const double2 v = w[i];
        t.x *= scale; t.y *= scale;
        w[i] = make_double2(v.x * t.x - v.y * t.y, v.x * t.y + v.y * t.x);
    });
}
// complex z-pencil array of the spectral decomposition, in place on the first nz planes of w
int zfourier_complex(pdo_spectral_s* s, double2* w, int which, cudaStream_t st) {
    if (int rc = spectral_ztables(s)) return rc;
    const long long cols = (long long)s->si.zsz[0] * s->si.zsz[1];
    if (int rc = fft3d_z_inplace(s->ft, w, -1, st)) return rc;
    if (int rc = ztable_multiply(w, cols, s->nz, ztable(s, which, false), 1.0 / (double)s->nz, st)) return rc;
    return fft3d_z_inplace(s->ft, w, +1, st);
}
// REAL z-pencil array of the physical decomposition: in(P, nz [+1]) -> out(P, nz).  The table with the oddball entry = 1 is
// conjugate-symmetric in k, so the operator maps real columns to real columns and is linear over C: two real columns a, b
// are transformed as ONE complex column a + i b and come back as a' + i b' (half the transform work of a zero-padded c2c,
// the same arithmetic as the reference's r2c / c2r pair up to rounding).  P odd: the last column is paired with zeros.
int zfourier_real(pdo_spectral_s* s, const double* in, double* out, int which, cudaStream_t st) {
    if (int rc = spectral_ztables(s)) return rc;
    const int nz = s->nz;
    const long long P = (long long)s->pi.zsz[0] * s->pi.zsz[1], Pc = (P + 1) / 2;
    const size_t need = sizeof(double2) * (size_t)Pc * nz;
    if (s->rz_cap < need) {
        if (s->rz_work) cudaFree(s->rz_work);
        s->rz_work = nullptr; s->rz_cap = 0;
        PDO_CUDA(cudaMalloc(&s->rz_work, need));
        s->rz_cap = need;
    }
    double2* w = s->rz_work;
    if (P & 1) PDO_CUDA(cudaMemsetAsync(w, 0, need, st));
    PDO_CUDA(cudaMemcpy2DAsync(w, sizeof(double2) * Pc, in, sizeof(double) * P, sizeof(double) * P, nz, cudaMemcpyDeviceToDevice, st));
    if (int rc = zcols_exec(&s->rz_plan, nz, Pc, w, -1, st)) return rc;
    if (int rc = ztable_multiply(w, Pc, nz, ztable(s, which, true), 1.0 / (double)nz, st)) return rc;
    if (int rc = zcols_exec(&s->rz_plan, nz, Pc, w, +1, st)) return rc;
    PDO_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * P, w, sizeof(double2) * Pc, sizeof(double) * P, nz, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// z-pencil array a(zsz0, zsz1, nz) *= gx(i) gy(j) gz(k) * scale
int spectral_mask_z(pdo_spectral_s* s, double2* a, double scale, cudaStream_t st) {
    const int n1 = s->si.zsz[0], n2 = s->si.zsz[1];
    const long long n = (long long)n1 * n2 * s->nz;
    const double *gx = s->gx, *gy = s->gyz, *gz = s->gz;
    return launch_ew(n, st, [=] __device__(long long i) {
        const int ii = (int)(i % n1);
        const long long t = i / n1;
        const int jj = (int)(t % n2);
        const int kk = (int)(t / n2);
        const double m = gx[ii] * gy[jj] * gz[kk] * scale;
        double2 v = a[i];
        v.x *= m; v.y *= m;
        a[i] = v;
    });
}

int spectral_dealias(pdo_spectral_s* s, double2* fhat, cudaStream_t st) {
    if (!s->periodicInZ) {  // 2-D mask (spectral.F90:329-338 with the table of :1147-1159)
        const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
        const double *gx = s->gx, *gy = s->gy;
        return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
            const double m = gx[(int)(i % n1)] * gy[(int)((i / n1) % n2)];
            double2 v = fhat[i];
            v.x *= m; v.y *= m;
            fhat[i] = v;
        });
    }
    pdo_decomp_t spec = fft3d_spec_decomp(s->ft);
    double2* work = fhat;  // one rank in the column communicator: y- and z-pencil layouts coincide
    if (s->p_col > 1) {
        work = s->ctmpz;
        if (int rc = decomp_transpose_device(spec, 2, (const double*)fhat, (double*)work, 2, st)) return rc;  // take_fftz
    }
    if (int rc = fft3d_z_inplace(s->ft, work, -1, st)) return rc;
    if (int rc = spectral_mask_z(s, work, s->normfactz, st)) return rc;
    if (int rc = fft3d_z_inplace(s->ft, work, +1, st)) return rc;  // take_ifftz
    if (s->p_col > 1) return decomp_transpose_device(spec, 3, (const double*)work, (double*)fhat, 2, st);
    return 0;
}

int spectral_dealias_edge(pdo_spectral_s* s, double2* fE, cudaStream_t st) {
    if (!s->periodicInZ) return 0;  // the reference does nothing on this branch (spectral.F90:348)
    if (int rc = fft3d_z_inplace(s->ft, fE, -1, st)) return rc;
    if (int rc = spectral_mask_z(s, fE, s->normfactz, st)) return rc;
    if (int rc = fft3d_z_inplace(s->ft, fE, +1, st)) return rc;
    const size_t plane = (size_t)s->si.zsz[0] * s->si.zsz[1];
    PDO_CUDA(cudaMemcpyAsync(fE + plane * s->nz, fE, sizeof(double2) * plane, cudaMemcpyDeviceToDevice, st));  // :361
    return 0;
}

}  // namespace

extern "C" {

int pdo_spectral_init(pdo_spectral_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col,
                      int fix_oddball, int init_periodic_in_z, double dealias_fact) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    if (nx < 2 || ny < 2 || nz < 1) return fail(PDO_E_BADARG, "bad sizes");
    if (init_periodic_in_z && (nz % 2) != 0)
        return fail(104, "You cannot initialize a periodic_inZ spectral type with an odd values nz");  // spectral.F90:773-775
    if (p_row == 0 && p_col == 0) { p_row = 1; p_col = pdo_comm_size(); }
    pdo_spectral_s* s = new (std::nothrow) pdo_spectral_s();
    if (!s) return fail(PDO_E_BADARG, "out of memory");
    s->nx = nx; s->ny = ny; s->nz = nz; s->nxh = nx / 2 + 1; s->dx = dx; s->dy = dy; s->dz = dz;
    s->p_row = p_row; s->p_col = p_col;
    s->periodicInZ = init_periodic_in_z != 0;
    int rc = pdo_fft3d_init(&s->ft, nx, ny, nz, dx, dy, dz, p_row, p_col);
    if (rc) { delete s; return rc; }
    pdo_fft3d_get_physical_info(s->ft, &s->pi);
    pdo_fft3d_get_spectral_info(s->ft, &s->si);
    // 1-D wavenumbers with the oddball sign flip (spectral.F90:1024-1032) and the optional fixOddball (:1189-1199)
    std::vector<double> k1 = wavenums(nx, dx), k2 = wavenums(ny, dy), k3 = wavenums(nz, dz);
    k1[nx / 2] = -k1[nx / 2];
    k2[ny / 2] = -k2[ny / 2];
    k3[nz / 2] = -k3[nz / 2];
    if (fix_oddball) { k1[nx / 2] = 0.0; k2[ny / 2] = 0.0; }
    s->h_k1.assign(k1.begin(), k1.begin() + s->nxh);
    s->h_k2 = k2;
    s->h_gx.assign(s->nxh, 1.0); s->h_gy.assign(ny, 1.0); s->h_gz.assign(nz, 1.0);
    if (s->periodicInZ) {  // zero where |k| >= f pi/d  (spectral.F90:785-814)
        const double kdx = dealias_fact * kPi / dx, kdy = dealias_fact * kPi / dy, kdz = dealias_fact * kPi / dz;
        for (int i = 0; i < s->nxh; ++i) if (std::fabs(s->h_k1[i]) >= kdx) s->h_gx[i] = 0.0;
        for (int j = 0; j < ny; ++j) if (std::fabs(k2[j]) >= kdy) s->h_gy[j] = 0.0;
        for (int k = 0; k < nz; ++k) if (std::fabs(k3[k]) >= kdz) s->h_gz[k] = 0.0;
        s->normfactz = 1.0 / (double)nz;
    } else {               // pass band |k| < (2/3) pi/d, factor hard-wired (spectral.F90:1147-1159)
        const double kdx = (2.0 / 3.0) * kPi / dx, kdy = (2.0 / 3.0) * kPi / dy;
        for (int i = 0; i < s->nxh; ++i) s->h_gx[i] = (std::fabs(s->h_k1[i]) < kdx) ? 1.0 : 0.0;
        for (int j = 0; j < ny; ++j) s->h_gy[j] = (std::fabs(k2[j]) < kdy) ? 1.0 : 0.0;
    }
    const int i0 = s->si.yst[0] - 1, ni = s->si.ysz[0];
    rc = upload(&s->k1y, s->h_k1, i0, ni);
    if (!rc) rc = upload(&s->k2, s->h_k2, 0, ny);
    if (!rc) rc = upload(&s->gx, s->h_gx, i0, ni);
    if (!rc) rc = upload(&s->gy, s->h_gy, 0, ny);
    if (!rc) rc = upload(&s->gyz, s->h_gy, s->si.zst[1] - 1, s->si.zsz[1]);
    if (!rc) rc = upload(&s->gz, s->h_gz, 0, nz);
    if (!rc && s->periodicInZ && p_col > 1) {
        cudaError_t e = cudaMalloc(&s->ctmpz, sizeof(double2) * (size_t)vol(s->si.zsz));
        if (e != cudaSuccess) rc = fail(PDO_E_CUDA, "spectral ctmpz: %s", cudaGetErrorString(e));
        else comm_register_buffer_quiet(s->ctmpz, sizeof(double2) * (size_t)vol(s->si.zsz));
    }
    if (!rc) {
        cudaError_t e = cudaMalloc(&s->partial, sizeof(double) * 2048);
        if (e != cudaSuccess) rc = fail(PDO_E_CUDA, "spectral scratch: %s", cudaGetErrorString(e));
    }
    if (rc) { pdo_spectral_destroy(s); return rc; }
    *h = s;
    return 0;
}

int pdo_spectral_destroy(pdo_spectral_t s) {
    if (!s) return 0;
    double* ptrs[] = {s->k1y, s->k2, s->gx, s->gy, s->gyz, s->gz, s->partial};
    for (double* p : ptrs) if (p) cudaFree(p);
    if (s->ctmpz) { comm_deregister_buffer(s->ctmpz); cudaFree(s->ctmpz); }
    if (s->ztab) cudaFree(s->ztab);
    if (s->rz_work) cudaFree(s->rz_work);
    zcols_destroy(&s->rz_plan);
    pdo_fft3d_destroy(s->ft);
    delete s;
    return 0;
}

int pdo_spectral_get_physical_info(pdo_spectral_t s, pdo_decomp_info* info) {
    if (!s || !info) return fail(PDO_E_BADARG, "null argument");
    *info = s->pi;
    return 0;
}
int pdo_spectral_get_spectral_info(pdo_spectral_t s, pdo_decomp_info* info) {
    if (!s || !info) return fail(PDO_E_BADARG, "null argument");
    *info = s->si;
    return 0;
}
int pdo_spectral_get_tables(pdo_spectral_t s, double* k1, double* k2, double* gx, double* gy, double* gz) {
    if (!s) return fail(PDO_E_BADARG, "null handle");
    if (k1) std::memcpy(k1, s->h_k1.data(), sizeof(double) * s->nxh);
    if (k2) std::memcpy(k2, s->h_k2.data(), sizeof(double) * s->ny);
    if (gx) std::memcpy(gx, s->h_gx.data(), sizeof(double) * s->nxh);
    if (gy) std::memcpy(gy, s->h_gy.data(), sizeof(double) * s->ny);
    if (gz) std::memcpy(gz, s->h_gz.data(), sizeof(double) * s->nz);
    return 0;
}

int pdo_spectral_fft(pdo_spectral_t s, const double* in, double* out, void* stream) {
    if (!s) return fail(PDO_E_BADARG, "null handle");
    return pdo_fft3d_fft2_x2y(s->ft, in, out, stream);
}
int pdo_spectral_ifft(pdo_spectral_t s, const double* in, double* out, int set_oddball, void* stream) {
    if (!s) return fail(PDO_E_BADARG, "null handle");
    return pdo_fft3d_ifft2_y2x(s->ft, in, out, set_oddball, stream);
}

static int spectral_ywise(pdo_spectral_t s, const double* in, double* out, void* stream, int op) {
    if (!s || !in || !out) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = sizeof(double2) * (size_t)vol(s->si.ysz);
    return with_device_views(in, bytes, out, bytes, st, [&](const void* di, void* d_o) -> int {
        if (op == 1 || op == 2) return spectral_mtimes(s, op, (const double2*)di, (double2*)d_o, st);
        if (di != d_o) PDO_CUDA(cudaMemcpyAsync(d_o, di, bytes, cudaMemcpyDeviceToDevice, st));
        return spectral_dealias(s, (double2*)d_o, st);
    });
}
int pdo_spectral_mtimes_ik1_oop(pdo_spectral_t s, const double* fin, double* fout, void* st) { return spectral_ywise(s, fin, fout, st, 1); }
int pdo_spectral_mtimes_ik2_oop(pdo_spectral_t s, const double* fin, double* fout, void* st) { return spectral_ywise(s, fin, fout, st, 2); }
int pdo_spectral_mtimes_ik1_ip(pdo_spectral_t s, double* f, void* st) { return spectral_ywise(s, f, f, st, 1); }
int pdo_spectral_mtimes_ik2_ip(pdo_spectral_t s, double* f, void* st) { return spectral_ywise(s, f, f, st, 2); }
int pdo_spectral_dealias(pdo_spectral_t s, double* fhat, void* st) { return spectral_ywise(s, fhat, fhat, st, 3); }

static int spectral_zwise(pdo_spectral_t s, double* a, void* stream, int op) {
    if (!s || !a) return fail(PDO_E_BADARG, "null argument");
    if (!s->periodicInZ) return fail(PDO_E_BADARG, "spectral type was not initialised with init_periodicInZ");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane = (size_t)s->si.zsz[0] * s->si.zsz[1];
    const size_t bytes = sizeof(double2) * plane * (size_t)(s->nz + (op == 0 ? 1 : 0));
    return with_device_views(a, bytes, a, bytes, st, [&](const void* di, void* d_o) -> int {
        if (di != d_o) PDO_CUDA(cudaMemcpyAsync(d_o, di, bytes, cudaMemcpyDeviceToDevice, st));
        double2* w = (double2*)d_o;
        if (op == 0) return spectral_dealias_edge(s, w, st);
        if (op == 1) return fft3d_z_inplace(s->ft, w, -1, st);
        if (int rc = fft3d_z_inplace(s->ft, w, +1, st)) return rc;
        const double nf = s->normfactz;
        return launch_ew((long long)plane * s->nz, st, [=] __device__(long long i) { double2 v = w[i]; v.x *= nf; v.y *= nf; w[i] = v; });
    });
}
int pdo_spectral_dealias_edgefield(pdo_spectral_t s, double* fE, void* st) { return spectral_zwise(s, fE, st, 0); }
int pdo_spectral_take_fft1d_z2z_ip(pdo_spectral_t s, double* a, void* st) { return spectral_zwise(s, a, st, 1); }
int pdo_spectral_take_ifft1d_z2z_ip(pdo_spectral_t s, double* a, void* st) { return spectral_zwise(s, a, st, 2); }

// ddz_C2C_complex_inplace (:528-547), shiftz_E2C / shiftz_C2E (:409-437): complex z-pencil arrays of the spectral decomposition
static int spectral_zcomplex(pdo_spectral_t s, double* a, void* stream, int op) {
    if (!s || !a) return fail(PDO_E_BADARG, "null argument");
    if (!s->periodicInZ) return fail(PDO_E_BADARG, "spectral type was not initialised with init_periodicInZ");
    cudaStream_t st = (cudaStream_t)stream;
    const long long cols = (long long)s->si.zsz[0] * s->si.zsz[1];
    const size_t bytes = sizeof(double2) * (size_t)cols * (size_t)s->nz;
    return with_device_views(a, bytes, a, bytes, st, [&](const void* di, void* d_o) -> int {
        if (di != d_o) PDO_CUDA(cudaMemcpyAsync(d_o, di, bytes, cudaMemcpyDeviceToDevice, st));
        double2* w = (double2*)d_o;
        if (op == 0) return zfourier_complex(s, w, ZT_K3_C2C, st);
        if (int rc = spectral_ztables(s)) return rc;
        return ztable_multiply(w, cols, s->nz, ztable(s, op == 1 ? ZT_E2C : ZT_C2E, false), 1.0, st);
    });
}
// test hook (host only, not in the public header): the z-Fourier tables as the kernels get them, out[2][6][nz] complex
int pdo_debug_ztables(int nz, double dz, double* out) {
    if (nz < 2 || (nz & 1) || !out) return fail(PDO_E_BADARG, "bad argument");
    std::vector<double2> t = build_ztables_host(nz, dz);
    std::memcpy(out, t.data(), sizeof(double2) * t.size());
    return 0;
}
int pdo_spectral_ddz_c2c_complex_ip(pdo_spectral_t s, double* a, void* st) { return spectral_zcomplex(s, a, st, 0); }
int pdo_spectral_shiftz_e2c(pdo_spectral_t s, double* a, void* st) { return spectral_zcomplex(s, a, st, 1); }
int pdo_spectral_shiftz_c2e(pdo_spectral_t s, double* a, void* st) { return spectral_zcomplex(s, a, st, 2); }
// ddz_C2C_real_inplace (:507-526): real z-pencil array of the physical decomposition; the oddball mode passes through
int pdo_spectral_ddz_c2c_real_ip(pdo_spectral_t s, double* a, void* stream) {
    if (!s || !a) return fail(PDO_E_BADARG, "null argument");
    if (!s->periodicInZ) return fail(PDO_E_BADARG, "spectral type was not initialised with init_periodicInZ");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = sizeof(double) * (size_t)vol(s->pi.zsz);
    return with_device_views(a, bytes, a, bytes, st, [&](const void* di, void* d_o) -> int {
        return zfourier_real(s, (const double*)di, (double*)d_o, ZT_K3_C2C, st);
    });
}

}  // extern "C"

// ================================================================================================
// Pade6stagg (periodic)
// ================================================================================================
struct pdo_pade6stagg_s {
    int gp_zsz[3], sp_zsz[3];
    double dz;
    int scheme;
    pdo_cd06stagg_t der = nullptr;
    pdo_spectral_t spectC = nullptr;   // scheme = fourierColl: the spectral type whose z transforms and tables are used (borrowed)
    // isPeriodic = .false., cd06: the nine wall handles derOO .. derSS (PadeDerOps.F90:92-110) at index 3 (bot + 1) + (top + 1),
    // bot / top = -1 odd, 0 one-sided, +1 even
    bool periodic = true;
    pdo_cd06stagg_t wall[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {
typedef int (*stagg_fn)(pdo_cd06stagg_t, const double*, double*, int, int, int, void*);

// Fourier collocation in z (spectral.F90:365-702): c2c-z forward on the first nz planes, multiply plane k by table(k), c2c-z
// backward, x 1/nz; edge outputs get plane nz+1 := plane 1.  Complex arrays live on the spectral z-pencil, real ones on the
// physical z-pencil (r2c / c2r in the reference, oddball mode untouched: zfourier_real).
// which: 0 ddz_E2C, 1 ddz_C2E, 2 interp_E2C, 3 interp_C2E, 4 d2dz2_C2C, 5 d2dz2_E2E
int pade_fourier(pdo_pade6stagg_s* p, int which, const double* in, double* out, int is_complex, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    pdo_spectral_s* s = p->spectC;
    const int nz = s->nz;
    const int* zs = is_complex ? s->si.zsz : s->pi.zsz;
    const size_t esz = is_complex ? sizeof(double2) : sizeof(double);
    const size_t plane = (size_t)zs[0] * zs[1];
    const bool edge_in = (which == 0 || which == 2 || which == 5), edge_out = (which == 1 || which == 3 || which == 5);
    const size_t bin = esz * plane * (size_t)(nz + (edge_in ? 1 : 0)), bout = esz * plane * (size_t)(nz + (edge_out ? 1 : 0));
    static const int tab_of[6] = {ZT_K3_E2C, ZT_K3_C2E, ZT_E2C, ZT_C2E, ZT_MK3SQ, ZT_MK3SQ};
    return with_device_views(in, bin, out, bout, st, [&](const void* di, void* d_o) -> int {
        if (is_complex) {
            double2* w = (double2*)d_o;
            if (di != d_o) PDO_CUDA(cudaMemcpyAsync(w, di, esz * plane * (size_t)nz, cudaMemcpyDeviceToDevice, st));
            if (int rc = zfourier_complex(s, w, tab_of[which], st)) return rc;
        } else {
            if (int rc = zfourier_real(s, (const double*)di, (double*)d_o, tab_of[which], st)) return rc;
        }
        if (edge_out) PDO_CUDA(cudaMemcpyAsync((char*)d_o + esz * plane * (size_t)nz, d_o, esz * plane, cudaMemcpyDeviceToDevice, st));
        return 0;
    });
}

int pade_apply(pdo_pade6stagg_s* p, stagg_fn fn, int which, const double* in, double* out, int is_complex, int bot, int top, void* st) {
    if (!p) return fail(PDO_E_BADARG, "null handle");
    const int* z = is_complex ? p->sp_zsz : p->gp_zsz;
    if (!p->periodic) {
        // PadeDerOps.F90:185-205, 449-482, ...: the first-order operators take bot, top in {-1, 0, +1}, the second derivatives
        // {-1, +1}; any other code gives output = 0
        const bool second = which >= 4;
        const bool ok = bot >= -1 && bot <= 1 && top >= -1 && top <= 1 && !(second && (bot == 0 || top == 0));
        if (!ok) {
            if (!out) return fail(PDO_E_BADARG, "null field pointer");
            const bool edge_out = (which == 1 || which == 3 || which == 5);
            const size_t bytes = sizeof(double) * (is_complex ? 2 : 1) * (size_t)z[0] * z[1] * (size_t)(z[2] + (edge_out ? 1 : 0));
            if (is_device_ptr(out)) PDO_CUDA(cudaMemsetAsync(out, 0, bytes, (cudaStream_t)st));
            else std::memset(out, 0, bytes);
            return 0;
        }
        return fn(p->wall[3 * (bot + 1) + (top + 1)], in, out, z[0], z[1], is_complex, st);
    }
    if (p->scheme == PDO_SCHEME_FOURIER) return pade_fourier(p, which, in, out, is_complex, st);
    return fn(p->der, in, out, z[0], z[1], is_complex, st);
}
}  // namespace

extern "C" {

int pdo_pade6stagg_init2(pdo_pade6stagg_t* h, const int gp_zsz[3], const int sp_zsz[3], double dz, int scheme, int is_periodic,
                         pdo_spectral_t spectC) {
    if (!h || !gp_zsz || !sp_zsz) return fail(PDO_E_BADARG, "null argument");
    *h = nullptr;
    if (scheme == PDO_SCHEME_FD02) return fail(PDO_E_UNSUPPORTED, "Pade6stagg: scheme fd02 is not built (cd06 and fourierColl are)");
    if (!is_periodic && scheme != PDO_SCHEME_CD06) return fail(323, "Invalid choice for numerical scheme in vertical direction");  // PadeDerOps.F90:121
    if (scheme != PDO_SCHEME_CD06 && scheme != PDO_SCHEME_FOURIER) return fail(434, "Invalid choice of numerical scheme in vertical");  // PadeDerOps.F90:84
    if (scheme == PDO_SCHEME_FOURIER) {
        if (!spectC) return fail(43, "You need to pass in a spectral derived type if you want to use Fourier differentiation in z");  // :77
        if (!spectC->periodicInZ) return fail(PDO_E_BADARG, "fourierColl needs a spectral type initialised with init_periodicInZ");
        if (spectC->si.zsz[0] != sp_zsz[0] || spectC->si.zsz[1] != sp_zsz[1] || spectC->si.zsz[2] != sp_zsz[2])
            return fail(PDO_E_BADARG, "spectral type and sp_gpC disagree on the z-pencil");
        if (spectC->pi.zsz[0] != gp_zsz[0] || spectC->pi.zsz[1] != gp_zsz[1] || spectC->pi.zsz[2] != gp_zsz[2])
            return fail(PDO_E_BADARG, "spectral type and gpC disagree on the z-pencil");
    }
    pdo_pade6stagg_s* p = new (std::nothrow) pdo_pade6stagg_s();
    if (!p) return fail(PDO_E_BADARG, "out of memory");
    std::memcpy(p->gp_zsz, gp_zsz, sizeof(int) * 3);
    std::memcpy(p->sp_zsz, sp_zsz, sizeof(int) * 3);
    p->dz = dz; p->scheme = scheme;
    p->periodic = is_periodic != 0;
    if (!p->periodic) {
        // derOO .. derSS (:92-110): the Even flag of a one-sided wall never reaches a row (the sided branch comes first)
        for (int bot = -1; bot <= 1; ++bot)
            for (int top = -1; top <= 1; ++top) {
                int rc = pdo_cd06stagg_init_nonperiodic(&p->wall[3 * (bot + 1) + (top + 1)], gp_zsz[2], dz, top == 1, bot == 1, top == 0, bot == 0);
                if (rc) { pdo_pade6stagg_destroy(p); return rc; }
            }
    } else if (scheme == PDO_SCHEME_CD06) {
        int rc = pdo_cd06stagg_init_periodic(&p->der, gp_zsz[2], dz);  // derPeriodic%init(gp%zsz(3), dz)  :79-80
        if (rc) { delete p; return rc; }
    } else {
        p->spectC = spectC;   // the tables are the spectral type's own (spectral.F90:843-856), built on first use
        if (int rc = spectral_ztables(spectC)) { delete p; return rc; }
    }
    *h = p;
    return 0;
}
int pdo_pade6stagg_init(pdo_pade6stagg_t* h, const int gp_zsz[3], const int sp_zsz[3], double dz, int scheme, int is_periodic) {
    return pdo_pade6stagg_init2(h, gp_zsz, sp_zsz, dz, scheme, is_periodic, nullptr);
}
int pdo_pade6stagg_destroy(pdo_pade6stagg_t p) {
    if (!p) return 0;
    pdo_cd06stagg_destroy(p->der);
    for (int i = 0; i < 9; ++i) pdo_cd06stagg_destroy(p->wall[i]);
    delete p;
    return 0;
}
#define PDO_PADE_FN(name, target, which)                                                                                   \
    int name(pdo_pade6stagg_t p, const double* in, double* out, int is_complex, int bot, int top, void* st) {              \
        return pade_apply(p, target, which, in, out, is_complex, bot, top, st);                                            \
    }
PDO_PADE_FN(pdo_pade6stagg_ddz_C2E, pdo_cd06stagg_ddz_C2E, 1)
PDO_PADE_FN(pdo_pade6stagg_ddz_E2C, pdo_cd06stagg_ddz_E2C, 0)
PDO_PADE_FN(pdo_pade6stagg_interpz_C2E, pdo_cd06stagg_interpz_C2E, 3)
PDO_PADE_FN(pdo_pade6stagg_interpz_E2C, pdo_cd06stagg_interpz_E2C, 2)
PDO_PADE_FN(pdo_pade6stagg_d2dz2_C2C, pdo_cd06stagg_d2dz2_C2C, 4)
PDO_PADE_FN(pdo_pade6stagg_d2dz2_E2E, pdo_cd06stagg_d2dz2_E2E, 5)

// getmodCD06stagg (PadeDerOps.F90:1034-1053)
int pdo_pade6stagg_get_modified_wavenumbers(pdo_pade6stagg_t p, const double* k, double* kp, int n) {
    if (!p || !k || !kp) return fail(PDO_E_BADARG, "null argument");
    if (p->scheme == PDO_SCHEME_FOURIER) {   // PadeDerOps.F90:1003-1004
        for (int i = 0; i < n; ++i) kp[i] = k[i];
        return 0;
    }
    const double alpha = 9.0 / 62.0, beta = 0.0, a = 63.0 / 62.0, b = 17.0 / 62.0, c = 0.0;
    for (int i = 0; i < n; ++i) {
        const double omega = k[i] * p->dz;
        double v = (2.0 * a * std::sin(omega / 2.0) + (2.0 / 3.0) * b * std::sin(3.0 * omega / 2.0) + (2.0 / 5.0) * c * std::sin(5.0 * omega / 2.0)) /
                   (1.0 + 2.0 * alpha * std::cos(omega) + 2.0 * beta * std::cos(2.0 * omega));
        kp[i] = v / p->dz;
    }
    return 0;
}

}  // extern "C"

// ================================================================================================
// padepoisson (periodic in z)
// ================================================================================================
struct pdo_padepoisson_s {
    pdo_spectral_t sp = nullptr, spE = nullptr;
    pdo_pade6stagg_t derivZ = nullptr;
    pdo_decomp_t dC = nullptr, dE = nullptr;  // spectral decompositions of the cell / edge grids (borrowed from sp / spE)
    pdo_decomp_info sC, sE;
    double *k1sq = nullptr, *k2sq = nullptr, *k3sq = nullptr;  // z-pencil slices of GetWaveNums(nx,dx)^2, (ny,dy)^2; k3mod^2
    double mfact = 1.0;
    double2 *f2d = nullptr, *f2dy = nullptr, *w2 = nullptr, *uhatInZ = nullptr, *dwdz = nullptr;
    double* div_tmp = nullptr;  // real x-pencil, used when the caller passes no divergence array
    bool alias = false;         // one rank in the column communicator: y- and z-pencil layouts coincide, transposes are skipped
    const double2* phat_y = nullptr;  // where the last projection left the pressure (y-pencil layout)
    // PeriodicInZ = .false. (walls; PadePoisson.F90:180-230, 459-623): even / odd extensions to 2 nz planes and their tables
    bool periodic_in_z = true;
    double2 *fext = nullptr, *wext = nullptr, *k3modcm = nullptr, *k3modcp = nullptr;
    double* k3sq_ext = nullptr;
    ZColsPlan ext_plan;
};

namespace {

// f2dy = i (k1 u + k2 v)   (PadePoisson.F90:392-401)
int poiss_div_xy(pdo_padepoisson_s* p, const double2* u, const double2* v, double2* out, cudaStream_t st) {
    const pdo_spectral_s* s = p->sp;
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const double *k1 = s->k1y, *k2 = s->k2;
    return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
        const double a = k1[(int)(i % n1)], b = k2[(int)((i / n1) % n2)];
        const double2 uu = u[i], vv = v[i];
        const double re = a * uu.x + b * vv.x, im = a * uu.y + b * vv.y;
        out[i] = make_double2(-im, re);
    });
}

// steps shared by PeriodicProjection / Periodic_getPressure*: leaves phat in f2d (z-pencil) and what in w2 (z-pencil)
int poiss_solve(pdo_padepoisson_s* p, const double2* uhat, const double2* vhat, const double2* what, cudaStream_t st) {
    if (int rc = poiss_div_xy(p, uhat, vhat, p->f2dy, st)) return rc;
    const double2 *uz = p->f2dy, *wz = what;
    if (!p->alias) {
        if (int rc = decomp_transpose_device(p->dC, 2, (const double*)p->f2dy, (double*)p->uhatInZ, 2, st)) return rc;
        if (int rc = decomp_transpose_device(p->dE, 2, (const double*)what, (double*)p->w2, 2, st)) return rc;
        uz = p->uhatInZ; wz = p->w2;
    }
    if (int rc = pdo_pade6stagg_ddz_E2C(p->derivZ, (const double*)wz, (double*)p->f2d, 1, 0, 0, st)) return rc;
    const long long n = vol(p->sC.zsz);
    double2* f2d = p->f2d;
    if (int rc = launch_ew(n, st, [=] __device__(long long i) { double2 a = f2d[i]; const double2 b = uz[i]; a.x += b.x; a.y += b.y; f2d[i] = a; })) return rc;
    if (int rc = fft3d_z_inplace(p->sp->ft, f2d, -1, st)) return rc;
    const int n1 = p->sC.zsz[0], n2 = p->sC.zsz[1];
    const double *k1sq = p->k1sq, *k2sq = p->k2sq, *k3sq = p->k3sq;
    const double mfact = p->mfact;
    if (int rc = launch_ew(n, st, [=] __device__(long long i) {  // f2d = -kradsq_inv f2d, mfact folded in (:413-415, 103-108)
            const int ii = (int)(i % n1);
            const long long t = i / n1;
            const int jj = (int)(t % n2), kk = (int)(t / n2);
            const double kradsq = k1sq[ii] + k2sq[jj] + k3sq[kk];
            const double m = (kradsq <= 1.e-14) ? 0.0 : -(1.0 / kradsq) * mfact;
            double2 a = f2d[i];
            a.x *= m; a.y *= m;
            f2d[i] = a;
        })) return rc;
    return fft3d_z_inplace(p->sp->ft, f2d, +1, st);
}

// w2 -= ddz_C2E(f2d); what <- w2; f2dy <- f2d; u -= i k1 p, v -= i k2 p   (:417-431)
int poiss_correct(pdo_padepoisson_s* p, double2* uhat, double2* vhat, double2* what, cudaStream_t st) {
    if (int rc = pdo_pade6stagg_ddz_C2E(p->derivZ, (const double*)p->f2d, (double*)p->dwdz, 1, 0, 0, st)) return rc;
    double2* w2 = p->alias ? what : p->w2;
    const double2* dw = p->dwdz;
    if (int rc = launch_ew(vol(p->sE.zsz), st, [=] __device__(long long i) { double2 a = w2[i]; const double2 b = dw[i]; a.x -= b.x; a.y -= b.y; w2[i] = a; })) return rc;
    const double2* ph = p->f2d;
    if (!p->alias) {
        if (int rc = decomp_transpose_device(p->dE, 3, (const double*)p->w2, (double*)what, 2, st)) return rc;
        if (int rc = decomp_transpose_device(p->dC, 3, (const double*)p->f2d, (double*)p->f2dy, 2, st)) return rc;
        ph = p->f2dy;
    }
    p->phat_y = ph;
    const pdo_spectral_s* s = p->sp;
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const double *k1 = s->k1y, *k2 = s->k2;
    return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
        const double a = k1[(int)(i % n1)], b = k2[(int)((i / n1) % n2)];
        const double2 q = ph[i];
        double2 uu = uhat[i], vv = vhat[i];
        uu.x += a * q.y; uu.y -= a * q.x;  // u - i k1 p
        vv.x += b * q.y; vv.y -= b * q.x;
        uhat[i] = uu; vhat[i] = vv;
    });
}

int poiss_divergence(pdo_padepoisson_s* p, const double2* uhat, const double2* vhat, const double2* what, double* div, cudaStream_t st) {
    const double2* wz = what;
    if (!p->alias) {
        if (int rc = decomp_transpose_device(p->dE, 2, (const double*)what, (double*)p->w2, 2, st)) return rc;
        wz = p->w2;
    }
    if (int rc = pdo_pade6stagg_ddz_E2C(p->derivZ, (const double*)wz, (double*)p->f2d, 1, -1, -1, st)) return rc;
    double2* f = p->f2d;
    if (!p->alias) {
        if (int rc = decomp_transpose_device(p->dC, 3, (const double*)p->f2d, (double*)p->f2dy, 2, st)) return rc;
        f = p->f2dy;
    }
    const pdo_spectral_s* s = p->sp;
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const double *k1 = s->k1y, *k2 = s->k2;
    if (int rc = launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {  // + i k1 u + i k2 v  (:1191-1200)
            const double a = k1[(int)(i % n1)], b = k2[(int)((i / n1) % n2)];
            const double2 uu = uhat[i], vv = vhat[i];
            double2 q = f[i];
            q.x += -a * uu.y - b * vv.y;
            q.y += a * uu.x + b * vv.x;
            f[i] = q;
        })) return rc;
    return fft3d_backward_yx(s->ft, f, div, false, st);
}

// p_maxval(maxval(a)) (use_abs = 0, as DivergenceCheck does) or of |a|
int global_max(pdo_spectral_s* s, const double* a, long long n, int use_abs, double* out, cudaStream_t st) {
    const int blocks = 1024;
    max_kernel<<<blocks, 256, 0, st>>>(a, n, use_abs, s->partial);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    double hpart[1024];
    PDO_CUDA(cudaMemcpyAsync(hpart, s->partial, sizeof(double) * blocks, cudaMemcpyDeviceToHost, st));
    PDO_CUDA(cudaStreamSynchronize(st));
    double m = -1.0e300;
    for (int i = 0; i < blocks; ++i) m = hpart[i] > m ? hpart[i] : m;
    return pdo_p_maxval(m, out);
}

// PressureProjection with walls, computeStokesPressure = .false. (PadePoisson.F90:459-623): the horizontal divergence is extended
// evenly and w oddly about both walls to 2 nz planes, one c2c-z pair solves and projects (the half-cell shifts ride on
// k3modcm / k3modcp), the upper halves come back and w is zero on both walls.
int poiss_wall_projection(pdo_padepoisson_s* p, double2* uhat, double2* vhat, double2* what, cudaStream_t st) {
    if (int rc = poiss_div_xy(p, uhat, vhat, p->f2dy, st)) return rc;                       // Step 1
    const double2 *uz = p->f2dy, *wz = what;
    if (!p->alias) {                                                                          // Step 2
        if (int rc = decomp_transpose_device(p->dC, 2, (const double*)p->f2dy, (double*)p->uhatInZ, 2, st)) return rc;
        if (int rc = decomp_transpose_device(p->dE, 2, (const double*)what, (double*)p->w2, 2, st)) return rc;
        uz = p->uhatInZ; wz = p->w2;
    }
    const int nz = p->sp->nz, n1 = p->sC.zsz[0], n2 = p->sC.zsz[1];
    const long long cols = (long long)n1 * n2, next = cols * 2 * nz;
    double2 *fe = p->fext, *we = p->wext;
    if (int rc = launch_ew(next, st, [=] __device__(long long i) {                            // Step 3
            const long long c = i % cols;
            const int kk = (int)(i / cols);
            fe[i] = kk < nz ? uz[c + cols * (nz - 1 - kk)] : uz[c + cols * (kk - nz)];
            if (kk < nz - 1) { const double2 a = wz[c + cols * (nz - 1 - kk)]; we[i] = make_double2(-a.x, -a.y); }
            else we[i] = wz[c + cols * (kk - (nz - 1))];
        })) return rc;
    if (int rc = zcols_exec(&p->ext_plan, 2 * nz, cols, fe, -1, st)) return rc;               // Step 4
    if (int rc = zcols_exec(&p->ext_plan, 2 * nz, cols, we, -1, st)) return rc;
    const double *k1sq = p->k1sq, *k2sq = p->k2sq, *k3sq = p->k3sq_ext;
    const double2 *cm = p->k3modcm, *cp = p->k3modcp;
    const double mfact = p->mfact;
    if (int rc = launch_ew(next, st, [=] __device__(long long i) {                            // Steps 5-6 (+ mfact of Step 7)
            const int ii = (int)(i % n1);
            const long long t = i / n1;
            const int jj = (int)(t % n2), kk = (int)(t / n2);
            const double kradsq = k1sq[ii] + k2sq[jj] + k3sq[kk];
            const double kinv = (kradsq <= 1.e-14) ? 0.0 : 1.0 / kradsq;
            double2 f = fe[i], w = we[i];
            const double2 a = cm[kk], b = cp[kk];
            // f = f + i cm w;  f = -f kinv
            f.x += -(a.x * w.y + a.y * w.x);
            f.y += a.x * w.x - a.y * w.y;
            f.x = -f.x * kinv; f.y = -f.y * kinv;
            // w = w - i cp f
            w.x -= -(b.x * f.y + b.y * f.x);
            w.y -= b.x * f.x - b.y * f.y;
            fe[i] = make_double2(f.x * mfact, f.y * mfact);
            we[i] = make_double2(w.x * mfact, w.y * mfact);
        })) return rc;
    if (int rc = zcols_exec(&p->ext_plan, 2 * nz, cols, fe, +1, st)) return rc;               // Step 7
    if (int rc = zcols_exec(&p->ext_plan, 2 * nz, cols, we, +1, st)) return rc;
    double2* f2d = p->f2d;
    double2* w2 = p->alias ? what : p->w2;
    if (int rc = launch_ew(cols * (nz + 1), st, [=] __device__(long long i) {
            const int kk = (int)(i / cols);
            if (kk < nz) f2d[i] = fe[i + cols * nz];
            w2[i] = (kk == 0 || kk == nz) ? make_double2(0.0, 0.0) : we[i + cols * (nz - 1)];
        })) return rc;
    const double2* ph = p->f2d;
    if (!p->alias) {                                                                          // Step 8
        if (int rc = decomp_transpose_device(p->dE, 3, (const double*)p->w2, (double*)what, 2, st)) return rc;
        if (int rc = decomp_transpose_device(p->dC, 3, (const double*)p->f2d, (double*)p->f2dy, 2, st)) return rc;
        ph = p->f2dy;
    }
    p->phat_y = ph;
    const pdo_spectral_s* s = p->sp;
    const int m1 = s->si.ysz[0], m2 = s->si.ysz[1];
    const double *k1 = s->k1y, *k2 = s->k2;
    return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {                        // Step 9
        const double a = k1[(int)(i % m1)], b = k2[(int)((i / m1) % m2)];
        const double2 q = ph[i];
        double2 uu = uhat[i], vv = vhat[i];
        uu.x += a * q.y; uu.y -= a * q.x;
        vv.x += b * q.y; vv.y -= b * q.x;
        uhat[i] = uu; vhat[i] = vv;
    });
}

int poiss_projection(pdo_padepoisson_s* p, double2* u, double2* v, double2* w, cudaStream_t st) {
    if (!p->periodic_in_z) return poiss_wall_projection(p, u, v, w, st);
    if (int rc = poiss_solve(p, u, v, w, st)) return rc;
    return poiss_correct(p, u, v, w, st);
}

int poiss_divergence_check(pdo_padepoisson_s* p, double2* u, double2* v, double2* w, double* div, bool fix, double* max_div, cudaStream_t st) {
    if (!div) div = p->div_tmp;
    const long long n = vol(p->sp->pi.xsz);
    if (int rc = poiss_divergence(p, u, v, w, div, st)) return rc;
    double md = 0.0;
    if (fix || max_div) { if (int rc = global_max(p->sp, div, n, 0, &md, st)) return rc; }
    if (fix && md > 1.e-13) {  // PadePoisson.F90:1209-1241
        if (int rc = poiss_projection(p, u, v, w, st)) return rc;
        if (int rc = poiss_divergence(p, u, v, w, div, st)) return rc;
        if (int rc = global_max(p->sp, div, n, 0, &md, st)) return rc;
        if (md > 1.e-10) { if (int rc = poiss_projection(p, u, v, w, st)) return rc; }
    }
    if (max_div) *max_div = md;
    return 0;
}

}  // namespace

extern "C" {

int pdo_padepoisson_init(pdo_padepoisson_t* h, double dx, double dy, double dz, pdo_spectral_t sp, pdo_spectral_t spE,
                         pdo_pade6stagg_t derivZ) {
    return pdo_padepoisson_init2(h, dx, dy, dz, sp, spE, derivZ, 1);
}
int pdo_padepoisson_init2(pdo_padepoisson_t* h, double dx, double dy, double dz, pdo_spectral_t sp, pdo_spectral_t spE,
                          pdo_pade6stagg_t derivZ, int periodic_in_z) {
    if (!h || !sp || !spE || !derivZ) return fail(PDO_E_BADARG, "null argument");
    *h = nullptr;
    if (spE->nz != sp->nz + 1 || spE->nx != sp->nx || spE->ny != sp->ny) return fail(PDO_E_BADARG, "spE must be the (nx, ny, nz+1) edge type of sp");
    if (periodic_in_z && !derivZ->periodic)
        return fail(PDO_E_BADARG, "padepoisson: PeriodicInZ = .true. needs a derivZ initialised with isPeriodic = .true.");
    if (!periodic_in_z && derivZ->periodic)
        return fail(PDO_E_BADARG, "padepoisson: PeriodicInZ = .false. needs a derivZ initialised with isPeriodic = .false.");
    // PadePoisson.F90:215-218 — the two decompositions must split x and y identically in the z-pencil
    if (sp->si.zst[0] != spE->si.zst[0] || sp->si.zst[1] != spE->si.zst[1])
        return fail(423, "Failed at initializing Padepoisson. sp_gp and sp_gpE have different x and y starts in z-decomp");
    pdo_padepoisson_s* p = new (std::nothrow) pdo_padepoisson_s();
    if (!p) return fail(PDO_E_BADARG, "out of memory");
    p->sp = sp; p->spE = spE; p->derivZ = derivZ;
    p->dC = fft3d_spec_decomp(sp->ft); p->dE = fft3d_spec_decomp(spE->ft);
    p->sC = sp->si; p->sE = spE->si;
    const int nz = sp->nz;
    // InitPeriodicPoissonSolver (:76-128): k1, k2 straight from GetWaveNums (no oddball flip), k3 through the z scheme's symbol
    std::vector<double> k1 = wavenums(sp->nx, dx), k2 = wavenums(sp->ny, dy), k3 = wavenums(nz, dz), k3m(nz);
    pdo_pade6stagg_get_modified_wavenumbers(derivZ, k3.data(), k3m.data(), nz);
    for (auto& v : k1) v = v * v;
    for (auto& v : k2) v = v * v;
    for (auto& v : k3m) v = v * v;
    int rc = upload(&p->k1sq, k1, p->sC.zst[0] - 1, p->sC.zsz[0]);
    if (!rc) rc = upload(&p->k2sq, k2, p->sC.zst[1] - 1, p->sC.zsz[1]);
    if (!rc) rc = upload(&p->k3sq, k3m, 0, nz);
    p->mfact = 1.0 / (double)nz;
    p->alias = (sp->p_col == 1);
    p->periodic_in_z = periodic_in_z != 0;
    cudaError_t e = cudaSuccess;
    if (!rc && !p->periodic_in_z) {
        // :183-210: k3 = GetWaveNums(2 nz, dz) through the z scheme's symbol; tfm / tfp = exp(-+ i dz/2 k3); mfact = 1 / (2 nz)
        const int nze = 2 * nz;
        std::vector<double> k3e = wavenums(nze, dz), k3me(nze), k3sq(nze);
        pdo_pade6stagg_get_modified_wavenumbers(derivZ, k3e.data(), k3me.data(), nze);
        std::vector<double2> cm(nze), cp(nze);
        for (int k = 0; k < nze; ++k) {
            const double ph = (dz / 2.0) * k3e[k];
            cm[k] = make_double2(k3me[k] * std::cos(ph), -k3me[k] * std::sin(ph));   // k3mod exp(-i dz/2 k3)
            cp[k] = make_double2(k3me[k] * std::cos(ph), k3me[k] * std::sin(ph));    // k3mod exp(+i dz/2 k3)
            k3sq[k] = k3me[k] * k3me[k];
        }
        p->mfact = 1.0 / (double)nze;
        rc = upload(&p->k3sq_ext, k3sq, 0, nze);
        const size_t ext = sizeof(double2) * (size_t)p->sC.zsz[0] * p->sC.zsz[1] * (size_t)nze;
        if (!rc) {
            e = cudaMalloc(&p->k3modcm, sizeof(double2) * nze);
            if (e == cudaSuccess) e = cudaMalloc(&p->k3modcp, sizeof(double2) * nze);
            if (e == cudaSuccess) e = cudaMemcpy(p->k3modcm, cm.data(), sizeof(double2) * nze, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(p->k3modcp, cp.data(), sizeof(double2) * nze, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMalloc(&p->fext, ext);
            if (e == cudaSuccess) e = cudaMalloc(&p->wext, ext);
            if (e != cudaSuccess) rc = fail(PDO_E_CUDA, "padepoisson wall buffers: %s", cudaGetErrorString(e));
        }
    }
    if (!rc) {
        e = cudaMalloc(&p->f2d, sizeof(double2) * (size_t)vol(p->sC.zsz));
        if (e == cudaSuccess) e = cudaMalloc(&p->uhatInZ, sizeof(double2) * (size_t)vol(p->sC.zsz));
        if (e == cudaSuccess) e = cudaMalloc(&p->f2dy, sizeof(double2) * (size_t)vol(p->sC.ysz));
        if (e == cudaSuccess) e = cudaMalloc(&p->w2, sizeof(double2) * (size_t)vol(p->sE.zsz));
        if (e == cudaSuccess) e = cudaMalloc(&p->dwdz, sizeof(double2) * (size_t)vol(p->sE.zsz));
        if (e == cudaSuccess) e = cudaMalloc(&p->div_tmp, sizeof(double) * (size_t)vol(sp->pi.xsz));
        if (e != cudaSuccess) rc = fail(PDO_E_CUDA, "padepoisson buffers: %s", cudaGetErrorString(e));
        if (!rc) {  // transpose destinations (collective, same order on every rank)
            comm_register_buffer_quiet(p->uhatInZ, sizeof(double2) * (size_t)vol(p->sC.zsz));
            comm_register_buffer_quiet(p->w2, sizeof(double2) * (size_t)vol(p->sE.zsz));
            comm_register_buffer_quiet(p->f2dy, sizeof(double2) * (size_t)vol(p->sC.ysz));
        }
    }
    if (rc) { pdo_padepoisson_destroy(p); return rc; }
    *h = p;
    return 0;
}

int pdo_padepoisson_destroy(pdo_padepoisson_t p) {
    if (!p) return 0;
    void* ptrs[] = {p->k1sq, p->k2sq, p->k3sq, p->f2d, p->f2dy, p->w2, p->uhatInZ, p->dwdz, p->div_tmp};
    for (void* q : ptrs) if (q) { comm_deregister_buffer(q); cudaFree(q); }
    void* wall[] = {p->fext, p->wext, p->k3modcm, p->k3modcp, p->k3sq_ext};
    for (void* q : wall) if (q) cudaFree(q);
    zcols_destroy(&p->ext_plan);
    delete p;
    return 0;
}

}  // extern "C"

namespace {
// Runs body(u, v, w) on device views of the three spectral arrays; host arrays are staged in and (when writable) out.
template <class Body>
int with_uvw(pdo_padepoisson_s* p, const double* u, const double* v, const double* w, bool writeback, cudaStream_t st, Body body) {
    const size_t bC = sizeof(double2) * (size_t)vol(p->sC.ysz), bE = sizeof(double2) * (size_t)vol(p->sE.ysz);
    const double* in[3] = {u, v, w};
    const size_t bytes[3] = {bC, bC, bE};
    double2* dev[3];
    bool staged[3];
    for (int i = 0; i < 3; ++i) {
        staged[i] = !is_device_ptr(in[i]);
        if (staged[i]) {
            PDO_CUDA(cudaMalloc(&dev[i], bytes[i]));
            PDO_CUDA(cudaMemcpyAsync(dev[i], in[i], bytes[i], cudaMemcpyHostToDevice, st));
        } else {
            dev[i] = (double2*)in[i];
        }
    }
    int rc = body(dev[0], dev[1], dev[2]);
    for (int i = 0; i < 3; ++i) {
        if (!staged[i]) continue;
        if (!rc && writeback) {
            if (cudaMemcpyAsync((void*)in[i], dev[i], bytes[i], cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = fail(PDO_E_CUDA, "D2H failed");
        }
        cudaStreamSynchronize(st);
        cudaFree(dev[i]);
    }
    return rc;
}
}  // namespace

extern "C" {

int pdo_padepoisson_pressure_projection(pdo_padepoisson_t p, double* uhat, double* vhat, double* what, void* stream) {
    if (!p || !uhat || !vhat || !what) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    return with_uvw(p, uhat, vhat, what, true, st, [&](double2* u, double2* v, double2* w) { return poiss_projection(p, u, v, w, st); });
}

int pdo_padepoisson_get_pressure(pdo_padepoisson_t p, const double* uhat, const double* vhat, const double* what, double* pressure,
                                 void* stream) {
    if (!p || !uhat || !vhat || !what || !pressure) return fail(PDO_E_BADARG, "null argument");
    if (!p->periodic_in_z) return fail(PDO_E_UNSUPPORTED, "padepoisson getPressure: only PressureProjection and DivergenceCheck are built for PeriodicInZ = .false.");
    cudaStream_t st = (cudaStream_t)stream;
    return with_uvw(p, uhat, vhat, what, false, st, [&](double2* u, double2* v, double2* w) -> int {
        if (int rc = poiss_solve(p, u, v, w, st)) return rc;
        const double2* ph = p->f2d;
        if (!p->alias) {
            if (int rc = decomp_transpose_device(p->dC, 3, (const double*)p->f2d, (double*)p->f2dy, 2, st)) return rc;
            ph = p->f2dy;
        }
        const size_t bytes = sizeof(double) * (size_t)vol(p->sp->pi.xsz);
        return with_device_views(pressure, 0, pressure, bytes, st, [&](const void*, void* d_o) {
            return fft3d_backward_yx(p->sp->ft, ph, (double*)d_o, false, st);
        });
    });
}

int pdo_padepoisson_get_pressure_and_update_rhs(pdo_padepoisson_t p, double* uhat, double* vhat, double* what, double* pressure,
                                                void* stream) {
    if (!p || !uhat || !vhat || !what || !pressure) return fail(PDO_E_BADARG, "null argument");
    if (!p->periodic_in_z) return fail(PDO_E_UNSUPPORTED, "padepoisson getPressureAndUpdateRHS: only PressureProjection and DivergenceCheck are built for PeriodicInZ = .false.");
    cudaStream_t st = (cudaStream_t)stream;
    return with_uvw(p, uhat, vhat, what, true, st, [&](double2* u, double2* v, double2* w) -> int {
        if (int rc = poiss_projection(p, u, v, w, st)) return rc;  // leaves phat at phat_y
        const size_t bytes = sizeof(double) * (size_t)vol(p->sp->pi.xsz);
        return with_device_views(pressure, 0, pressure, bytes, st, [&](const void*, void* d_o) {
            return fft3d_backward_yx(p->sp->ft, p->phat_y, (double*)d_o, false, st);
        });
    });
}

int pdo_padepoisson_divergence_check(pdo_padepoisson_t p, double* uhat, double* vhat, double* what, double* divergence, int fix_div,
                                     double* max_div, void* stream) {
    if (!p || !uhat || !vhat || !what) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    return with_uvw(p, uhat, vhat, what, fix_div != 0, st, [&](double2* u, double2* v, double2* w) -> int {
        if (!divergence) return poiss_divergence_check(p, u, v, w, nullptr, fix_div != 0, max_div, st);
        const size_t bytes = sizeof(double) * (size_t)vol(p->sp->pi.xsz);
        return with_device_views(divergence, 0, divergence, bytes, st, [&](const void*, void* d_o) {
            return poiss_divergence_check(p, u, v, w, (double*)d_o, fix_div != 0, max_div, st);
        });
    });
}

}  // extern "C"

// ================================================================================================
// forcingmod::HIT_shell_forcing (incompressible/forcingIsotropic.F90:45-314)
//
// The reference z-transforms three whole fields, touches Nwaves modes, and inverse-transforms three whole (almost empty)
// fields: 6 transposes + 6 FFT passes per right-hand side for O(Nwaves) numbers.  Here the forcing is evaluated where it
// lives: a direct DFT of the Nwaves columns (x, y) = (kx, ky) at the single wavenumber kz each (O(Nwaves nz) work, one small
// reduction kernel + one allreduce of 3 Nwaves complex numbers when z is distributed), and the inverse transform of a
// single-mode spectrum is the plane wave itself, added to the right-hand side in place by a second small kernel.
// Same arithmetic per mode (den, fac, conjg, the E2C / C2E shifts of w); sums over z instead of an FFT: rounding-level
// differences.
// ================================================================================================
struct pdo_hit_forcing_s {
    pdo_spectral_t spC = nullptr, spE = nullptr;   // borrowed
    double kmin = 2.0, kmax = 10.0, eps = 0.1, normfact = 1.0;
    int nwaves = 0;
    long long seed0 = 0, seed1 = 0, seed2 = 0, seed3 = 0;
    std::vector<int> waves;      // wave_x[n], wave_y[n], wave_z[n]
    bool have_waves = false, waves_dirty = false;   // dirty: the host copy is newer than d_waves
    int* d_waves = nullptr;
    double2* d_part = nullptr;   // (U, V, Wraw) per wave
};

namespace {

void hit_update_seeds(pdo_hit_forcing_s* f) {   // :122-128
    auto ab = [](long long v) { return v < 0 ? -v : v; };
    f->seed0 = ab(f->seed0 + 2223345);
    f->seed1 = ab(f->seed0 + 1423246);
    f->seed2 = ab(f->seed0 + 8723446);
    f->seed3 = ab(f->seed0 + 3423444);
}
// `count` doubles in [0, 1): SplitMix64.  Fortran's random_seed(put) / random_number (utilities/random.F90:154-174) is
// compiler-specific, so the stream is this documented generator; the reference's own draw can be
// injected with pdo_hit_forcing_set_wavenumbers.
void hit_uniform(double* out, int count, double left, double right, long long seed) {
    unsigned long long state = (unsigned long long)seed;
    const double diff = right - left;
    for (int i = 0; i < count; ++i) {
        state += 0x9E3779B97F4A7C15ULL;
        unsigned long long z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        double a = (double)(z >> 11) * (1.0 / 9007199254740992.0);
        a = diff * a;          // "array = diff*array; array = array + left"
        out[i] = a + left;
    }
}
void hit_waves_from_samples(pdo_hit_forcing_s* f, const double* kabs, const double* zeta, const double* theta) {   // :137-147
    const int n = f->nwaves;
    for (int i = 0; i < n; ++i) {
        double t = kabs[i] * std::sqrt(1 - zeta[i] * zeta[i]) * std::cos(theta[i]);
        f->waves[i] = (int)std::ceil(std::fabs(t));
        t = kabs[i] * std::sqrt(1 - zeta[i] * zeta[i]) * std::sin(theta[i]);
        f->waves[n + i] = (int)std::ceil(std::fabs(t));
        t = kabs[i] * zeta[i];
        f->waves[2 * n + i] = (int)std::ceil(std::fabs(t));
    }
    f->have_waves = true;
    f->waves_dirty = true;
}

struct HitGeom {
    int nxh, ny, nz;
    int x0, nxl;          // first global x index (0-based) and count of the spectral y-pencil
    int zc0, nzc;         // cell planes of this rank
    int ze0, nze;         // edge planes of this rank (global indices 0 .. nz)
};

__device__ __forceinline__ double2 hit_phase(int kz, int zg, int nz, double sign) {
    const int m = (int)(((long long)kz * zg) % nz);
    double sn, cs;
    sincospi(sign * 2.0 * (double)m / (double)nz, &sn, &cs);
    return make_double2(cs, sn);
}
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// block w: (U, V, Wraw)(w) = sum over the local planes of (u, v, w)_hat(kx, ky, z) e^{-2 pi i kz z / nz}
__global__ void __launch_bounds__(128) hit_reduce_kernel(const double2* __restrict__ uh, const double2* __restrict__ vh, const double2* __restrict__ wh,
                                                         const int* __restrict__ waves, int nwaves, HitGeom g, double2* __restrict__ part) {
    const int w = blockIdx.x;
    const int kx = waves[w], ky = waves[nwaves + w], kz = waves[2 * nwaves + w];
    double2 a[3] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    const int il = kx - g.x0;
    if (il >= 0 && il < g.nxl && ky >= 0 && ky < g.ny && kz >= 0 && kz < g.nz) {
        for (int zl = threadIdx.x; zl < g.nzc; zl += blockDim.x) {
            const double2 ph = hit_phase(kz, g.zc0 + zl, g.nz, -1.0);
            const size_t idx = (size_t)il + (size_t)g.nxl * ((size_t)ky + (size_t)g.ny * zl);
            const double2 pu = cmul(uh[idx], ph), pv = cmul(vh[idx], ph);
            a[0].x += pu.x; a[0].y += pu.y; a[1].x += pv.x; a[1].y += pv.y;
        }
        for (int zl = threadIdx.x; zl < g.nze; zl += blockDim.x) {
            const int zg = g.ze0 + zl;
            if (zg >= g.nz) continue;      // "this%what = this%cbuffzE(:,:,1:nz)"
            const double2 pw = cmul(wh[(size_t)il + (size_t)g.nxl * ((size_t)ky + (size_t)g.ny * zl)], hit_phase(kz, zg, g.nz, -1.0));
            a[2].x += pw.x; a[2].y += pw.y;
        }
    }
    __shared__ double sm[128][6];
    for (int c = 0; c < 3; ++c) { sm[threadIdx.x][2 * c] = a[c].x; sm[threadIdx.x][2 * c + 1] = a[c].y; }
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if (threadIdx.x < s) for (int c = 0; c < 6; ++c) sm[threadIdx.x][c] += sm[threadIdx.x + s][c];
        __syncthreads();
    }
    if (threadIdx.x < 3) part[3 * w + threadIdx.x] = make_double2(sm[0][2 * threadIdx.x], sm[0][2 * threadIdx.x + 1]);
}

// thread t: cell plane t and edge plane t of this rank; the waves are applied one after another (embed_forcing_mode :215-251,
// then the inverse z transform of each single-mode spectrum: normfactz x the plane wave)
__global__ void __launch_bounds__(128) hit_apply_kernel(double2* __restrict__ ur, double2* __restrict__ vr, double2* __restrict__ wr,
                                                        const int* __restrict__ waves, int nwaves, HitGeom g, const double2* __restrict__ part,
                                                        const double2* __restrict__ e2c, const double2* __restrict__ c2e, double normfact,
                                                        double eps) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool doC = t < g.nzc, doE = t < g.nze;
    if (!doC && !doE) return;
    const double nwr = (double)nwaves, nfz = 1.0 / (double)g.nz;
    for (int w = 0; w < nwaves; ++w) {
        const int kx = waves[w], ky = waves[nwaves + w], kz = waves[2 * nwaves + w];
        const int il = kx - g.x0;
        if (!(il >= 0 && il < g.nxl && ky >= 0 && ky < g.ny && kz >= 0 && kz < g.nz)) continue;
        const double2 U = part[3 * w], V = part[3 * w + 1], W = cmul(part[3 * w + 2], e2c[kz]);   // shiftz_E2C(what)
        const double den = (U.x * U.x + U.y * U.y) + (V.x * V.x + V.y * V.y) + (W.x * W.x + W.y * W.y) + 1.0e-14;
        const double fac = normfact * eps / den / nwr;
        if (doC) {
            const double2 ph = hit_phase(kz, g.zc0 + t, g.nz, +1.0);
            const size_t idx = (size_t)il + (size_t)g.nxl * ((size_t)ky + (size_t)g.ny * t);
            const double2 fu = cmul(make_double2(fac * U.x * nfz, -fac * U.y * nfz), ph);
            const double2 fv = cmul(make_double2(fac * V.x * nfz, -fac * V.y * nfz), ph);
            double2 a = ur[idx]; a.x += fu.x; a.y += fu.y; ur[idx] = a;
            double2 b = vr[idx]; b.x += fv.x; b.y += fv.y; vr[idx] = b;
        }
        if (doE) {
            const int zg = g.ze0 + t;
            const double2 ph = hit_phase(kz, zg >= g.nz ? 0 : zg, g.nz, +1.0);     // plane nz+1 := plane 1
            const double2 fz = cmul(make_double2(fac * W.x, -fac * W.y), c2e[kz]);  // shiftz_C2E(fzhat)
            const double2 fw = cmul(make_double2(fz.x * nfz, fz.y * nfz), ph);
            const size_t idx = (size_t)il + (size_t)g.nxl * ((size_t)ky + (size_t)g.ny * t);
            double2 a = wr[idx]; a.x += fw.x; a.y += fw.y; wr[idx] = a;
        }
    }
}

// device pointers, y-pencils of the spectral decompositions
int hit_get_rhs_dev(pdo_hit_forcing_s* f, double2* ur, double2* vr, double2* wr, const double2* uh, const double2* vh, const double2* wh,
                    bool new_timestep, cudaStream_t st) {
    const int n = f->nwaves;
    if (new_timestep) {     // :265-268
        std::vector<double> a(3 * (size_t)n);
        hit_uniform(a.data(), n, f->kmin, f->kmax, f->seed1);
        hit_uniform(a.data() + n, n, -1.0, 1.0, f->seed2);
        hit_uniform(a.data() + 2 * n, n, 0.0, 2.0 * kPi, f->seed3);
        hit_waves_from_samples(f, a.data(), a.data() + n, a.data() + 2 * n);
        hit_update_seeds(f);
    }
    if (!f->have_waves) return fail(PDO_E_BADARG, "HIT forcing: no wavenumbers yet (newTimestep was never true and none were set)");
    if (f->waves_dirty) {   // once per time step: stream-ordered copy from pageable memory (staged before the call returns)
        PDO_CUDA(cudaMemcpyAsync(f->d_waves, f->waves.data(), sizeof(int) * 3 * n, cudaMemcpyHostToDevice, st));
        f->waves_dirty = false;
    }
    if (int rc = spectral_ztables(f->spC)) return rc;
    pdo_spectral_s *C = f->spC, *E = f->spE;
    HitGeom g;
    g.nxh = C->nxh; g.ny = C->ny; g.nz = C->nz;
    g.x0 = C->si.yst[0] - 1; g.nxl = C->si.ysz[0];
    g.zc0 = C->si.yst[2] - 1; g.nzc = C->si.ysz[2];
    g.ze0 = E->si.yst[2] - 1; g.nze = E->si.ysz[2];
    hit_reduce_kernel<<<n, 128, 0, st>>>(uh, vh, wh, f->d_waves, n, g, f->d_part);
    PDO_CUDA(cudaGetLastError());
    if (int rc = comm_allreduce_sum((double*)f->d_part, 6 * n, st)) return rc;
    const int planes = g.nzc > g.nze ? g.nzc : g.nze;
    hit_apply_kernel<<<(planes + 127) / 128, 128, 0, st>>>(ur, vr, wr, f->d_waves, n, g, f->d_part, ztable(C, ZT_E2C, false), ztable(C, ZT_C2E, false),
                                                            f->normfact, f->eps);
    PDO_CUDA(cudaGetLastError());
    g_launches += 2;
    return 0;
}

}  // namespace

extern "C" {

/* hitforce%init(inputfile, sp_gpC, sp_gpE, spectC, ...) :45-120; the &HIT_Forcing namelist enters as arguments */
int pdo_hit_forcing_init(pdo_hit_forcing_t* h, pdo_spectral_t spectC, pdo_spectral_t spectE, double kmin, double kmax, int nwaves,
                         double eps_amplitude, int tid_start, int rand_seed_to_add) {
    if (!h || !spectC || !spectE) return fail(PDO_E_BADARG, "null argument");
    *h = nullptr;
    if (nwaves < 1) return fail(PDO_E_BADARG, "HIT forcing: Nwaves must be positive");
    if (!spectC->periodicInZ) return fail(PDO_E_BADARG, "HIT forcing needs a spectral type initialised with init_periodicInZ");
    if (spectE->nz != spectC->nz + 1 || spectE->nx != spectC->nx || spectE->ny != spectC->ny || spectE->si.yst[0] != spectC->si.yst[0] ||
        spectE->si.ysz[0] != spectC->si.ysz[0])
        return fail(PDO_E_BADARG, "spectE must be the (nx, ny, nz+1) edge type of spectC on the same process grid");
    pdo_hit_forcing_s* f = new (std::nothrow) pdo_hit_forcing_s();
    if (!f) return fail(PDO_E_BADARG, "out of memory");
    f->spC = spectC; f->spE = spectE;
    f->kmin = kmin; f->kmax = kmax; f->eps = eps_amplitude; f->nwaves = nwaves;
    f->seed0 = (long long)tid_start + (long long)rand_seed_to_add;   // :86
    hit_update_seeds(f);
    const double n3 = (double)spectC->nx * (double)spectC->ny * (double)spectC->nz;
    f->normfact = n3 * n3;                                             // :89
    f->waves.assign(3 * (size_t)nwaves, 0);
    cudaError_t e = cudaMalloc(&f->d_waves, sizeof(int) * 3 * nwaves);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_part, sizeof(double2) * 3 * nwaves);
    if (e != cudaSuccess) { pdo_hit_forcing_destroy(f); return fail(PDO_E_CUDA, "HIT forcing init: %s", cudaGetErrorString(e)); }
    *h = f;
    return 0;
}
int pdo_hit_forcing_destroy(pdo_hit_forcing_t f) {
    if (!f) return 0;
    if (f->d_waves) cudaFree(f->d_waves);
    if (f->d_part) cudaFree(f->d_part);
    delete f;
    return 0;
}
/* the draw of the CURRENT step, e.g. the reference RNG's wave_x / wave_y / wave_z for an A/B run */
int pdo_hit_forcing_set_wavenumbers(pdo_hit_forcing_t f, const int* wave_x, const int* wave_y, const int* wave_z) {
    if (!f || !wave_x || !wave_y || !wave_z) return fail(PDO_E_BADARG, "null argument");
    const int n = f->nwaves;
    for (int i = 0; i < n; ++i) { f->waves[i] = wave_x[i]; f->waves[n + i] = wave_y[i]; f->waves[2 * n + i] = wave_z[i]; }
    f->have_waves = true;
    f->waves_dirty = true;
    return 0;
}
int pdo_hit_forcing_get_wavenumbers(pdo_hit_forcing_t f, int* wave_x, int* wave_y, int* wave_z) {
    if (!f || !wave_x || !wave_y || !wave_z) return fail(PDO_E_BADARG, "null argument");
    const int n = f->nwaves;
    for (int i = 0; i < n; ++i) { wave_x[i] = f->waves[i]; wave_y[i] = f->waves[n + i]; wave_z[i] = f->waves[2 * n + i]; }
    return 0;
}
/* getRHS_HITforcing(urhs_xy, vrhs_xy, wrhs_xy, uhat_xy, vhat_xy, what_xy, newTimestep) :254-311; DEVICE pointers (y-pencils) */
int pdo_hit_forcing_get_rhs(pdo_hit_forcing_t f, double* urhs, double* vrhs, double* wrhs, const double* uhat, const double* vhat,
                            const double* what, int new_timestep, void* stream) {
    if (!f || !urhs || !vrhs || !wrhs || !uhat || !vhat || !what) return fail(PDO_E_BADARG, "null argument");
    for (const void* p : {(const void*)urhs, (const void*)vrhs, (const void*)wrhs, (const void*)uhat, (const void*)vhat, (const void*)what})
        if (!is_device_ptr(p)) return fail(PDO_E_BADARG, "HIT forcing works on device-resident right-hand sides");
    return hit_get_rhs_dev(f, (double2*)urhs, (double2*)vrhs, (double2*)wrhs, (const double2*)uhat, (const double2*)vhat, (const double2*)what,
                           new_timestep != 0, (cudaStream_t)stream);
}
/* test hook (host only, not in the public header): seeds after init + `updates` further update_seeds, the draw for the current seeds */
int pdo_debug_hit_draw(double kmin, double kmax, int nwaves, int tid_start, int rand_seed_to_add, int updates, long long seeds[4], int* wx,
                       int* wy, int* wz) {
    pdo_hit_forcing_s f;
    f.kmin = kmin; f.kmax = kmax; f.nwaves = nwaves;
    f.seed0 = (long long)tid_start + rand_seed_to_add;
    hit_update_seeds(&f);
    for (int i = 0; i < updates; ++i) hit_update_seeds(&f);
    f.waves.assign(3 * (size_t)nwaves, 0);
    std::vector<double> a(3 * (size_t)nwaves);
    hit_uniform(a.data(), nwaves, kmin, kmax, f.seed1);
    hit_uniform(a.data() + nwaves, nwaves, -1.0, 1.0, f.seed2);
    hit_uniform(a.data() + 2 * nwaves, nwaves, 0.0, 2.0 * kPi, f.seed3);
    hit_waves_from_samples(&f, a.data(), a.data() + nwaves, a.data() + 2 * nwaves);
    seeds[0] = f.seed0; seeds[1] = f.seed1; seeds[2] = f.seed2; seeds[3] = f.seed3;
    for (int i = 0; i < nwaves; ++i) { wx[i] = f.waves[i]; wy[i] = f.waves[nwaves + i]; wz[i] = f.waves[2 * nwaves + i]; }
    return 0;
}

}  // extern "C"

// ================================================================================================
// igrid
// ================================================================================================
struct pdo_igrid_s {
    pdo_igrid_params prm;
    double dx, dy, dz;
    pdo_spectral_t spC = nullptr, spE = nullptr;
    pdo_pade6stagg_t ops = nullptr;
    pdo_padepoisson_t poiss = nullptr;
    pdo_decomp_t dC = nullptr, dE = nullptr;  // spectral decompositions (cell / edge)
    pdo_decomp_info gC, gE, sC, sE;
    long long nRC, nRE, nYC, nYE, nZC, nZE;   // element counts: real x-pencils, complex y- and z-pencils
    int step = 0;
    double tsim = 0.0, dt = 0.0;
    // PeriodicInZ = .false.: (bottom, top) stencil codes of get_boundary_conditions_stencil (igrid.F90:5148-5204); ignored by the
    // periodic operators
    int bc[12][2] = {};
    pdo_hit_forcing_t hit = nullptr;   // useHITForcing
    // useSGS: eddy-viscosity model with a global constant (sgsmod_igrid.F90); nu on cells / edges, real work pencils for the
    // cell -> edge interpolation of nu on decomposed grids
    bool sgs_on = false, sgs_explicit_edge = false;
    SgsConst sgs{};
    double *sgs_nuC = nullptr, *sgs_nuE = nullptr, *sgs_ry = nullptr, *sgs_rzC = nullptr, *sgs_rzE = nullptr;
    bool new_timestep = true;          // igrid.F90:137, 1128, 2075
    bool alias = false;  // p_col == 1: y- and z-pencil layouts coincide
    std::vector<void*> allocs;
    // physical fields
    double *u, *v, *wC, *w, *uE, *vE, *divergence;
    double *gradC[9], *gradE[9];  // duidxjC / duidxjE in the reference's order (unused slots stay null)
    double *rbC[2], *rbE[2];
    // spectral state: S[slot][component]; slot 0 = SfieldsC/E(:,:,:,1..), 1..3 = stage arrays; R = rhs, RX = uRHSExtra
    double2 *S[4][3], *R[3], *RX[3];
    double2 *cur[3];
    double2 *whatC, *uEhat, *vEhat, *d2u, *d2v, *d2w;
    double2 *yC[2], *yE[2], *zC[2], *zE[2];
};

namespace {

int ig_alloc(pdo_igrid_s* g, void** p, size_t bytes) {
    PDO_CUDA(cudaMalloc(p, bytes ? bytes : 8));
    g->allocs.push_back(*p);
    return 0;
}
template <class T>
int ig_alloc_n(pdo_igrid_s* g, T** p, long long count) {
    if (int rc = ig_alloc(g, (void**)p, sizeof(T) * (size_t)count)) return rc;
    // every spectral (complex) array can be the destination of a y<->z transpose: make it peer-writable
    if (sizeof(T) == sizeof(double2)) comm_register_buffer_quiet(*p, sizeof(T) * (size_t)count);
    return 0;
}

inline int fftC(pdo_igrid_s* g, const double* in, double2* out, cudaStream_t st) { return fft3d_forward_xy(g->spC->ft, in, out, st); }
inline int fftE(pdo_igrid_s* g, const double* in, double2* out, cudaStream_t st) { return fft3d_forward_xy(g->spE->ft, in, out, st); }
inline int ifftC(pdo_igrid_s* g, const double2* in, double* out, cudaStream_t st) { return fft3d_backward_yx(g->spC->ft, in, out, false, st); }
inline int ifftE(pdo_igrid_s* g, const double2* in, double* out, cudaStream_t st) { return fft3d_backward_yx(g->spE->ft, in, out, false, st); }
// When the column communicator has one rank the y- and z-pencils of a spectral array are the same memory layout, so a
// "transpose" is the identity: z-operators then read the y-pencil array directly (zview*) and write straight into the
// y-pencil destination (ztarget* / zcommit*) instead of paying two device copies per visit to z.
inline int y2zC(pdo_igrid_s* g, const double2* s, double2* d, cudaStream_t st) { return decomp_transpose_device(g->dC, 2, (const double*)s, (double*)d, 2, st); }
inline int z2yC(pdo_igrid_s* g, const double2* s, double2* d, cudaStream_t st) { return decomp_transpose_device(g->dC, 3, (const double*)s, (double*)d, 2, st); }
inline int y2zE(pdo_igrid_s* g, const double2* s, double2* d, cudaStream_t st) { return decomp_transpose_device(g->dE, 2, (const double*)s, (double*)d, 2, st); }
inline int z2yE(pdo_igrid_s* g, const double2* s, double2* d, cudaStream_t st) { return decomp_transpose_device(g->dE, 3, (const double*)s, (double*)d, 2, st); }

inline int zviewC(pdo_igrid_s* g, const double2* s, double2* buf, const double2** out, cudaStream_t st) {
    if (g->alias) { *out = s; return 0; }
    *out = buf;
    return y2zC(g, s, buf, st);
}
inline int zviewE(pdo_igrid_s* g, const double2* s, double2* buf, const double2** out, cudaStream_t st) {
    if (g->alias) { *out = s; return 0; }
    *out = buf;
    return y2zE(g, s, buf, st);
}
inline double2* ztarget(pdo_igrid_s* g, double2* ydst, double2* buf) { return g->alias ? ydst : buf; }
inline int zcommitC(pdo_igrid_s* g, const double2* z, double2* ydst, cudaStream_t st) { return g->alias ? 0 : z2yC(g, z, ydst, st); }
inline int zcommitE(pdo_igrid_s* g, const double2* z, double2* ydst, cudaStream_t st) { return g->alias ? 0 : z2yE(g, z, ydst, st); }

#define IG(expr) do { if (int _rc = (expr)) return _rc; } while (0)
#define ZOP(fn, in, out) IG(fn(g->ops, (const double*)(in), (double*)(out), 1, 0, 0, st))
#define ZOPB(fn, in, out, q) IG(fn(g->ops, (const double*)(in), (double*)(out), 1, g->bc[q][0], g->bc[q][1], st))
enum { BC_W = 0, BC_U, BC_V, BC_WdUdz, BC_WdVdz, BC_WdWdz, BC_WW, BC_UW, BC_VW, BC_dUdz, BC_dVdz, BC_dWdz };

// out = a*b (+ c*d)
int mul2(double* out, const double* a, const double* b, const double* c, const double* d, long long n, cudaStream_t st) {
    if (c) return launch_ew(n, st, [=] __device__(long long i) { out[i] = a[i] * b[i] + c[i] * d[i]; });
    return launch_ew(n, st, [=] __device__(long long i) { out[i] = a[i] * b[i]; });
}
// out = (a - b) * c  (+ (d - e) * f): the rotational form's products (igrid.F90:1527-1549: "T = dvdx - dudy; T = T*v")
int muldiff(double* out, const double* a, const double* b, const double* c, const double* d, const double* e, const double* f, long long n,
            cudaStream_t st) {
    if (d) return launch_ew(n, st, [=] __device__(long long i) {
        const double t1 = (a[i] - b[i]) * c[i];
        const double t2 = (d[i] - e[i]) * f[i];
        out[i] = t1 + t2;
    });
    return launch_ew(n, st, [=] __device__(long long i) { out[i] = (a[i] - b[i]) * c[i]; });
}
// dst += src (complex arrays viewed as doubles)
int cadd(double2* dst, const double2* src, long long n, cudaStream_t st) {
    double* d = (double*)dst;
    const double* s = (const double*)src;
    return launch_ew(2 * n, st, [=] __device__(long long i) { d[i] += s[i]; });
}
// dst += i k f  with k along index 1 (which = 1) or 2 (which = 2) of a y-pencil (mTimes_ik*_ip followed by the add)
int cadd_ik(pdo_spectral_s* s, int which, double2* dst, const double2* f, cudaStream_t st) {
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const double* k = which == 1 ? s->k1y : s->k2;
    const int w = which;
    return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
        const double kv = (w == 1) ? k[(int)(i % n1)] : k[(int)((i / n1) % n2)];
        const double2 q = f[i];
        double2 a = dst[i];
        a.x += -kv * q.y; a.y += kv * q.x;
        dst[i] = a;
    });
}
// out = sum_i c_i x_i over complex arrays (as doubles); up to five terms, out may alias any x_i
struct Lin5 { double c[5]; const double* x[5]; int n; };
int lincomb(double2* out, const Lin5& L, long long ncplx, cudaStream_t st) {
    double* o = (double*)out;
    const Lin5 l = L;
    return launch_ew(2 * ncplx, st, [=] __device__(long long i) {
        double acc = l.c[0] * l.x[0][i];
        for (int t = 1; t < l.n; ++t) acc += l.c[t] * l.x[t][i];
        o[i] = acc;
    });
}

// ---- igrid.F90:1020-1037
int ig_dealias_fields(pdo_igrid_s* g, cudaStream_t st) {
    IG(spectral_dealias(g->spC, g->cur[0], st));
    IG(spectral_dealias(g->spC, g->cur[1], st));
    if (g->prm.wall_bounded) return spectral_dealias(g->spE, g->cur[2], st);   // igrid.F90:1029-1031: spectE's 2-D mask
    double2* we = ztarget(g, g->cur[2], g->zE[0]);
    if (!g->alias) IG(y2zE(g, g->cur[2], we, st));
    IG(spectral_dealias_edge(g->spC, we, st));
    return zcommitE(g, we, g->cur[2], st);
}

// ---- igrid.F90:1423-1447
int ig_interp_primitive(pdo_igrid_s* g, cudaStream_t st) {
    const double2* z = nullptr;
    IG(zviewE(g, g->cur[2], g->zE[0], &z, st));
    double2* t = ztarget(g, g->whatC, g->zC[0]);
    ZOPB(pdo_pade6stagg_interpz_E2C, z, t, BC_W);
    IG(zcommitC(g, t, g->whatC, st));
    IG(ifftC(g, g->whatC, g->wC, st));
    for (int c = 0; c < 2; ++c) {
        double2* eh = c == 0 ? g->uEhat : g->vEhat;
        IG(zviewC(g, g->cur[c], g->zC[0], &z, st));
        t = ztarget(g, eh, g->zE[0]);
        ZOPB(pdo_pade6stagg_interpz_C2E, z, t, c == 0 ? BC_U : BC_V);
        IG(zcommitE(g, t, eh, st));
        IG(ifftE(g, eh, c == 0 ? g->uE : g->vE, st));
    }
    return 0;
}

// ---- igrid.F90:2553-2683.  Slots: 0 dudx 1 dudy 2 dudz 3 dvdx 4 dvdy 5 dvdz 6 dwdx 7 dwdy 8 dwdz (C: cell values, E: edge values)
int ig_compute_duidxj(pdo_igrid_s* g, cudaStream_t st) {
    pdo_spectral_s *C = g->spC, *E = g->spE;
    const bool visc = !g->prm.is_inviscid;
    // i k f goes into a scratch array that nobody reads again: the 1/(nx ny) of the inverse is folded into that pass and
    // the inverse transform consumes the scratch directly (no intent(in) staging copy)
    const double nf2 = 1.0 / ((double)g->prm.nx * (double)g->prm.ny);
    auto dC = [&](int which, const double2* fhat, double* out) -> int {
        if (!out) return 0;
        IG(spectral_mtimes(C, which, fhat, g->yC[0], st, nf2));
        return fft3d_backward_yx_scratch(C->ft, g->yC[0], out, st);
    };
    auto dE = [&](int which, const double2* fhat, double* out) -> int {
        if (!out) return 0;
        IG(spectral_mtimes(E, which, fhat, g->yE[0], st, nf2));
        return fft3d_backward_yx_scratch(E->ft, g->yE[0], out, st);
    };
    IG(dC(1, g->cur[0], g->gradC[0])); IG(dE(1, g->uEhat, g->gradE[0]));
    IG(dC(2, g->cur[0], g->gradC[1])); IG(dE(2, g->uEhat, g->gradE[1]));
    IG(dC(1, g->cur[1], g->gradC[3])); IG(dE(1, g->vEhat, g->gradE[3]));
    IG(dC(2, g->cur[1], g->gradC[4])); IG(dE(2, g->vEhat, g->gradE[4]));
    IG(dC(1, g->whatC, g->gradC[6])); IG(dE(1, g->cur[2], g->gradE[6]));
    IG(dC(2, g->whatC, g->gradC[7])); IG(dE(2, g->cur[2], g->gradE[7]));
    // dwdz (and its edge interpolant), d2wdz2
    const double2* wz = nullptr;
    IG(zviewE(g, g->cur[2], g->zE[0], &wz, st));
    double2* dwz = ztarget(g, g->yC[0], g->zC[0]);
    ZOPB(pdo_pade6stagg_ddz_E2C, wz, dwz, BC_W);
    IG(zcommitC(g, dwz, g->yC[0], st));
    IG(ifftC(g, g->yC[0], g->gradC[8], st));
    if (g->gradE[8]) {
        double2* t = ztarget(g, g->yE[0], g->zE[1]);
        ZOPB(pdo_pade6stagg_interpz_C2E, dwz, t, BC_dWdz);
        IG(zcommitE(g, t, g->yE[0], st));
        IG(ifftE(g, g->yE[0], g->gradE[8], st));
    }
    if (visc) {
        double2* t = ztarget(g, g->d2w, g->zE[1]);
        ZOPB(pdo_pade6stagg_d2dz2_E2E, wz, t, BC_W);
        IG(zcommitE(g, t, g->d2w, st));
    }
    // dudz / dvdz on edges, their cell interpolants, and the viscous second derivatives
    for (int c = 0; c < 2; ++c) {
        const double2* fz = nullptr;
        IG(zviewC(g, g->cur[c], g->zC[0], &fz, st));
        double2* te = ztarget(g, g->yE[0], g->zE[0]);
        ZOPB(pdo_pade6stagg_ddz_C2E, fz, te, c == 0 ? BC_U : BC_V);
        IG(zcommitE(g, te, g->yE[0], st));
        IG(ifftE(g, g->yE[0], g->gradE[2 + 3 * c], st));
        if (visc) {
            double2* d2 = c == 0 ? g->d2u : g->d2v;
            double2* td = ztarget(g, d2, g->zC[1]);
            if (g->prm.use_d2dz2_c2c) {
                ZOPB(pdo_pade6stagg_d2dz2_C2C, fz, td, c == 0 ? BC_U : BC_V);
            } else {
                ZOPB(pdo_pade6stagg_ddz_C2E, fz, g->zE[1], c == 0 ? BC_U : BC_V);
                ZOPB(pdo_pade6stagg_ddz_E2C, g->zE[1], td, c == 0 ? BC_dUdz : BC_dVdz);
            }
            IG(zcommitC(g, td, d2, st));
        }
        if (g->gradC[2 + 3 * c]) {
            double2* tc = ztarget(g, g->yC[0], g->zC[0]);
            ZOPB(pdo_pade6stagg_interpz_E2C, te, tc, c == 0 ? BC_dUdz : BC_dVdz);
            IG(zcommitC(g, tc, g->yC[0], st));
            IG(ifftC(g, g->yC[0], g->gradC[2 + 3 * c], st));
        }
    }
    return 0;
}

// ---- igrid.F90:1572-1679 into (ru, rv, rw)
int ig_nonlinear_skew(pdo_igrid_s* g, double2* ru, double2* rv, double2* rw, cudaStream_t st) {
    pdo_spectral_s *C = g->spC, *E = g->spE;
    double *T1C = g->rbC[0], *T1E = g->rbE[0];
    double2 *fT1C = g->yC[0], *fT1E = g->yE[0], *fT2E = g->yE[1];
    double **GC = g->gradC, **GE = g->gradE;
    const double2* z = nullptr;
    double2* t = nullptr;
    // u_rhs = interp_E2C(fft(dudz w)) + fft(dudx u + dudy v); same for v
    for (int c = 0; c < 2; ++c) {
        double2* r = c == 0 ? ru : rv;
        IG(mul2(T1C, GC[3 * c + 0], g->u, GC[3 * c + 1], g->v, g->nRC, st));
        IG(mul2(T1E, GE[3 * c + 2], g->w, nullptr, nullptr, g->nRE, st));
        IG(fftC(g, T1C, fT1C, st));
        IG(fftE(g, T1E, fT1E, st));
        IG(zviewE(g, fT1E, g->zE[0], &z, st));
        t = ztarget(g, r, g->zC[0]);
        ZOPB(pdo_pade6stagg_interpz_E2C, z, t, c == 0 ? BC_WdUdz : BC_WdVdz);
        IG(zcommitC(g, t, r, st));
        IG(cadd(r, fT1C, g->nYC, st));
    }
    // w_rhs = interp_C2E(fft(dwdz wC)) + fft(dwdx uE + dwdy vE)
    IG(mul2(T1E, GE[6], g->uE, GE[7], g->vE, g->nRE, st));
    IG(fftE(g, T1E, fT2E, st));
    IG(mul2(T1C, GC[8], g->wC, nullptr, nullptr, g->nRC, st));
    IG(fftC(g, T1C, fT1C, st));
    IG(zviewC(g, fT1C, g->zC[0], &z, st));
    t = ztarget(g, rw, g->zE[0]);
    ZOPB(pdo_pade6stagg_interpz_C2E, z, t, BC_WdWdz);
    IG(zcommitE(g, t, rw, st));
    IG(cadd(rw, fT2E, g->nYE, st));
    // conservative half: d(uu)/dx, d(vv)/dy, d(wC wC)/dz, d(uv)/dy & /dx, d(uE w)/dz & /dx, d(vE w)/dz & /dy
    IG(mul2(T1C, g->u, g->u, nullptr, nullptr, g->nRC, st));
    IG(fftC(g, T1C, fT1C, st));
    IG(cadd_ik(C, 1, ru, fT1C, st));
    IG(mul2(T1C, g->v, g->v, nullptr, nullptr, g->nRC, st));
    IG(fftC(g, T1C, fT1C, st));
    IG(cadd_ik(C, 2, rv, fT1C, st));
    IG(mul2(T1C, g->wC, g->wC, nullptr, nullptr, g->nRC, st));
    IG(fftC(g, T1C, fT1C, st));
    IG(zviewC(g, fT1C, g->zC[0], &z, st));
    t = ztarget(g, fT1E, g->zE[0]);
    ZOPB(pdo_pade6stagg_ddz_C2E, z, t, BC_WW);
    IG(zcommitE(g, t, fT1E, st));
    IG(cadd(rw, fT1E, g->nYE, st));
    IG(mul2(T1C, g->u, g->v, nullptr, nullptr, g->nRC, st));
    IG(fftC(g, T1C, fT1C, st));
    IG(cadd_ik(C, 2, ru, fT1C, st));
    IG(cadd_ik(C, 1, rv, fT1C, st));
    for (int c = 0; c < 2; ++c) {
        IG(mul2(T1E, c == 0 ? g->uE : g->vE, g->w, nullptr, nullptr, g->nRE, st));
        IG(fftE(g, T1E, fT1E, st));
        IG(zviewE(g, fT1E, g->zE[0], &z, st));
        t = ztarget(g, fT1C, g->zC[0]);
        ZOPB(pdo_pade6stagg_ddz_E2C, z, t, c == 0 ? BC_UW : BC_VW);
        IG(zcommitC(g, t, fT1C, st));
        IG(cadd(c == 0 ? ru : rv, fT1C, g->nYC, st));
        IG(cadd_ik(E, c == 0 ? 1 : 2, rw, fT1E, st));
    }
    return 0;
}

// ---- igrid.F90:1527-1555 into (ru, rv, rw): u x omega; the products with w live on the edge grid and come back through
// interpz_E2C.  Gradient slots: C 1 dudy, 3 dvdx; E 2 dudz, 5 dvdz, 6 dwdx, 7 dwdy.
int ig_nonlinear_rot(pdo_igrid_s* g, double2* ru, double2* rv, double2* rw, cudaStream_t st) {
    double *T1C = g->rbC[0], *T1E = g->rbE[0];
    double2 *fT1C = g->yC[0], *fT1E = g->yE[0];
    double **GC = g->gradC, **GE = g->gradE;
    const double2* z = nullptr;
    double2* t = nullptr;
    for (int c = 0; c < 2; ++c) {
        double2* r = c == 0 ? ru : rv;
        // c = 0: (dvdx - dudy) v and (dwdx - dudz) w;   c = 1: (dudy - dvdx) u and (dwdy - dvdz) w
        if (c == 0) IG(muldiff(T1C, GC[3], GC[1], g->v, nullptr, nullptr, nullptr, g->nRC, st));
        else IG(muldiff(T1C, GC[1], GC[3], g->u, nullptr, nullptr, nullptr, g->nRC, st));
        IG(fftC(g, T1C, fT1C, st));
        if (c == 0) IG(muldiff(T1E, GE[6], GE[2], g->w, nullptr, nullptr, nullptr, g->nRE, st));
        else IG(muldiff(T1E, GE[7], GE[5], g->w, nullptr, nullptr, nullptr, g->nRE, st));
        IG(fftE(g, T1E, fT1E, st));
        IG(zviewE(g, fT1E, g->zE[0], &z, st));
        t = ztarget(g, r, g->zC[0]);
        ZOP(pdo_pade6stagg_interpz_E2C, z, t);
        IG(zcommitC(g, t, r, st));
        IG(cadd(r, fT1C, g->nYC, st));
    }
    // w_rhs = fft((dudz - dwdx) uE + (dvdz - dwdy) vE)
    IG(muldiff(T1E, GE[2], GE[6], g->uE, GE[5], GE[7], g->vE, g->nRE, st));
    return fftE(g, T1E, rw, st);
}

// rhs = -half*rhs (skew-symmetric form only), then addViscousTerm (igrid.F90:1663-1665, 1914-1941), one pass per component
int ig_finish_rhs(pdo_igrid_s* g, double2* ru, double2* rv, double2* rw, cudaStream_t st) {
    const bool visc = !g->prm.is_inviscid;
    const double scale = g->prm.rotational_advection ? 1.0 : -0.5;
    const double oneByRe = visc ? 1.0 / g->prm.Re : 0.0;
    for (int c = 0; c < 3; ++c) {
        pdo_spectral_s* s = c < 2 ? g->spC : g->spE;
        double2* r = c == 0 ? ru : (c == 1 ? rv : rw);
        const double2* f = g->cur[c];
        const double2* d2 = c == 0 ? g->d2u : (c == 1 ? g->d2v : g->d2w);
        const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
        const double *k1 = s->k1y, *k2 = s->k2;
        IG(launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
            double2 a = r[i];
            a.x = scale * a.x; a.y = scale * a.y;
            if (visc) {
                const double ka = k1[(int)(i % n1)], kb = k2[(int)((i / n1) % n2)];
                const double ksq = ka * ka + kb * kb;  // kabs_sq = k1**2 + k2**2 (spectral.F90:1093-1099)
                const double2 q = f[i], dd = d2[i];
                a.x += oneByRe * (-ksq * q.x + dd.x);
                a.y += oneByRe * (-ksq * q.y + dd.y);
            }
            r[i] = a;
        }));
    }
    return 0;
}

// ---- Step 6 of populate_rhs: the SGS term (sgsmod_igrid.F90:156-268), eddy-viscosity models with a global constant ----
struct Grad9 { const double* p[9]; };

// nu = cmodel_global * kernel(duidxj)   (get_SGS_kernel + multiply_by_model_constant, eddyViscosity.F90:38-95)
int sgs_nu(const SgsConst& c, const Grad9& G, double* nu, long long n, cudaStream_t st) {
    return launch_ew(n, st, [=] __device__(long long i) {
        double d[9], S[6];
#pragma unroll
        for (int k = 0; k < 9; ++k) d[k] = G.p[k][i];
        sgs_sij(d, S);
        nu[i] = c.cmodel * sgs_kernel_point(c, d, S);
    });
}
// tau = -2 nu S: S = a (diagonal components) or 0.5 (a + b)
int sgs_tau(double* tau, const double* nu, const double* a, const double* b, long long n, cudaStream_t st) {
    if (b) return launch_ew(n, st, [=] __device__(long long i) { tau[i] = -2.0 * nu[i] * (0.5 * (a[i] + b[i])); });
    return launch_ew(n, st, [=] __device__(long long i) { tau[i] = -2.0 * nu[i] * a[i]; });
}
// dst -= src (complex arrays viewed as doubles)
int csub(double2* dst, const double2* src, long long n, cudaStream_t st) {
    double* d = (double*)dst;
    const double* sp = (const double*)src;
    return launch_ew(2 * n, st, [=] __device__(long long i) { d[i] -= sp[i]; });
}
// dst -= i k f  (mTimes_ik*_ip / _oop followed by "rhs = rhs - cbuffy")
int csub_ik(pdo_spectral_s* s, int which, double2* dst, const double2* f, cudaStream_t st) {
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const double* k = which == 1 ? s->k1y : s->k2;
    const int w = which;
    return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
        const double kv = (w == 1) ? k[(int)(i % n1)] : k[(int)((i / n1) % n2)];
        const double2 q = f[i];
        double2 a = dst[i];
        a.x -= -kv * q.y; a.y -= kv * q.x;
        dst[i] = a;
    });
}
// interpolate_eddy_viscosity(.true.) (eddyViscosity.F90:97-113): x -> y -> z on gpC, interpz_C2E on the REAL array, z -> y -> x on
// gpE, negative values clipped; a transpose inside a 1-rank group is the identity and is skipped
int sgs_interp_nu(pdo_igrid_s* g, cudaStream_t st) {
    pdo_decomp_t pC = fft3d_phys_decomp(g->spC->ft), pE = fft3d_phys_decomp(g->spE->ft);
    const bool tx = g->spC->p_row > 1, tz = g->spC->p_col > 1;
    const double* a = g->sgs_nuC;
    // (tx && !tz): the y-pencil IS the z-pencil; it lands in rzC so that the edge result can take ry
    double* ybuf = (tx && !tz) ? g->sgs_rzC : g->sgs_ry;
    if (tx) { IG(decomp_transpose_device(pC, 0, a, ybuf, 1, st)); a = ybuf; }
    if (tz) { IG(decomp_transpose_device(pC, 2, a, g->sgs_rzC, 1, st)); a = g->sgs_rzC; }
    double* zout = tz ? g->sgs_rzE : (tx ? g->sgs_ry : g->sgs_nuE);
    IG(pdo_pade6stagg_interpz_C2E(g->ops, a, zout, 0, 0, 0, st));
    const double* b = zout;
    if (tz) { double* yd = tx ? g->sgs_ry : g->sgs_nuE; IG(decomp_transpose_device(pE, 3, b, yd, 1, st)); b = yd; }
    if (tx) IG(decomp_transpose_device(pE, 1, b, g->sgs_nuE, 1, st));
    double* nuE = g->sgs_nuE;
    return launch_ew(g->nRE, st, [=] __device__(long long i) { if (nuE[i] < 0.0) nuE[i] = 0.0; });
}

int ig_sgs_rhs(pdo_igrid_s* g, double2* ru, double2* rv, double2* rw, cudaStream_t st) {
    pdo_spectral_s *C = g->spC, *E = g->spE;
    Grad9 GC, GE;
    for (int k = 0; k < 9; ++k) { GC.p[k] = g->gradC[k]; GE.p[k] = g->gradE[k]; }
    double **dC = g->gradC, **dE = g->gradE;
    // getTauSGS :156-203
    IG(sgs_nu(g->sgs, GC, g->sgs_nuC, g->nRC, st));
    if (g->sgs_explicit_edge) IG(sgs_nu(g->sgs, GE, g->sgs_nuE, g->nRE, st));
    else IG(sgs_interp_nu(g, st));
    double *TC = g->rbC[0], *TE = g->rbE[0];
    double2 *fC = g->yC[0], *fE = g->yE[0], *gC2 = g->yC[1], *gE2 = g->yE[1];
    const double2* z = nullptr;
    double2* t = nullptr;
    // ddx(tau11) -> urhs
    IG(sgs_tau(TC, g->sgs_nuC, dC[0], nullptr, g->nRC, st));
    IG(fftC(g, TC, fC, st));
    IG(csub_ik(C, 1, ru, fC, st));
    // ddy(tau22) -> vrhs
    IG(sgs_tau(TC, g->sgs_nuC, dC[4], nullptr, g->nRC, st));
    IG(fftC(g, TC, fC, st));
    IG(csub_ik(C, 2, rv, fC, st));
    // ddz(tau33) -> wrhs
    IG(sgs_tau(TC, g->sgs_nuC, dC[8], nullptr, g->nRC, st));
    IG(fftC(g, TC, fC, st));
    IG(zviewC(g, fC, g->zC[0], &z, st));
    t = ztarget(g, gE2, g->zE[0]);
    ZOP(pdo_pade6stagg_ddz_C2E, z, t);
    IG(zcommitE(g, t, gE2, st));
    IG(csub(rw, gE2, g->nYE, st));
    // tau12: ddx -> vrhs, ddy -> urhs
    IG(sgs_tau(TC, g->sgs_nuC, dC[1], dC[3], g->nRC, st));
    IG(fftC(g, TC, fC, st));
    IG(csub_ik(C, 1, rv, fC, st));
    IG(csub_ik(C, 2, ru, fC, st));
    // tau13 (edges): ddz -> urhs, ddx -> wrhs;  tau23: ddz -> vrhs, ddy -> wrhs
    for (int c = 0; c < 2; ++c) {
        IG(sgs_tau(TE, g->sgs_nuE, c == 0 ? dE[2] : dE[5], c == 0 ? dE[6] : dE[7], g->nRE, st));
        IG(fftE(g, TE, fE, st));
        IG(zviewE(g, fE, g->zE[0], &z, st));
        t = ztarget(g, gC2, g->zC[0]);
        ZOP(pdo_pade6stagg_ddz_E2C, z, t);
        IG(zcommitC(g, t, gC2, st));
        IG(csub(c == 0 ? ru : rv, gC2, g->nYC, st));
        IG(csub_ik(E, c == 0 ? 1 : 2, rw, fE, st));
    }
    return 0;
}

int ig_populate_rhs(pdo_igrid_s* g, double2** r, cudaStream_t st) {
    if (g->prm.rotational_advection) IG(ig_nonlinear_rot(g, r[0], r[1], r[2], st));
    else IG(ig_nonlinear_skew(g, r[0], r[1], r[2], st));
    IG(ig_finish_rhs(g, r[0], r[1], r[2], st));
    if (g->sgs_on) IG(ig_sgs_rhs(g, r[0], r[1], r[2], st));   // Step 6 (igrid.F90:1866-1871)
    if (g->hit)   // Step 8 (igrid.F90:1907-1910)
        IG(hit_get_rhs_dev(g->hit, r[0], r[1], r[2], g->cur[0], g->cur[1], g->cur[2], g->new_timestep, st));
    g->new_timestep = false;
    return 0;
}

// ---- igrid.F90:1961-1990
int ig_project_and_prep(pdo_igrid_s* g, bool already_projected, cudaStream_t st) {
    IG(ig_dealias_fields(g, st));
    if (!already_projected) {
        IG(poiss_projection(g->poiss, g->cur[0], g->cur[1], g->cur[2], st));
        if (g->prm.t_divergence_check > 0 && g->step % g->prm.t_divergence_check == 0)
            IG(poiss_divergence_check(g->poiss, g->cur[0], g->cur[1], g->cur[2], g->divergence, true, nullptr, st));
    }
    IG(ifftC(g, g->cur[0], g->u, st));
    IG(ifftC(g, g->cur[1], g->v, st));
    IG(ifftE(g, g->cur[2], g->w, st));
    IG(ig_interp_primitive(g, st));
    return ig_compute_duidxj(g, st);
}

int ig_stage_update(pdo_igrid_s* g, int dst_slot, int nterms, const double* coef, double2* const (*terms)[3], cudaStream_t st) {
    for (int c = 0; c < 3; ++c) {
        Lin5 L;
        L.n = nterms;
        for (int t = 0; t < nterms; ++t) { L.c[t] = coef[t]; L.x[t] = (const double*)terms[t][c]; }
        IG(lincomb(g->S[dst_slot][c], L, c < 2 ? g->nYC : g->nYE, st));
    }
    for (int c = 0; c < 3; ++c) g->cur[c] = g->S[dst_slot][c];
    return 0;
}

// ---- igrid.F90:1105-1173
int ig_tvd_rk3(pdo_igrid_s* g, double dt, cudaStream_t st) {
    IG(ig_populate_rhs(g, g->R, st));
    { double c[2] = {1.0, dt}; double2* const t[2][3] = {{g->S[0][0], g->S[0][1], g->S[0][2]}, {g->R[0], g->R[1], g->R[2]}};
      IG(ig_stage_update(g, 1, 2, c, t, st)); }
    IG(ig_project_and_prep(g, false, st));
    IG(ig_populate_rhs(g, g->R, st));
    { double c[3] = {3.0 / 4.0, 1.0 / 4.0, (1.0 / 4.0) * dt};
      double2* const t[3][3] = {{g->S[0][0], g->S[0][1], g->S[0][2]}, {g->S[1][0], g->S[1][1], g->S[1][2]}, {g->R[0], g->R[1], g->R[2]}};
      IG(ig_stage_update(g, 1, 3, c, t, st)); }
    IG(ig_project_and_prep(g, false, st));
    IG(ig_populate_rhs(g, g->R, st));
    { double c[3] = {1.0 / 3.0, 2.0 / 3.0, (2.0 / 3.0) * dt};
      double2* const t[3][3] = {{g->S[0][0], g->S[0][1], g->S[0][2]}, {g->S[1][0], g->S[1][1], g->S[1][2]}, {g->R[0], g->R[1], g->R[2]}};
      IG(ig_stage_update(g, 0, 3, c, t, st)); }
    return ig_project_and_prep(g, false, st);
}

// ---- igrid.F90:1176-1299
int ig_ssp_rk45(pdo_igrid_s* g, double dt, cudaStream_t st) {
    const double b01 = 0.39175222657189, b12 = 0.368410593050371, b23 = 0.25189177427169, b34 = 0.54497475022852;
    const double b35 = 0.06369246866629, b45 = 0.22600748323690;
    const double a20 = 0.444370493651235, a21 = 0.555629506348765;
    const double a30 = 0.620101851488403, a32 = 0.379898148511597;
    const double a40 = 0.17807995439313, a43 = 0.821920045606868;
    const double a52 = 0.517231671970585, a53 = 0.096059710526147, a54 = 0.386708617503269;
#define SL(s) {g->S[s][0], g->S[s][1], g->S[s][2]}
#define RR {g->R[0], g->R[1], g->R[2]}
#define RXX {g->RX[0], g->RX[1], g->RX[2]}
    IG(ig_populate_rhs(g, g->R, st));
    { double c[2] = {1.0, b01 * dt}; double2* const t[2][3] = {SL(0), RR}; IG(ig_stage_update(g, 1, 2, c, t, st)); }
    IG(ig_project_and_prep(g, false, st));
    IG(ig_populate_rhs(g, g->R, st));
    { double c[3] = {a20, a21, b12 * dt}; double2* const t[3][3] = {SL(0), SL(1), RR}; IG(ig_stage_update(g, 2, 3, c, t, st)); }
    IG(ig_project_and_prep(g, false, st));
    IG(ig_populate_rhs(g, g->R, st));
    { double c[3] = {a30, a32, b23 * dt}; double2* const t[3][3] = {SL(0), SL(2), RR}; IG(ig_stage_update(g, 3, 3, c, t, st)); }
    IG(ig_project_and_prep(g, false, st));
    IG(ig_populate_rhs(g, g->R, st));
    { double c[3] = {a40, a43, b34 * dt}; double2* const t[3][3] = {SL(0), SL(3), RR}; IG(ig_stage_update(g, 0, 3, c, t, st)); }
    IG(ig_project_and_prep(g, false, st));
    IG(ig_populate_rhs(g, g->RX, st));
    { double c[5] = {a52, a53, b35 * dt, a54, b45 * dt}; double2* const t[5][3] = {SL(2), SL(3), RR, SL(0), RXX};
      IG(ig_stage_update(g, 0, 5, c, t, st)); }
#undef SL
#undef RR
#undef RXX
    return ig_project_and_prep(g, false, st);
}

int copy_in(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    PDO_CUDA(cudaMemcpyAsync(dst, src, bytes, is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    return 0;
}

}  // namespace

extern "C" {

int pdo_igrid_destroy(pdo_igrid_t g) {
    if (!g) return 0;
    for (void* p : g->allocs) { comm_deregister_buffer(p); cudaFree(p); }
    pdo_padepoisson_destroy(g->poiss);
    pdo_hit_forcing_destroy(g->hit);
    pdo_pade6stagg_destroy(g->ops);
    pdo_spectral_destroy(g->spE);
    pdo_spectral_destroy(g->spC);
    delete g;
    return 0;
}

// From (u, v, w) on the x-pencils to a consistent state (igrid.F90:625-655): transforms, dealiasing, projection, back to
// physical space, interpolations and the velocity gradients.  u, v, w: host or device.
static int ig_set_fields(pdo_igrid_s* g, const double* u, const double* v, const double* w, cudaStream_t st) {
    IG(copy_in(g->u, u, sizeof(double) * g->nRC, st));
    IG(copy_in(g->v, v, sizeof(double) * g->nRC, st));
    IG(copy_in(g->w, w, sizeof(double) * g->nRE, st));
    IG(fftC(g, g->u, g->cur[0], st));
    IG(fftC(g, g->v, g->cur[1], st));
    IG(fftE(g, g->w, g->cur[2], st));
    IG(ig_dealias_fields(g, st));
    IG(poiss_divergence_check(g->poiss, g->cur[0], g->cur[1], g->cur[2], g->divergence, false, nullptr, st));
    IG(poiss_projection(g->poiss, g->cur[0], g->cur[1], g->cur[2], st));
    IG(ifftC(g, g->cur[0], g->u, st));
    IG(ifftC(g, g->cur[1], g->v, st));
    IG(ifftE(g, g->cur[2], g->w, st));
    IG(ig_interp_primitive(g, st));
    IG(ig_compute_duidxj(g, st));
    PDO_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int pdo_igrid_init(pdo_igrid_t* h, const pdo_igrid_params* p, const double* u, const double* v, const double* w) {
    if (!h || !p || !u || !v || !w) return fail(PDO_E_BADARG, "null argument");
    *h = nullptr;
    if ((p->nx % 2) || (p->ny % 2) || (p->nz % 2))
        return fail(423, "The code hasn't been tested for odd values of Nx, Ny or Nz");  // igrid.F90:437-445
    if (p->time_stepping_scheme != 1 && p->time_stepping_scheme != 2)
        return fail(PDO_E_UNSUPPORTED, "TimeSteppingScheme must be 1 (TVD-RK3) or 2 (SSP-RK45); Adams-Bashforth is out of scope");
    if (p->wall_bounded) {
        if (p->fourier_collocation_z) return fail(123, "If you use Fourier Collocation in Z, the problem must be periodic in Z.");   // igrid.F90:458-460
        if (p->bot_wall == 3 || p->top_wall == 3) return fail(PDO_E_UNSUPPORTED, "wall-model walls (3) are out of scope: no-slip (1) and slip (2) are built");
        if (p->bot_wall != 1 && p->bot_wall != 2) return fail(423, "Invalid choice for BOTTOM WALL BCs");                          // :5180
        if (p->top_wall != 1 && p->top_wall != 2) return fail(13, "Invalid choice for TOP WALL BCs");                               // :5204
    }
    pdo_igrid_s* g = new (std::nothrow) pdo_igrid_s();
    if (!g) return fail(PDO_E_BADARG, "out of memory");
    std::memset(g->gradC, 0, sizeof(g->gradC));
    std::memset(g->gradE, 0, sizeof(g->gradE));
    g->prm = *p;
    if (g->prm.dealias_fact <= 0.0) g->prm.dealias_fact = 2.0 / 3.0;
    if (p->wall_bounded) {   // get_boundary_conditions_stencil (igrid.F90:5148-5204); index 0 bottom, 1 top
        const int walls[2] = {p->bot_wall, p->top_wall};
        for (int sd = 0; sd < 2; ++sd) {
            g->bc[BC_W][sd] = -1; g->bc[BC_WdWdz][sd] = -1; g->bc[BC_WW][sd] = +1; g->bc[BC_dWdz][sd] = 0;
            if (walls[sd] == 1) {        // no-slip: w = 0 and dwdz = 0, so w is extended evenly
                g->bc[BC_U][sd] = -1; g->bc[BC_V][sd] = -1; g->bc[BC_dUdz][sd] = 0; g->bc[BC_dVdz][sd] = 0;
                g->bc[BC_WdUdz][sd] = 0; g->bc[BC_WdVdz][sd] = 0; g->bc[BC_UW][sd] = +1; g->bc[BC_VW][sd] = +1;
                g->bc[BC_W][sd] = +1; g->bc[BC_WdWdz][sd] = -1; g->bc[BC_WW][sd] = +1; g->bc[BC_dWdz][sd] = -1;
            } else {                     // slip
                g->bc[BC_U][sd] = +1; g->bc[BC_V][sd] = +1; g->bc[BC_dUdz][sd] = -1; g->bc[BC_dVdz][sd] = -1;
                g->bc[BC_WdUdz][sd] = +1; g->bc[BC_WdVdz][sd] = +1; g->bc[BC_UW][sd] = -1; g->bc[BC_VW][sd] = -1;
            }
        }
        g->prm.use_d2dz2_c2c = 1;   // uBC, vBC are +-1 for these walls (:2653, 2671)
    }
    g->dx = p->Lx / p->nx; g->dy = p->Ly / p->ny; g->dz = p->Lz / p->nz;
    cudaStream_t st = nullptr;
    const int perz = p->wall_bounded ? 0 : 1;
    int rc = pdo_spectral_init(&g->spC, p->nx, p->ny, p->nz, g->dx, g->dy, g->dz, p->p_row, p->p_col, 0, perz, g->prm.dealias_fact);
    if (!rc) rc = pdo_spectral_init(&g->spE, p->nx, p->ny, p->nz + 1, g->dx, g->dy, g->dz, p->p_row, p->p_col, 0, 0, g->prm.dealias_fact);
    if (!rc) {
        g->gC = g->spC->pi; g->gE = g->spE->pi; g->sC = g->spC->si; g->sE = g->spE->si;
        rc = pdo_pade6stagg_init2(&g->ops, g->gC.zsz, g->sC.zsz, g->dz, p->fourier_collocation_z ? PDO_SCHEME_FOURIER : PDO_SCHEME_CD06, perz,
                                  g->spC);   // igrid.F90:500
    }
    if (!rc) rc = pdo_padepoisson_init2(&g->poiss, g->dx, g->dy, g->dz, g->spC, g->spE, g->ops, perz);   // :585-586
    if (rc) { pdo_igrid_destroy(g); return rc; }
    g->dC = fft3d_spec_decomp(g->spC->ft); g->dE = fft3d_spec_decomp(g->spE->ft);
    g->alias = (g->spC->p_col == 1);
    g->nRC = vol(g->gC.xsz); g->nRE = vol(g->gE.xsz);
    g->nYC = vol(g->sC.ysz); g->nYE = vol(g->sE.ysz);
    g->nZC = vol(g->sC.zsz); g->nZE = vol(g->sE.zsz);
    const bool all = p->compute_all_gradients != 0, visc = !p->is_inviscid;
#define AL(ptr, cnt) do { if (!rc) rc = ig_alloc_n(g, &(ptr), (cnt)); } while (0)
    AL(g->u, g->nRC); AL(g->v, g->nRC); AL(g->wC, g->nRC); AL(g->divergence, g->nRC);
    AL(g->w, g->nRE); AL(g->uE, g->nRE); AL(g->vE, g->nRE);
    const bool needC[9] = {true, true, all, true, true, all, all, all, true};
    const bool needE[9] = {all, all, true, all, all, true, true, true, all};
    for (int i = 0; i < 9; ++i) { if (needC[i]) AL(g->gradC[i], g->nRC); if (needE[i]) AL(g->gradE[i], g->nRE); }
    for (int i = 0; i < 2; ++i) { AL(g->rbC[i], g->nRC); AL(g->rbE[i], g->nRE); AL(g->yC[i], g->nYC); AL(g->yE[i], g->nYE); AL(g->zC[i], g->nZC); AL(g->zE[i], g->nZE); }
    const int nslots = p->time_stepping_scheme == 1 ? 2 : 4;
    for (int s = 0; s < 4; ++s)
        for (int c = 0; c < 3; ++c) { g->S[s][c] = nullptr; if (s < nslots) AL(g->S[s][c], c < 2 ? g->nYC : g->nYE); }
    for (int c = 0; c < 3; ++c) { AL(g->R[c], c < 2 ? g->nYC : g->nYE); g->RX[c] = nullptr; if (p->time_stepping_scheme == 2) AL(g->RX[c], c < 2 ? g->nYC : g->nYE); }
    AL(g->whatC, g->nYC); AL(g->uEhat, g->nYE); AL(g->vEhat, g->nYE);
    g->d2u = g->d2v = g->d2w = nullptr;
    if (visc) { AL(g->d2u, g->nYC); AL(g->d2v, g->nYC); AL(g->d2w, g->nYE); }
#undef AL
    if (rc) { pdo_igrid_destroy(g); return rc; }
    for (int c = 0; c < 3; ++c) g->cur[c] = g->S[0][c];
    // igrid.F90:625-655
    rc = ig_set_fields(g, u, v, w, st);
    if (rc) { pdo_igrid_destroy(g); return rc; }
    *h = g;
    return 0;
}

int pdo_igrid_time_advance(pdo_igrid_t g, double dt, void* stream) {
    if (!g) return fail(PDO_E_BADARG, "null handle");
    cudaStream_t st = (cudaStream_t)stream;
    g->dt = dt;
    int rc = g->prm.time_stepping_scheme == 1 ? ig_tvd_rk3(g, dt, st) : ig_ssp_rk45(g, dt, st);
    if (rc) return rc;
    g->step += 1;       // wrapup_timestep (igrid.F90:2067-2075)
    g->tsim += dt;
    g->new_timestep = true;
    return 0;
}

int pdo_igrid_get_field(pdo_igrid_t g, int which, double* out, void* stream) {
    if (!g || !out) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const void* src = nullptr;
    size_t bytes = 0;
    switch (which) {
        case 0: src = g->u; bytes = sizeof(double) * g->nRC; break;
        case 1: src = g->v; bytes = sizeof(double) * g->nRC; break;
        case 2: src = g->w; bytes = sizeof(double) * g->nRE; break;
        case 3: src = g->wC; bytes = sizeof(double) * g->nRC; break;
        case 4: src = g->uE; bytes = sizeof(double) * g->nRE; break;
        case 5: src = g->vE; bytes = sizeof(double) * g->nRE; break;
        case 6: src = g->divergence; bytes = sizeof(double) * g->nRC; break;
        case 10: src = g->cur[0]; bytes = sizeof(double2) * g->nYC; break;
        case 11: src = g->cur[1]; bytes = sizeof(double2) * g->nYC; break;
        case 12: src = g->cur[2]; bytes = sizeof(double2) * g->nYE; break;
        default: return fail(PDO_E_BADARG, "unknown field id %d", which);
    }
    PDO_CUDA(cudaMemcpyAsync(out, src, bytes, is_device_ptr(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    PDO_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int pdo_igrid_get_decomp_info(pdo_igrid_t g, int which, pdo_decomp_info* info) {
    if (!g || !info) return fail(PDO_E_BADARG, "null argument");
    *info = which == 0 ? g->gC : which == 1 ? g->gE : which == 2 ? g->sC : g->sE;
    return 0;
}

/* useSGS = .true. with the &SGS_MODEL entries in scope: SGSModelID 0 Smagorinsky / 1 sigma / 2 AMD, Csgs, explicitCalcEdgeEddyViscosity;
   global model constant (no wall damping, no dynamic procedure, no wall model).  Needs compute_all_gradients. */
int pdo_igrid_enable_sgs(pdo_igrid_t g, int sgs_model_id, double csgs, int explicit_calc_edge_eddy_viscosity) {
    if (!g) return fail(PDO_E_BADARG, "null handle");
    if (sgs_model_id < 0 || sgs_model_id > 2) return fail(213, "Incorrect choice for SGS model ID.");   // init_destroy_sgs_igrid.F90:165
    if (!g->prm.compute_all_gradients) return fail(PDO_E_BADARG, "the SGS models need all eighteen velocity gradients: init with compute_all_gradients");
    if (g->prm.wall_bounded) return fail(PDO_E_UNSUPPORTED, "SGS with walls (wall damping / wall models, non-periodic filter width) is out of scope");
    const double dx = g->dx, dy = g->dy, dz = g->dz;
    SgsConst c{};
    c.mid = sgs_model_id;
    const double deltaLES = std::pow(1.5 * dx * 1.5 * dy * 1.5 * dz, 1.0 / 3.0);   // isPeriodic branch (smagorinsky.F90:17, sigma.F90:15)
    if (sgs_model_id < 2) c.cmodel = (csgs * deltaLES) * (csgs * deltaLES);
    else {                                                                          // AMD.F90:1-14
        const double poincare = g->prm.fourier_collocation_z ? 1.0 / std::sqrt(12.0) : 1.0 / std::sqrt(10.0);   // PadeDerOps.F90:1018-1031
        c.cx = csgs * dx * std::sqrt(1.0 / 12.0);
        c.cy = csgs * dy * std::sqrt(1.0 / 12.0);
        c.cz = csgs * dz * poincare;
        c.cmodel = 1.0;
    }
    g->sgs = c;
    g->sgs_explicit_edge = explicit_calc_edge_eddy_viscosity != 0;
    int rc = 0;
    if (!g->sgs_nuC) rc = ig_alloc_n(g, &g->sgs_nuC, g->nRC);
    if (!rc && !g->sgs_nuE) rc = ig_alloc_n(g, &g->sgs_nuE, g->nRE);
    const bool tx = g->spC->p_row > 1, tz = g->spC->p_col > 1;
    if (!rc && !g->sgs_explicit_edge && (tx || tz)) {
        const long long ny = std::max(vol(g->gC.ysz), vol(g->gE.ysz));
        if (!g->sgs_ry) rc = ig_alloc_n(g, &g->sgs_ry, ny);
        if (!rc && !g->sgs_rzC) rc = ig_alloc_n(g, &g->sgs_rzC, std::max(vol(g->gC.zsz), vol(g->gC.ysz)));
        if (!rc && !g->sgs_rzE) rc = ig_alloc_n(g, &g->sgs_rzE, vol(g->gE.zsz));
    }
    if (rc) return rc;
    g->sgs_on = true;
    return 0;
}
/* test hook (host only): nu = cmodel * kernel for one point, and S_ij */
int pdo_debug_sgs_point(int mid, double cmodel, double cx, double cy, double cz, const double* d9, double* nu, double* S6) {
    SgsConst c{mid, cmodel, cx, cy, cz};
    sgs_sij(d9, S6);
    *nu = c.cmodel * sgs_kernel_point(c, d9, S6);
    return 0;
}
int pdo_igrid_enable_hit_forcing(pdo_igrid_t g, double kmin, double kmax, int nwaves, double eps_amplitude, int rand_seed_to_add) {
    if (!g) return fail(PDO_E_BADARG, "null handle");
    pdo_hit_forcing_destroy(g->hit);
    g->hit = nullptr;
    return pdo_hit_forcing_init(&g->hit, g->spC, g->spE, kmin, kmax, nwaves, eps_amplitude, g->step, rand_seed_to_add);   // tidStart = this%step
}
int pdo_igrid_get_state(pdo_igrid_t g, int* step, double* tsim) {
    if (!g) return fail(PDO_E_BADARG, "null handle");
    if (step) *step = g->step;
    if (tsim) *tsim = g->tsim;
    return 0;
}

// ---- restart / field files in the reference's format (igrid.F90:2719-2823; 2DECOMP io_write_one.f90) ----
static std::string restart_name(const char* dir, int rid, const char* tag, int tid) {
    char name[64];
    std::snprintf(name, sizeof(name), "RESTART_Run%02d_%s.%06d", rid, tag, tid);   // "(A7,A4,I2.2,A3,I6.6)"
    return std::string(dir ? dir : ".") + "/" + name;
}
/* dumpRestartFile :2763-2799: u, v on gpC and w on gpE as flat global arrays + the info file with tsim in g15.5, at this%step */
int pdo_igrid_dump_restart(pdo_igrid_t g, const char* outputdir, int run_id) {
    if (!g) return fail(PDO_E_BADARG, "null handle");
    PDO_CUDA(cudaDeviceSynchronize());
    pdo_decomp_t dC = fft3d_phys_decomp(g->spC->ft), dE = fft3d_phys_decomp(g->spE->ft);
    if (int rc = pdo_decomp_write_one(dC, 1, g->u, 1, restart_name(outputdir, run_id, "u", g->step).c_str())) return rc;
    if (int rc = pdo_decomp_write_one(dC, 1, g->v, 1, restart_name(outputdir, run_id, "v", g->step).c_str())) return rc;
    if (int rc = pdo_decomp_write_one(dE, 1, g->w, 1, restart_name(outputdir, run_id, "w", g->step).c_str())) return rc;
    int rc = 0;
    if (pdo_comm_rank() == 0) {
        char line[16];
        pdo_io_format_g15_5(g->tsim, line);
        FILE* f = std::fopen(restart_name(outputdir, run_id, "info", g->step).c_str(), "w");
        if (!f) rc = fail(PDO_E_BADARG, "cannot write the restart info file in '%s'", outputdir ? outputdir : ".");
        else { std::fprintf(f, "%s\n", line); std::fclose(f); }
    }
    double dummy;
    if (int b = pdo_p_sum(0.0, &dummy)) return b;   // mpi_barrier :2801
    return rc;
}
/* readRestartFile :2719-2761 followed by what init does with freshly read fields (:589-591, 625-655): step = tid, tsim from the
   info file (rank 0 reads, everyone gets it), fields projected and all dependent state rebuilt */
int pdo_igrid_read_restart(pdo_igrid_t g, const char* inputdir, int run_id, int tid) {
    if (!g) return fail(PDO_E_BADARG, "null handle");
    cudaStream_t st = nullptr;
    pdo_decomp_t dC = fft3d_phys_decomp(g->spC->ft), dE = fft3d_phys_decomp(g->spE->ft);
    // rbC / rbE scratch pencils receive the file contents; ig_set_fields copies them into u, v, w
    double *ru = g->rbC[0], *rv = g->rbC[1], *rw = g->rbE[0];
    if (int rc = pdo_decomp_read_one(dC, 1, ru, 1, restart_name(inputdir, run_id, "u", tid).c_str())) return rc;
    if (int rc = pdo_decomp_read_one(dC, 1, rv, 1, restart_name(inputdir, run_id, "v", tid).c_str())) return rc;
    if (int rc = pdo_decomp_read_one(dE, 1, rw, 1, restart_name(inputdir, run_id, "w", tid).c_str())) return rc;
    double tsim = 0.0;
    int bad = 0;
    if (pdo_comm_rank() == 0) {
        FILE* f = std::fopen(restart_name(inputdir, run_id, "info", tid).c_str(), "r");
        if (!f || std::fscanf(f, "%lf", &tsim) != 1) bad = 1;
        if (f) std::fclose(f);
    }
    // mpi_bcast(tsim) from rank 0: the other ranks contribute zero to a sum
    double tsum = 0.0, badsum = 0.0;
    if (int b = pdo_p_sum(pdo_comm_rank() == 0 ? tsim : 0.0, &tsum)) return b;
    if (int b = pdo_p_sum((double)bad, &badsum)) return b;
    if (badsum != 0.0) return fail(PDO_E_BADARG, "cannot read the restart info file in '%s'", inputdir ? inputdir : ".");
    for (int c = 0; c < 3; ++c) g->cur[c] = g->S[0][c];
    if (int rc = ig_set_fields(g, ru, rv, rw, st)) return rc;
    g->tsim = tsum;
    g->step = tid;
    return 0;
}
/* dumpFullField(arr, label, gp2use) :2806-2823 for the fields the handle owns (ids of pdo_igrid_get_field: 0 u, 1 v, 2 w, 3 wC,
   4 uE, 5 vE, 6 divergence): "Run<rid>_<label>_t<step>.out", x-pencil of gpC, or of gpE for the edge fields */
int pdo_igrid_dump_full_field(pdo_igrid_t g, int which, const char* label4, const char* outputdir, int run_id) {
    if (!g || !label4) return fail(PDO_E_BADARG, "null argument");
    const double* src = nullptr;
    bool edge = false;
    switch (which) {
        case 0: src = g->u; break;
        case 1: src = g->v; break;
        case 2: src = g->w; edge = true; break;
        case 3: src = g->wC; break;
        case 4: src = g->uE; edge = true; break;
        case 5: src = g->vE; edge = true; break;
        case 6: src = g->divergence; break;
        default: return fail(PDO_E_BADARG, "unknown field id %d", which);
    }
    PDO_CUDA(cudaDeviceSynchronize());
    char name[64];
    std::snprintf(name, sizeof(name), "Run%02d_%.4s_t%06d.out", run_id, label4, g->step);   // "(A3,I2.2,A1,A4,A2,I6.6,A4)"
    const std::string fname = std::string(outputdir ? outputdir : ".") + "/" + name;
    return pdo_decomp_write_one(fft3d_phys_decomp((edge ? g->spE : g->spC)->ft), 1, src, 1, fname.c_str());
}

int pdo_igrid_compute_delta_t(pdo_igrid_t g, double cfl, double* dt, void* stream) {
    if (!g || !dt) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    double* rb = g->rbC[0];
    const double *u = g->u, *v = g->v, *wC = g->wC;
    const double ox = 1.0 / g->dx, oy = 1.0 / g->dy, oz = 1.0 / g->dz;
    IG(launch_ew(g->nRC, st, [=] __device__(long long i) { rb[i] = fabs(ox * u[i]) + fabs(oy * v[i]) + fabs(oz * wC[i]); }));
    double tsmax = 0.0;
    IG(global_max(g->spC, rb, g->nRC, 0, &tsmax, st));
    double d = cfl / tsmax;
    if (!g->prm.is_inviscid) {
        double m = g->dx < g->dy ? g->dx : g->dy;
        m = m < g->dz ? m : g->dz;
        const double tv = cfl * g->prm.Re * (m * m);
        d = d < tv ? d : tv;
    }
    *dt = d;
    return 0;
}

int pdo_igrid_max_divergence(pdo_igrid_t g, double* max_div, void* stream) {
    if (!g || !max_div) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    IG(poiss_divergence(g->poiss, g->cur[0], g->cur[1], g->cur[2], g->divergence, st));
    return global_max(g->spC, g->divergence, g->nRC, 1, max_div, st);
}

}  // extern "C"


// ================================================================================================
// igrid_Operators_Periodic::Ops_Periodic (igrid_operators_periodic.F90:13-161): Fourier operators on x-pencil fields of a
// triply periodic box — compositions of the spectral type's transforms, its pointwise passes and PoissonPeriodic
// ================================================================================================
struct pdo_ops_periodic_s {
    pdo_spectral_t spect = nullptr;
    pdo_poisson_t poiss = nullptr;
    double2* cbuffy1 = nullptr;                     // spectral y-pencil
    double *rbuffy = nullptr, *rbuffz1 = nullptr;   // physical y- / z-pencils (allocated only where they differ from the x- / y-pencil)
};

extern "C" {

/* init(nx, ny, nz, dx, dy, dz, gp, InputDir, OutputDir) :86-109; gp enters as its process grid (0, 0 = 1 x nproc) */
int pdo_ops_periodic_init(pdo_ops_periodic_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    pdo_ops_periodic_s* o = new (std::nothrow) pdo_ops_periodic_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    // spect%init("x", nx, ny, nz, dx, dy, dz, "four", "2/3rd", 2, fixOddball=.false., init_periodicInZ=.TRUE., dealiasF=2/3)  :94-95
    int rc = pdo_spectral_init(&o->spect, nx, ny, nz, dx, dy, dz, p_row, p_col, 0, 1, 2.0 / 3.0);
    // poiss%init(dx, dy, dz, gp, 1, .true., GetKmod_Fourier x 3): the spectral wavenumbers themselves  :107-108
    if (!rc) rc = pdo_poisson_init(&o->poiss, nx, ny, nz, dx, dy, dz, o->spect->p_row, o->spect->p_col, 1, nullptr, nullptr, nullptr);
    if (!rc) {
        pdo_spectral_s* s = o->spect;
        cudaError_t e = cudaMalloc(&o->cbuffy1, sizeof(double2) * (size_t)vol(s->si.ysz));
        if (e == cudaSuccess && s->p_row > 1) e = cudaMalloc(&o->rbuffy, sizeof(double) * (size_t)vol(s->pi.ysz));
        if (e == cudaSuccess && s->p_col > 1) e = cudaMalloc(&o->rbuffz1, sizeof(double) * (size_t)vol(s->pi.zsz));
        if (e != cudaSuccess) rc = fail(PDO_E_CUDA, "Ops_Periodic buffers: %s", cudaGetErrorString(e));
    }
    if (rc) { pdo_ops_periodic_destroy(o); return rc; }
    *h = o;
    return 0;
}
int pdo_ops_periodic_destroy(pdo_ops_periodic_t o) {
    if (!o) return 0;
    if (o->cbuffy1) cudaFree(o->cbuffy1);
    if (o->rbuffy) cudaFree(o->rbuffy);
    if (o->rbuffz1) cudaFree(o->rbuffz1);
    pdo_poisson_destroy(o->poiss);
    pdo_spectral_destroy(o->spect);
    delete o;
    return 0;
}
/* link_spect :46-52 */
pdo_spectral_t pdo_ops_periodic_spect(pdo_ops_periodic_t o) { return o ? o->spect : nullptr; }

// ddx :117-125, ddy :127-135 (which = 1, 2), dealiasField :56-62 (which = 0): fft, one pointwise pass, ifft
static int ops_periodic_xy(pdo_ops_periodic_t o, int which, const double* f, double* out, void* stream) {
    if (!o || !f || !out) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    pdo_spectral_s* s = o->spect;
    const size_t bytes = sizeof(double) * (size_t)vol(s->pi.xsz);
    return with_device_views(f, bytes, out, bytes, st, [&](const void* di, void* d_o) -> int {
        if (int rc = fft3d_forward_xy(s->ft, (const double*)di, o->cbuffy1, st)) return rc;
        if (which == 0) { if (int rc = spectral_dealias(s, o->cbuffy1, st)) return rc; }
        else if (int rc = spectral_mtimes(s, which, o->cbuffy1, o->cbuffy1, st)) return rc;
        return fft3d_backward_yx(s->ft, o->cbuffy1, (double*)d_o, false, st);
    });
}
int pdo_ops_periodic_ddx(pdo_ops_periodic_t o, const double* f, double* dfdx, void* st) { return ops_periodic_xy(o, 1, f, dfdx, st); }
int pdo_ops_periodic_ddy(pdo_ops_periodic_t o, const double* f, double* dfdy, void* st) { return ops_periodic_xy(o, 2, f, dfdy, st); }
int pdo_ops_periodic_dealias_field(pdo_ops_periodic_t o, double* f, void* st) { return ops_periodic_xy(o, 0, f, f, st); }

/* ddz :149-160: x -> y -> z, spect%ddz_C2C_real_inplace, z -> y -> x (a transpose inside a 1-rank group is the identity and is skipped) */
int pdo_ops_periodic_ddz(pdo_ops_periodic_t o, const double* f, double* dfdz, void* stream) {
    if (!o || !f || !dfdz) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    pdo_spectral_s* s = o->spect;
    pdo_decomp_t gp = fft3d_phys_decomp(s->ft);
    const size_t bytes = sizeof(double) * (size_t)vol(s->pi.xsz);
    return with_device_views(f, bytes, dfdz, bytes, st, [&](const void* di, void* d_o) -> int {
        const double* a = (const double*)di;
        double* out = (double*)d_o;
        const bool tx = s->p_row > 1, tz = s->p_col > 1;
        double* ydst = tx ? o->rbuffy : out;          // where the y-pencil result lives
        if (tx) { if (int rc = decomp_transpose_device(gp, 0, a, o->rbuffy, 1, st)) return rc; a = o->rbuffy; }
        if (tz) {
            if (int rc = decomp_transpose_device(gp, 2, a, o->rbuffz1, 1, st)) return rc;
            if (int rc = zfourier_real(s, o->rbuffz1, o->rbuffz1, ZT_K3_C2C, st)) return rc;
            if (int rc = decomp_transpose_device(gp, 3, o->rbuffz1, ydst, 1, st)) return rc;
        } else if (int rc = zfourier_real(s, a, ydst, ZT_K3_C2C, st)) return rc;
        if (tx) return decomp_transpose_device(gp, 1, o->rbuffy, out, 1, st);
        return 0;
    });
}
/* ddz_cmplx2cmplx :137-145: complex y-pencil of the spectral decomposition, in place */
int pdo_ops_periodic_ddz_cmplx2cmplx(pdo_ops_periodic_t o, double* fhat, void* stream) {
    if (!o || !fhat) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    pdo_spectral_s* s = o->spect;
    pdo_decomp_t spec = fft3d_spec_decomp(s->ft);
    const size_t bytes = sizeof(double2) * (size_t)vol(s->si.ysz);
    return with_device_views(fhat, bytes, fhat, bytes, st, [&](const void* di, void* d_o) -> int {
        if (di != d_o) PDO_CUDA(cudaMemcpyAsync(d_o, di, bytes, cudaMemcpyDeviceToDevice, st));
        double2* w = (double2*)d_o;
        if (s->p_col == 1) return zfourier_complex(s, w, ZT_K3_C2C, st);
        if (int rc = decomp_transpose_device(spec, 2, (const double*)w, (double*)s->ctmpz, 2, st)) return rc;
        if (int rc = zfourier_complex(s, s->ctmpz, ZT_K3_C2C, st)) return rc;
        return decomp_transpose_device(spec, 3, (const double*)s->ctmpz, (double*)w, 2, st);
    });
}
/* ReadField3D :162-187 / WriteField3D :189-205: "<dir>/Run<runID>_<label>_t<tidx>.out", x-pencil of gp through decomp_2d_io */
static std::string ops_periodic_fname(const char* dir, const char* label4, int tidx, int run_id) {
    char name[64];
    std::snprintf(name, sizeof(name), "Run%02d_%.4s_t%06d.out", run_id, label4, tidx);   // "(A3,I2.2,A1,A4,A2,I6.6,A4)"
    return std::string(dir ? dir : ".") + "/" + name;
}
int pdo_ops_periodic_write_field3d(pdo_ops_periodic_t o, const double* field, const char* label4, int tidx, int run_id, const char* outputdir) {
    if (!o || !field || !label4) return fail(PDO_E_BADARG, "null argument");
    return pdo_decomp_write_one(fft3d_phys_decomp(o->spect->ft), 1, field, 1, ops_periodic_fname(outputdir, label4, tidx, run_id).c_str());
}
int pdo_ops_periodic_read_field3d(pdo_ops_periodic_t o, double* field, const char* label4, int tidx, int run_id, const char* inputdir) {
    if (!o || !field || !label4) return fail(PDO_E_BADARG, "null argument");
    return pdo_decomp_read_one(fft3d_phys_decomp(o->spect->ft), 1, field, 1, ops_periodic_fname(inputdir, label4, tidx, run_id).c_str());   // missing file -> 321
}
/* SolvePoisson_oop :70-76 (p != rhs), SolvePoisson_ip :78-84 (p == rhs) */
int pdo_ops_periodic_solve_poisson(pdo_ops_periodic_t o, const double* rhs, double* p, void* stream) {
    if (!o) return fail(PDO_E_BADARG, "null handle");
    return pdo_poisson_solve(o->poiss, rhs, p, stream);
}

}  // extern "C"
