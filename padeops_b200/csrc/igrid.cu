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
//
// One translation unit, several files: the sections are kept in ig_*.inc.cuh and textually included below in dependency order
// (they share the file-local helpers that follow and each other's internals):
//   ig_spectral.inc.cuh      spectral type            ig_pade6stagg.inc.cuh   Pade6stagg dispatch
//   ig_padepoisson.inc.cuh   projection / pressure    ig_hit_forcing.inc.cuh  HIT shell forcing
//   ig_sgs.inc.cuh           SGS term of the RHS      ig_ops_periodic.inc.cuh Ops_Periodic
// This file keeps the helpers and the igrid handle itself (state, RK stages, right-hand side, restart files).
#include <algorithm>
#include <cmath>
#include <cstdlib>
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

#include "ig_spectral.inc.cuh"

#include "ig_pade6stagg.inc.cuh"

#include "ig_padepoisson.inc.cuh"

#include "ig_hit_forcing.inc.cuh"

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
    // decomposed column communicator (not alias), periodic in z: z-pencil copies of cur[0..2].  project_and_prep takes the three
    // fields to the z-pencil ONCE, dealiases and projects them there and brings them back once; the interpolations and the
    // z-derivatives that follow read these copies instead of transposing the same arrays again (16 -> 6 transposes)
    double2 *zU = nullptr, *zV = nullptr, *zW = nullptr;
    bool zviews_valid = false;
    double2* zAcc[3] = {nullptr, nullptr, nullptr};   // z-pencil sums of the two z-operator terms of ru, rv (cell), rw (edge)
    // terms of the advection right-hand side, each in its own y-pencil array until ONE assembly pass per component sums them,
    // applies the -1/2 and adds the viscous term (replaces the reference's chain of in-place adds, igrid.F90:1572-1679, 1914-1941).
    // cell: 0 A_u 1 P_u 2 F_uu 3 B_u 4 A_v 5 P_v 6 F_vv 7 B_v 8 F_uv; edge: 0 A_w 1 P_w 2 B_w 3 F_uEw 4 F_vEw
    double2 *TC[9] = {}, *TE[5] = {};
    struct AsmTerm { const double2* p; int kind; };   // kind 0: + p, 1: + i k1 p, 2: + i k2 p
    struct AsmDesc { AsmTerm t[5]; int n; } asmd[3];
};

namespace {

int ig_alloc(pdo_igrid_s* g, void** p, size_t bytes) {
    PDO_CUDA(cudaMalloc(p, bytes ? bytes : 8));
    g->allocs.push_back(*p);
    return 0;
}
template <class T>
int ig_alloc_n(pdo_igrid_s* g, T** p, long long count) {
    // every spectral (complex) array can be the destination of a y<->z transpose: peer-writable memory
    if (sizeof(T) == sizeof(double2)) {
        if (int rc = comm_shared_malloc((void**)p, sizeof(T) * (size_t)count)) return rc;
        g->allocs.push_back(*p);
        return 0;
    }
    return ig_alloc(g, (void**)p, sizeof(T) * (size_t)count);
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

// fft(a*b (+ c*d)) on the cell (E = false) or edge grid: the product is formed on the first load of the x pass
int fft_mul2(pdo_igrid_s* g, bool edge, double2* out, const double* a, const double* b, const double* c, const double* d, cudaStream_t st) {
    RealPro rp;
    rp.p[0] = a; rp.p[1] = b; rp.p[2] = c; rp.p[3] = d;
    rp.mode = c ? 2 : 1;
    return fft3d_forward_xy_pro(edge ? g->spE->ft : g->spC->ft, rp, out, st);
}
// fft((a - b)*c (+ (d - e)*f))
int fft_muldiff(pdo_igrid_s* g, bool edge, double2* out, const double* a, const double* b, const double* c, const double* d, const double* e,
                const double* f, cudaStream_t st) {
    RealPro rp;
    rp.p[0] = a; rp.p[1] = b; rp.p[2] = c; rp.p[3] = d; rp.p[4] = e; rp.p[5] = f;
    rp.mode = d ? 4 : 3;
    return fft3d_forward_xy_pro(edge ? g->spE->ft : g->spC->ft, rp, out, st);
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
    if (g->zviews_valid) z = g->zW;
    else IG(zviewE(g, g->cur[2], g->zE[0], &z, st));
    double2* t = ztarget(g, g->whatC, g->zC[0]);
    ZOPB(pdo_pade6stagg_interpz_E2C, z, t, BC_W);
    IG(zcommitC(g, t, g->whatC, st));
    IG(ifftC(g, g->whatC, g->wC, st));
    for (int c = 0; c < 2; ++c) {
        double2* eh = c == 0 ? g->uEhat : g->vEhat;
        if (g->zviews_valid) z = c == 0 ? g->zU : g->zV;
        else IG(zviewC(g, g->cur[c], g->zC[0], &z, st));
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
    // i k f and the 1/(nx ny) of the inverse ride on the first load of the inverse y pass (fft2d.cu); on the cuFFT path they are the
    // one pointwise pass that stages the intent(in) input
    const double nf2 = 1.0 / ((double)g->prm.nx * (double)g->prm.ny);
    auto dC = [&](int which, const double2* fhat, double* out) -> int {
        if (!out) return 0;
        return fft3d_backward_yx_mul(C->ft, fhat, which, which == 1 ? C->k1y : C->k2, nf2, false, out, st);
    };
    auto dE = [&](int which, const double2* fhat, double* out) -> int {
        if (!out) return 0;
        return fft3d_backward_yx_mul(E->ft, fhat, which, which == 1 ? E->k1y : E->k2, nf2, false, out, st);
    };
    IG(dC(1, g->cur[0], g->gradC[0])); IG(dE(1, g->uEhat, g->gradE[0]));
    IG(dC(2, g->cur[0], g->gradC[1])); IG(dE(2, g->uEhat, g->gradE[1]));
    IG(dC(1, g->cur[1], g->gradC[3])); IG(dE(1, g->vEhat, g->gradE[3]));
    IG(dC(2, g->cur[1], g->gradC[4])); IG(dE(2, g->vEhat, g->gradE[4]));
    IG(dC(1, g->whatC, g->gradC[6])); IG(dE(1, g->cur[2], g->gradE[6]));
    IG(dC(2, g->whatC, g->gradC[7])); IG(dE(2, g->cur[2], g->gradE[7]));
    // dwdz (and its edge interpolant), d2wdz2
    const double2* wz = nullptr;
    if (g->zviews_valid) wz = g->zW;
    else IG(zviewE(g, g->cur[2], g->zE[0], &wz, st));
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
        if (g->zviews_valid) fz = c == 0 ? g->zU : g->zV;
        else IG(zviewC(g, g->cur[c], g->zC[0], &fz, st));
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

// dst += src on z-pencil arrays (complex viewed as doubles)
int zadd(double2* dst, const double2* src, long long n, cudaStream_t st) {
    double* d = (double*)dst;
    const double* q = (const double*)src;
    return launch_ew(2 * n, st, [=] __device__(long long i) { d[i] += q[i]; });
}

// ---- igrid.F90:1572-1679: the terms of (ru, rv, rw), left in TC / TE with the assembly recipes in g->asmd.
// Each component has TWO terms that come out of a staggered z-operator (A = interp(fft(d./dz w)), B = ddz(fft(. w))); they are summed
// first.  On a decomposed column communicator that sum is formed in the z-pencil (g->zAcc) and travels back as ONE array: three
// transposes less per substep; on one column the assembly pass forms the same sum (same operations, bit-identical results).
int ig_nonlinear_skew(pdo_igrid_s* g, cudaStream_t st) {
    double2 *fT1C = g->yC[0], *fT1E = g->yE[0];
    double2 **TC = g->TC, **TE = g->TE;
    double **GC = g->gradC, **GE = g->gradE;
    const bool zsum = !g->alias && g->zAcc[0];
    const double2* z = nullptr;
    double2* t = nullptr;
    // u_rhs = interp_E2C(fft(dudz w)) + fft(dudx u + dudy v); same for v
    for (int c = 0; c < 2; ++c) {
        IG(fft_mul2(g, false, TC[4 * c + 1], GC[3 * c + 0], g->u, GC[3 * c + 1], g->v, st));
        IG(fft_mul2(g, true, fT1E, GE[3 * c + 2], g->w, nullptr, nullptr, st));
        IG(zviewE(g, fT1E, g->zE[0], &z, st));
        t = zsum ? g->zAcc[c] : ztarget(g, TC[4 * c + 0], g->zC[0]);
        ZOPB(pdo_pade6stagg_interpz_E2C, z, t, c == 0 ? BC_WdUdz : BC_WdVdz);
        if (!zsum) IG(zcommitC(g, t, TC[4 * c + 0], st));
    }
    // w_rhs = interp_C2E(fft(dwdz wC)) + fft(dwdx uE + dwdy vE)
    IG(fft_mul2(g, true, TE[1], GE[6], g->uE, GE[7], g->vE, st));
    IG(fft_mul2(g, false, fT1C, GC[8], g->wC, nullptr, nullptr, st));
    IG(zviewC(g, fT1C, g->zC[0], &z, st));
    t = zsum ? g->zAcc[2] : ztarget(g, TE[0], g->zE[0]);
    ZOPB(pdo_pade6stagg_interpz_C2E, z, t, BC_WdWdz);
    if (!zsum) IG(zcommitE(g, t, TE[0], st));
    // conservative half: d(uu)/dx, d(vv)/dy, d(wC wC)/dz, d(uv)/dy & /dx, d(uE w)/dz & /dx, d(vE w)/dz & /dy
    IG(fft_mul2(g, false, TC[2], g->u, g->u, nullptr, nullptr, st));
    IG(fft_mul2(g, false, TC[6], g->v, g->v, nullptr, nullptr, st));
    IG(fft_mul2(g, false, fT1C, g->wC, g->wC, nullptr, nullptr, st));
    IG(zviewC(g, fT1C, g->zC[0], &z, st));
    t = ztarget(g, TE[2], g->zE[0]);
    ZOPB(pdo_pade6stagg_ddz_C2E, z, t, BC_WW);
    if (zsum) {
        IG(zadd(g->zAcc[2], t, g->nZE, st));
        IG(z2yE(g, g->zAcc[2], TE[0], st));
    } else IG(zcommitE(g, t, TE[2], st));
    IG(fft_mul2(g, false, TC[8], g->u, g->v, nullptr, nullptr, st));
    for (int c = 0; c < 2; ++c) {
        IG(fft_mul2(g, true, TE[3 + c], c == 0 ? g->uE : g->vE, g->w, nullptr, nullptr, st));
        IG(zviewE(g, TE[3 + c], g->zE[0], &z, st));
        t = ztarget(g, TC[4 * c + 3], g->zC[0]);
        ZOPB(pdo_pade6stagg_ddz_E2C, z, t, c == 0 ? BC_UW : BC_VW);
        if (zsum) {
            IG(zadd(g->zAcc[c], t, g->nZC, st));
            IG(z2yC(g, g->zAcc[c], TC[4 * c + 0], st));
        } else IG(zcommitC(g, t, TC[4 * c + 3], st));
    }
    if (zsum) {
        g->asmd[0] = {{{TC[0], 0}, {TC[1], 0}, {TC[2], 1}, {TC[8], 2}}, 4};
        g->asmd[1] = {{{TC[4], 0}, {TC[5], 0}, {TC[6], 2}, {TC[8], 1}}, 4};
        g->asmd[2] = {{{TE[0], 0}, {TE[1], 0}, {TE[3], 1}, {TE[4], 2}}, 4};
    } else {
        g->asmd[0] = {{{TC[0], 0}, {TC[3], 0}, {TC[1], 0}, {TC[2], 1}, {TC[8], 2}}, 5};
        g->asmd[1] = {{{TC[4], 0}, {TC[7], 0}, {TC[5], 0}, {TC[6], 2}, {TC[8], 1}}, 5};
        g->asmd[2] = {{{TE[0], 0}, {TE[2], 0}, {TE[1], 0}, {TE[3], 1}, {TE[4], 2}}, 5};
    }
    return 0;
}

// ---- igrid.F90:1527-1555: u x omega; the products with w live on the edge grid and come back through interpz_E2C.
// Gradient slots: C 1 dudy, 3 dvdx; E 2 dudz, 5 dvdz, 6 dwdx, 7 dwdy.
int ig_nonlinear_rot(pdo_igrid_s* g, cudaStream_t st) {
    double2* fT1E = g->yE[0];
    double2 **TC = g->TC, **TE = g->TE;
    double **GC = g->gradC, **GE = g->gradE;
    const double2* z = nullptr;
    double2* t = nullptr;
    for (int c = 0; c < 2; ++c) {
        // c = 0: (dvdx - dudy) v and (dwdx - dudz) w;   c = 1: (dudy - dvdx) u and (dwdy - dvdz) w
        if (c == 0) IG(fft_muldiff(g, false, TC[1], GC[3], GC[1], g->v, nullptr, nullptr, nullptr, st));
        else IG(fft_muldiff(g, false, TC[5], GC[1], GC[3], g->u, nullptr, nullptr, nullptr, st));
        if (c == 0) IG(fft_muldiff(g, true, fT1E, GE[6], GE[2], g->w, nullptr, nullptr, nullptr, st));
        else IG(fft_muldiff(g, true, fT1E, GE[7], GE[5], g->w, nullptr, nullptr, nullptr, st));
        IG(zviewE(g, fT1E, g->zE[0], &z, st));
        t = ztarget(g, TC[4 * c + 0], g->zC[0]);
        ZOP(pdo_pade6stagg_interpz_E2C, z, t);
        IG(zcommitC(g, t, TC[4 * c + 0], st));
    }
    // w_rhs = fft((dudz - dwdx) uE + (dvdz - dwdy) vE)
    IG(fft_muldiff(g, true, TE[1], GE[2], GE[6], g->uE, GE[5], GE[7], g->vE, st));
    g->asmd[0] = {{{TC[0], 0}, {TC[1], 0}}, 2};
    g->asmd[1] = {{{TC[4], 0}, {TC[5], 0}}, 2};
    g->asmd[2] = {{{TE[1], 0}}, 1};
    return 0;
}

// rhs = -half * (sum of the terms) (skew-symmetric form; +1 for the rotational one), then addViscousTerm (igrid.F90:1663-1665,
// 1914-1941): ONE pass per component over its terms
// With `upd` the stage update that consumes the right-hand side runs in the same pass (S_dst = sum_t c_t x_t, the term x_t == r
// taken from the register; same order of operations as the stand-alone pass); the right-hand side itself is stored only if keep_r.
struct StageUpd { double c[5]; const double2* x[5]; int n; int ridx; double2* dst; };
int ig_finish_rhs(pdo_igrid_s* g, double2* ru, double2* rv, double2* rw, cudaStream_t st, const StageUpd* upd = nullptr, bool keep_r = true) {
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
        const pdo_igrid_s::AsmDesc D = g->asmd[c];
        StageUpd U{};
        const bool fuse = upd != nullptr;
        if (fuse) U = upd[c];
        const bool keep = keep_r || !fuse;
        IG(launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
            const double ka = k1[(int)(i % n1)], kb = k2[(int)((i / n1) % n2)];
            double2 a = D.t[0].p[i];
#pragma unroll
            for (int t = 1; t < 5; ++t) {
                if (t < D.n) {
                    const double2 q = D.t[t].p[i];
                    if (D.t[t].kind == 0) { a.x += q.x; a.y += q.y; }
                    else {
                        const double kv = D.t[t].kind == 1 ? ka : kb;
                        a.x += -kv * q.y; a.y += kv * q.x;
                    }
                }
            }
            a.x = scale * a.x; a.y = scale * a.y;
            if (visc) {
                const double ksq = ka * ka + kb * kb;  // kabs_sq = k1**2 + k2**2 (spectral.F90:1093-1099)
                const double2 q = f[i], dd = d2[i];
                a.x += oneByRe * (-ksq * q.x + dd.x);
                a.y += oneByRe * (-ksq * q.y + dd.y);
            }
            if (keep) r[i] = a;
            if (fuse) {
                double2 x0 = U.ridx == 0 ? a : U.x[0][i];
                double2 acc = make_double2(U.c[0] * x0.x, U.c[0] * x0.y);
#pragma unroll
                for (int t = 1; t < 5; ++t) {
                    if (t < U.n) {
                        const double2 xt = U.ridx == t ? a : U.x[t][i];
                        acc.x += U.c[t] * xt.x; acc.y += U.c[t] * xt.y;
                    }
                }
                U.dst[i] = acc;
            }
        }));
    }
    return 0;
}

#include "ig_sgs.inc.cuh"

// PopulateRHS (igrid.F90:1793-1911) followed by the Runge-Kutta stage update that consumes it:
// S[dst_slot] = sum_t coef_t terms_t, with r among the terms.
int ig_stage_update(pdo_igrid_s* g, int dst_slot, int nterms, const double* coef, double2* const (*terms)[3], cudaStream_t st);
int ig_rhs_and_update(pdo_igrid_s* g, double2** r, int dst_slot, int nterms, const double* coef, double2* const (*terms)[3], bool keep_r,
                      cudaStream_t st) {
    if (g->prm.rotational_advection) IG(ig_nonlinear_rot(g, st));
    else IG(ig_nonlinear_skew(g, st));
    if (!g->sgs_on && !g->hit) {   // nothing else touches the right-hand side: assemble it and update the stage in one pass
        StageUpd U[3];
        for (int c = 0; c < 3; ++c) {
            U[c].n = nterms; U[c].ridx = -1; U[c].dst = g->S[dst_slot][c];
            for (int t = 0; t < nterms; ++t) {
                U[c].c[t] = coef[t]; U[c].x[t] = terms[t][c];
                if (terms[t][c] == r[c]) U[c].ridx = t;
            }
        }
        IG(ig_finish_rhs(g, r[0], r[1], r[2], st, U, keep_r));
        for (int c = 0; c < 3; ++c) g->cur[c] = g->S[dst_slot][c];
        g->new_timestep = false;
        return 0;
    }
    IG(ig_finish_rhs(g, r[0], r[1], r[2], st));
    if (g->sgs_on) IG(ig_sgs_rhs(g, r[0], r[1], r[2], st));   // Step 6 (igrid.F90:1866-1871)
    if (g->hit)   // Step 8 (igrid.F90:1907-1910)
        IG(hit_get_rhs_dev(g->hit, r[0], r[1], r[2], g->cur[0], g->cur[1], g->cur[2], g->new_timestep, st));
    g->new_timestep = false;
    return ig_stage_update(g, dst_slot, nterms, coef, terms, st);
}

// ---- igrid.F90:1961-1990
int ig_project_and_prep(pdo_igrid_s* g, bool already_projected, cudaStream_t st) {
    g->zviews_valid = false;
    const bool check_now = g->prm.t_divergence_check > 0 && g->step % g->prm.t_divergence_check == 0;
    static const bool kz_on = [] { const char* e = std::getenv("PDO_IG_KZ"); return !(e && e[0] == '0'); }();
    const bool kz = kz_on && g->poiss->symE2C && !g->prm.wall_bounded;   // dealias + project in kz-space (poiss_dealias_project_kz)
    if (g->zU && !already_projected && !check_now) {
        // z-resident form: to the z-pencil once, dealias (spectral.F90:343-363) and project (PadePoisson.F90:386-432) there, back once
        IG(y2zC(g, g->cur[0], g->zU, st));
        IG(y2zC(g, g->cur[1], g->zV, st));
        IG(y2zE(g, g->cur[2], g->zW, st));
        if (kz) IG(poiss_dealias_project_kz(g->poiss, g->zU, g->zV, g->zW, st));
        else {
            IG(spectral_dealias_zwork(g->spC, g->zU, st));
            IG(spectral_dealias_zwork(g->spC, g->zV, st));
            IG(spectral_dealias_edge(g->spC, g->zW, st));
            IG(poiss_projection_z(g->poiss, g->zU, g->zV, g->zW, st));
        }
        IG(z2yC(g, g->zU, g->cur[0], st));
        IG(z2yC(g, g->zV, g->cur[1], st));
        IG(z2yE(g, g->zW, g->cur[2], st));
        g->zviews_valid = true;
        IG(ifftC(g, g->cur[0], g->u, st));
        IG(ifftC(g, g->cur[1], g->v, st));
        IG(ifftE(g, g->cur[2], g->w, st));
        IG(ig_interp_primitive(g, st));
        const int rc = ig_compute_duidxj(g, st);
        g->zviews_valid = false;   // the next stage update overwrites cur
        return rc;
    }
    if (g->alias && kz && !already_projected && !check_now) {
        // one column: y- and z-pencils coincide, the same pass works on the stage arrays directly
        IG(poiss_dealias_project_kz(g->poiss, g->cur[0], g->cur[1], g->cur[2], st));
        IG(ifftC(g, g->cur[0], g->u, st));
        IG(ifftC(g, g->cur[1], g->v, st));
        IG(ifftE(g, g->cur[2], g->w, st));
        IG(ig_interp_primitive(g, st));
        return ig_compute_duidxj(g, st);
    }
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
    { double c[2] = {1.0, dt}; double2* const t[2][3] = {{g->S[0][0], g->S[0][1], g->S[0][2]}, {g->R[0], g->R[1], g->R[2]}};
      IG(ig_rhs_and_update(g, g->R, 1, 2, c, t, false, st)); }
    IG(ig_project_and_prep(g, false, st));
    { double c[3] = {3.0 / 4.0, 1.0 / 4.0, (1.0 / 4.0) * dt}; double2* const t[3][3] = {{g->S[0][0], g->S[0][1], g->S[0][2]}, {g->S[1][0], g->S[1][1], g->S[1][2]}, {g->R[0], g->R[1], g->R[2]}};
      IG(ig_rhs_and_update(g, g->R, 1, 3, c, t, false, st)); }
    IG(ig_project_and_prep(g, false, st));
    { double c[3] = {1.0 / 3.0, 2.0 / 3.0, (2.0 / 3.0) * dt}; double2* const t[3][3] = {{g->S[0][0], g->S[0][1], g->S[0][2]}, {g->S[1][0], g->S[1][1], g->S[1][2]}, {g->R[0], g->R[1], g->R[2]}};
      IG(ig_rhs_and_update(g, g->R, 0, 3, c, t, false, st)); }
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
    { double c[2] = {1.0, b01 * dt}; double2* const t[2][3] = {SL(0), RR};
      IG(ig_rhs_and_update(g, g->R, 1, 2, c, t, false, st)); }
    IG(ig_project_and_prep(g, false, st));
    { double c[3] = {a20, a21, b12 * dt}; double2* const t[3][3] = {SL(0), SL(1), RR};
      IG(ig_rhs_and_update(g, g->R, 2, 3, c, t, false, st)); }
    IG(ig_project_and_prep(g, false, st));
    { double c[3] = {a30, a32, b23 * dt}; double2* const t[3][3] = {SL(0), SL(2), RR};
      IG(ig_rhs_and_update(g, g->R, 3, 3, c, t, false, st)); }
    IG(ig_project_and_prep(g, false, st));
    { double c[3] = {a40, a43, b34 * dt}; double2* const t[3][3] = {SL(0), SL(3), RR};
      IG(ig_rhs_and_update(g, g->R, 0, 3, c, t, true, st)); }   // this right-hand side enters the last stage as well
    IG(ig_project_and_prep(g, false, st));
    { double c[5] = {a52, a53, b35 * dt, a54, b45 * dt}; double2* const t[5][3] = {SL(2), SL(3), RR, SL(0), RXX};
      IG(ig_rhs_and_update(g, g->RX, 0, 5, c, t, false, st)); }
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
    for (void* p : g->allocs) comm_shared_free(p);
    pdo_padepoisson_destroy(g->poiss);
    pdo_hit_forcing_destroy(g->hit);
    pdo_pade6stagg_destroy(g->ops);
    pdo_spectral_destroy(g->spE);
    pdo_spectral_destroy(g->spC);
    delete g;
    return 0;
}

// get_boundary_conditions_stencil (igrid.F90:5148-5204): (bottom, top) stencil codes per quantity; -1 odd, +1 even, 0 one-sided
static void ig_fill_bcs(int bc[12][2], int bot_wall, int top_wall) {
    const int walls[2] = {bot_wall, top_wall};
    for (int sd = 0; sd < 2; ++sd) {
        bc[BC_W][sd] = -1; bc[BC_WdWdz][sd] = -1; bc[BC_WW][sd] = +1; bc[BC_dWdz][sd] = 0;
        if (walls[sd] == 1) {        // no-slip: w = 0 and dwdz = 0, so w is extended evenly
            bc[BC_U][sd] = -1; bc[BC_V][sd] = -1; bc[BC_dUdz][sd] = 0; bc[BC_dVdz][sd] = 0;
            bc[BC_WdUdz][sd] = 0; bc[BC_WdVdz][sd] = 0; bc[BC_UW][sd] = +1; bc[BC_VW][sd] = +1;
            bc[BC_W][sd] = +1; bc[BC_WdWdz][sd] = -1; bc[BC_WW][sd] = +1; bc[BC_dWdz][sd] = -1;
        } else {                     // slip
            bc[BC_U][sd] = +1; bc[BC_V][sd] = +1; bc[BC_dUdz][sd] = -1; bc[BC_dVdz][sd] = -1;
            bc[BC_WdUdz][sd] = +1; bc[BC_WdVdz][sd] = +1; bc[BC_UW][sd] = -1; bc[BC_VW][sd] = -1;
        }
    }
}
/* test hook (host only): the stencil codes for (botWall, topWall), order w u v WdUdz WdVdz WdWdz WW UW VW dUdz dVdz dWdz, (bottom, top) each */
}  // extern "C"
namespace pdo { namespace hooks {
int igrid_bcs(int bot_wall, int top_wall, int* out24) {
    int bc[12][2] = {};
    ig_fill_bcs(bc, bot_wall, top_wall);
    for (int q = 0; q < 12; ++q) { out24[2 * q] = bc[q][0]; out24[2 * q + 1] = bc[q][1]; }
    return 0;
}
}}  // namespace pdo::hooks
extern "C" {

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
    if (p->wall_bounded) {
        ig_fill_bcs(g->bc, p->bot_wall, p->top_wall);
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
    if (!rc) rc = pdo_padepoisson_init3(&g->poiss, g->dx, g->dy, g->dz, g->spC, g->spE, g->ops, perz,
                                        p->wall_bounded && !p->no_stokes_pressure, p->Lz);   // :585-586 (ComputeStokesPressure, default .true.)
    if (rc) { pdo_igrid_destroy(g); return rc; }
    g->dC = fft3d_spec_decomp(g->spC->ft); g->dE = fft3d_spec_decomp(g->spE->ft);
    g->alias = (g->spC->p_col == 1);
    if (const char* e = std::getenv("PDO_IG_FORCE_TRANSPOSES")) {
        // test switch: run the decomposed-grid code path (explicit y <-> z transposes, z-resident projection) on one rank,
        // where every transpose is a device copy
        if (e[0] == '1') g->alias = false;
    }
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
    {
        const char* e = std::getenv("PDO_IG_ZRESIDENT");   // "0": the reference's pass structure (A/B measurements)
        if (!g->alias && !p->wall_bounded && !(e && e[0] == '0')) { AL(g->zU, g->nZC); AL(g->zV, g->nZC); AL(g->zW, g->nZE); }
        if (!g->alias && !p->rotational_advection && !(e && e[0] == '0')) { AL(g->zAcc[0], g->nZC); AL(g->zAcc[1], g->nZC); AL(g->zAcc[2], g->nZE); }
    }
    {   // right-hand-side terms (see TC / TE): the rotational form has two per horizontal component and one for w
        const bool rot = p->rotational_advection != 0;
        const bool needTC[9] = {true, true, !rot, !rot, true, true, !rot, !rot, !rot};
        const bool needTE[5] = {!rot, true, !rot, !rot, !rot};
        for (int i = 0; i < 9; ++i) if (needTC[i]) AL(g->TC[i], g->nYC);
        for (int i = 0; i < 5; ++i) if (needTE[i]) AL(g->TE[i], g->nYE);
    }
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
}  // extern "C"
namespace pdo { namespace hooks {
int sgs_point(int mid, double cmodel, double cx, double cy, double cz, const double* d9, double* nu, double* S6) {
    SgsConst c{mid, cmodel, cx, cy, cz};
    sgs_sij(d9, S6);
    *nu = c.cmodel * sgs_kernel_point(c, d9, S6);
    return 0;
}
}}  // namespace pdo::hooks
extern "C" {
/* the forcing object of a handle with useHITForcing (borrowed; e.g. for pdo_hit_forcing_set_wavenumbers before a time step) */
pdo_hit_forcing_t pdo_igrid_hit_forcing(pdo_igrid_t g) { return g ? g->hit : nullptr; }
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
/* A real as Fortran list / G editing writes it: "0.12346E+02", "1.5", and — when the exponent needs three digits — the form
   without the exponent letter, "0.12346+123" / "0.12346-123" (what pdo_io_format_g15_5 and gfortran emit); 'D' is accepted too. */
static bool parse_fortran_real(const char* tok, double* out) {
    std::string t(tok);
    for (auto& c : t) if (c == 'D' || c == 'd') c = 'E';
    if (t.find('E') == std::string::npos && t.find('e') == std::string::npos) {
        // a sign after the first character that does not follow an exponent letter starts the exponent
        for (size_t i = 1; i < t.size(); ++i)
            if ((t[i] == '+' || t[i] == '-') && (t[i - 1] == '.' || (t[i - 1] >= '0' && t[i - 1] <= '9'))) { t.insert(i, "E"); break; }
    }
    char* end = nullptr;
    const double v = std::strtod(t.c_str(), &end);
    if (end == t.c_str() || *end != 0) return false;
    *out = v;
    return true;
}
int pdo_igrid_read_restart(pdo_igrid_t g, const char* inputdir, int run_id, int tid) {
    if (!g) return fail(PDO_E_BADARG, "null handle");
    cudaStream_t st = nullptr;
    pdo_decomp_t dC = fft3d_phys_decomp(g->spC->ft), dE = fft3d_phys_decomp(g->spE->ft);
    // rbC / rbE scratch pencils receive the file contents; ig_set_fields copies them into u, v, w
    double *ru = g->rbC[0], *rv = g->rbC[1], *rw = g->rbE[0];
    // pdo_decomp_read_one is rank-local (a short file fails only on the ranks whose sub-box reaches past its end): every rank
    // goes through all three reads and the collectives below, and the outcome is agreed on before anyone returns
    int rc_local = pdo_decomp_read_one(dC, 1, ru, 1, restart_name(inputdir, run_id, "u", tid).c_str());
    if (!rc_local) rc_local = pdo_decomp_read_one(dC, 1, rv, 1, restart_name(inputdir, run_id, "v", tid).c_str());
    if (!rc_local) rc_local = pdo_decomp_read_one(dE, 1, rw, 1, restart_name(inputdir, run_id, "w", tid).c_str());
    const std::string read_err = rc_local ? std::string(pdo_last_error()) : std::string();
    double tsim = 0.0;
    int bad = 0;
    if (pdo_comm_rank() == 0) {
        FILE* f = std::fopen(restart_name(inputdir, run_id, "info", tid).c_str(), "r");
        char tok[64] = {0};
        if (!f || std::fscanf(f, "%63s", tok) != 1 || !parse_fortran_real(tok, &tsim)) bad = 1;
        if (f) std::fclose(f);
    }
    // mpi_bcast(tsim) from rank 0: the other ranks contribute zero to a sum
    double tsum = 0.0, badsum = 0.0, rcsum = 0.0;
    if (int b = pdo_p_sum(pdo_comm_rank() == 0 ? tsim : 0.0, &tsum)) return b;
    if (int b = pdo_p_sum((double)bad, &badsum)) return b;
    if (int b = pdo_p_sum(rc_local ? 1.0 : 0.0, &rcsum)) return b;
    if (rcsum != 0.0)
        return rc_local ? fail(rc_local, "%s", read_err.c_str())
                        : fail(PDO_E_BADARG, "restart files in '%s' could not be read on %d other rank(s)", inputdir ? inputdir : ".", (int)rcsum);
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


#include "ig_ops_periodic.inc.cuh"
