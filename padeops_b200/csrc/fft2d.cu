// fft2d.cu — hand-written FFT passes (FP64, powers of two) for the spectral path; see fft2d.cuh for what they replace.
//
// One algorithm for every pass: Stockham autosort, decimation in frequency, radix 8 with one leading radix 2 / 4 / 8 stage, eight
// complex points per thread per stage.  Stage i (radix R, s = product of the radices before it) takes butterfly b = q + s p:
//     reads  x[b + (N/R) r],  r < R                 (consecutive b -> consecutive addresses: coalesced / conflict-free)
//     writes y[q + s (R p + k)] = (sum_r x_r w_R^{rk}) w_N^{s p k}
// so the first stage reads straight from global memory, the last one writes natural order straight to global memory, and only the
// exchanges in between go through shared memory.
//   * strided pass (y or z): a tile = XT adjacent columns x the whole line, thread (column, t); shared layout [row][XT] — lanes run
//     along the columns, every access is a contiguous 16 B x XT segment, no bank conflicts whatever the stage's stride;
//   * contiguous pass (x): a real line of nx points is ONE complex transform of M = nx/2 points on z_p = x_2p + i x_2p+1 with the
//     usual split / merge (E_k, O_k) step after (r2c) or before (c2r) it; T = M/8 threads per line (a warp per 512-point line,
//     __syncwarp between stages), shared index padded p + (p >> 3), which makes every stage's writes conflict-free as well.
// HBM traffic per pass is the algorithmic one (read once, write once); the callers' pointwise work rides on the first load.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "fft2d.cuh"

namespace pdo {
namespace {

template <int LOG2N>
struct Plan {
    static constexpr int N = 1 << LOG2N;
    static constexpr int NS = (LOG2N + 2) / 3;                   // stages
    static constexpr int R0 = 1 << (LOG2N - 3 * (NS - 1));       // leading radix 2 / 4 / 8
    static constexpr int T = N / 8;                              // threads per transform
    __host__ __device__ static constexpr int radix(int i) { return i == 0 ? R0 : 8; }
    __host__ __device__ static constexpr int stride(int i) { return i == 0 ? 1 : (R0 << (3 * (i - 1))); }
};

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 w) { return make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
// a * (SGN i)
template <int SGN>
__device__ __forceinline__ double2 mul_si(double2 a) { return SGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }

// y_k = sum_r a_r w^{rk}, w = exp(SGN 2 pi i / R), in place, outputs in natural order
template <int R, int SGN>
__device__ __forceinline__ void bfly(double2* a) {
    if (R == 2) {
        const double2 t = cadd(a[0], a[1]);
        a[1] = csub(a[0], a[1]);
        a[0] = t;
    } else if (R == 4) {
        const double2 d0 = cadd(a[0], a[2]), d1 = csub(a[0], a[2]), d2 = cadd(a[1], a[3]), d3 = mul_si<SGN>(csub(a[1], a[3]));
        a[0] = cadd(d0, d2); a[1] = cadd(d1, d3); a[2] = csub(d0, d2); a[3] = csub(d1, d3);
    } else {
        constexpr double h = 0.70710678118654752440084436210484903928;
        const double2 b0 = cadd(a[0], a[4]), b1 = cadd(a[1], a[5]), b2 = cadd(a[2], a[6]), b3 = cadd(a[3], a[7]);
        const double2 c0 = csub(a[0], a[4]);
        const double2 e1 = csub(a[1], a[5]), e2 = csub(a[2], a[6]), e3 = csub(a[3], a[7]);
        // w8 = (1 + SGN i)/sqrt2, w8^2 = SGN i, w8^3 = (-1 + SGN i)/sqrt2
        const double2 c1 = make_double2(h * (e1.x - SGN * e1.y), h * (e1.y + SGN * e1.x));
        const double2 c2 = mul_si<SGN>(e2);
        const double2 c3 = make_double2(h * (-e3.x - SGN * e3.y), h * (-e3.y + SGN * e3.x));
        {   // even outputs: radix 4 on b
            const double2 d0 = cadd(b0, b2), d1 = csub(b0, b2), d2 = cadd(b1, b3), d3 = mul_si<SGN>(csub(b1, b3));
            a[0] = cadd(d0, d2); a[2] = cadd(d1, d3); a[4] = csub(d0, d2); a[6] = csub(d1, d3);
        }
        {   // odd outputs: radix 4 on c
            const double2 d0 = cadd(c0, c2), d1 = csub(c0, c2), d2 = cadd(c1, c3), d3 = mul_si<SGN>(csub(c1, c3));
            a[1] = cadd(d0, d2); a[3] = cadd(d1, d3); a[5] = csub(d0, d2); a[7] = csub(d1, d3);
        }
    }
}

// Stage I of the N-point transform, thread t of its T: load through ld(position), butterflies, twiddles.  W = exp(-2 pi i j / Nw)
// with Nw = wmul * N.
template <int LOG2N, int I, int SGN, class LD>
__device__ __forceinline__ void stage_compute(int t, double2 (&v)[8], const double2* __restrict__ W, int wmul, LD ld) {
    using P = Plan<LOG2N>;
    constexpr int N = P::N, R = P::radix(I), s = P::stride(I), T = P::T, NB = 8 / R;
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        const int b = t + T * u;
#pragma unroll
        for (int r = 0; r < R; ++r) v[u * R + r] = ld(b + (N / R) * r);
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) bfly<R, SGN>(&v[u * R]);
    if (I < P::NS - 1) {
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int b = t + T * u;
            const int e = (b / s) * s * wmul;   // s p
#pragma unroll
            for (int k = 1; k < R; ++k) {
                double2 w = __ldg(W + e * k);
                if (SGN > 0) w.y = -w.y;
                v[u * R + k] = cmul(v[u * R + k], w);
            }
        }
    }
}
template <int LOG2N, int I, class ST>
__device__ __forceinline__ void stage_store(int t, double2 (&v)[8], ST stf) {
    using P = Plan<LOG2N>;
    constexpr int R = P::radix(I), s = P::stride(I), T = P::T, NB = 8 / R;
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        const int b = t + T * u;
        const int q = b % s, p = b / s;
#pragma unroll
        for (int k = 0; k < R; ++k) stf(q + s * (R * p + k), v[u * R + k]);
    }
}

// All stages of one transform.  ld0 feeds the first stage, stl takes the last stage's output, exchanges go through (lds, sts);
// sync() separates a stage's shared reads from its shared writes and the writes from the next stage's reads.  FIRST_SM / LAST_SM say
// that ld0 / stl are shared-memory accessors themselves, i.e. that the first / last stage needs the separating sync as well.
template <int LOG2N, int SGN, bool FIRST_SM, bool LAST_SM, class LD0, class STL, class LDS, class STS, class SYNC>
__device__ __forceinline__ void transform(int t, const double2* __restrict__ W, int wmul, LD0 ld0, STL stl, LDS lds, STS sts, SYNC sync) {
    using P = Plan<LOG2N>;
    constexpr int NS = P::NS;
    double2 v[8];
    if constexpr (NS == 1) {
        stage_compute<LOG2N, 0, SGN>(t, v, W, wmul, ld0);
        if (FIRST_SM && LAST_SM) sync();
        stage_store<LOG2N, 0>(t, v, stl);
        return;
    }
    stage_compute<LOG2N, 0, SGN>(t, v, W, wmul, ld0);
    if (FIRST_SM) sync();
    stage_store<LOG2N, 0>(t, v, sts);
    sync();
    if constexpr (NS >= 3) {
        stage_compute<LOG2N, 1, SGN>(t, v, W, wmul, lds);
        sync();
        stage_store<LOG2N, 1>(t, v, sts);
        sync();
    }
    if constexpr (NS >= 4) {
        stage_compute<LOG2N, 2, SGN>(t, v, W, wmul, lds);
        sync();
        stage_store<LOG2N, 2>(t, v, sts);
        sync();
    }
    stage_compute<LOG2N, NS - 1, SGN>(t, v, W, wmul, lds);
    if (LAST_SM) sync();
    stage_store<LOG2N, NS - 1>(t, v, stl);
}

// ---------------------------------------------------------------------------------------------------------------------------
// strided pass
// ---------------------------------------------------------------------------------------------------------------------------
template <int LOG2N> struct ColsCfg {
    static constexpr int XT = LOG2N <= 5 ? 32 : (LOG2N <= 7 ? 16 : 8);
    static constexpr int THREADS = XT * (1 << LOG2N) / 8;
    static constexpr int MINB = LOG2N <= 8 ? 4 : (LOG2N == 9 ? 2 : 1);
    static constexpr size_t SMEM = sizeof(double2) * XT * (size_t)(1 << LOG2N);
};

template <int LOG2N, int SGN>
__global__ void __launch_bounds__(ColsCfg<LOG2N>::THREADS, ColsCfg<LOG2N>::MINB)
fft_cols_kernel(const double2* __restrict__ in, double2* __restrict__ out, long long ncols, long long nplanes, long long row_stride,
                long long plane_stride, FftPro pro, const double2* __restrict__ W) {
    constexpr int XT = ColsCfg<LOG2N>::XT;
    extern __shared__ double2 sm[];
    const int c = threadIdx.x % XT, t = threadIdx.x / XT;
    const long long cblocks = (ncols + XT - 1) / XT;
    const long long ntiles = cblocks * nplanes;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long pl = tile / cblocks, cb = tile - pl * cblocks;
        const long long col = cb * XT + c;
        const bool ok = col < ncols;
        const double2* src = in + pl * plane_stride + col;
        double2* dst = out + pl * plane_stride + col;
        double colf = pro.scale;
        bool zero = false;
        if (pro.active && ok) {
            const long long c1 = col / pro.n1;
            const int c0 = (int)(col - c1 * pro.n1);
            if (pro.A) colf *= __ldg(pro.A + c0);
            if (pro.B) colf *= __ldg(pro.B + c1);
            zero = (c0 == pro.nyq);
        }
        auto ldg = [&](int row) -> double2 {
            if (!ok) return make_double2(0.0, 0.0);
            double2 x = src[(long long)row * row_stride];
            if (pro.active) {
                double m = colf;
                if (pro.C) m *= __ldg(pro.C + row);
                x.x *= m; x.y *= m;
                if (pro.times_i) x = make_double2(-x.y, x.x);
                if (zero) x = make_double2(0.0, 0.0);
            }
            return x;
        };
        auto stg = [&](int row, double2 x) { if (ok) dst[(long long)row * row_stride] = x; };
        auto lds = [&](int pos) -> double2 { return sm[pos * XT + c]; };
        auto sts = [&](int pos, double2 x) { sm[pos * XT + c] = x; };
        auto sync = [] { __syncthreads(); };
        transform<LOG2N, SGN, false, false>(t, W, 1, ldg, stg, lds, sts, sync);
        if (Plan<LOG2N>::NS > 1) __syncthreads();   // the last stage's shared reads before the next tile's first writes
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// contiguous pass: real line of n = 2M points <-> M + 1 complex modes
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int padx(int p) { return p + (p >> 3); }

template <int LOG2M> struct LineCfg {
    static constexpr int M = 1 << LOG2M;
    static constexpr int T = M / 8;
    static constexpr int THREADS = 256;
    static constexpr int LPB = THREADS / T;          // lines per block and iteration
    static constexpr int LS = M + M / 8;             // padded shared line
    static constexpr bool WARP = T <= 32;            // a line's threads sit in one warp
};

__device__ __forceinline__ double2 ld2(const double* p, long long i) { return *reinterpret_cast<const double2*>(p + i); }

__device__ __forceinline__ double2 real_pro_load(const RealPro& pro, long long i) {
    const double2 a = ld2(pro.p[0], i);
    switch (pro.mode) {
        default: return a;
        case 1: { const double2 b = ld2(pro.p[1], i); return make_double2(a.x * b.x, a.y * b.y); }
        case 2: {
            const double2 b = ld2(pro.p[1], i), c = ld2(pro.p[2], i), d = ld2(pro.p[3], i);
            return make_double2(a.x * b.x + c.x * d.x, a.y * b.y + c.y * d.y);
        }
        case 3: { const double2 b = ld2(pro.p[1], i), c = ld2(pro.p[2], i); return make_double2((a.x - b.x) * c.x, (a.y - b.y) * c.y); }
        case 4: {
            const double2 b = ld2(pro.p[1], i), c = ld2(pro.p[2], i), d = ld2(pro.p[3], i), e = ld2(pro.p[4], i), f = ld2(pro.p[5], i);
            const double t1x = (a.x - b.x) * c.x, t1y = (a.y - b.y) * c.y;
            const double t2x = (d.x - e.x) * f.x, t2y = (d.y - e.y) * f.y;
            return make_double2(t1x + t2x, t1y + t2y);
        }
    }
}

template <int LOG2M>
__global__ void __launch_bounds__(256, 4) fft_r2c_kernel(RealPro pro, double2* __restrict__ out, long long nlines, const double2* __restrict__ W) {
    using C = LineCfg<LOG2M>;
    constexpr int M = C::M, T = C::T, LPB = C::LPB, LS = C::LS;
    __shared__ double2 sm[LPB * LS];
    const int l = threadIdx.x / T, t = threadIdx.x % T;
    double2* z = sm + l * LS;
    const long long ngroups = (nlines + LPB - 1) / LPB;
    for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const long long line = g * LPB + l;
        const bool ok = line < nlines;
        const long long ibase = line * (2LL * M);
        double2* dst = out + line * (M + 1);
        auto ldg = [&](int p) -> double2 { return ok ? real_pro_load(pro, ibase + 2 * p) : make_double2(0.0, 0.0); };
        auto lds = [&](int p) -> double2 { return z[padx(p)]; };
        auto sts = [&](int p, double2 x) { z[padx(p)] = x; };
        auto sync = [] { if (C::WARP) __syncwarp(); else __syncthreads(); };
        transform<LOG2M, -1, false, true>(t, W, 2, ldg, sts, lds, sts, sync);
        sync();
        // merge: X_k = E_k + w^k O_k, X_{M-k} = conj(E_k - w^k O_k), E = (Z_k + conj Z_{M-k})/2, O = -i (Z_k - conj Z_{M-k})/2
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = t + T * u;
            if (k == 0) {
                const double2 z0 = z[0];
                if (ok) { dst[0] = make_double2(z0.x + z0.y, 0.0); dst[M] = make_double2(z0.x - z0.y, 0.0); }
            } else {
                const double2 A = z[padx(k)], B = z[padx(M - k)];
                const double2 E = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y - B.y));
                const double2 O = make_double2(0.5 * (A.y + B.y), -0.5 * (A.x - B.x));
                const double2 G = cmul(O, __ldg(W + k));
                if (ok) { dst[k] = cadd(E, G); dst[M - k] = make_double2(E.x - G.x, -(E.y - G.y)); }
            }
        }
        if (t == 0 && ok) { const double2 A = z[padx(M / 2)]; dst[M / 2] = make_double2(A.x, -A.y); }
        sync();   // merge reads before the next group's first writes
    }
}

template <int LOG2M>
__global__ void __launch_bounds__(256, 4) fft_c2r_kernel(const double2* __restrict__ in, double* __restrict__ out, long long nlines, const double2* __restrict__ W) {
    using C = LineCfg<LOG2M>;
    constexpr int M = C::M, T = C::T, LPB = C::LPB, LS = C::LS;
    __shared__ double2 sm[LPB * LS];
    const int l = threadIdx.x / T, t = threadIdx.x % T;
    double2* z = sm + l * LS;
    const long long ngroups = (nlines + LPB - 1) / LPB;
    for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const long long line = g * LPB + l;
        const bool ok = line < nlines;
        const double2* src = in + line * (M + 1);
        double* dst = out + line * (2LL * M);
        auto sync = [] { if (C::WARP) __syncwarp(); else __syncthreads(); };
        // split: Z_k = (X_k + conj X_{M-k}) + i conj(w)^k (X_k - conj X_{M-k}),  Z_{M-k} = conj(X_k + conj X_{M-k}) + i conj(G)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = t + T * u;   // 4 T = M / 2 values
            if (k == 0) {
                const double2 x0 = ok ? src[0] : make_double2(0.0, 0.0), xm = ok ? src[M] : make_double2(0.0, 0.0);
                z[0] = make_double2(x0.x + xm.x, x0.x - xm.x);
            } else {
                const double2 A = ok ? src[k] : make_double2(0.0, 0.0), B = ok ? src[M - k] : make_double2(0.0, 0.0);
                const double2 Ze = make_double2(A.x + B.x, A.y - B.y), D = make_double2(A.x - B.x, A.y + B.y);
                double2 w = __ldg(W + k);
                w.y = -w.y;
                const double2 G = cmul(D, w);
                z[padx(k)] = make_double2(Ze.x - G.y, Ze.y + G.x);
                z[padx(M - k)] = make_double2(Ze.x + G.y, -Ze.y + G.x);
            }
        }
        if (t == 0) { const double2 A = ok ? src[M / 2] : make_double2(0.0, 0.0); z[padx(M / 2)] = make_double2(2.0 * A.x, -2.0 * A.y); }
        sync();
        auto lds = [&](int p) -> double2 { return z[padx(p)]; };
        auto sts = [&](int p, double2 x) { z[padx(p)] = x; };
        auto stg = [&](int p, double2 x) { if (ok) *reinterpret_cast<double2*>(dst + 2 * p) = x; };
        transform<LOG2M, +1, true, false>(t, W, 2, lds, stg, lds, sts, sync);
        sync();   // last stage's shared reads before the next group's split writes
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------
std::mutex g_tw_mutex;
std::map<int, double2*> g_tw;

// exp(-2 pi i j / n), j < n, from long double sines of the first octant
int twiddles(int n, const double2** out) {
    std::lock_guard<std::mutex> lk(g_tw_mutex);
    auto it = g_tw.find(n);
    if (it != g_tw.end()) { *out = it->second; return 0; }
    std::vector<double2> h((size_t)n);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int j = 0; j < n; ++j) {
        // reduce to the first octant so that symmetric entries are exact mirror images
        int jj = j % n;
        const int q = (int)((8LL * jj) / n);      // octant
        long double c, s;
        auto cs = [&](long long num, long double& cc, long double& ss) {   // angle = 2 pi num / n, 0 <= angle <= pi/4
            const long double a = two_pi * (long double)num / (long double)n;
            cc = cosl(a); ss = sinl(a);
        };
        switch (q) {
            case 0: cs(jj, c, s); break;
            case 1: { long double a, b; cs(n / 4 - jj, a, b); c = b; s = a; break; }
            case 2: { long double a, b; cs(jj - n / 4, a, b); c = -b; s = a; break; }
            case 3: { long double a, b; cs(n / 2 - jj, a, b); c = -a; s = b; break; }
            case 4: { long double a, b; cs(jj - n / 2, a, b); c = -a; s = -b; break; }
            case 5: { long double a, b; cs(3 * n / 4 - jj, a, b); c = -b; s = -a; break; }
            case 6: { long double a, b; cs(jj - 3 * n / 4, a, b); c = b; s = -a; break; }
            default: { long double a, b; cs(n - jj, a, b); c = a; s = -b; break; }
        }
        h[(size_t)j] = make_double2((double)c, (double)(-s));
    }
    double2* d = nullptr;
    PDO_CUDA(cudaMalloc(&d, sizeof(double2) * (size_t)n));
    PDO_CUDA(cudaMemcpy(d, h.data(), sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice));
    g_tw[n] = d;
    *out = d;
    return 0;
}

int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
bool pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int LOG2N, int SGN>
int launch_cols(long long ncols, long long nplanes, long long row_stride, long long plane_stride, const double2* in, double2* out,
                const FftPro& pro, const double2* W, cudaStream_t st) {
    using Cc = ColsCfg<LOG2N>;
    static bool attr_done = false;
    if (!attr_done) {
        PDO_CUDA(cudaFuncSetAttribute(fft_cols_kernel<LOG2N, SGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cc::SMEM));
        attr_done = true;
    }
    const long long ntiles = ((ncols + Cc::XT - 1) / Cc::XT) * nplanes;
    long long grid = (long long)sm_count() * Cc::MINB;
    if (grid > ntiles) grid = ntiles;
    fft_cols_kernel<LOG2N, SGN><<<(unsigned)grid, Cc::THREADS, Cc::SMEM, st>>>(in, out, ncols, nplanes, row_stride, plane_stride, pro, W);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return 0;
}

template <int LOG2M>
int launch_r2c(long long nlines, const RealPro& pro, double2* out, const double2* W, cudaStream_t st) {
    using C = LineCfg<LOG2M>;
    const long long ngroups = (nlines + C::LPB - 1) / C::LPB;
    long long grid = (long long)sm_count() * 4;
    if (grid > ngroups) grid = ngroups;
    fft_r2c_kernel<LOG2M><<<(unsigned)grid, 256, 0, st>>>(pro, out, nlines, W);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return 0;
}
template <int LOG2M>
int launch_c2r(long long nlines, const double2* in, double* out, const double2* W, cudaStream_t st) {
    using C = LineCfg<LOG2M>;
    const long long ngroups = (nlines + C::LPB - 1) / C::LPB;
    long long grid = (long long)sm_count() * 4;
    if (grid > ngroups) grid = ngroups;
    fft_c2r_kernel<LOG2M><<<(unsigned)grid, 256, 0, st>>>(in, out, nlines, W);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return 0;
}

}  // namespace

bool fft2d_x_ok(int nx) { return pow2(nx) && nx >= 16 && nx <= 2048; }
bool fft2d_cols_ok(int n) { return pow2(n) && n >= 16 && n <= 1024; }
bool fft2d_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = std::getenv("PDO_FFT");
        on = (e && std::strcmp(e, "cufft") == 0) ? 0 : 1;
    }
    return on == 1;
}

int fft2d_cols(int n, long long ncols, long long nplanes, long long row_stride, long long plane_stride, const double2* in, double2* out,
               int dir, const FftPro& pro, cudaStream_t st) {
    if (!fft2d_cols_ok(n)) return fail(PDO_E_BADARG, "fft2d_cols: n = %d is not covered", n);
    if (ncols <= 0 || nplanes <= 0) return 0;
    const double2* W = nullptr;
    if (int rc = twiddles(n, &W)) return rc;
#define PDO_COLS_CASE(L)                                                                                                         \
    case L:                                                                                                                      \
        return dir < 0 ? launch_cols<L, -1>(ncols, nplanes, row_stride, plane_stride, in, out, pro, W, st)                       \
                       : launch_cols<L, +1>(ncols, nplanes, row_stride, plane_stride, in, out, pro, W, st);
    switch (ilog2(n)) {
        PDO_COLS_CASE(4) PDO_COLS_CASE(5) PDO_COLS_CASE(6) PDO_COLS_CASE(7) PDO_COLS_CASE(8) PDO_COLS_CASE(9) PDO_COLS_CASE(10)
    }
#undef PDO_COLS_CASE
    return fail(PDO_E_BADARG, "fft2d_cols: n = %d", n);
}

int fft2d_r2c_lines(int nx, long long nlines, const RealPro& pro, double2* out, cudaStream_t st) {
    if (!fft2d_x_ok(nx)) return fail(PDO_E_BADARG, "fft2d_r2c: nx = %d is not covered", nx);
    if (nlines <= 0) return 0;
    const double2* W = nullptr;
    if (int rc = twiddles(nx, &W)) return rc;
    switch (ilog2(nx / 2)) {
        case 3: return launch_r2c<3>(nlines, pro, out, W, st);
        case 4: return launch_r2c<4>(nlines, pro, out, W, st);
        case 5: return launch_r2c<5>(nlines, pro, out, W, st);
        case 6: return launch_r2c<6>(nlines, pro, out, W, st);
        case 7: return launch_r2c<7>(nlines, pro, out, W, st);
        case 8: return launch_r2c<8>(nlines, pro, out, W, st);
        case 9: return launch_r2c<9>(nlines, pro, out, W, st);
        case 10: return launch_r2c<10>(nlines, pro, out, W, st);
    }
    return fail(PDO_E_BADARG, "fft2d_r2c: nx = %d", nx);
}

int fft2d_c2r_lines(int nx, long long nlines, const double2* in, double* out, cudaStream_t st) {
    if (!fft2d_x_ok(nx)) return fail(PDO_E_BADARG, "fft2d_c2r: nx = %d is not covered", nx);
    if (nlines <= 0) return 0;
    const double2* W = nullptr;
    if (int rc = twiddles(nx, &W)) return rc;
    switch (ilog2(nx / 2)) {
        case 3: return launch_c2r<3>(nlines, in, out, W, st);
        case 4: return launch_c2r<4>(nlines, in, out, W, st);
        case 5: return launch_c2r<5>(nlines, in, out, W, st);
        case 6: return launch_c2r<6>(nlines, in, out, W, st);
        case 7: return launch_c2r<7>(nlines, in, out, W, st);
        case 8: return launch_c2r<8>(nlines, in, out, W, st);
        case 9: return launch_c2r<9>(nlines, in, out, W, st);
        case 10: return launch_c2r<10>(nlines, in, out, W, st);
    }
    return fail(PDO_E_BADARG, "fft2d_c2r: nx = %d", nx);
}

}  // namespace pdo
