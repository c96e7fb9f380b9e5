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

template <int LOG2N, int LOGE = 3>
struct Plan {
    static constexpr int N = 1 << LOG2N;
    static constexpr int E = 1 << LOGE;                              // complex points per thread per stage (8 or 16)
    static constexpr int NS = (LOG2N + LOGE - 1) / LOGE;             // stages
    static constexpr int R0 = 1 << (LOG2N - LOGE * (NS - 1));        // leading radix
    static constexpr int T = N / E;                                  // threads per transform
    __host__ __device__ static constexpr int radix(int i) { return i == 0 ? R0 : E; }
    __host__ __device__ static constexpr int stride(int i) { return i == 0 ? 1 : (R0 << (LOGE * (i - 1))); }
    // stage twiddles, one table per stage laid out [k - 1][p] (p = b / s, N / (R s) distinct values): lanes of a warp read
    // consecutive (or equal) entries instead of gathering w^{s p k} from a flat table
    __host__ __device__ static constexpr int npts(int i) { return N / (radix(i) * stride(i)); }
    __host__ __device__ static constexpr int toff(int i) { return i == 0 ? 0 : toff(i - 1) + (radix(i - 1) - 1) * npts(i - 1); }
};

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 w) { return make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
// a * (SGN i)
template <int SGN>
__device__ __forceinline__ double2 mul_si(double2 a) { return SGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }

// y_k = sum_r a_r w^{rk}, w = exp(SGN 2 pi i / R), in place, outputs in natural order
template <int SGN>
__device__ __forceinline__ void bfly4(double2& a0, double2& a1, double2& a2, double2& a3) {
    const double2 d0 = cadd(a0, a2), d1 = csub(a0, a2), d2 = cadd(a1, a3), d3 = mul_si<SGN>(csub(a1, a3));
    a0 = cadd(d0, d2); a1 = cadd(d1, d3); a2 = csub(d0, d2); a3 = csub(d1, d3);
}
// x * (c + SGN i s)
template <int SGN>
__device__ __forceinline__ double2 mulc(double2 x, double c, double sn) {
    return make_double2(x.x * c - SGN * (x.y * sn), x.y * c + SGN * (x.x * sn));
}
template <int R, int SGN>
__device__ __forceinline__ void bfly(double2* a) {
    if (R == 16) {
        // r = r0 + 4 r1, k = 4 k0 + k1: radix 4 over r1, twiddle w16^{r0 k1}, radix 4 over r0
        constexpr double h = 0.70710678118654752440084436210484903928;
        constexpr double c1 = 0.92387953251128675612818318939678828682, s1 = 0.38268343236508977172845998403039886676;
#pragma unroll
        for (int r0 = 0; r0 < 4; ++r0) bfly4<SGN>(a[r0], a[r0 + 4], a[r0 + 8], a[r0 + 12]);   // a[r0 + 4 k1] = B[r0][k1]
        a[1 + 4] = mulc<SGN>(a[1 + 4], c1, s1);       // w16^1
        a[1 + 8] = mulc<SGN>(a[1 + 8], h, h);         // w16^2
        a[1 + 12] = mulc<SGN>(a[1 + 12], s1, c1);     // w16^3
        a[2 + 4] = mulc<SGN>(a[2 + 4], h, h);         // w16^2
        a[2 + 8] = mul_si<SGN>(a[2 + 8]);             // w16^4
        a[2 + 12] = mulc<SGN>(a[2 + 12], -h, h);      // w16^6
        a[3 + 4] = mulc<SGN>(a[3 + 4], s1, c1);       // w16^3
        a[3 + 8] = mulc<SGN>(a[3 + 8], -h, h);        // w16^6
        a[3 + 12] = mulc<SGN>(a[3 + 12], -c1, -s1);   // w16^9
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) bfly4<SGN>(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);   // -> y[4 k0 + k1] at a[4 k1 + k0]
        // natural order: y[4 k0 + k1] currently sits at a[4 k1 + k0] — a 4 x 4 transpose in registers
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = i + 1; j < 4; ++j) { const double2 t = a[4 * i + j]; a[4 * i + j] = a[4 * j + i]; a[4 * j + i] = t; }
    } else if (R == 2) {
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

// Stage I of the N-point transform, thread t of its T: raw loads through ld(position) — all of them issued before anything
// depends on one — then fix(position, value) (the caller's pointwise work on the first stage), butterflies, twiddles from the
// stage tables TW.
template <int LOG2N, int I, int SGN, int LOGE, class LD, class FX>
__device__ __forceinline__ void stage_compute(int t, double2 (&v)[1 << LOGE], const double2* __restrict__ TW, LD ld, FX fix) {
    using P = Plan<LOG2N, LOGE>;
    constexpr int N = P::N, R = P::radix(I), s = P::stride(I), T = P::T, NB = P::E / R;
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        const int b = t + T * u;
#pragma unroll
        for (int r = 0; r < R; ++r) v[u * R + r] = ld(b + (N / R) * r);
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        const int b = t + T * u;
#pragma unroll
        for (int r = 0; r < R; ++r) v[u * R + r] = fix(b + (N / R) * r, v[u * R + r]);
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) bfly<R, SGN>(&v[u * R]);
    if (I < P::NS - 1) {
        constexpr int NP = P::npts(I);
        const double2* tw = TW + P::toff(I);
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int p = (t + T * u) / s;
#pragma unroll
            for (int k = 1; k < R; ++k) {
                double2 w = __ldg(tw + (k - 1) * NP + p);
                if (SGN > 0) w.y = -w.y;
                v[u * R + k] = cmul(v[u * R + k], w);
            }
        }
    }
}
struct NoFix { __device__ __forceinline__ double2 operator()(int, double2 x) const { return x; } };
template <int LOG2N, int I, int LOGE, class ST>
__device__ __forceinline__ void stage_store(int t, double2 (&v)[1 << LOGE], ST stf) {
    using P = Plan<LOG2N, LOGE>;
    constexpr int R = P::radix(I), s = P::stride(I), T = P::T, NB = P::E / R;
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        const int b = t + T * u;
        const int q = b % s, p = b / s;
#pragma unroll
        for (int k = 0; k < R; ++k) stf(q + s * (R * p + k), v[u * R + k]);
    }
}

// All stages of one transform.  ld0 (+ fix0) feeds the first stage, stl takes the last stage's output, exchanges go through
// (lds, sts); sync() separates a stage's shared reads from its shared writes and the writes from the next stage's reads.
// FIRST_SM / LAST_SM say that ld0 / stl are shared-memory accessors themselves, i.e. that the first / last stage needs the
// separating sync as well.
template <int LOG2N, int SGN, bool FIRST_SM, bool LAST_SM, int LOGE = 3, class LD0, class FX0, class STL, class LDS, class STS, class SYNC>
__device__ __forceinline__ void transform(int t, const double2* __restrict__ TW, LD0 ld0, FX0 fix0, STL stl, LDS lds, STS sts, SYNC sync) {
    using P = Plan<LOG2N, LOGE>;
    constexpr int NS = P::NS;
    double2 v[P::E];
    if constexpr (NS == 1) {
        stage_compute<LOG2N, 0, SGN, LOGE>(t, v, TW, ld0, fix0);
        if (FIRST_SM && LAST_SM) sync();
        stage_store<LOG2N, 0, LOGE>(t, v, stl);
        return;
    }
    stage_compute<LOG2N, 0, SGN, LOGE>(t, v, TW, ld0, fix0);
    if (FIRST_SM) sync();
    stage_store<LOG2N, 0, LOGE>(t, v, sts);
    sync();
    if constexpr (NS >= 3) {
        stage_compute<LOG2N, 1, SGN, LOGE>(t, v, TW, lds, NoFix());
        sync();
        stage_store<LOG2N, 1, LOGE>(t, v, sts);
        sync();
    }
    if constexpr (NS >= 4) {
        stage_compute<LOG2N, 2, SGN, LOGE>(t, v, TW, lds, NoFix());
        sync();
        stage_store<LOG2N, 2, LOGE>(t, v, sts);
        sync();
    }
    stage_compute<LOG2N, NS - 1, SGN, LOGE>(t, v, TW, lds, NoFix());
    if (LAST_SM) sync();
    stage_store<LOG2N, NS - 1, LOGE>(t, v, stl);
}

// the first-load work of a strided pass (FftPro) for one thread's column
struct ColFix {
    const FftPro& pro;
    double colf = 1.0, b1 = 0.0, ku = 0.0, kv = 0.0;
    bool zero = false;
    __device__ __forceinline__ explicit ColFix(const FftPro& p) : pro(p) {}
    __device__ __forceinline__ void setup(int col, bool ok) {
        colf = pro.scale;
        zero = false;
        if (pro.active && ok) {
            const int c1 = col / pro.n1, c0 = col - c1 * pro.n1;
            if (pro.poisson) { colf = __ldg(pro.A + c0); b1 = __ldg(pro.B + c1); }
            else {
                if (pro.A) colf *= __ldg(pro.A + c0);
                if (pro.B) colf *= __ldg(pro.B + c1);
            }
            if (pro.U) { ku = __ldg(pro.KU + c0); kv = __ldg(pro.KV + c1); }
            zero = (c0 == pro.nyq);
        }
    }
    // idx: element index of (col, row, plane) in the input's layout
    __device__ __forceinline__ double2 operator()(int row, double2 x, long long idx) const {
        if (pro.active) {
            if (pro.U) {
                const double2 uu = pro.U[idx], vv = pro.V[idx];
                const double re = ku * uu.x + kv * vv.x, im = ku * uu.y + kv * vv.y;
                x.x += -im; x.y += re;
            }
            if (pro.poisson) {
                const double kradsq = colf + b1 + __ldg(pro.C + row);
                const double m = (kradsq <= 1.e-14) ? 0.0 : -(1.0 / kradsq) * pro.scale;
                x.x *= m; x.y *= m;
            } else {
                double m = colf;
                if (pro.C) m *= __ldg(pro.C + row);
                x.x *= m; x.y *= m;
            }
            if (pro.times_i) x = make_double2(-x.y, x.x);
            if (zero) x = make_double2(0.0, 0.0);
        }
        return x;
    }
};

template <int LOG2N> struct ColsCfg {
    static constexpr int XT = LOG2N <= 5 ? 32 : (LOG2N <= 7 ? 16 : 8);
    static constexpr int THREADS = XT * (1 << LOG2N) / 8;
    static constexpr int MINB = LOG2N <= 8 ? 4 : (LOG2N == 9 ? 2 : 1);
    static constexpr size_t SMEM = sizeof(double2) * XT * (size_t)(1 << LOG2N);
};

template <int LOG2N, int SGN>
__global__ void __launch_bounds__(ColsCfg<LOG2N>::THREADS, ColsCfg<LOG2N>::MINB)
fft_cols_kernel(const double2* __restrict__ in, double2* __restrict__ out, int ncols, int nplanes, long long row_stride,
                long long plane_stride, FftPro pro, const double2* __restrict__ TW) {
    constexpr int XT = ColsCfg<LOG2N>::XT;
    extern __shared__ double2 sm[];
    const int c = threadIdx.x % XT, t = threadIdx.x / XT;
    const int cblocks = (ncols + XT - 1) / XT;
    const int ntiles = cblocks * nplanes;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int pl = tile / cblocks, cb = tile - pl * cblocks;
        const int col = cb * XT + c;
        const bool ok = col < ncols;
        const double2* src = in + pl * plane_stride + col;
        double2* dst = out + pl * plane_stride + col;
        ColFix cf(pro);
        cf.setup(col, ok);
        const long long ebase = pl * plane_stride + col;
        auto ldg = [&](int row) -> double2 { return ok ? src[(long long)row * row_stride] : make_double2(0.0, 0.0); };
        auto fix = [&](int row, double2 x) -> double2 { return ok ? cf(row, x, ebase + (long long)row * row_stride) : x; };
        auto stg = [&](int row, double2 x) { if (ok) dst[(long long)row * row_stride] = x; };
        auto lds = [&](int pos) -> double2 { return sm[pos * XT + c]; };
        auto sts = [&](int pos, double2 x) { sm[pos * XT + c] = x; };
        auto sync = [] { __syncthreads(); };
        transform<LOG2N, SGN, false, false>(t, TW, ldg, fix, stg, lds, sts, sync);
        if (Plan<LOG2N>::NS > 1) __syncthreads();   // the last stage's shared reads before the next tile's first writes
    }
}

// ---- pipelined form of the strided pass (N = 128, 256, 512) ----
// Persistent CTAs, three tile buffers: every thread copies the eight points ITS first stage will read with cp.async (16 B each,
// no registers held) two tiles ahead, so two tiles' worth of loads are in flight per CTA while a third is transformed.  The
// stages run in place (decimation in frequency, same butterflies and the same stage tables as above: butterfly b = hi S + lo
// touches hi R S + r S + lo, output k is scaled by w_N^{lo P k}); a thread's stage reads and writes hit the same slots, so one
// barrier per exchange is enough, and the first stage needs none (it reads only what the thread copied itself).  Results leave
// in digit-reversed slot order, which costs nothing here: every row of a tile is its own contiguous segment.
template <int LOG2N> struct PipeCfg {
    static constexpr int XT = 8, NBUF = 3, N = 1 << LOG2N;
    static constexpr int THREADS = XT * N / 8;
    static constexpr int TILE = XT * N;                         // complex numbers per buffer
    static constexpr size_t SMEM = sizeof(double2) * NBUF * (size_t)TILE;
    static constexpr int MINB = LOG2N >= 9 ? 1 : (LOG2N == 8 ? 2 : 4);
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// in-place stage I: slots of butterfly (t, u)
template <int LOG2N, int I>
__device__ __forceinline__ void ip_slots(int t, int u, int& base, int& lo) {
    using P = Plan<LOG2N>;
    constexpr int R = P::radix(I), S = P::npts(I), T = P::T;
    const int b = t + T * u, hi = b / S;
    lo = b - hi * S;
    base = hi * R * S + lo;
}
template <int LOG2N, int I, int SGN, class FX>
__device__ __forceinline__ void ip_stage(int t, double2 (&v)[8], const double2* __restrict__ TW, const double2* buf, int c, FX fix) {
    using P = Plan<LOG2N>;
    constexpr int R = P::radix(I), S = P::npts(I), NB = 8 / R, XT = PipeCfg<LOG2N>::XT;
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        int base, lo;
        ip_slots<LOG2N, I>(t, u, base, lo);
#pragma unroll
        for (int r = 0; r < R; ++r) v[u * R + r] = buf[(base + r * S) * XT + c];
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        int base, lo;
        ip_slots<LOG2N, I>(t, u, base, lo);
#pragma unroll
        for (int r = 0; r < R; ++r) v[u * R + r] = fix(base + r * S, v[u * R + r]);
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) bfly<R, SGN>(&v[u * R]);
    if (I < P::NS - 1) {
        const double2* tw = TW + P::toff(I);
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            int base, lo;
            ip_slots<LOG2N, I>(t, u, base, lo);
#pragma unroll
            for (int k = 1; k < R; ++k) {
                double2 w = __ldg(tw + (k - 1) * S + lo);
                if (SGN > 0) w.y = -w.y;
                v[u * R + k] = cmul(v[u * R + k], w);
            }
        }
    }
}
template <int LOG2N, int I>
__device__ __forceinline__ void ip_store(int t, double2 (&v)[8], double2* buf, int c) {
    using P = Plan<LOG2N>;
    constexpr int R = P::radix(I), S = P::npts(I), NB = 8 / R, XT = PipeCfg<LOG2N>::XT;
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        int base, lo;
        ip_slots<LOG2N, I>(t, u, base, lo);
#pragma unroll
        for (int k = 0; k < R; ++k) buf[(base + k * S) * XT + c] = v[u * R + k];
    }
}
// output index held by slot `pos` after the last in-place stage
template <int LOG2N>
__device__ __forceinline__ int ip_freq_of(int pos) {
    using P = Plan<LOG2N>;
    int rem = pos, f = 0;
#pragma unroll
    for (int i = 0; i < P::NS; ++i) {
        const int S = P::npts(i), d = rem / S;
        rem -= d * S;
        f += d * P::stride(i);
    }
    return f;
}

template <int LOG2N, int SGN>
__global__ void __launch_bounds__(PipeCfg<LOG2N>::THREADS, PipeCfg<LOG2N>::MINB)
fft_cols_pipe_kernel(const double2* __restrict__ in, double2* __restrict__ out, int ncols, int nplanes, long long row_stride,
                     long long plane_stride, FftPro pro, const double2* __restrict__ TW) {
    using P = Plan<LOG2N>;
    using Cc = PipeCfg<LOG2N>;
    constexpr int XT = Cc::XT, N = P::N, NS = P::NS, R0 = P::R0, T = P::T;
    static_assert(NS >= 2 && NS <= 4, "pipelined strided pass: 16 <= N <= 2048");
    extern __shared__ double2 sm[];
    const int c = threadIdx.x % XT, t = threadIdx.x / XT;
    const int cblocks = (ncols + XT - 1) / XT;
    const int ntiles = cblocks * nplanes;
    const int G = gridDim.x;
    const int f0 = ip_freq_of<LOG2N>(8 * t);   // last stage: butterfly t owns slots 8 t + k, outputs f0 + (N / 8) k

    auto prefetch = [&](int tile, double2* buf) {
        if (tile < ntiles) {
            const int pl = tile / cblocks, cb = tile - pl * cblocks, col = cb * XT + c;
            if (col < ncols) {
                const double2* src = in + pl * plane_stride + col;
#pragma unroll
                for (int u = 0; u < 8 / R0; ++u)
#pragma unroll
                    for (int r = 0; r < R0; ++r) {
                        const int row = t + T * u + (N / R0) * r;   // first stage: hi = 0, lo = b
                        cp_async16(buf + row * XT + c, src + (long long)row * row_stride);
                    }
            }
        }
        cp_async_commit();
    };

    prefetch(blockIdx.x, sm);
    prefetch(blockIdx.x + G, sm + Cc::TILE);
    int slot = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += G) {
        double2* buf = sm + slot * Cc::TILE;
        const int pl = tile / cblocks, cb = tile - pl * cblocks, col = cb * XT + c;
        const bool ok = col < ncols;
        double2* dst = out + pl * plane_stride + col;
        ColFix cf(pro);
        cf.setup(col, ok);
        const long long ebase = pl * plane_stride + col;
        auto fix = [&](int row, double2 x) -> double2 { return ok ? cf(row, x, ebase + (long long)row * row_stride) : x; };
        double2 v[8];
        cp_async_wait<1>();   // this tile's copies (mine) have landed; the next tile's may still be in flight
        ip_stage<LOG2N, 0, SGN>(t, v, TW, buf, c, fix);
        ip_store<LOG2N, 0>(t, v, buf, c);
        __syncthreads();
        // everybody is past the previous tile's last reads: its buffer takes the tile after next
        const int nslot = slot == 0 ? 2 : slot - 1;
        prefetch(tile + 2 * G, sm + nslot * Cc::TILE);
        if constexpr (NS >= 3) {
            ip_stage<LOG2N, 1, SGN>(t, v, TW, buf, c, NoFix());
            ip_store<LOG2N, 1>(t, v, buf, c);
            __syncthreads();
        }
        if constexpr (NS >= 4) {
            ip_stage<LOG2N, 2, SGN>(t, v, TW, buf, c, NoFix());
            ip_store<LOG2N, 2>(t, v, buf, c);
            __syncthreads();
        }
        ip_stage<LOG2N, NS - 1, SGN>(t, v, TW, buf, c, NoFix());
        if (ok) {
#pragma unroll
            for (int k = 0; k < 8; ++k) dst[(long long)(f0 + (N / 8) * k) * row_stride] = v[k];
        }
        slot = slot == 2 ? 0 : slot + 1;
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------------------------
// contiguous pass: real line of n = 2M points <-> M + 1 complex modes
// ---------------------------------------------------------------------------------------------------------------------------
template <int LOG2M, int LOGE> struct LineCfg {
    using P = Plan<LOG2M, LOGE>;
    static constexpr int M = 1 << LOG2M;
    static constexpr int T = P::T;
    static constexpr int THREADS = LOGE == 3 ? 256 : 128;          // 64 / 128 registers per thread, four blocks per SM
    static constexpr int MINB = 4;
    static constexpr int LPB = THREADS / T;                         // lines per block and iteration
    // shared index p + (p >> PS): one pad per leading-radix group makes the first stage's stride-R0 writes conflict-free
    static constexpr int PS = P::R0 >= 16 ? 4 : 3;
    static constexpr int LS = M + (M >> PS) + 1;                    // padded shared line
    static constexpr bool WARP = T <= 32;                           // a line's threads sit in one warp
    __device__ static __forceinline__ int pad(int p) { return p + (p >> PS); }
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

template <int LOG2M, int LOGE>
__global__ void __launch_bounds__(LineCfg<LOG2M, LOGE>::THREADS, LineCfg<LOG2M, LOGE>::MINB)
fft_r2c_kernel(RealPro pro, double2* __restrict__ out, long long nlines, const double2* __restrict__ TW, const double2* __restrict__ W) {
    using C = LineCfg<LOG2M, LOGE>;
    constexpr int M = C::M, T = C::T, LPB = C::LPB, LS = C::LS, E = 1 << LOGE;
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
        auto lds = [&](int p) -> double2 { return z[C::pad(p)]; };
        auto sts = [&](int p, double2 x) { z[C::pad(p)] = x; };
        auto sync = [] { if (C::WARP) __syncwarp(); else __syncthreads(); };
        transform<LOG2M, -1, false, true, LOGE>(t, TW, ldg, NoFix(), sts, lds, sts, sync);
        sync();
        // merge: X_k = E_k + w^k O_k, X_{M-k} = conj(E_k - w^k O_k), E = (Z_k + conj Z_{M-k})/2, O = -i (Z_k - conj Z_{M-k})/2
#pragma unroll
        for (int u = 0; u < E / 2; ++u) {
            const int k = t + T * u;   // (E / 2) T = M / 2 values
            if (k == 0) {
                const double2 z0 = z[0];
                if (ok) { dst[0] = make_double2(z0.x + z0.y, 0.0); dst[M] = make_double2(z0.x - z0.y, 0.0); }
            } else {
                const double2 A = z[C::pad(k)], B = z[C::pad(M - k)];
                const double2 Ev = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y - B.y));
                const double2 O = make_double2(0.5 * (A.y + B.y), -0.5 * (A.x - B.x));
                const double2 G = cmul(O, __ldg(W + k));
                if (ok) { dst[k] = cadd(Ev, G); dst[M - k] = make_double2(Ev.x - G.x, -(Ev.y - G.y)); }
            }
        }
        if (t == 0 && ok) { const double2 A = z[C::pad(M / 2)]; dst[M / 2] = make_double2(A.x, -A.y); }
        sync();   // merge reads before the next group's first writes
    }
}

template <int LOG2M, int LOGE>
__global__ void __launch_bounds__(LineCfg<LOG2M, LOGE>::THREADS, LineCfg<LOG2M, LOGE>::MINB)
fft_c2r_kernel(const double2* __restrict__ in, double* __restrict__ out, long long nlines, const double2* __restrict__ TW,
               const double2* __restrict__ W) {
    using C = LineCfg<LOG2M, LOGE>;
    constexpr int M = C::M, T = C::T, LPB = C::LPB, LS = C::LS, E = 1 << LOGE, NP = E / 2;
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
        {
            const double2 zz = make_double2(0.0, 0.0);
            double2 A[NP], B[NP];
#pragma unroll
            for (int u = 0; u < NP; ++u) {   // k = t + T u, NP T = M / 2 values; k = 0 pairs modes 0 and M
                const int k = t + T * u;
                A[u] = ok ? src[k] : zz;
                B[u] = ok ? src[M - k] : zz;
            }
            const double2 Ah = (t == 0 && ok) ? src[M / 2] : zz;
#pragma unroll
            for (int u = 0; u < NP; ++u) {
                const int k = t + T * u;
                if (k == 0) {
                    z[0] = make_double2(A[u].x + B[u].x, A[u].x - B[u].x);
                } else {
                    const double2 Ze = make_double2(A[u].x + B[u].x, A[u].y - B[u].y), D = make_double2(A[u].x - B[u].x, A[u].y + B[u].y);
                    double2 w = __ldg(W + k);
                    w.y = -w.y;
                    const double2 G = cmul(D, w);
                    z[C::pad(k)] = make_double2(Ze.x - G.y, Ze.y + G.x);
                    z[C::pad(M - k)] = make_double2(Ze.x + G.y, -Ze.y + G.x);
                }
            }
            if (t == 0) z[C::pad(M / 2)] = make_double2(2.0 * Ah.x, -2.0 * Ah.y);
        }
        sync();
        auto lds = [&](int p) -> double2 { return z[C::pad(p)]; };
        auto sts = [&](int p, double2 x) { z[C::pad(p)] = x; };
        auto stg = [&](int p, double2 x) { if (ok) *reinterpret_cast<double2*>(dst + 2 * p) = x; };
        transform<LOG2M, +1, true, false, LOGE>(t, TW, lds, NoFix(), stg, lds, sts, sync);
        sync();   // last stage's shared reads before the next group's split writes
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------
std::mutex g_tw_mutex;
std::map<int, double2*> g_tw;

// exp(-2 pi i j / n), j < n, from long double sines of the first octant (host)
void twiddles_host(int n, std::vector<double2>& h) {
    h.resize((size_t)n);
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
}
int twiddles(int n, const double2** out) {
    std::lock_guard<std::mutex> lk(g_tw_mutex);
    auto it = g_tw.find(n);
    if (it != g_tw.end()) { *out = it->second; return 0; }
    std::vector<double2> h;
    twiddles_host(n, h);
    double2* d = nullptr;
    PDO_CUDA(cudaMalloc(&d, sizeof(double2) * (size_t)n));
    PDO_CUDA(cudaMemcpy(d, h.data(), sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice));
    g_tw[n] = d;
    *out = d;
    return 0;
}

// the stage tables of Plan<log2 n> (see Plan::toff / npts), built from the flat table
template <int LOG2N, int LOGE = 3>
void fill_stage_tables(const std::vector<double2>& flat, std::vector<double2>& out) {
    using P = Plan<LOG2N, LOGE>;
    out.assign((size_t)(P::toff(P::NS - 1) > 0 ? P::toff(P::NS - 1) : 1), make_double2(1.0, 0.0));
    for (int i = 0; i + 1 < P::NS; ++i) {
        const int R = P::radix(i), s = P::stride(i), NP = P::npts(i);
        for (int k = 1; k < R; ++k)
            for (int p = 0; p < NP; ++p) out[(size_t)P::toff(i) + (size_t)(k - 1) * NP + p] = flat[(size_t)((long long)s * p * k) % P::N];
    }
}
// host: the stage tables of the plan (n, points per thread 2^loge); false when no such plan is compiled
bool stage_tables_host(int n, int loge, const std::vector<double2>& flat, std::vector<double2>& tab) {
    if (loge == 4) {
        if (n == 128) fill_stage_tables<7, 4>(flat, tab);
        else if (n == 256) fill_stage_tables<8, 4>(flat, tab);
        else return false;
        return true;
    }
    switch (n) {
        case 8: fill_stage_tables<3>(flat, tab); break;
        case 16: fill_stage_tables<4>(flat, tab); break;
        case 32: fill_stage_tables<5>(flat, tab); break;
        case 64: fill_stage_tables<6>(flat, tab); break;
        case 128: fill_stage_tables<7>(flat, tab); break;
        case 256: fill_stage_tables<8>(flat, tab); break;
        case 512: fill_stage_tables<9>(flat, tab); break;
        case 1024: fill_stage_tables<10>(flat, tab); break;
        default: return false;
    }
    return true;
}
std::map<int, double2*> g_stw;
int stage_twiddles(int n, const double2** out, int loge = 3) {
    const int key = n * 8 + loge;
    {
        std::lock_guard<std::mutex> lk(g_tw_mutex);
        auto it = g_stw.find(key);
        if (it != g_stw.end()) { *out = it->second; return 0; }
    }
    std::vector<double2> flat, tab;
    twiddles_host(n, flat);
    if (!stage_tables_host(n, loge, flat, tab)) return fail(PDO_E_BADARG, "stage_twiddles: no plan for n = %d, %d points per thread", n, 1 << loge);
    double2* d = nullptr;
    PDO_CUDA(cudaMalloc(&d, sizeof(double2) * tab.size()));
    PDO_CUDA(cudaMemcpy(d, tab.data(), sizeof(double2) * tab.size(), cudaMemcpyHostToDevice));
    std::lock_guard<std::mutex> lk(g_tw_mutex);
    g_stw[key] = d;
    *out = d;
    return 0;
}

// host: the plan constants a kernel instance is compiled with
template <int LOG2N, int LOGE>
void plan_constants(int* out) {
    using P = Plan<LOG2N, LOGE>;
    out[0] = P::NS; out[1] = P::R0; out[2] = P::T; out[3] = P::E;
    for (int i = 0; i < 4; ++i) {
        out[4 + i] = i < P::NS ? P::radix(i) : 0;
        out[8 + i] = i < P::NS ? P::stride(i) : 0;
        out[12 + i] = i < P::NS ? P::npts(i) : 0;
        out[16 + i] = i < P::NS ? P::toff(i) : 0;
    }
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
                const FftPro& pro, const double2* TW, cudaStream_t st) {
    using Cc = ColsCfg<LOG2N>;
    static bool attr_done = false;
    if (!attr_done) {
        PDO_CUDA(cudaFuncSetAttribute(fft_cols_kernel<LOG2N, SGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cc::SMEM));
        attr_done = true;
    }
    const long long ntiles = ((ncols + Cc::XT - 1) / Cc::XT) * nplanes;
    long long grid = (long long)sm_count() * Cc::MINB;
    if (grid > ntiles) grid = ntiles;
    if (ncols > 0x7fffffffLL || ntiles > 0x7fffffffLL) return fail(PDO_E_BADARG, "fft2d_cols: too many columns");
    fft_cols_kernel<LOG2N, SGN><<<(unsigned)grid, Cc::THREADS, Cc::SMEM, st>>>(in, out, (int)ncols, (int)nplanes, row_stride, plane_stride, pro, TW);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return 0;
}

// Radix 16 (sixteen points per thread, one exchange less through shared memory, 128 registers) for 256- and 512-point lines.
// Measured at 512^3 (profiles/r02w): c2r 383 vs 409 us, r2c 442 vs 416 us — the lower occupancy costs the r2c pass more than the
// lighter shared-memory traffic gives back, so only c2r takes it by default.  PDO_FFT_X = r8 / r16 forces one form for both.
bool x_radix16(int nx, bool c2r) {
    static int mode = -1;   // 0 default, 1 radix 8 everywhere, 2 radix 16 where a plan exists
    if (mode < 0) {
        const char* e = std::getenv("PDO_FFT_X");
        mode = (e && std::strcmp(e, "r8") == 0) ? 1 : ((e && std::strcmp(e, "r16") == 0) ? 2 : 0);
    }
    if (nx != 256 && nx != 512) return false;
    return mode == 2 || (mode == 0 && c2r);
}
bool cols_pipe_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = std::getenv("PDO_FFT_COLS");
        on = (e && std::strcmp(e, "simple") == 0) ? 0 : 1;
    }
    return on == 1;
}
template <int LOG2N, int SGN>
int launch_cols_pipe(long long ncols, long long nplanes, long long row_stride, long long plane_stride, const double2* in, double2* out,
                     const FftPro& pro, const double2* TW, cudaStream_t st) {
    using Cc = PipeCfg<LOG2N>;
    static bool attr_done = false;
    if (!attr_done) {
        PDO_CUDA(cudaFuncSetAttribute(fft_cols_pipe_kernel<LOG2N, SGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cc::SMEM));
        attr_done = true;
    }
    const long long ntiles = ((ncols + Cc::XT - 1) / Cc::XT) * nplanes;
    if (ncols > 0x7fffffffLL || ntiles > 0x3fffffffLL) return fail(PDO_E_BADARG, "fft2d_cols: too many columns");
    long long grid = (long long)sm_count() * Cc::MINB;
    if (grid > ntiles) grid = ntiles;
    fft_cols_pipe_kernel<LOG2N, SGN><<<(unsigned)grid, Cc::THREADS, Cc::SMEM, st>>>(in, out, (int)ncols, (int)nplanes, row_stride, plane_stride, pro, TW);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return 0;
}

template <int LOG2M, int LOGE = 3>
int launch_r2c(long long nlines, const RealPro& pro, double2* out, const double2* TW, const double2* W, cudaStream_t st) {
    using C = LineCfg<LOG2M, LOGE>;
    const long long ngroups = (nlines + C::LPB - 1) / C::LPB;
    long long grid = (long long)sm_count() * C::MINB;
    if (grid > ngroups) grid = ngroups;
    fft_r2c_kernel<LOG2M, LOGE><<<(unsigned)grid, C::THREADS, 0, st>>>(pro, out, nlines, TW, W);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return 0;
}
template <int LOG2M, int LOGE = 3>
int launch_c2r(long long nlines, const double2* in, double* out, const double2* TW, const double2* W, cudaStream_t st) {
    using C = LineCfg<LOG2M, LOGE>;
    const long long ngroups = (nlines + C::LPB - 1) / C::LPB;
    long long grid = (long long)sm_count() * C::MINB;
    if (grid > ngroups) grid = ngroups;
    fft_c2r_kernel<LOG2M, LOGE><<<(unsigned)grid, C::THREADS, 0, st>>>(in, out, nlines, TW, W);
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

// Tables are built (device allocation + synchronous copy) here, at plan time, so that the passes themselves only launch kernels:
// stream-ordered from the first call on and safe under stream capture.
int fft2d_prepare_x(int nx) {
    if (!fft2d_x_ok(nx)) return 0;
    const double2* t = nullptr;
    if (int rc = twiddles(nx, &t)) return rc;
    if (int rc = stage_twiddles(nx / 2, &t)) return rc;
    if (nx == 256 || nx == 512) return stage_twiddles(nx / 2, &t, 4);
    return 0;
}
int fft2d_prepare_cols(int n) {
    if (!fft2d_cols_ok(n)) return 0;
    const double2* t = nullptr;
    return stage_twiddles(n, &t);
}

int fft2d_cols(int n, long long ncols, long long nplanes, long long row_stride, long long plane_stride, const double2* in, double2* out,
               int dir, const FftPro& pro, cudaStream_t st) {
    if (!fft2d_cols_ok(n)) return fail(PDO_E_BADARG, "fft2d_cols: n = %d is not covered", n);
    if (ncols <= 0 || nplanes <= 0) return 0;
    const double2* W = nullptr;
    if (int rc = stage_twiddles(n, &W)) return rc;
#define PDO_COLS_CASE(L)                                                                                                         \
    case L:                                                                                                                      \
        return dir < 0 ? launch_cols<L, -1>(ncols, nplanes, row_stride, plane_stride, in, out, pro, W, st)                       \
                       : launch_cols<L, +1>(ncols, nplanes, row_stride, plane_stride, in, out, pro, W, st);
#define PDO_PIPE_CASE(L)                                                                                                         \
    case L:                                                                                                                      \
        return dir < 0 ? launch_cols_pipe<L, -1>(ncols, nplanes, row_stride, plane_stride, in, out, pro, W, st)                  \
                       : launch_cols_pipe<L, +1>(ncols, nplanes, row_stride, plane_stride, in, out, pro, W, st);
    if (cols_pipe_enabled()) {
        switch (ilog2(n)) { PDO_PIPE_CASE(7) PDO_PIPE_CASE(8) PDO_PIPE_CASE(9) default: break; }
    }
#undef PDO_PIPE_CASE
    switch (ilog2(n)) {
        PDO_COLS_CASE(4) PDO_COLS_CASE(5) PDO_COLS_CASE(6) PDO_COLS_CASE(7) PDO_COLS_CASE(8) PDO_COLS_CASE(9) PDO_COLS_CASE(10)
    }
#undef PDO_COLS_CASE
    return fail(PDO_E_BADARG, "fft2d_cols: n = %d", n);
}

int fft2d_r2c_lines(int nx, long long nlines, const RealPro& pro, double2* out, cudaStream_t st) {
    if (!fft2d_x_ok(nx)) return fail(PDO_E_BADARG, "fft2d_r2c: nx = %d is not covered", nx);
    if (nlines <= 0) return 0;
    const double2 *W = nullptr, *TW = nullptr;
    if (int rc = twiddles(nx, &W)) return rc;
    if (x_radix16(nx, false)) {
        if (int rc = stage_twiddles(nx / 2, &TW, 4)) return rc;
        return nx == 256 ? launch_r2c<7, 4>(nlines, pro, out, TW, W, st) : launch_r2c<8, 4>(nlines, pro, out, TW, W, st);
    }
    if (int rc = stage_twiddles(nx / 2, &TW)) return rc;
    switch (ilog2(nx / 2)) {
        case 3: return launch_r2c<3>(nlines, pro, out, TW, W, st);
        case 4: return launch_r2c<4>(nlines, pro, out, TW, W, st);
        case 5: return launch_r2c<5>(nlines, pro, out, TW, W, st);
        case 6: return launch_r2c<6>(nlines, pro, out, TW, W, st);
        case 7: return launch_r2c<7>(nlines, pro, out, TW, W, st);
        case 8: return launch_r2c<8>(nlines, pro, out, TW, W, st);
        case 9: return launch_r2c<9>(nlines, pro, out, TW, W, st);
        case 10: return launch_r2c<10>(nlines, pro, out, TW, W, st);
    }
    return fail(PDO_E_BADARG, "fft2d_r2c: nx = %d", nx);
}

int fft2d_c2r_lines(int nx, long long nlines, const double2* in, double* out, cudaStream_t st) {
    if (!fft2d_x_ok(nx)) return fail(PDO_E_BADARG, "fft2d_c2r: nx = %d is not covered", nx);
    if (nlines <= 0) return 0;
    const double2 *W = nullptr, *TW = nullptr;
    if (int rc = twiddles(nx, &W)) return rc;
    if (x_radix16(nx, true)) {
        if (int rc = stage_twiddles(nx / 2, &TW, 4)) return rc;
        return nx == 256 ? launch_c2r<7, 4>(nlines, in, out, TW, W, st) : launch_c2r<8, 4>(nlines, in, out, TW, W, st);
    }
    if (int rc = stage_twiddles(nx / 2, &TW)) return rc;
    switch (ilog2(nx / 2)) {
        case 3: return launch_c2r<3>(nlines, in, out, TW, W, st);
        case 4: return launch_c2r<4>(nlines, in, out, TW, W, st);
        case 5: return launch_c2r<5>(nlines, in, out, TW, W, st);
        case 6: return launch_c2r<6>(nlines, in, out, TW, W, st);
        case 7: return launch_c2r<7>(nlines, in, out, TW, W, st);
        case 8: return launch_c2r<8>(nlines, in, out, TW, W, st);
        case 9: return launch_c2r<9>(nlines, in, out, TW, W, st);
        case 10: return launch_c2r<10>(nlines, in, out, TW, W, st);
    }
    return fail(PDO_E_BADARG, "fft2d_c2r: nx = %d", nx);
}

// test hooks (host only): the plan a transform length is compiled with, its stage tables and the flat twiddle table, so that the
// CPU tests can re-enact the kernels' index algebra with the product's own constants (tests/test_fft_plan_cpu.py)
namespace hooks {
int fft_plan(int log2n, int loge, int* out20) {
    if (!out20) return fail(PDO_E_BADARG, "null argument");
    if (loge == 4) {
        if (log2n == 7) { plan_constants<7, 4>(out20); return 0; }
        if (log2n == 8) { plan_constants<8, 4>(out20); return 0; }
        return fail(PDO_E_BADARG, "no radix-16 plan for 2^%d points", log2n);
    }
    switch (log2n) {
        case 3: plan_constants<3, 3>(out20); return 0;
        case 4: plan_constants<4, 3>(out20); return 0;
        case 5: plan_constants<5, 3>(out20); return 0;
        case 6: plan_constants<6, 3>(out20); return 0;
        case 7: plan_constants<7, 3>(out20); return 0;
        case 8: plan_constants<8, 3>(out20); return 0;
        case 9: plan_constants<9, 3>(out20); return 0;
        case 10: plan_constants<10, 3>(out20); return 0;
    }
    return fail(PDO_E_BADARG, "no plan for 2^%d points", log2n);
}
// out: n flat twiddles then the stage tables (complex as pairs of doubles); returns the number of complex entries written
int fft_tables(int n, int loge, double* out, int capacity) {
    std::vector<double2> flat, tab;
    twiddles_host(n, flat);
    if (!stage_tables_host(n, loge, flat, tab)) return fail(PDO_E_BADARG, "no plan for n = %d", n);
    const int total = (int)(flat.size() + tab.size());
    if (!out || capacity < total) return fail(PDO_E_BADARG, "capacity %d < %d", capacity, total);
    std::memcpy(out, flat.data(), sizeof(double2) * flat.size());
    std::memcpy(out + 2 * flat.size(), tab.data(), sizeof(double2) * tab.size());
    return total;
}
}  // namespace hooks
}  // namespace pdo
