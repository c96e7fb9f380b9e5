// np_chunk.cu — chunked fast-path kernels of the NON-PERIODIC closures: CD10 d1 / d2, CF90 and CD06 d1
// with boundary codes (bc1, bcn), one fused pass per call (16 B per point) instead of the RHS pass + one-thread-per-line sweeps
// of the correctness path (nonperiodic.cu, ~48 B per point).
//
// Reference: derivatives/cd10.F90:429-709 (ComputePenta1/2), 823-1095 (SolveXPenta*), 1143-1262 / 1636-1731 (non-periodic RHS);
// filters/cf90.F90:276-418, 532-608, 672-801; derivatives/cd06.F90:264-327, 432-490, 551-590.
//
// Algebra (np_chunk_tables.cpp; the non-cyclic counterpart of tables.cpp): the line is cut into chunks of 32 rows whose last
// two rows are separators.  Three table sets — first / mid / last chunk, because the interior blocks of the end chunks contain
// boundary rows — give z = T^-1 r per chunk in registers; the reduced right-hand sides h_p = gA_p + gB_{p+1} meet in shared
// memory; the separator system is block TRIDIAGONAL here (not circulant), so its truncated inverse depends on the position:
// s_p = sum_d G[p][d] h_{p-W+d} with G read from global memory; x = z - V s_{p-1} - U s_p.  The right-hand side uses the
// correctness path's own per-point routine (np_rhs_point) for the one-sided rows and the interior stencil on the even / odd
// reflection of the line otherwise — the reflection is resolved when the chunk window is loaded, so the stencil itself is
// branch-free.  The mid-chunk code path reads its factors as immediate constant-bank operands (the tables travel as a
// __grid_constant__ kernel parameter, 7 KB); only the two end chunks of a line take the copies compiled for their sets.
//
// Thread (xi, p) owns chunk p of line / column xi; a CTA holds whole lines (P = n / 32 chunks, XT lines or columns, XT * P <= 512):
//   np_chunk_strided_kernel   solve axis y or z: XT contiguous x-columns, every global access a coalesced row segment
//   np_chunk_x_kernel         solve axis x: L whole lines staged through shared memory with coalesced 16-byte accesses
#include <cstdint>
#include <cstring>
#include <vector>

#include "nonperiodic.cuh"
#include "np_chunk_tables.h"

namespace pdo {

namespace {

constexpr int kM = 32;          // chunk length
constexpr int kH = 4;           // window halo on each side (CF90's 9-point stencil; CD10 uses 3, CD06 2)

struct NpSets3 { NpChunkSet first, mid, last; };

// interior stencil on a window whose out-of-line entries already hold the reflected values (w points at the centre)
template <int KIND>
__device__ __forceinline__ double np_interior(const double* w, const NpCoefs& c) {
    if (KIND == NP_CD10_D1) {
        return c.in[0] * (w[1] - w[-1]) + c.in[1] * (w[2] - w[-2]) + c.in[2] * (w[3] - w[-3]);
    } else if (KIND == NP_CD06_D1) {
        return c.in[1] * (w[2] - w[-2]) + c.in[0] * (w[1] - w[-1]);
    } else if (KIND == NP_CD10_D2) {
        const double f0 = w[0];
        return c.in[0] * (w[1] - 2.0 * f0 + w[-1]) + c.in[1] * (w[2] - 2.0 * f0 + w[-2]) + c.in[2] * (w[3] - 2.0 * f0 + w[-3]);
    } else {
        return c.in[0] * (w[0]) + c.in[1] * (w[1] + w[-1]) + c.in[2] * (w[2] + w[-2]) + c.in[3] * (w[3] + w[-3]) + c.in[4] * (w[4] + w[-4]);
    }
}

// The row a window slot stands for and its sign: row r (0-based, may lie outside [0, n)) of the reflected line G.
// G(j) = F(j) inside, bc1 * F(2 - j) below the first node, bcn * F(2n - j) above the last (1-based j; nonperiodic.cuh).
__device__ __forceinline__ int np_reflect(int r, int n, int bc1, int bcn, double* sign) {
    if (r < 0) { *sign = (double)bc1; return -r; }
    if (r >= n) { *sign = (double)bcn; return 2 * n - 2 - r; }
    *sign = 1.0;
    return r;
}

// right-hand side of my chunk from its window v[0 .. kM + 2 kH): v[kH + i] = row p*kM + i
template <int KIND>
__device__ __forceinline__ void np_chunk_rhs(const double (&v)[kM + 2 * kH], double (&r)[kM], int p, int P, int n, int bc1, int bcn,
                                             const NpCoefs& co) {
#pragma unroll
    for (int i = 0; i < kM; ++i) r[i] = np_interior<KIND>(&v[kH + i], co);
    constexpr int NB = (KIND == NP_CD10_D2 || KIND == NP_CD06_D1) ? 3 : 4;   // rows owned by the one-sided closure at each end
    if (p == 0 && bc1 == 0) {
        auto F = [&](int j) -> double { return v[kH + j - 1]; };              // 1-based node j of the line = window slot kH + j - 1
#pragma unroll
        for (int i = 0; i < NB; ++i) r[i] = np_rhs_point<KIND>(i + 1, kM, 0, 1, co, F);
    }
    (void)n;
    if (p == P - 1 && bcn == 0) {
        // the closure only looks at the last seven nodes: number them inside the chunk (a line of kM nodes ending at the wall), so
        // that every window index is a compile-time constant and v[] stays in registers
        auto F = [&](int j) -> double { return v[kH + j - 1]; };
#pragma unroll
        for (int i = 0; i < NB; ++i) r[kM - NB + i] = np_rhs_point<KIND>(kM - NB + 1 + i, kM, 1, 0, co, F);
    }
}

// interior sweeps with one table set: z in r[0 .. kM-2), reduced pieces gA (own separator rows) and gB (previous chunk's)
__device__ __forceinline__ void np_interior_solve(double (&r)[kM], const NpChunkSet& t, double (&gA)[2], double (&gB)[2]) {
    constexpr int mi = kM - 2;
    r[1] = __fma_rn(-t.l1[1], r[0], r[1]);
#pragma unroll
    for (int i = 2; i < mi; ++i) r[i] = __fma_rn(-t.l1[i], r[i - 1], __fma_rn(-t.l2[i], r[i - 2], r[i]));
    r[mi - 1] = r[mi - 1] * t.ginv[mi - 1];
    r[mi - 2] = __fma_rn(-t.ug[mi - 2], r[mi - 1], r[mi - 2] * t.ginv[mi - 2]);
#pragma unroll
    for (int i = mi - 3; i >= 0; --i) r[i] = __fma_rn(-t.ug[i], r[i + 1], __fma_rn(-t.bg[i], r[i + 2], r[i] * t.ginv[i]));
    gA[0] = r[kM - 2] - t.cA[0] * r[mi - 2] - t.cA[1] * r[mi - 1];
    gA[1] = r[kM - 1] - t.cA[2] * r[mi - 1];
    gB[0] = -t.cB[0] * r[0];
    gB[1] = -t.cB[1] * r[0] - t.cB[2] * r[1];
}
__device__ __forceinline__ void np_finish(double (&r)[kM], const NpChunkSet& t, double s0, double s1, double sp0, double sp1) {
    constexpr int mi = kM - 2;
#pragma unroll
    for (int i = 0; i < mi; ++i) r[i] = r[i] - t.V[i][0] * sp0 - t.V[i][1] * sp1 - t.U[i][0] * s0 - t.U[i][1] * s1;
    r[mi] = s0;
    r[mi + 1] = s1;
}

// steps 1-3 for the chunk in r; exchange through sm = gA[2][slots] gB[2][slots] s[2][slots]; `slot(q)` = slot of chunk q of my line
template <class SlotFn>
__device__ __forceinline__ void np_chunk_solve(double (&r)[kM], const NpSets3& tabs, const double* __restrict__ G, int P, int W, int p,
                                               bool active, double* __restrict__ sm, int slots, SlotFn slot) {
    double a_[2] = {0.0, 0.0}, b_[2] = {0.0, 0.0};
    // three copies of the unrolled sweeps, each with its set's factors as immediate constant operands
    if (p == 0) np_interior_solve(r, tabs.first, a_, b_);
    else if (p == P - 1) np_interior_solve(r, tabs.last, a_, b_);
    else np_interior_solve(r, tabs.mid, a_, b_);
    double* gA = sm;
    double* gB = sm + 2 * slots;
    double* sS = sm + 4 * slots;
    const int me = slot(p);
    gA[me] = a_[0]; gA[slots + me] = a_[1];
    gB[me] = b_[0]; gB[slots + me] = b_[1];
    __syncthreads();
    double s0 = 0.0, s1 = 0.0;
    if (active) {
        const double* g = G + (size_t)p * (2 * W + 1) * 4;
        for (int d = 0; d <= 2 * W; ++d) {
            const int q = p - W + d;
            if (q < 0 || q >= P) continue;           // zero blocks outside the matrix
            const int a = slot(q);
            double h0 = gA[a], h1 = gA[slots + a];
            if (q + 1 < P) { const int b = slot(q + 1); h0 += gB[b]; h1 += gB[slots + b]; }   // gB_P = 0
            s0 += __ldg(g + 4 * d + 0) * h0 + __ldg(g + 4 * d + 1) * h1;
            s1 += __ldg(g + 4 * d + 2) * h0 + __ldg(g + 4 * d + 3) * h1;
        }
    }
    sS[me] = s0; sS[slots + me] = s1;
    __syncthreads();
    double sp0 = 0.0, sp1 = 0.0;
    if (p > 0) { const int pm = slot(p - 1); sp0 = sS[pm]; sp1 = sS[slots + pm]; }
    if (p == 0) np_finish(r, tabs.first, s0, s1, sp0, sp1);
    else if (p == P - 1) np_finish(r, tabs.last, s0, s1, sp0, sp1);
    else np_finish(r, tabs.mid, s0, s1, sp0, sp1);
}

// ------------------------------------------------------------------------------------------------
// strided axes: f(n1, n, n3); THREADS = XT * P
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(512)
np_chunk_strided_kernel(const double* __restrict__ f, double* __restrict__ out, long long n1, int n, int tiles_x, int XT, int bc1, int bcn,
                        const double* __restrict__ G, int W, const __grid_constant__ NpCoefs co, const __grid_constant__ NpSets3 tabs) {
    extern __shared__ __align__(16) double sm[];
    const int P = n / kM;
    const int tid = threadIdx.x;
    const int xi = tid % XT, p = tid / XT;
    const long long tile = blockIdx.x;
    const long long k = tile / tiles_x;
    const long long x = (tile - k * tiles_x) * XT + xi;
    const bool active = x < n1;
    const double* fin = f + k * n1 * n + (active ? x : 0);
    double v[kM + 2 * kH];
    {
        const double* pr = fin + (long long)(p * kM) * n1;
#pragma unroll
        for (int j = 0; j < kM; ++j) { v[kH + j] = __ldg(pr); pr += n1; }
#pragma unroll
        for (int j = 0; j < kH; ++j) {
            double sg;
            const int q = np_reflect(p * kM - kH + j, n, bc1, bcn, &sg);
            v[j] = sg * __ldg(fin + (long long)q * n1);
            const int q2 = np_reflect((p + 1) * kM + j, n, bc1, bcn, &sg);
            v[kH + kM + j] = sg * __ldg(fin + (long long)q2 * n1);
        }
    }
    double r[kM];
    np_chunk_rhs<KIND>(v, r, p, P, n, bc1, bcn, co);
    const int slots = P * XT;
    np_chunk_solve(r, tabs, G, P, W, p, true, sm, slots, [&](int q) { return q * XT + xi; });
    if (active) {
        double* po = out + k * n1 * n + x + (long long)(p * kM) * n1;
#pragma unroll
        for (int i = 0; i < kM; ++i) { *po = r[i]; po += n1; }
    }
}

// Persistent, software-pipelined form of the strided kernel (the non-periodic sibling of banded.cu's chunk_strided_pipe_kernel):
// one CTA per SM walks tiles of XT columns x the whole line; the NEXT tile streams into shared memory with cp.async while the
// current one is solved out of registers, so load, FP64 work and store of successive tiles overlap.  The reflected / clamped
// window rows are read from the staged tile (no second global access).
__device__ __forceinline__ void np_cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void np_cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void np_issue_tile(double* buf, const double* __restrict__ f, long long tile, int tiles_x, int XT, int xt_shift,
                                              int n, long long n1, bool vec16) {
    const long long k = tile / tiles_x;
    const long long x0 = (tile - k * tiles_x) * XT;
    const double* base = f + k * n1 * n + x0;
    const int tid = threadIdx.x;
    if (vec16) {
        const int sh = xt_shift - 1;                 // log2(16-byte units per row)
        const int c = (tid & ((1 << sh) - 1)) * 2;
        int row = tid >> sh;
        const int rstep = blockDim.x >> sh;
        if (x0 + c < n1) {
            const double* src = base + (long long)row * n1 + c;
            double* dst = buf + row * XT + c;
            const long long sstep = (long long)rstep * n1;
            const int dstep = rstep * XT;
            for (; row < n; row += rstep) { np_cp_async16(dst, src); src += sstep; dst += dstep; }
        }
    } else {
        const int c = tid & (XT - 1);
        int row = tid >> xt_shift;
        const int rstep = blockDim.x >> xt_shift;
        if (x0 + c < n1) {
            const double* src = base + (long long)row * n1 + c;
            double* dst = buf + row * XT + c;
            const long long sstep = (long long)rstep * n1;
            const int dstep = rstep * XT;
            for (; row < n; row += rstep) { np_cp_async8(dst, src); src += sstep; dst += dstep; }
        }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

template <int KIND>
__global__ void __launch_bounds__(512, 1)
np_chunk_pipe_kernel(const double* __restrict__ f, double* __restrict__ out, long long n1, int n, int tiles_x, int XT, long long ntiles,
                     int vec16, int bc1, int bcn, const double* __restrict__ G, int W, const __grid_constant__ NpCoefs co,
                     const __grid_constant__ NpSets3 tabs) {
    extern __shared__ __align__(16) double sm[];
    const int P = n / kM;
    double* buf = sm;                                  // [n][XT]
    double* sm_g = sm + (size_t)n * XT;
    const int tid = threadIdx.x;
    const int xt_shift = __ffs(XT) - 1;
    const int xi = tid & (XT - 1), p = tid >> xt_shift;
    long long tile = blockIdx.x;
    if (tile < ntiles) np_issue_tile(buf, f, tile, tiles_x, XT, xt_shift, n, n1, vec16 != 0);
    for (; tile < ntiles; tile += gridDim.x) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        double v[kM + 2 * kH];
        {
            const double* b = buf + (p * kM) * XT + xi;
#pragma unroll
            for (int j = 0; j < kM; ++j) { v[kH + j] = *b; b += XT; }
#pragma unroll
            for (int j = 0; j < kH; ++j) {
                double sg;
                const int q = np_reflect(p * kM - kH + j, n, bc1, bcn, &sg);
                v[j] = sg * buf[q * XT + xi];
                const int q2 = np_reflect((p + 1) * kM + j, n, bc1, bcn, &sg);
                v[kH + kM + j] = sg * buf[q2 * XT + xi];
            }
        }
        double r[kM];
        np_chunk_rhs<KIND>(v, r, p, P, n, bc1, bcn, co);
        __syncthreads();                               // everyone has consumed its rows: the buffer may be refilled
        const long long next = tile + gridDim.x;
        if (next < ntiles) np_issue_tile(buf, f, next, tiles_x, XT, xt_shift, n, n1, vec16 != 0);
        np_chunk_solve(r, tabs, G, P, W, p, true, sm_g, P * XT, [&](int q) { return q * XT + xi; });
        const long long k = tile / tiles_x;
        const long long x = (tile - k * tiles_x) * XT + xi;
        if (x < n1) {
            double* po = out + k * n1 * n + x + (long long)(p * kM) * n1;
#pragma unroll
            for (int i = 0; i < kM; ++i) { *po = r[i]; po += n1; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// contiguous axis: f(n, nlines); a CTA stages L whole lines (one padding double per chunk: conflict-free chunk reads)
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256)
np_chunk_x_kernel(const double* __restrict__ f, double* __restrict__ out, long long nlines, int n, int L, int bc1, int bcn,
                  const double* __restrict__ G, int W, const __grid_constant__ NpCoefs co, const __grid_constant__ NpSets3 tabs) {
    extern __shared__ __align__(16) double sm[];
    const int P = n / kM;
    const int pitch = n + P;                         // chunk q of a line starts at q * (kM + 1)
    double* tile = sm;
    double* sm_g = sm + (size_t)L * pitch;
    const int slots = blockDim.x + P;
    const int tid = threadIdx.x;
    const long long line0 = (long long)blockIdx.x * L;
    const int nl = (int)min((long long)L, nlines - line0);
    const double* fbase = f + line0 * n;
    double* obase = out + line0 * n;
    const int tot = nl * n;
    const bool vec = ((reinterpret_cast<uintptr_t>(fbase) | reinterpret_cast<uintptr_t>(obase)) & 15) == 0;
    if (vec) {   // all global loads of a batch are issued before the first shared store: 8 independent 16-byte loads per thread
        const double2* f2 = reinterpret_cast<const double2*>(fbase);
        constexpr int U = 8;
        for (int g0 = 0; g0 < tot; g0 += 2 * (int)blockDim.x * U) {
            double2 val[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int g = g0 + 2 * (tid + u * (int)blockDim.x);
                if (g < tot) val[u] = __ldg(f2 + (g >> 1));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int g = g0 + 2 * (tid + u * (int)blockDim.x);
                if (g < tot) { const int pos = g + (g >> 5); tile[pos] = val[u].x; tile[pos + 1] = val[u].y; }
            }
        }
    } else {
        for (int g = tid; g < tot; g += blockDim.x) tile[g + (g >> 5)] = __ldg(fbase + g);
    }
    __syncthreads();
    const int p = tid % P, ln = tid / P;
    const bool active = (tid < L * P) && (ln < nl);
    const double* row = tile + (active ? ln : 0) * pitch;
    double v[kM + 2 * kH];
    {
#pragma unroll
        for (int j = 0; j < kM; ++j) v[kH + j] = row[p * (kM + 1) + j];
#pragma unroll
        for (int j = 0; j < kH; ++j) {
            double sg;
            const int q = np_reflect(p * kM - kH + j, n, bc1, bcn, &sg);
            v[j] = sg * row[q + (q >> 5)];
            const int q2 = np_reflect((p + 1) * kM + j, n, bc1, bcn, &sg);
            v[kH + kM + j] = sg * row[q2 + (q2 >> 5)];
        }
    }
    double r[kM];
    np_chunk_rhs<KIND>(v, r, p, P, n, bc1, bcn, co);
    np_chunk_solve(r, tabs, G, P, W, p, active, sm_g, slots, [&](int q) { return tid - p + q; });
    // (np_chunk_solve's barriers also guarantee every thread finished reading `tile`)
    if (active) {
        double* wrow = tile + ln * pitch + p * (kM + 1);
#pragma unroll
        for (int i = 0; i < kM; ++i) wrow[i] = r[i];
    }
    __syncthreads();
    if (vec) {
        double2* o2 = reinterpret_cast<double2*>(obase);
        for (int g = 2 * tid; g < tot; g += 2 * (int)blockDim.x) {
            const int pos = g + (g >> 5);
            o2[g >> 1] = make_double2(tile[pos], tile[pos + 1]);
        }
    } else {
        for (int g = tid; g < tot; g += blockDim.x) obase[g] = tile[g + (g >> 5)];
    }
}

template <int KIND>
cudaError_t np_fast_launch(const NpFast& t, const NpCoefs& co, int n, int axis, const double* f, double* out, long long n1, long long n3,
                           int bc1, int bcn, cudaStream_t st) {
    const int P = n / kM;
    NpSets3 tabs;
    std::memcpy(&tabs, t.sets, sizeof(tabs));
    if (axis == 0) {
        const int threads = P <= 128 ? (P > 64 ? 256 : 128) : 0;
        if (!threads) return cudaErrorInvalidConfiguration;
        const int L = threads / P;
        const size_t smem = sizeof(double) * ((size_t)L * (n + P) + 6 * (size_t)(threads + P));
        auto kern = np_chunk_x_kernel<KIND>;
        static bool attr_done = false;
        const size_t cap = 100 * 1024;
        if (smem > cap) return cudaErrorInvalidConfiguration;
        if (!attr_done) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
            if (e != cudaSuccess) return e;
            attr_done = true;
        }
        const long long grid = (n3 + L - 1) / L;
        kern<<<(unsigned)grid, threads, smem, st>>>(f, out, n3, n, L, bc1, bcn, t.d_G, t.W, co, tabs);
        return cudaGetLastError();
    }
    if (P > 512) return cudaErrorInvalidConfiguration;
    int XT = 1;
    while (XT * 2 * P <= 512) XT *= 2;
    while (XT > 1 && XT / 2 >= n1) XT /= 2;
    const int tiles_x = (int)((n1 + XT - 1) / XT);
    const long long ntiles = (long long)tiles_x * n3;
    if (ntiles >= (1LL << 31)) return cudaErrorInvalidConfiguration;
    const size_t smem = sizeof(double) * 6 * (size_t)P * XT;
    const size_t smem_pipe = sizeof(double) * (size_t)n * XT + smem;
    if (smem_pipe <= 200 * 1024 && ntiles >= 148 && XT >= 2) {   // enough tiles for a persistent grid: the pipelined kernel
        auto kern = np_chunk_pipe_kernel<KIND>;
        static bool attr_done = false;
        if (!attr_done) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return e;
            attr_done = true;
        }
        const int vec16 = (XT % 2 == 0) && (n1 % 2 == 0) && ((reinterpret_cast<uintptr_t>(f) & 15) == 0);
        kern<<<148, XT * P, smem_pipe, st>>>(f, out, n1, n, tiles_x, XT, ntiles, vec16, bc1, bcn, t.d_G, t.W, co, tabs);
        return cudaGetLastError();
    }
    np_chunk_strided_kernel<KIND><<<(unsigned)ntiles, XT * P, smem, st>>>(f, out, n1, n, tiles_x, XT, bc1, bcn, t.d_G, t.W, co, tabs);
    return cudaGetLastError();
}

}  // namespace

// Builds the chunk tables of one (bc1, bcn) system; leaves t->ok = false when the line is not chunkable (n % 32, < 2 chunks,
// separator reach beyond the table) — callers then keep the correctness path.
cudaError_t np_fast_create(NpFast* t, int kind, int n, int bc1, int bcn) {
    t->ok = false; t->d_G = nullptr; t->P = 0; t->W = 0;
    if (n % kM != 0 || n / kM < 2 || n / kM > 128) return cudaSuccess;
    std::vector<double> rows(5 * (size_t)n);
    if (np_build_rows(kind, n, bc1, bcn, rows.data()) != 0) return cudaSuccess;
    NpChunkTables ct;
    if (build_np_chunk_tables(n, kM, rows.data(), &ct) != 0) return cudaSuccess;
    static_assert(sizeof(NpFast::sets) == 3 * sizeof(NpChunkSet), "three table sets");
    std::memcpy(&t->sets[0], &ct.first, sizeof(NpChunkSet));
    std::memcpy(&t->sets[1], &ct.mid, sizeof(NpChunkSet));
    std::memcpy(&t->sets[2], &ct.last, sizeof(NpChunkSet));
    cudaError_t e = cudaMalloc(&t->d_G, sizeof(double) * ct.G.size());
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(t->d_G, ct.G.data(), sizeof(double) * ct.G.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(t->d_G); t->d_G = nullptr; return e; }
    t->P = ct.P; t->W = ct.W; t->ok = true;
    return cudaSuccess;
}
void np_fast_destroy(NpFast* t) {
    if (t->d_G) cudaFree(t->d_G);
    t->d_G = nullptr; t->ok = false;
}

cudaError_t np_fast_apply(const NpFast& t, int kind, const NpCoefs& co, int n, int axis, const double* f, double* out, long long n1,
                          long long n3, int bc1, int bcn, cudaStream_t st) {
    if (!t.ok) return cudaErrorInvalidConfiguration;
    switch (kind) {
        case NP_CD10_D1: return np_fast_launch<NP_CD10_D1>(t, co, n, axis, f, out, n1, n3, bc1, bcn, st);
        case NP_CD10_D2: return np_fast_launch<NP_CD10_D2>(t, co, n, axis, f, out, n1, n3, bc1, bcn, st);
        case NP_CF90: return np_fast_launch<NP_CF90>(t, co, n, axis, f, out, n1, n3, bc1, bcn, st);
        case NP_CD06_D1: return np_fast_launch<NP_CD06_D1>(t, co, n, axis, f, out, n1, n3, bc1, bcn, st);
        default: return cudaErrorInvalidConfiguration;
    }
}

}  // namespace pdo
