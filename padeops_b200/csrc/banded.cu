// banded.cu — hand-written sm_100a kernels for the periodic compact line operators.
//
// Replaces, per call, the reference's three sweeps (RHS stencil, forward elimination with running
// corner sums, back substitution):
//   CD10 d1/d2  derivatives/cd10.F90:1098-1142, 1265-1308, 1428-1468, 1587-1635 (RHS) + 712-1013 (Solve*LU*)
//   CD06 d1     derivatives/cd06.F90:515-550, 596-630, 674-704 (RHS) + 345-429 (Solve*LU1)
//   CF90        filters/cf90.F90:610-669 (+ CF90_files/ComputeZRHS_common.F90) + 421-530
//   Gaussian    filters/gaussian.F90:137-187 (explicit stencil, no solve)
//   CD06 stagg  derivatives/cd06stagg.F90:301-629 (RHS) + 248-299 (SolveZLU_*)
// with ONE fused pass: every grid point is read from HBM once and written once (16 B/pt).
//
// Two kernel families share one register-resident "chunk engine" (see tables.h for the algebra):
//   chunk_strided_kernel : solve axis is y or z.  A CTA owns XT contiguous x-columns x the whole line;
//                          thread (xi,p) owns chunk p (M consecutive points along the line) of column xi,
//                          so every global access is a coalesced row segment of XT doubles.
//   chunk_x_kernel       : solve axis is x (contiguous).  A CTA stages L whole lines in shared memory with
//                          coalesced 16-byte accesses (one padding double per chunk → conflict-free), thread
//                          (line,p) pulls its chunk into registers, solves, and the tile goes back the same way.
// LU factors / spike tables arrive as a __grid_constant__ struct, i.e. in the constant bank.
// Generic any-n kernels (one thread per line, tables in global memory) cover line lengths that are not a
// multiple of 8.  There is no CPU path.
#include <cuda.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "banded.cuh"

namespace pdo {

// ------------------------------------------------------------------------------------------------
// RHS stencils
// ------------------------------------------------------------------------------------------------
template <int RK> struct Halo;
template <> struct Halo<RK_D1_7> { static constexpr int L = 3, R = 3; };
template <> struct Halo<RK_D2_7> { static constexpr int L = 3, R = 3; };
template <> struct Halo<RK_D1_5> { static constexpr int L = 2, R = 2; };
template <> struct Halo<RK_SYM_9> { static constexpr int L = 4, R = 4; };
template <> struct Halo<RK_D2_5> { static constexpr int L = 2, R = 2; };
template <> struct Halo<RK_STAG_E2C> { static constexpr int L = 1, R = 2; };
template <> struct Halo<RK_STAG_C2E> { static constexpr int L = 2, R = 1; };

// w points at the centre value; operand order follows the Fortran expressions.
template <int RK>
__device__ __forceinline__ double rhs_eval(const double* w, const OpParams& op) {
    if (RK == RK_D1_7) {
        return op.co[0] * (w[1] - w[-1]) + op.co[1] * (w[2] - w[-2]) + op.co[2] * (w[3] - w[-3]);
    } else if (RK == RK_D2_7) {
        return op.co[0] * (w[1] - 2.0 * w[0] + w[-1]) + op.co[1] * (w[2] - 2.0 * w[0] + w[-2]) +
               op.co[2] * (w[3] - 2.0 * w[0] + w[-3]);
    } else if (RK == RK_D1_5) {
        return op.co[0] * (w[1] - w[-1]) + op.co[1] * (w[2] - w[-2]);
    } else if (RK == RK_SYM_9) {
        return op.co[0] * (w[0]) + op.co[1] * (w[1] + w[-1]) + op.co[2] * (w[2] + w[-2]) + op.co[3] * (w[3] + w[-3]) +
               op.co[4] * (w[4] + w[-4]);
    } else if (RK == RK_D2_5) {
        return op.co[0] * (w[1] - 2.0 * w[0] + w[-1]) + op.co[1] * (w[2] - 2.0 * w[0] + w[-2]);
    } else if (RK == RK_STAG_E2C) {
        return op.co[0] * (w[1] + op.co[2] * w[0]) + op.co[1] * (w[2] + op.co[2] * w[-1]);
    } else {  // RK_STAG_C2E
        return op.co[0] * (w[0] + op.co[2] * w[-1]) + op.co[1] * (w[1] + op.co[2] * w[-2]);
    }
}

// ------------------------------------------------------------------------------------------------
// Chunk engine: r[0..M) (RHS of this thread's chunk, in registers) → x[0..M) in place.
// smem: gA[BW][slots], gB[BW][slots], s[BW][slots]; `slot(q)` maps chunk q of MY line to its slot.
// Two __syncthreads; every thread of the CTA must call this.
// The sweeps are written so that the loop-carried dependency is ONE fma per row: the term that does not
// depend on the previous row is folded in first, and 1/g is pre-multiplied into the factors.
// ------------------------------------------------------------------------------------------------
// step 1: interior sweeps (register-only).  Leaves z in r[0..mi) and returns the reduced-RHS pieces.
template <int BW, int M>
__device__ __forceinline__ void chunk_interior(double (&r)[M], const ChunkTables& t, double (&gA)[2], double (&gB)[2]) {
    constexpr int mi = M - BW;
    // forward elimination on the interior block: y_i = (r_i - l2_i y_{i-2}) - l1_i y_{i-1}
    r[1] = __fma_rn(-t.l1[1], r[0], r[1]);
#pragma unroll
    for (int i = 2; i < mi; ++i) {
        if (BW == 2) r[i] = __fma_rn(-t.l1[i], r[i - 1], __fma_rn(-t.l2[i], r[i - 2], r[i]));
        else r[i] = __fma_rn(-t.l1[i], r[i - 1], r[i]);
    }
    // back substitution: z_i = (y_i ginv_i - bg_i z_{i+2}) - ug_i z_{i+1}
    r[mi - 1] = r[mi - 1] * t.ginv[mi - 1];
    if (BW == 2) {
        r[mi - 2] = __fma_rn(-t.ug[mi - 2], r[mi - 1], r[mi - 2] * t.ginv[mi - 2]);
#pragma unroll
        for (int i = mi - 3; i >= 0; --i)
            r[i] = __fma_rn(-t.ug[i], r[i + 1], __fma_rn(-t.bg[i], r[i + 2], r[i] * t.ginv[i]));
    } else {
#pragma unroll
        for (int i = mi - 2; i >= 0; --i) r[i] = __fma_rn(-t.ug[i], r[i + 1], r[i] * t.ginv[i]);
    }
    // reduced right-hand side pieces: gA from my chunk's tail rows, gB = what my head rows contribute to the
    // separator rows of the PREVIOUS chunk.
    if (BW == 2) {
        gA[0] = r[M - 2] - t.b2 * r[mi - 2] - t.b1 * r[mi - 1];
        gA[1] = r[M - 1] - t.b2 * r[mi - 1];
        gB[0] = -t.b2 * r[0];
        gB[1] = -t.b1 * r[0] - t.b2 * r[1];
    } else {
        gA[0] = r[M - 1] - t.b1 * r[mi - 1];
        gB[0] = -t.b1 * r[0];
        gA[1] = gB[1] = 0.0;
    }
}

// step 3: spikes.  s = own separator values, sp = previous chunk's.
template <int BW, int M>
__device__ __forceinline__ void chunk_finish(double (&r)[M], const ChunkTables& t, double s0, double s1, double sp0,
                                             double sp1) {
    constexpr int mi = M - BW;
#pragma unroll
    for (int i = 0; i < mi; ++i) {
        if (BW == 2) r[i] = r[i] - t.V[i][0] * sp0 - t.V[i][1] * sp1 - t.U[i][0] * s0 - t.U[i][1] * s1;
        else r[i] = r[i] - t.V[i][0] * sp0 - t.U[i][0] * s0;
    }
    r[mi] = s0;
    if (BW == 2) r[mi + 1] = s1;
}

// steps 1-3 with the separator exchange (step 2) through this CTA's shared memory.
template <int BW, int M, class SlotFn>
__device__ __forceinline__ void chunk_solve(double (&r)[M], const ChunkTables& t, double* __restrict__ sm_g,
                                            int slots, int p, SlotFn slot) {
    double a_[2], b_[2];
    chunk_interior<BW, M>(r, t, a_, b_);
    double* gA = sm_g;
    double* gB = sm_g + BW * slots;
    double* sS = sm_g + 2 * BW * slots;
    const int me = slot(p);
    gA[me] = a_[0];
    gB[me] = b_[0];
    if (BW == 2) { gA[slots + me] = a_[1]; gB[slots + me] = b_[1]; }
    __syncthreads();
    // separator solve: s_p = sum_d G[d] (gA_{p+d} + gB_{p+d+1})
    const int P = t.P, W = t.W;
    double s0 = 0.0, s1 = 0.0;
    int q = p - W;
    q %= P;
    if (q < 0) q += P;
    for (int d = 0; d <= 2 * W; ++d) {
        int q1 = q + 1;
        if (q1 == P) q1 = 0;
        const int a = slot(q), b = slot(q1);
        if (BW == 2) {
            const double h0 = gA[a] + gB[b];
            const double h1 = gA[slots + a] + gB[slots + b];
            s0 += t.G[d][0] * h0 + t.G[d][1] * h1;
            s1 += t.G[d][2] * h0 + t.G[d][3] * h1;
        } else {
            s0 += t.G[d][0] * (gA[a] + gB[b]);
        }
        q = q1;
    }
    sS[me] = s0;
    if (BW == 2) sS[slots + me] = s1;
    __syncthreads();
    const int pm = slot(p == 0 ? P - 1 : p - 1);
    const double sp0 = sS[pm];
    const double sp1 = (BW == 2) ? sS[slots + pm] : 0.0;
    chunk_finish<BW, M>(r, t, s0, s1, sp0, sp1);
}

// ------------------------------------------------------------------------------------------------
// Strided (y / z) kernels.  THREADS = XT * P threads per CTA (XT contiguous columns x P chunks).
//   chunk_strided_kernel       one tile per CTA, loads straight into registers (halo rows re-read via L1/L2)
//   chunk_strided_pipe_kernel  persistent CTAs; the NEXT tile streams into shared memory with cp.async while
//                              the current one is being solved out of registers; halo rows come from the tile
// ------------------------------------------------------------------------------------------------
template <int RK, int BW, int M, int THREADS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
chunk_strided_kernel(const double* __restrict__ f, double* __restrict__ out, long long n1, int n, long long in_slab,
                     long long out_slab, int tiles_x, int XT, const __grid_constant__ ChunkTables tab,
                     const __grid_constant__ OpParams op) {
    constexpr int HL = Halo<RK>::L, HR = Halo<RK>::R;
    extern __shared__ __align__(16) double sm_g[];
    const int tid = threadIdx.x;
    const int xi = tid % XT, p = tid / XT;
    const int P = n / M;
    const long long tile = blockIdx.x;
    const long long k = tile / tiles_x;
    const long long x = (tile - k * tiles_x) * XT + xi;
    const bool active = x < n1;
    const double* fin = f + k * in_slab + (active ? x : 0);
    double* fo = out + k * out_slab + (active ? x : 0);

    double v[M + HL + HR];
    const int nwrap = op.edge_in ? n + 1 : n;  // first logical position that wraps back by n
    {
        // running pointers: one 64-bit add per row instead of a 64-bit multiply
        const double* pr = fin + (long long)(p * M) * n1;
#pragma unroll
        for (int j = 0; j < M; ++j) { v[HL + j] = __ldg(pr); pr += n1; }
#pragma unroll
        for (int j = 0; j < HL; ++j) {
            int q = p * M - HL + j;
            if (q < 0) q += n;
            v[j] = __ldg(fin + (long long)q * n1);
        }
#pragma unroll
        for (int j = 0; j < HR; ++j) {
            int q = (p + 1) * M + j;
            if (q >= nwrap) q -= n;
            v[HL + M + j] = __ldg(fin + (long long)q * n1);
        }
    }
    double r[M];
#pragma unroll
    for (int i = 0; i < M; ++i) r[i] = rhs_eval<RK>(&v[i + HL], op);

    if constexpr (BW > 0) {
        const int slots = P * XT;
        chunk_solve<BW, M>(r, tab, sm_g, slots, p, [&](int q) { return q * XT + xi; });
    }
    if (active) {
        double* po = fo + (long long)(p * M) * n1;
#pragma unroll
        for (int i = 0; i < M; ++i) { *po = r[i]; po += n1; }
        if (op.edge_out && p == 0) fo[(long long)n * n1] = r[0];
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Stream tile `tile` (rows_in rows of XT doubles starting at column x0 of slab k) into smem `buf[row][XT]`.
// XT is a power of two and blockDim.x = XT*P, so a thread keeps ONE column and walks rows with pointer increments.
__device__ __forceinline__ void pipe_issue_tile(double* buf, const double* __restrict__ f, long long tile, int tiles_x,
                                                int XT, int xt_shift, int rows_in, long long n1, long long in_slab,
                                                bool vec16) {
    const long long k = tile / tiles_x;
    const long long x0 = (tile - k * tiles_x) * XT;
    const double* base = f + k * in_slab + x0;
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
            for (; row < rows_in; row += rstep) { cp_async16(dst, src); src += sstep; dst += dstep; }
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
            for (; row < rows_in; row += rstep) { cp_async8(dst, src); src += sstep; dst += dstep; }
        }
    }
    cp_async_commit();
}

template <int RK, int BW, int M, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
chunk_strided_pipe_kernel(const double* __restrict__ f, double* __restrict__ out, long long n1, int n, long long in_slab,
                          long long out_slab, int tiles_x, int XT, long long ntiles, int vec16,
                          const __grid_constant__ ChunkTables tab, const __grid_constant__ OpParams op) {
    constexpr int HL = Halo<RK>::L, HR = Halo<RK>::R;
    extern __shared__ __align__(16) double sm[];
    const int P = n / M;
    const int rows_in = op.edge_in ? n + 1 : n;
    double* buf = sm;                                   // [rows_in][XT]
    double* sm_g = sm + (((size_t)rows_in * XT + 1) & ~(size_t)1);
    const int tid = threadIdx.x;
    const int xt_shift = __ffs(XT) - 1;
    const int xi = tid & (XT - 1), p = tid >> xt_shift;
    const int nwrap = rows_in;

    long long tile = blockIdx.x;
    if (tile < ntiles) pipe_issue_tile(buf, f, tile, tiles_x, XT, xt_shift, rows_in, n1, in_slab, vec16 != 0);
    for (; tile < ntiles; tile += gridDim.x) {
        cp_async_wait_all();
        __syncthreads();
        double v[M + HL + HR];
        {
            const double* b = buf + (p * M) * XT + xi;
#pragma unroll
            for (int j = 0; j < M; ++j) { v[HL + j] = *b; b += XT; }
#pragma unroll
            for (int j = 0; j < HL; ++j) {
                int q = p * M - HL + j;
                if (q < 0) q += n;
                v[j] = buf[q * XT + xi];
            }
#pragma unroll
            for (int j = 0; j < HR; ++j) {
                int q = (p + 1) * M + j;
                if (q >= nwrap) q -= n;
                v[HL + M + j] = buf[q * XT + xi];
            }
        }
        // the stencil is evaluated before the buffer is released: its FP64 work overlaps the shared loads above
        double r[M];
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = rhs_eval<RK>(&v[i + HL], op);
        __syncthreads();  // everyone has consumed its rows: the buffer may be refilled
        const long long next = tile + gridDim.x;
        if (next < ntiles) pipe_issue_tile(buf, f, next, tiles_x, XT, xt_shift, rows_in, n1, in_slab, vec16 != 0);
        if constexpr (BW > 0) {
            const int slots = P * XT;
            chunk_solve<BW, M>(r, tab, sm_g, slots, p, [&](int q) { return q * XT + xi; });
        }
        const long long k = tile / tiles_x;
        const long long x = (tile - k * tiles_x) * XT + xi;
        if (x < n1) {
            double* fo = out + k * out_slab + x;
            double* po = fo + (long long)(p * M) * n1;
#pragma unroll
            for (int i = 0; i < M; ++i) { *po = r[i]; po += n1; }
            if (op.edge_out && p == 0) fo[(long long)n * n1] = r[0];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Cluster variant of the strided kernel: a line is split over the C CTAs of a thread-block cluster.
// CTA = 256 threads = PC (8) chunks x 32 columns, so every warp moves whole 256-byte row segments and two
// CTAs fit on an SM (their load / solve / store phases overlap).  The only coupling between the chunks of a
// line is the separator exchange, which goes through distributed shared memory (each CTA publishes its
// gA/gB/s, neighbours read them with mapa'd pointers) between cluster barriers.
// ------------------------------------------------------------------------------------------------
constexpr int kClThreads = 256;  // = PC chunks x XT columns; PC in {8, 4} -> XT in {32, 64}

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// generic address of `smem_ptr` in CTA `rank` of this cluster
__device__ __forceinline__ const double* cluster_map(const double* smem_ptr, unsigned rank) {
    unsigned long long in = (unsigned long long)smem_ptr, out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(in), "r"(rank));
    return (const double*)out;
}

constexpr int kClMaxHW = 8;  // halo chunks gathered from each neighbour CTA (needs W + 1 <= 8)

template <int RK, int BW, int M, int PC>
__global__ void __launch_bounds__(kClThreads, 2)
chunk_strided_cluster_kernel(const double* __restrict__ f, double* __restrict__ out, long long n1, int n, long long in_slab,
                             long long out_slab, int tiles_x, int C, const __grid_constant__ ChunkTables tab,
                             const __grid_constant__ OpParams op) {
    constexpr int HL = Halo<RK>::L, HR = Halo<RK>::R;
    constexpr int XT = kClThreads / PC;
    constexpr int PSH = (PC == 8) ? 3 : 2, XSH = (PC == 8) ? 5 : 6;
    constexpr int BWc = (BW > 0 ? BW : 1);
    constexpr int EC = PC + 2 * kClMaxHW;  // extended chunk slots: [halo left | own PC chunks | halo right]
    constexpr int ESL = EC * XT;
    // eA/eB[comp][extended chunk][column]; sS[comp][own chunk][column]
    __shared__ __align__(16) double eA[BWc * ESL];
    __shared__ __align__(16) double eB[BWc * ESL];
    __shared__ __align__(16) double sS[BWc * PC * XT];
    const int tid = threadIdx.x;
    const int xi = tid & (XT - 1), pl = tid >> XSH;
    const unsigned rank = (unsigned)(blockIdx.x % C);  // == %cluster_ctarank for the 1-D clusters launched here; explicit
    const int p0 = (int)rank * PC;                      // stencils (BW == 0) need no exchange and are launched unclustered
    const int p = p0 + pl;
    const int P = n / M;
    const long long cl = blockIdx.x / C;
    const long long k = cl / tiles_x;
    const long long x = (cl - k * tiles_x) * XT + xi;
    const bool active = x < n1;
    const double* fin = f + k * in_slab + (active ? x : 0);
    double* fo = out + k * out_slab + (active ? x : 0);

    double v[M + HL + HR];
    const int nwrap = op.edge_in ? n + 1 : n;
    {
        const double* pr = fin + (long long)(p * M) * n1;
#pragma unroll
        for (int j = 0; j < M; ++j) { v[HL + j] = __ldg(pr); pr += n1; }
#pragma unroll
        for (int j = 0; j < HL; ++j) {
            int q = p * M - HL + j;
            if (q < 0) q += n;
            v[j] = __ldg(fin + (long long)q * n1);
        }
#pragma unroll
        for (int j = 0; j < HR; ++j) {
            int q = (p + 1) * M + j;
            if (q >= nwrap) q -= n;
            v[HL + M + j] = __ldg(fin + (long long)q * n1);
        }
    }
    double r[M];
#pragma unroll
    for (int i = 0; i < M; ++i) r[i] = rhs_eval<RK>(&v[i + HL], op);

    if constexpr (BW > 0) {
        double a_[2], b_[2];
        chunk_interior<BW, M>(r, tab, a_, b_);
        const int W = tab.W, HW = W + 1;
        // publish into the middle region of the extended arrays (this is what neighbours read)
        const int me = (kClMaxHW + pl) * XT + xi;
        eA[me] = a_[0];
        eB[me] = b_[0];
        if (BW == 2) { eA[ESL + me] = a_[1]; eB[ESL + me] = b_[1]; }
        if (C > 1) cluster_sync_all(); else __syncthreads();
        // gather the HW halo chunks on each side from whoever owns them (DSMEM; all loads independent)
        for (int idx = tid; idx < 2 * HW * XT; idx += kClThreads) {
            const int hc = idx >> XSH, xc = idx & (XT - 1);
            const int e = (hc < HW) ? (kClMaxHW - HW + hc) : (kClMaxHW + PC + (hc - HW));  // my extended slot
            int q = p0 + (e - kClMaxHW);                                                  // global chunk
            q %= P;
            if (q < 0) q += P;
            const unsigned rq = (unsigned)(q >> PSH);
            const int src = (kClMaxHW + (q & (PC - 1))) * XT + xc;
            const double* A = (C > 1 && rq != rank) ? cluster_map(eA, rq) : eA;
            const double* B = (C > 1 && rq != rank) ? cluster_map(eB, rq) : eB;
            const double a0 = A[src], b0 = B[src];
            double a1 = 0.0, b1 = 0.0;
            if (BW == 2) { a1 = A[ESL + src]; b1 = B[ESL + src]; }
            eA[e * XT + xc] = a0;
            eB[e * XT + xc] = b0;
            if (BW == 2) { eA[ESL + e * XT + xc] = a1; eB[ESL + e * XT + xc] = b1; }
        }
        // my remote reads are done once this arrive executes; nobody leaves before everyone has arrived (wait below)
        if (C > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        __syncthreads();
        // separator solve from local shared memory: s_p = sum_d G[d] (gA_{p+d} + gB_{p+d+1})
        double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
        {
            int e = (kClMaxHW + pl - W) * XT + xi;
            for (int d = 0; d <= 2 * W; ++d) {
                if (BW == 2) {
                    const double h0 = eA[e] + eB[e + XT];
                    const double h1 = eA[ESL + e] + eB[ESL + e + XT];
                    s0 += tab.G[d][0] * h0 + tab.G[d][1] * h1;
                    s1 += tab.G[d][2] * h0 + tab.G[d][3] * h1;
                } else {
                    s0 += tab.G[d][0] * (eA[e] + eB[e + XT]);
                }
                e += XT;
            }
        }
        double sp0, sp1 = 0.0;
        if (pl == 0) {  // previous chunk lives in another CTA: recompute its separator values from the halo
            int e = (kClMaxHW - 1 - W) * XT + xi;
            for (int d = 0; d <= 2 * W; ++d) {
                if (BW == 2) {
                    const double h0 = eA[e] + eB[e + XT];
                    const double h1 = eA[ESL + e] + eB[ESL + e + XT];
                    t0 += tab.G[d][0] * h0 + tab.G[d][1] * h1;
                    t1 += tab.G[d][2] * h0 + tab.G[d][3] * h1;
                } else {
                    t0 += tab.G[d][0] * (eA[e] + eB[e + XT]);
                }
                e += XT;
            }
        }
        sS[pl * XT + xi] = s0;
        if (BW == 2) sS[PC * XT + pl * XT + xi] = s1;
        __syncthreads();
        if (pl == 0) { sp0 = t0; sp1 = t1; }
        else { sp0 = sS[(pl - 1) * XT + xi]; if (BW == 2) sp1 = sS[PC * XT + (pl - 1) * XT + xi]; }
        chunk_finish<BW, M>(r, tab, s0, s1, sp0, sp1);
    }
    if (active) {
        double* po = fo + (long long)(p * M) * n1;
#pragma unroll
        for (int i = 0; i < M; ++i) { *po = r[i]; po += n1; }
        if (op.edge_out && p == 0) fo[(long long)n * n1] = r[0];
    }
    if constexpr (BW > 0) {
        if (C > 1) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Cluster + pipeline variant ("cpipe"): the persistent cp.async pipeline of chunk_strided_pipe_kernel with the
// line split over the C CTAs of a cluster as in chunk_strided_cluster_kernel.  A CTA owns PC = 16 chunks
// (512 rows at M = 32) of XT = 32 columns, so every row segment it touches is 256 contiguous bytes whatever
// the line length (the single-CTA pipeline shrinks to 128 B at n = 1024 and 64 B at n = 2048): half the
// TLB / DRAM-page visits per byte when the row stride is megabytes (solve axis outermost, ddz).  The RHS
// halo rows of the neighbouring CTA are fetched with the tile (HL + HR extra rows, ~1 % extra L2 reads);
// the only cross-CTA traffic is the separator exchange through distributed shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kCpXT = 32;
constexpr int kCpPC = 16;
constexpr int kCpBoxRows = 256;
constexpr int kCpMaxHW = 8;

// --- distributed-shared-memory push with transaction barriers (no fences, no cluster barrier per tile) ---
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_map_u32(unsigned saddr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(mbar), "r"(parity) : "memory");
}
// 8-byte store into a peer CTA's shared memory that signals `remote_mbar` (in the same peer) with 8 transaction bytes
__device__ __forceinline__ void st_async_f64(unsigned remote_addr, double v, unsigned remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "l"(__double_as_longlong(v)), "r"(remote_mbar) : "memory");
}

// --- bulk-copy (TMA) and transaction-barrier helpers shared by the TMA-staged kernels ---
constexpr int kXtBuf = 3;
constexpr int kXTma = 1000, kXTma16 = 1016;  // x-kernel variant codes next to the CTA sizes 128 / 256

__device__ __forceinline__ void mbar_wait_or_trap(unsigned mbar, unsigned parity) {
    // bounded spin: a protocol error traps (the launch fails loudly) instead of hanging the device
    for (unsigned spin = 0; spin < (1u << 20); ++spin) {
        unsigned ok;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}" : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(unsigned smem_dst, const void* gsrc, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, unsigned smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(unsigned smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, unsigned smem_src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_src) : "memory");
}


template <int RK, int BW, int M>
__global__ void __launch_bounds__(kCpPC * kCpXT, 1)
chunk_strided_cpipe_kernel(const double* __restrict__ f, double* __restrict__ out, long long n1, int n, long long in_slab,
                           long long out_slab, int tiles_x, long long ntiles, int C, int vec16, int tma_in,
                           const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_halo,
                           const __grid_constant__ ChunkTables tab, const __grid_constant__ OpParams op) {
    constexpr int HL = Halo<RK>::L, HR = Halo<RK>::R;
    constexpr int XT = kCpXT, PC = kCpPC, THREADS = PC * XT;
    constexpr int RC = PC * M;            // rows this CTA solves
    constexpr int RB = RC + HL + HR;      // rows it buffers (own rows + stencil halo)
    constexpr int BWc = (BW > 0 ? BW : 1);
    constexpr int EC = PC + 2 * kCpMaxHW;  // extended chunk slots: [halo left | own PC chunks | halo right]
    constexpr int ESL = EC * XT;
    extern __shared__ __align__(128) double smt[];
    double* sm = smt;
    double* buf = sm;                     // [RB][XT]
    double* ex = sm + RB * XT;            // 2 x { eA[BWc][EC][XT], eB[BWc][EC][XT] }: ping-pong per tile (a neighbour may run
    double* sS = ex + 4 * BWc * ESL;      // one tile ahead and push into the other half); sS[BWc][PC][XT]
    __shared__ __align__(8) unsigned long long full_bar[2];  // transaction barriers: "halo of parity b has landed"
    __shared__ __align__(8) unsigned long long in_bar;       // tma_in: "the tile has landed" (one phase per tile)
    const int tid = threadIdx.x;
    const int xi = tid & (XT - 1), pl = tid >> 5;
    const unsigned rank = (C > 1) ? cluster_ctarank() : 0u;
    const int p0 = (int)rank * PC;
    const int nwrap = op.edge_in ? n + 1 : n;
    const long long ncl = gridDim.x / C;
    long long tile = blockIdx.x / C;

    auto issue = [&](long long t) {
        const long long k = t / tiles_x;
        const long long x0 = (t - k * tiles_x) * XT;
        if (tma_in) {  // three tensor-map boxes issued by one thread: halo above, own rows, halo below (HL == HR here)
            if (tid == 0) {
                const unsigned bar = smem_u32(&in_bar);
                mbar_arrive_expect_tx(bar, (unsigned)(RB * XT * sizeof(double)));
                const int rw0 = p0 * M;
                const int rl = rw0 == 0 ? n - HL : rw0 - HL;
                const int rr = rw0 + RC == n ? 0 : rw0 + RC;
                tma_load_3d(smem_u32(buf), &tm_halo, (int)x0, rl, (int)k, bar);
                for (int r = 0; r < RC; r += kCpBoxRows)   // a tensor-map box is at most 256 rows
                    tma_load_3d(smem_u32(buf + (HL + r) * XT), &tm_in, (int)x0, rw0 + r, (int)k, bar);
                tma_load_3d(smem_u32(buf + (HL + RC) * XT), &tm_halo, (int)x0, rr, (int)k, bar);
            }
            return;
        }
        const double* base = f + k * in_slab + x0;
        const int q0 = p0 * M - HL;  // logical row of buffer row 0
        if (vec16) {
            const int c = (tid & 15) * 2;
            if (x0 + c < n1) {
                for (int l = tid >> 4; l < RB; l += THREADS / 16) {
                    int q = q0 + l;
                    q = q < 0 ? q + n : (q >= nwrap ? q - n : q);
                    cp_async16(buf + l * XT + c, base + (long long)q * n1 + c);
                }
            }
        } else {
            if (x0 + xi < n1) {
                for (int l = pl; l < RB; l += PC) {
                    int q = q0 + l;
                    q = q < 0 ? q + n : (q >= nwrap ? q - n : q);
                    cp_async8(buf + l * XT + xi, base + (long long)q * n1 + xi);
                }
            }
        }
        cp_async_commit();
    };

    if (tma_in) {
        if (tid == 0) {
            mbar_init(smem_u32(&in_bar), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    if constexpr (BW > 0) {
        if (C > 1) {
            if (tid == 0) {
                mbar_init(smem_u32(&full_bar[0]), 1);
                mbar_init(smem_u32(&full_bar[1]), 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            cluster_sync_all();  // peers' barriers exist before anyone pushes
        }
    }
    if (tile < ntiles) issue(tile);
    int par = 0;
    unsigned it = 0;
    for (; tile < ntiles; tile += ncl, par ^= 1, ++it) {
        if (tma_in) {
            mbar_wait_or_trap(smem_u32(&in_bar), it & 1u);
        } else {
            cp_async_wait_all();
            __syncthreads();
        }
        double v[M + HL + HR];
        {
            const double* b = buf + (pl * M) * XT + xi;
#pragma unroll
            for (int j = 0; j < M + HL + HR; ++j) v[j] = b[j * XT];
        }
        // the stencil is evaluated before the buffer is released: its FP64 work overlaps the shared loads above and
        // the window v[] is dead by the time the solve needs registers
        double r[M];
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = rhs_eval<RK>(&v[i + HL], op);
        __syncthreads();  // everyone has consumed its rows: the buffer may be refilled
        const long long next = tile + ncl;
        if (next < ntiles) issue(next);

        if constexpr (BW > 0) {
            double a_[2], b_[2];
            chunk_interior<BW, M>(r, tab, a_, b_);
            const int W = tab.W, HW = W + 1;
            double* eA = ex + par * (2 * BWc * ESL);  // [BWc][EC][XT]
            double* eB = eA + BWc * ESL;
            const int me = (kCpMaxHW + pl) * XT + xi;
            eA[me] = a_[0];
            eB[me] = b_[0];
            if (BW == 2) { eA[ESL + me] = a_[1]; eB[ESL + me] = b_[1]; }
            // Push model: the chunks within HW of a CTA edge are what the neighbouring CTA needs as halo; their owner
            // stores them straight into the neighbour's extended slots (st.async, completion counted on the neighbour's
            // transaction barrier).  No cluster barrier and no fence sit on the per-tile path, so the prefetch issued
            // above and the previous tile's stores stay in flight.
            {
                const bool to_right = pl >= PC - HW;  // my tail chunks are the right neighbour's left halo
                const bool to_left = pl < HW;         // my head chunks are the left neighbour's right halo
                const int er = (kCpMaxHW + pl - PC) * XT + xi;
                const int el = (kCpMaxHW + PC + pl) * XT + xi;
                if (C > 1) {
                    const unsigned bar = smem_u32(&full_bar[par]);
                    if (tid == 0) mbar_arrive_expect_tx(bar, (unsigned)(2 * HW * XT * 2 * BW * sizeof(double)));
                    const unsigned aA = smem_u32(eA), aB = smem_u32(eB);
                    if (to_right) {
                        const unsigned rr = rank + 1 == (unsigned)C ? 0u : rank + 1;
                        const unsigned rb = cluster_map_u32(bar, rr);
                        const unsigned rA = cluster_map_u32(aA + er * 8, rr), rB = cluster_map_u32(aB + er * 8, rr);
                        st_async_f64(rA, a_[0], rb);
                        st_async_f64(rB, b_[0], rb);
                        if (BW == 2) { st_async_f64(rA + ESL * 8, a_[1], rb); st_async_f64(rB + ESL * 8, b_[1], rb); }
                    }
                    if (to_left) {
                        const unsigned rl = rank == 0 ? (unsigned)C - 1 : rank - 1;
                        const unsigned rb = cluster_map_u32(bar, rl);
                        const unsigned rA = cluster_map_u32(aA + el * 8, rl), rB = cluster_map_u32(aB + el * 8, rl);
                        st_async_f64(rA, a_[0], rb);
                        st_async_f64(rB, b_[0], rb);
                        if (BW == 2) { st_async_f64(rA + ESL * 8, a_[1], rb); st_async_f64(rB + ESL * 8, b_[1], rb); }
                    }
                } else {  // the line lives in this CTA: the halo is a periodic copy of my own edge chunks
                    if (to_right) {
                        eA[er] = a_[0]; eB[er] = b_[0];
                        if (BW == 2) { eA[ESL + er] = a_[1]; eB[ESL + er] = b_[1]; }
                    }
                    if (to_left) {
                        eA[el] = a_[0]; eB[el] = b_[0];
                        if (BW == 2) { eA[ESL + el] = a_[1]; eB[ESL + el] = b_[1]; }
                    }
                }
            }
            __syncthreads();
            if (C > 1) mbar_wait(smem_u32(&full_bar[par]), (it >> 1) & 1u);
            // separator solve from local shared memory: s_p = sum_d G[d] (gA_{p+d} + gB_{p+d+1})
            double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
            {
                int e = (kCpMaxHW + pl - W) * XT + xi;
                for (int d = 0; d <= 2 * W; ++d) {
                    if (BW == 2) {
                        const double h0 = eA[e] + eB[e + XT];
                        const double h1 = eA[ESL + e] + eB[ESL + e + XT];
                        s0 += tab.G[d][0] * h0 + tab.G[d][1] * h1;
                        s1 += tab.G[d][2] * h0 + tab.G[d][3] * h1;
                    } else {
                        s0 += tab.G[d][0] * (eA[e] + eB[e + XT]);
                    }
                    e += XT;
                }
            }
            double sp0, sp1 = 0.0;
            if (pl == 0) {  // previous chunk lives in another CTA: recompute its separator values from the halo
                int e = (kCpMaxHW - 1 - W) * XT + xi;
                for (int d = 0; d <= 2 * W; ++d) {
                    if (BW == 2) {
                        const double h0 = eA[e] + eB[e + XT];
                        const double h1 = eA[ESL + e] + eB[ESL + e + XT];
                        t0 += tab.G[d][0] * h0 + tab.G[d][1] * h1;
                        t1 += tab.G[d][2] * h0 + tab.G[d][3] * h1;
                    } else {
                        t0 += tab.G[d][0] * (eA[e] + eB[e + XT]);
                    }
                    e += XT;
                }
            }
            sS[pl * XT + xi] = s0;
            if (BW == 2) sS[PC * XT + pl * XT + xi] = s1;
            __syncthreads();
            if (pl == 0) { sp0 = t0; sp1 = t1; }
            else { sp0 = sS[(pl - 1) * XT + xi]; if (BW == 2) sp1 = sS[PC * XT + (pl - 1) * XT + xi]; }
            chunk_finish<BW, M>(r, tab, s0, s1, sp0, sp1);
        }
        const long long k = tile / tiles_x;
        const long long x = (tile - k * tiles_x) * XT + xi;
        if (x < n1) {
            double* fo = out + k * out_slab + x;
            double* po = fo + (long long)((p0 + pl) * M) * n1;
#pragma unroll
            for (int i = 0; i < M; ++i) { *po = r[i]; po += n1; }
            if (op.edge_out && p0 + pl == 0) fo[(long long)n * n1] = r[0];
        }
    }
    // Exit is safe without a barrier: nobody reads a peer's shared memory, and every push aimed at this CTA was
    // awaited by its last mbar_wait.
}

// ------------------------------------------------------------------------------------------------
// Contiguous (x) kernel
// ------------------------------------------------------------------------------------------------
constexpr int kXPairsPerThread = 16;  // a full tile is L*n = 32 doubles per thread = 16 double2 per thread

// kXThreads threads per CTA, 512 / kXThreads CTAs per SM: smaller CTAs put more independent load / solve / store
// phases in flight on an SM.
template <int RK, int BW, int M, int kXThreads>
__global__ void __launch_bounds__(kXThreads, 512 / kXThreads)
chunk_x_kernel(const double* __restrict__ f, double* __restrict__ out, long long nlines, int n, int L,
               const __grid_constant__ ChunkTables tab, const __grid_constant__ OpParams op) {
    constexpr int HL = Halo<RK>::L, HR = Halo<RK>::R;
    extern __shared__ __align__(16) double sm[];
    const int P = n / M;
    const int pitch = n + P;  // one padding double per chunk: chunk p starts at p*(M+1)
    double* tile = sm;
    double* sm_g = sm + (size_t)L * pitch;
    const int slots = kXThreads + P;  // threads beyond L*P own no line but still index slot(q) in chunk_solve
    const int tid = threadIdx.x;
    const long long line0 = (long long)blockIdx.x * L;
    const int nl = (int)min((long long)L, nlines - line0);  // lines present in this tile
    const double* fbase = f + line0 * n;
    double* obase = out + line0 * n;

    // ---- coalesced tile load: all global loads of a thread are issued before the first shared store.
    // The padded layout puts one pad after every M doubles and lines are multiples of M long, so the shared
    // position of flat tile element g is simply g + g/M — no line bookkeeping.
    const long long tot = (long long)nl * n;
    const bool vec = ((reinterpret_cast<uintptr_t>(fbase) | reinterpret_cast<uintptr_t>(obase)) & 15) == 0;
    constexpr int MSH = (M == 32) ? 5 : (M == 16) ? 4 : 3;
    if (vec) {
        const double2* f2 = reinterpret_cast<const double2*>(fbase);
        for (int g0 = 0; g0 < tot; g0 += 2 * kXThreads * kXPairsPerThread) {
            double2 val[kXPairsPerThread];
#pragma unroll
            for (int u = 0; u < kXPairsPerThread; ++u) {
                const int g = g0 + 2 * (tid + u * kXThreads);
                if (g < tot) val[u] = __ldg(f2 + (g >> 1));
            }
#pragma unroll
            for (int u = 0; u < kXPairsPerThread; ++u) {
                const int g = g0 + 2 * (tid + u * kXThreads);
                if (g < tot) {
                    const int pos = g + (g >> MSH);
                    tile[pos] = val[u].x;
                    tile[pos + 1] = val[u].y;
                }
            }
        }
    } else {
        for (int g = tid; g < tot; g += kXThreads) tile[g + (g >> MSH)] = __ldg(fbase + g);
    }
    __syncthreads();

    const int p = tid % P, ln = tid / P;
    const bool active = (tid < L * P) && (ln < nl);
    const double* row = tile + (active ? ln : 0) * pitch;
    double v[M + HL + HR];
    {
        const int pl = (p == 0 ? P - 1 : p - 1) * (M + 1) + M - HL;  // left halo lives at the tail of chunk p-1
        const int pr = (p == P - 1 ? 0 : p + 1) * (M + 1);           // right halo at the head of chunk p+1
        const int pc = p * (M + 1);
#pragma unroll
        for (int j = 0; j < HL; ++j) v[j] = row[pl + j];
#pragma unroll
        for (int j = 0; j < M; ++j) v[HL + j] = row[pc + j];
#pragma unroll
        for (int j = 0; j < HR; ++j) v[HL + M + j] = row[pr + j];
    }
    double r[M];
#pragma unroll
    for (int i = 0; i < M; ++i) r[i] = rhs_eval<RK>(&v[i + HL], op);

    if constexpr (BW > 0) {
        chunk_solve<BW, M>(r, tab, sm_g, slots, p, [&](int q) { return tid - p + q; });
    } else {
        __syncthreads();  // everyone has read the tile
    }
    // (chunk_solve's barriers also guarantee every thread finished reading `tile`)
    if (active) {
        double* wrow = tile + ln * pitch + p * (M + 1);
#pragma unroll
        for (int i = 0; i < M; ++i) wrow[i] = r[i];
    }
    __syncthreads();
    // ---- coalesced tile store ----
    if (vec) {
        double2* o2 = reinterpret_cast<double2*>(obase);
        for (int g = 2 * tid; g < tot; g += 2 * kXThreads) {
            const int pos = g + (g >> MSH);
            o2[g >> 1] = make_double2(tile[pos], tile[pos + 1]);
        }
    } else {
        for (int g = tid; g < tot; g += kXThreads) obase[g] = tile[g + (g >> MSH)];
    }
}

// ------------------------------------------------------------------------------------------------
// Contiguous (x) kernel, TMA-staged persistent pipeline ("xtma")
//
// One persistent CTA per SM walks tiles of L whole lines.  Every line moves HBM -> shared and shared -> HBM as ONE
// 1-D bulk copy (cp.async.bulk, SASS UBLKCP) issued by a lane of warp 0: no thread stages data through registers.
// Three tile buffers rotate through the states {landing, being solved, draining}; a transaction mbarrier per buffer
// says when a tile has landed, a bulk group per issuing lane says when a drain has finished reading its buffer.
// Lines sit in shared memory with a pitch of n+2 doubles (rows stay 16-byte aligned) and lanes of a quarter-warp
// own the SAME chunk of 8 DIFFERENT lines (ln = tid % L), so every 128-bit shared access of a quarter-warp hits 8
// distinct 16-byte bank groups ((n+2)/2 is odd): conflict-free without per-chunk padding, which a bulk copy could
// not produce.  The solved chunk goes back in place (all reads of the tile precede chunk_solve's first barrier).
// ------------------------------------------------------------------------------------------------
template <int RK, int BW, int M, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
chunk_x_tma_kernel(const double* __restrict__ f, double* __restrict__ out, long long nlines, int n, int L, long long ntiles,
                   const __grid_constant__ ChunkTables tab, const __grid_constant__ OpParams op) {
    constexpr int HL = Halo<RK>::L, HR = Halo<RK>::R;
    static_assert(HL <= 4 && HR <= 4 && M % 2 == 0, "halo window is 4 doubles each side");
    extern __shared__ __align__(16) double sm[];
    __shared__ __align__(8) unsigned long long full_bar[kXtBuf];
    const int P = n / M;
    const int pitch = n + 2;
    const size_t tile_elems = (size_t)L * pitch;
    double* sm_g = sm + kXtBuf * tile_elems;
    const int tid = threadIdx.x;
    const int ln = tid % L;
    const int pq = tid / L;
    const bool active = pq < P;
    const int p = active ? pq : 0;
    const int slots = THREADS;
    const unsigned line_bytes = (unsigned)n * (unsigned)sizeof(double);

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < kXtBuf; ++b) mbar_init(smem_u32(&full_bar[b]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // warp 0 is the copy issuer: lane i moves lines i, i+32, ... of a tile
    auto issue_load = [&](long long tile, int b) {
        const long long line0 = tile * L;
        const int nl = (int)min((long long)L, nlines - line0);
        const unsigned bar = smem_u32(&full_bar[b]);
        if (tid == 0) mbar_arrive_expect_tx(bar, (unsigned)nl * line_bytes);
        __syncwarp();
        for (int i = tid; i < nl; i += 32)
            bulk_g2s(smem_u32(sm + b * tile_elems + (size_t)i * pitch), f + (line0 + i) * n, line_bytes, bar);
    };

    const long long t0 = blockIdx.x;
    if (tid < 32) {
        if (t0 < ntiles) issue_load(t0, 0);
        if (t0 + gridDim.x < ntiles) issue_load(t0 + gridDim.x, 1);
    }

    int it = 0;
    for (long long tile = t0; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it % kXtBuf;
        const long long line0 = tile * L;
        const int nl = (int)min((long long)L, nlines - line0);
        double* tb = sm + b * tile_elems;
        mbar_wait_or_trap(smem_u32(&full_bar[b]), (unsigned)(it / kXtBuf) & 1u);

        // ---- my chunk plus a 4-double window on each side, 128-bit shared loads ----
        const double* row = tb + (size_t)ln * pitch;
        double w[M + 8];
        {
            const int c0 = p * M;
            const int cl = (p == 0) ? n - 4 : c0 - 4;
            const int cr = (p == P - 1) ? 0 : c0 + M;
            const double2 a0 = *reinterpret_cast<const double2*>(row + cl);
            const double2 a1 = *reinterpret_cast<const double2*>(row + cl + 2);
            w[0] = a0.x; w[1] = a0.y; w[2] = a1.x; w[3] = a1.y;
#pragma unroll
            for (int j = 0; j < M; j += 2) {
                const double2 c = *reinterpret_cast<const double2*>(row + c0 + j);
                w[4 + j] = c.x; w[5 + j] = c.y;
            }
            const double2 b0 = *reinterpret_cast<const double2*>(row + cr);
            const double2 b1 = *reinterpret_cast<const double2*>(row + cr + 2);
            w[M + 4] = b0.x; w[M + 5] = b0.y; w[M + 6] = b1.x; w[M + 7] = b1.y;
        }
        double r[M];
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = rhs_eval<RK>(&w[4 + i], op);

        if constexpr (BW > 0) {
            chunk_solve<BW, M>(r, tab, sm_g, slots, p, [&](int q) { return active ? q * L + ln : tid; });
        } else {
            __syncthreads();  // everyone has read the tile
        }
        if (active) {
            double* wrow = tb + (size_t)ln * pitch + p * M;
#pragma unroll
            for (int j = 0; j < M; j += 2) *reinterpret_cast<double2*>(wrow + j) = make_double2(r[j], r[j + 1]);
        }
        fence_proxy_async_smem();  // my generic-proxy writes become visible to the bulk-copy (async) proxy
        __syncthreads();
        if (tid < 32) {
            for (int i = tid; i < nl; i += 32)
                bulk_s2g(out + (line0 + i) * n, smem_u32(tb + (size_t)i * pitch), line_bytes);
            bulk_commit();
            // buffer (it+2)%3 was drained by the group committed one iteration ago: wait until that one has read its
            // shared source (the group just committed may stay in flight), then refill it with tile it+2
            bulk_wait_read<1>();
            __syncwarp();
            const long long nxt = tile + 2LL * gridDim.x;
            if (nxt < ntiles) issue_load(nxt, (it + 2) % kXtBuf);
        }
    }
    if (tid < 32) bulk_wait_read<0>();  // shared memory must outlive the last drains
}

// ------------------------------------------------------------------------------------------------
// Strided (y / z) kernel, TMA-staged persistent pipeline ("stma")
//
// Same rotation of three tile buffers as chunk_x_tma_kernel, but a tile is XT contiguous x-columns x the whole line
// and moves as a few 3-D tensor-map boxes {XT, BR, 1} (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG): the TMA unit
// walks the n1-strided rows, zero-fills / clips columns past n1, and no thread issues a per-row copy or store (the
// register-staged pipelines spend most of their stall cycles in the LSU queue doing exactly that).  Thread (xi, p)
// owns chunk p of column xi: a half-warp reads XT >= 8 adjacent doubles of one row, conflict-free.
// ------------------------------------------------------------------------------------------------
template <int RK, int BW, int M>
__global__ void __launch_bounds__(512, 1)
chunk_strided_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, int n, int XT,
                         int BR, int nbox_in, int nbox_out, int tiles_x, long long ntiles,
                         const __grid_constant__ ChunkTables tab, const __grid_constant__ OpParams op) {
    constexpr int HL = Halo<RK>::L, HR = Halo<RK>::R;
    extern __shared__ __align__(128) double smt[];
    __shared__ __align__(8) unsigned long long full_bar[kXtBuf];
    const int P = n / M;
    const int rows_in = op.edge_in ? n + 1 : n;
    const int nbox_max = nbox_in > nbox_out ? nbox_in : nbox_out;
    const size_t tile_elems = (size_t)nbox_max * BR * XT;
    const unsigned box_bytes = (unsigned)(BR * XT) * (unsigned)sizeof(double);
    double* sm_g = smt + kXtBuf * tile_elems;
    const int tid = threadIdx.x;
    const int xt_shift = __ffs(XT) - 1;
    const int xi = tid & (XT - 1), p = tid >> xt_shift;

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < kXtBuf; ++b) mbar_init(smem_u32(&full_bar[b]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_load = [&](long long tile, int b) {  // thread 0 only
        const int k = (int)(tile / tiles_x);
        const int x0 = (int)(tile - (long long)k * tiles_x) * XT;
        const unsigned bar = smem_u32(&full_bar[b]);
        mbar_arrive_expect_tx(bar, box_bytes * (unsigned)nbox_in);
        for (int i = 0; i < nbox_in; ++i)
            tma_load_3d(smem_u32(smt + b * tile_elems + (size_t)i * BR * XT), &tm_in, x0, i * BR, k, bar);
    };

    const long long t0 = blockIdx.x;
    if (tid == 0) {
        if (t0 < ntiles) issue_load(t0, 0);
        if (t0 + gridDim.x < ntiles) issue_load(t0 + gridDim.x, 1);
    }

    int it = 0;
    for (long long tile = t0; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it % kXtBuf;
        double* buf = smt + b * tile_elems;
        mbar_wait_or_trap(smem_u32(&full_bar[b]), (unsigned)(it / kXtBuf) & 1u);

        double v[M + HL + HR];
        {
            const double* c = buf + (p * M) * XT + xi;
#pragma unroll
            for (int j = 0; j < M; ++j) { v[HL + j] = *c; c += XT; }
#pragma unroll
            for (int j = 0; j < HL; ++j) {
                int q = p * M - HL + j;
                if (q < 0) q += n;
                v[j] = buf[q * XT + xi];
            }
#pragma unroll
            for (int j = 0; j < HR; ++j) {
                int q = (p + 1) * M + j;
                if (q >= rows_in) q -= n;  // the staggered edge plane n is read un-wrapped
                v[HL + M + j] = buf[q * XT + xi];
            }
        }
        double r[M];
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = rhs_eval<RK>(&v[i + HL], op);

        if constexpr (BW > 0) {
            chunk_solve<BW, M>(r, tab, sm_g, P * XT, p, [&](int q) { return q * XT + xi; });
        } else {
            __syncthreads();  // everyone has read the tile
        }
        {
            double* c = buf + (p * M) * XT + xi;
#pragma unroll
            for (int i = 0; i < M; ++i) { *c = r[i]; c += XT; }
            if (op.edge_out && p == 0) buf[n * XT + xi] = r[0];
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            const int k = (int)(tile / tiles_x);
            const int x0 = (int)(tile - (long long)k * tiles_x) * XT;
            for (int i = 0; i < nbox_out; ++i) tma_store_3d(&tm_out, x0, i * BR, k, smem_u32(buf + (size_t)i * BR * XT));
            bulk_commit();
            bulk_wait_read<1>();
            const long long nxt = tile + 2LL * gridDim.x;
            if (nxt < ntiles) issue_load(nxt, (it + 2) % kXtBuf);
        }
    }
    if (tid == 0) bulk_wait_read<0>();
}

// Arguments of the z-slab (distributed line) mode of the cluster + TMA kernel.  The periodic line of tab.n points is cut
// across GPUs into slabs; this GPU solves rows [0, n_local).  `tm_planes` views the received stencil halo rows
// [2*HB][n1] (HB rows below the slab, then HB rows above it); glo / ghi hold the reduced-system pieces (gA, gB) of the
// HW = W+1 chunks next to the slab on the lower / upper GPU, laid out [A|B][BW][HW][n1].
struct ZSlabArgs {
    CUtensorMap tm_planes;
    const double* glo;
    const double* ghi;
    long long n1;
    int enabled;
};

// ------------------------------------------------------------------------------------------------
// Strided (y / z) kernel, cluster + TMA ("ctma"): for row strides of megabytes (solve axis outermost) what decides the
// rate is how many bytes every visited row contributes, so the tile must be WIDE (XT = 32 / 64 columns = 256 / 512
// byte row segments) and a whole line of such a tile no longer fits in one CTA.  A thread-block cluster of C CTAs
// shares the line: CTA `rank` owns PC = P/C consecutive chunks (RPC = PC*M rows) of the tile.  Per tile and CTA:
//   TMA in   one box {XT, RPC} plus two halo boxes of HB rows (periodic wrap resolved in the box coordinates)
//   solve    interior sweeps in registers; gA/gB go into the CTA's own extended array and the chunks within W+1 of
//            a CTA edge are PUSHED into the neighbour's halo slots (st.async, counted on the neighbour's transaction
//            barrier), so the separator solve reads local shared memory only; the chunk before a CTA's first one is
//            recomputed from the same window instead of being exchanged a second time.  The extended array is single
//            buffered: a split-phase cluster barrier (arrive after my gather, wait before my next push) keeps a
//            neighbour from overwriting slots I am still reading, and its wait lands after a whole tile's worth of
//            independent work
//   TMA out  one box {XT, RPC}, written in place of the input rows
// with the same three-buffer rotation as the other TMA kernels.
// ------------------------------------------------------------------------------------------------
template <int RK, int BW, int M>
__global__ void __launch_bounds__(512, 1)
chunk_strided_ctma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_halo,
                          const __grid_constant__ CUtensorMap tm_out, int n, int XT, int C, int PC, int tiles_x, long long ntiles,
                          const __grid_constant__ ChunkTables tab, const __grid_constant__ OpParams op,
                          // z-slab (distributed line) mode: this GPU holds rows [0, n) of a longer periodic line; what lies
                          // beyond either end was received from the neighbouring GPUs before the launch
                          const __grid_constant__ ZSlabArgs zs, int nbuf) {
    constexpr int HL = Halo<RK>::L, HR = Halo<RK>::R;
    constexpr int HB = HL > HR ? HL : HR;           // rows of a halo box
    constexpr int BWc = BW > 0 ? BW : 1;
    extern __shared__ __align__(128) double smt[];
    __shared__ __align__(8) unsigned long long full_bar[kXtBuf];
    __shared__ __align__(8) unsigned long long ex_bar[2];   // "the halo of parity b has landed" (pushed by the neighbour CTAs)
    const int RPC = PC * M;
    const size_t tile_elems = (size_t)(RPC + 2 * HB) * XT;
    const int W = tab.W, HW = W + 1;
    const int EC = PC + 2 * HW;                     // extended chunk slots: [HW left halo][PC own][HW right halo]
    const int ESL = EC * XT;
    // 2 x { eA[BWc][EC][XT], eB[BWc][EC][XT] }: ping-pong per tile.  A neighbour CTA can run at most one tile ahead (its next
    // separator solve needs MY next push), so it only ever pushes into the half I am not reading: no cluster barrier, no
    // fence on the per-tile path, and the CTAs of a cluster drift freely within that one-tile window.
    double* ex = smt + nbuf * tile_elems;         // nbuf = 3 rotating tile buffers (2 when the exchange arrays are large: CF90)
    const int tid = threadIdx.x;
    const int xt_shift = __ffs(XT) - 1;
    const int xi = tid & (XT - 1), pl = tid >> xt_shift;
    const unsigned rank = (C > 1) ? cluster_ctarank() : 0u;
    const int p0 = (int)rank * PC;
    const int r0 = p0 * M;                          // first row of this CTA
    const long long ncl = gridDim.x / C;
    const unsigned load_bytes = (unsigned)tile_elems * (unsigned)sizeof(double);
    const bool dist = zs.enabled != 0;
    const bool ext_left = dist && rank == 0, ext_right = dist && rank == (unsigned)C - 1;  // halo from another GPU on that side

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < kXtBuf; ++b) mbar_init(smem_u32(&full_bar[b]), 1);
        mbar_init(smem_u32(&ex_bar[0]), 1);
        mbar_init(smem_u32(&ex_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (BW > 0 && C > 1) cluster_sync_all();        // peers' barriers exist before anyone pushes
    else __syncthreads();

    auto issue_load = [&](long long tile, int b) {  // thread 0 only
        const int k = (int)(tile / tiles_x);
        const int x0 = (int)(tile - (long long)k * tiles_x) * XT;
        const unsigned bar = smem_u32(&full_bar[b]);
        double* buf = smt + b * tile_elems;
        mbar_arrive_expect_tx(bar, load_bytes);
        const int rl = r0 == 0 ? n - HB : r0 - HB;  // periodic wrap: whole halo boxes wrap, never straddle
        const int rr = r0 + RPC == n ? 0 : r0 + RPC;
        if (dist && r0 == 0) tma_load_3d(smem_u32(buf), &zs.tm_planes, x0, 0, 0, bar);        // rows -HB..-1 of the lower GPU
        else tma_load_3d(smem_u32(buf), &tm_halo, x0, rl, k, bar);
        tma_load_3d(smem_u32(buf + (size_t)HB * XT), &tm_in, x0, r0, k, bar);
        if (dist && r0 + RPC == n) tma_load_3d(smem_u32(buf + (size_t)(HB + RPC) * XT), &zs.tm_planes, x0, HB, 0, bar);  // rows n.. of the upper GPU
        else tma_load_3d(smem_u32(buf + (size_t)(HB + RPC) * XT), &tm_halo, x0, rr, k, bar);
    };

    const long long t0 = blockIdx.x / C;
    if (tid == 0) {
        if (t0 < ntiles) issue_load(t0, 0);
        if (t0 + ncl < ntiles) issue_load(t0 + ncl, 1);
    }

    int it = 0;
    for (long long tile = t0; tile < ntiles; tile += ncl, ++it) {
        const int b = it % nbuf;
        double* buf = smt + b * tile_elems;
        // z-slab mode: the edge CTAs' halo pieces come from global memory; the loads are issued here, a whole tile's worth of
        // work before their values are stored into the extended arrays, so their latency is off the cluster's critical path
        double gl_[2 * BWc];
        int gslot = -1;
        if (dist) {
            const bool lo = ext_left && pl < HW, hi = ext_right && pl >= PC - HW;
            if (lo || hi) {
                const long long x = (long long)(tile - (tile / tiles_x) * tiles_x) * XT + xi;
                const int j = lo ? pl : pl - (PC - HW);
                gslot = lo ? pl : HW + PC + j;
                const double* g = lo ? zs.glo : zs.ghi;
#pragma unroll
                for (int c = 0; c < BWc; ++c) {
                    gl_[c] = 0.0; gl_[BWc + c] = 0.0;
                    if (x < zs.n1) {
                        asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(gl_[c]) : "l"(g + ((long long)(0 * BWc + c) * HW + j) * zs.n1 + x));
                        asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(gl_[BWc + c]) : "l"(g + ((long long)(1 * BWc + c) * HW + j) * zs.n1 + x));
                    }
                }
            }
        }
        mbar_wait_or_trap(smem_u32(&full_bar[b]), (unsigned)(it / nbuf) & 1u);

        double v[M + HL + HR];
        {
            const double* c = buf + (size_t)(HB + pl * M - HL) * XT + xi;
#pragma unroll
            for (int j = 0; j < M + HL + HR; ++j) v[j] = c[j * XT];
        }
        double r[M];
#pragma unroll
        for (int i = 0; i < M; ++i) r[i] = rhs_eval<RK>(&v[i + HL], op);

        if constexpr (BW > 0) {
            double a_[2], b_[2];
            chunk_interior<BW, M>(r, tab, a_, b_);
            const int par = it & 1;
            double* eA = ex + par * (2 * BWc * ESL);
            double* eB = eA + BWc * ESL;
            const int me = (HW + pl) * XT + xi;
            eA[me] = a_[0];
            eB[me] = b_[0];
            if (BW == 2) { eA[ESL + me] = a_[1]; eB[ESL + me] = b_[1]; }
            {
                const bool to_right = pl >= PC - HW && !ext_right;  // my tail chunks are the right neighbour's left halo
                const bool to_left = pl < HW && !ext_left;          // my head chunks are the left neighbour's right halo
                const int er = (HW + pl - PC) * XT + xi;
                const int el = (HW + PC + pl) * XT + xi;
                if (gslot >= 0) {  // halo chunks that live on another GPU (values prefetched at the top of the iteration)
#pragma unroll
                    for (int c = 0; c < BW; ++c) {
                        eA[c * ESL + gslot * XT + xi] = gl_[c];
                        eB[c * ESL + gslot * XT + xi] = gl_[BWc + c];
                    }
                }
                if (C > 1) {
                    const unsigned bar = smem_u32(&ex_bar[par]);
                    const unsigned sides_in = (ext_left ? 0u : 1u) + (ext_right ? 0u : 1u);
                    if (tid == 0) mbar_arrive_expect_tx(bar, sides_in * (unsigned)(HW * XT * 2 * BW * sizeof(double)));
                    const unsigned aA = smem_u32(eA), aB = smem_u32(eB);
                    if (to_right) {
                        const unsigned rr = rank + 1 == (unsigned)C ? 0u : rank + 1;
                        const unsigned rb = cluster_map_u32(bar, rr);
                        const unsigned rA = cluster_map_u32(aA + er * 8, rr), rB = cluster_map_u32(aB + er * 8, rr);
                        st_async_f64(rA, a_[0], rb);
                        st_async_f64(rB, b_[0], rb);
                        if (BW == 2) { st_async_f64(rA + ESL * 8, a_[1], rb); st_async_f64(rB + ESL * 8, b_[1], rb); }
                    }
                    if (to_left) {
                        const unsigned rl = rank == 0 ? (unsigned)C - 1 : rank - 1;
                        const unsigned rb = cluster_map_u32(bar, rl);
                        const unsigned rA = cluster_map_u32(aA + el * 8, rl), rB = cluster_map_u32(aB + el * 8, rl);
                        st_async_f64(rA, a_[0], rb);
                        st_async_f64(rB, b_[0], rb);
                        if (BW == 2) { st_async_f64(rA + ESL * 8, a_[1], rb); st_async_f64(rB + ESL * 8, b_[1], rb); }
                    }
                } else if (!dist) {  // the line lives in this CTA: the halo is a periodic copy of my own edge chunks
                    if (to_right) {
                        eA[er] = a_[0]; eB[er] = b_[0];
                        if (BW == 2) { eA[ESL + er] = a_[1]; eB[ESL + er] = b_[1]; }
                    }
                    if (to_left) {
                        eA[el] = a_[0]; eB[el] = b_[0];
                        if (BW == 2) { eA[ESL + el] = a_[1]; eB[ESL + el] = b_[1]; }
                    }
                }
            }
            __syncthreads();
            if (C > 1) mbar_wait_or_trap(smem_u32(&ex_bar[par]), (unsigned)(it >> 1) & 1u);
            // separator solve from local shared memory: s_p = sum_d G[d] (gA_{p+d} + gB_{p+d+1})
            // s of my chunk and of the chunk before it (whose separator values close my left side), both from the same window
            // of 2W + 2 reduced right-hand sides: nothing is exchanged a second time and no barrier sits between the two
            double s0 = 0.0, s1 = 0.0, sp0 = 0.0, sp1 = 0.0;
            {
                int e = (HW + pl - 1 - W) * XT + xi;
                double hp0 = eA[e] + eB[e + XT], hp1 = 0.0;
                if (BW == 2) hp1 = eA[ESL + e] + eB[ESL + e + XT];
                for (int d = 0; d <= 2 * W; ++d) {
                    e += XT;
                    const double h0 = eA[e] + eB[e + XT];
                    if (BW == 2) {
                        const double h1 = eA[ESL + e] + eB[ESL + e + XT];
                        sp0 += tab.G[d][0] * hp0 + tab.G[d][1] * hp1;
                        sp1 += tab.G[d][2] * hp0 + tab.G[d][3] * hp1;
                        s0 += tab.G[d][0] * h0 + tab.G[d][1] * h1;
                        s1 += tab.G[d][2] * h0 + tab.G[d][3] * h1;
                        hp1 = h1;
                    } else {
                        sp0 += tab.G[d][0] * hp0;
                        s0 += tab.G[d][0] * h0;
                    }
                    hp0 = h0;
                }
            }
            chunk_finish<BW, M>(r, tab, s0, s1, sp0, sp1);
        } else {
            __syncthreads();  // everyone has read the tile (neighbours' halo rows overlap my rows)
        }
        {
            double* c = buf + (size_t)(HB + pl * M) * XT + xi;
#pragma unroll
            for (int i = 0; i < M; ++i) { *c = r[i]; c += XT; }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            const int k = (int)(tile / tiles_x);
            const int x0 = (int)(tile - (long long)k * tiles_x) * XT;
            tma_store_3d(&tm_out, x0, r0, k, smem_u32(buf + (size_t)HB * XT));
            bulk_commit();
            // the buffer tile it+2 lands in was drained by the store committed one iteration ago (three buffers) or just now (two)
            if (nbuf == 3) bulk_wait_read<1>(); else bulk_wait_read<0>();
            const long long nxt = tile + 2 * ncl;
            if (nxt < ntiles) issue_load(nxt, (it + 2) % nbuf);
        }
    }
    if (tid == 0) bulk_wait_read<0>();
    // Exit is safe without a barrier: nobody reads a peer's shared memory, and every push aimed at this CTA was awaited by
    // its last wait on ex_bar.
}

// ------------------------------------------------------------------------------------------------
// Generic any-n kernels (tables in global memory).  f(n1, n, n3): es = n1 is the element stride along
// the line.  Pass 1: pointwise RHS.  Pass 2: one thread per line, in place on `out`.
// ------------------------------------------------------------------------------------------------
template <int RK>
__global__ void rhs_generic_kernel(const double* __restrict__ f, double* __restrict__ out, long long n1, int n,
                                   long long n3, long long in_slab, long long out_slab, OpParams op) {
    constexpr int HL = Halo<RK>::L, HR = Halo<RK>::R;
    const long long tot = n1 * n * n3;
    const int nwrap = op.edge_in ? n + 1 : n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx % n1;
        const long long t = idx / n1;
        const int j = (int)(t % n);
        const long long k = t / n;
        const double* base = f + k * in_slab + i;
        double w[HL + HR + 1];
#pragma unroll
        for (int o = -HL; o <= HR; ++o) {
            int q = j + o;
            if (q < 0) q += n;
            else if (q >= nwrap) q -= n;
            w[o + HL] = base[(long long)q * n1];
        }
        out[k * out_slab + (long long)j * n1 + i] = rhs_eval<RK>(&w[HL], op);
    }
}

template <int BW>
__global__ void line_solve_generic_kernel(double* __restrict__ y, long long n1, int n, long long n3, long long slab,
                                          const double* __restrict__ T, int edge_out) {
    const long long nlines = n1 * n3;
    const long long ln = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (ln >= nlines) return;
    const long long i = ln % n1, k = ln / n1;
    double* Y = y + k * slab + i;
    const long long es = n1;
    const int mi = n - BW;
    const double *l1 = T, *l2 = T + n, *ginv = T + 2 * n, *u1 = T + 3 * n, *VU = T + 4 * n, *G0 = T + 6 * (long long)n;
    // forward
    double ym1 = Y[0], ym2 = 0.0;
    for (int r = 1; r < mi; ++r) {
        double v = Y[r * es] - l1[r] * ym1;
        if (BW == 2 && r >= 2) v -= l2[r] * ym2;
        Y[r * es] = v;
        ym2 = ym1;
        ym1 = v;
    }
    // backward
    const double B2 = VU[2 * (long long)n - 1];  // b2 stashed in the last VU slot by the host (see create)
    const double B1 = VU[2 * (long long)n - 2];
    double zp1 = 0.0, zp2 = 0.0;
    for (int r = mi - 1; r >= 0; --r) {
        double v = Y[r * es];
        if (r + 1 < mi) v -= u1[r] * zp1;
        if (BW == 2 && r + 2 < mi) v -= B2 * zp2;
        v *= ginv[r];
        Y[r * es] = v;
        zp2 = zp1;
        zp1 = v;
    }
    // separator values (one chunk: previous == own)
    double s0, s1 = 0.0;
    if (BW == 2) {
        const double z0 = Y[0], z1 = Y[es], za = Y[(mi - 2) * es], zb = Y[(mi - 1) * es];
        const double g0 = Y[(n - 2) * es] - B2 * za - B1 * zb - B2 * z0;
        const double g1 = Y[(n - 1) * es] - B2 * zb - B1 * z0 - B2 * z1;
        s0 = G0[0] * g0 + G0[1] * g1;
        s1 = G0[2] * g0 + G0[3] * g1;
    } else {
        const double g0 = Y[(n - 1) * es] - B1 * Y[(mi - 1) * es] - B1 * Y[0];
        s0 = G0[0] * g0;
    }
    for (int r = 0; r < mi; ++r) {
        double v = Y[r * es] - VU[2 * r] * s0;
        if (BW == 2) v -= VU[2 * r + 1] * s1;
        Y[r * es] = v;
    }
    Y[mi * es] = s0;
    if (BW == 2) Y[(mi + 1) * es] = s1;
    if (edge_out) Y[(long long)n * es] = Y[0];
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
cudaError_t banded_op_create(BandedOp* h, int n, int rk, int bw, double b1, double b2, const OpParams& op) {
    h->n = n; h->rk = rk; h->bw = bw; h->op = op; h->M = 0; h->d_line = nullptr;
    if (n <= 1) return cudaSuccess;
    const int cands[3] = {32, 16, 8};
    for (int c = 0; c < 3; ++c) {
        const int M = cands[c];
        if (n % M != 0) continue;
        const int P = n / M;
        // whole-line kernels hold a line's chunks in one CTA / cluster: at most 128 chunks.  Longer lines keep their
        // 32-point tables for the z-slab (distributed) mode, where only the chunks of the local slab meet in a cluster,
        // and take the any-n kernels when a whole line does arrive on one GPU.
        if (P > 128 && (M != 32 || bw == 0)) continue;
        if (bw == 0) {
            std::memset(&h->tab, 0, sizeof(h->tab));
            h->tab.n = n; h->tab.M = M; h->tab.P = P;
            h->M = M;
            break;
        }
        if (build_chunk_tables(n, M, bw, b1, b2, &h->tab) == 0) { h->M = M; break; }
    }
    h->has_tab16 = 0;
    if (h->M == 32 && n % 16 == 0 && n / 16 <= 512) {
        if (bw == 0) {
            std::memset(&h->tab16, 0, sizeof(h->tab16));
            h->tab16.n = n; h->tab16.M = 16; h->tab16.P = n / 16;
            h->has_tab16 = 1;
        } else if (build_chunk_tables(n, 16, bw, b1, b2, &h->tab16) == 0) {
            h->has_tab16 = 1;
        }
    }
    if (bw > 0) {
        LineTablesHost lt{};
        if (build_line_tables(n, bw, b1, b2, &lt) != 0) return cudaErrorInvalidValue;
        lt.data[6 * (size_t)n - 2] = b1;  // VU rows >= n-BW are unused: stash the off-diagonals there
        lt.data[6 * (size_t)n - 1] = (bw == 2) ? b2 : 0.0;
        cudaError_t e = cudaMalloc(&h->d_line, sizeof(double) * (6 * (size_t)n + 4));
        if (e != cudaSuccess) { std::free(lt.data); return e; }
        e = cudaMemcpy(h->d_line, lt.data, sizeof(double) * (6 * (size_t)n + 4), cudaMemcpyHostToDevice);
        std::free(lt.data);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

void banded_op_destroy(BandedOp* h) {
    if (h->d_line) cudaFree(h->d_line);
    h->d_line = nullptr;
}

namespace {

// PDO_STRIDED_MODE = auto | t512 | t256 | cluster | cluster4 | cpipe | pipe1 selects the strided kernel variant
// (default auto: cpipe for megabyte row strides and lines of more than 32 chunks, the single-CTA pipeline otherwise)
int g_strided_mode = -1;
int strided_mode() {
    if (g_strided_mode < 0) {
        const char* e = std::getenv("PDO_STRIDED_MODE");
        int mode = 0;
        if (e && std::strcmp(e, "t512") == 0) mode = 1;
        if (e && std::strcmp(e, "t256") == 0) mode = 2;
        if (e && std::strcmp(e, "cluster") == 0) mode = 3;
        if (e && std::strcmp(e, "cluster4") == 0) mode = 4;
        if (e && std::strcmp(e, "cpipe") == 0) mode = 5;
        if (e && std::strcmp(e, "pipe1") == 0) mode = 6;
        if (e && std::strcmp(e, "stma") == 0) mode = 7;   // TMA tensor-map pipeline
        if (e && std::strcmp(e, "ctma64") == 0) mode = 8; // cluster + TMA, 64-column tiles
        if (e && std::strcmp(e, "ctma32") == 0) mode = 9; // cluster + TMA, 32-column tiles
        if (e && std::strcmp(e, "ctma32s") == 0) mode = 10; // same, 4 chunks per CTA, two CTAs per SM
        if (e && std::strcmp(e, "cpipe_t") == 0) mode = 11; // cpipe with tensor-map tile loads  // single-CTA pipeline even where cpipe is the default
        g_strided_mode = mode;
    }
    return g_strided_mode;
}

// PDO_X_THREADS = 256 | 128: CTA size of the contiguous-axis kernel; 1000 | 1016: the TMA pipeline (own M | M=16)
int g_x_threads = -1;
int x_threads() {  // 0 = auto
    int& v = g_x_threads;
    if (v < 0) {
        const char* e = std::getenv("PDO_X_THREADS");
        const int a = e ? std::atoi(e) : 0;
        v = e ? ((a == 256 || a == kXTma || a == kXTma16) ? a : 128) : 0;
    }
    return v;
}

// PDO_TUNE=1 lets the FIRST large call of an operator on a shape time the candidates (the round-1 behaviour).  Off by default:
// a user call never synchronises, never breaks stream ordering or graph capture and picks the same kernel on every box; timing
// is an explicit, separate step (banded_op_plan / pdo_*_plan).
int g_tuning = -1;
bool tuning_enabled() {
    int& v = g_tuning;
    if (v < 0) {
        const char* e = std::getenv("PDO_TUNE");
        v = (e && std::atoi(e) == 1) ? 1 : 0;
    }
    return v != 0;
}
// PDO_XTMA=0 keeps the TMA x kernels out of the planner's candidate list
bool xtma_in_planner() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("PDO_XTMA");
        v = (e && std::atoi(e) == 0) ? 0 : 1;
    }
    return v != 0;
}
std::mutex g_plan_mutex;
int g_last_variant = 0;  // what the most recent chunked launch ran: 128/256 (x kernel) or 1 t512, 2 t256, 3 cluster, 5 cpipe, 6 pipe

template <int RK, int BW, int M, int XTH>
cudaError_t launch_x(const BandedOp* h, const double* f, double* out, long long nlines, cudaStream_t st) {
    static bool attr_done = false;
    const int n = h->n, P = n / M;
    const int L = XTH / P > 0 ? XTH / P : 1;
    const size_t smem = sizeof(double) * ((size_t)L * (n + P) + 3 * (BW > 0 ? BW : 1) * (size_t)(XTH + P));
    auto kern = chunk_x_kernel<RK, BW, M, XTH>;
    const size_t cap = (size_t)(200 * 1024) / (512 / XTH);
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    if (smem > cap || P > XTH) return cudaErrorInvalidConfiguration;
    const long long grid = (nlines + L - 1) / L;
    kern<<<(unsigned)grid, XTH, smem, st>>>(f, out, nlines, n, L, h->tab, h->op);
    return cudaGetLastError();
}

// TMA-staged persistent x kernel.  `tab` is the chunk table for this M (the operator's own, or its M=16 alternate).
template <int RK, int BW, int M, int THREADS>
cudaError_t launch_xtma(const BandedOp* h, const ChunkTables& tab, const double* f, double* out, long long nlines, cudaStream_t st) {
    static bool attr_done = false;
    const int n = h->n, P = n / M;
    if (n % M != 0 || P < 1 || P > THREADS) return cudaErrorInvalidConfiguration;
    if (((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return cudaErrorInvalidConfiguration;
    const int L = THREADS / P;
    const size_t smem = sizeof(double) * ((size_t)kXtBuf * L * (n + 2) + 3 * (BW > 0 ? BW : 1) * (size_t)THREADS);
    constexpr size_t cap = 227 * 1024 - 256;  // static mbarriers + padding up to the 128-byte aligned dynamic base
    if (smem > cap) return cudaErrorInvalidConfiguration;
    auto kern = chunk_x_tma_kernel<RK, BW, M, THREADS>;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const long long ntiles = (nlines + L - 1) / L;
    const long long grid = ntiles < 148 ? ntiles : 148;
    kern<<<(unsigned)grid, THREADS, smem, st>>>(f, out, nlines, n, L, ntiles, tab, h->op);
    return cudaGetLastError();
}

// Driver entry point for tensor-map encoding, resolved through the runtime (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// f64 field f(n1, rows, n3) as a 3-D tensor, box {XT, BR, 1}, no swizzle (rows land densely as [row][XT]).
bool encode_field_map(CUtensorMap* tm, const double* base, long long n1, long long rows, long long n3, int XT, int BR) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)n1, (cuuint64_t)rows, (cuuint64_t)n3};
    const cuuint64_t strides[2] = {(cuuint64_t)n1 * sizeof(double), (cuuint64_t)n1 * (cuuint64_t)rows * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)XT, (cuuint32_t)BR, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int kSTma = 7;     // strided-mode code of the TMA pipeline
constexpr int kCpipeT = 11;  // cpipe with tensor-map tile loads

template <int RK, int BW, int M>
cudaError_t launch_stma(const BandedOp* h, const double* f, double* out, long long n1, long long n3, long long in_slab,
                        long long out_slab, cudaStream_t st) {
    static bool attr_done = false;
    const int n = h->n, P = n / M;
    if (n % M != 0 || (n1 & 1) || n1 >= (1LL << 31) || n3 >= (1LL << 31)) return cudaErrorInvalidConfiguration;
    if (((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return cudaErrorInvalidConfiguration;
    const int rows_in = n + (h->op.edge_in ? 1 : 0), rows_out = n + (h->op.edge_out ? 1 : 0);
    const int rows_max = rows_in > rows_out ? rows_in : rows_out;
    constexpr size_t cap = 227 * 1024 - 256;  // static mbarriers + padding up to the 128-byte aligned dynamic base
    // widest power-of-two column count whose three tiles fit: XT * P threads (<= 512), XT >= 8 (64-byte row segments)
    for (int XT = 64; XT >= 8; XT >>= 1) {
        if (XT * P > 512 || XT * P < 64) continue;
        if (XT / 2 >= n1 && XT > 8) continue;
        const int nbox = (rows_max + 255) / 256;
        int BR = (rows_max + nbox - 1) / nbox;
        const int ralign = (128 / (XT * 8)) > 1 ? 128 / (XT * 8) : 1;  // boxes start 128-byte aligned in shared memory
        BR = (BR + ralign - 1) / ralign * ralign;
        if (BR > 256) continue;
        const int nbox_in = (rows_in + BR - 1) / BR, nbox_out = (rows_out + BR - 1) / BR;
        const int nbox_max = nbox_in > nbox_out ? nbox_in : nbox_out;
        const size_t smem = sizeof(double) * ((size_t)kXtBuf * nbox_max * BR * XT + 3 * (BW > 0 ? BW : 1) * (size_t)(XT * P));
        if (smem > cap) continue;
        CUtensorMap tm_in, tm_out;
        if (!encode_field_map(&tm_in, f, n1, in_slab / n1, n3, XT, BR) || !encode_field_map(&tm_out, out, n1, out_slab / n1, n3, XT, BR))
            return cudaErrorInvalidConfiguration;
        auto kern = chunk_strided_tma_kernel<RK, BW, M>;
        if (!attr_done) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
            if (e != cudaSuccess) return e;
            attr_done = true;
        }
        const int tiles_x = (int)((n1 + XT - 1) / XT);
        const long long ntiles = (long long)tiles_x * n3;
        const long long grid = ntiles < 148 ? ntiles : 148;
        g_last_variant = kSTma;
        kern<<<(unsigned)grid, XT * P, smem, st>>>(tm_in, tm_out, n, XT, BR, nbox_in, nbox_out, tiles_x, ntiles, h->tab, h->op);
        return cudaGetLastError();
    }
    return cudaErrorInvalidConfiguration;
}

constexpr int kCTma64 = 8, kCTma32 = 9, kCTma32s = 10;  // strided-mode codes of the cluster + TMA kernel (512- / 256-byte row segments)

// z-slab mode, host side (see ZSlabArgs)
struct ZSlabHost { int n_local; const double* planes; const double* glo; const double* ghi; };

// Edge pass of the z-slab mode: the reduced-system pieces (gA, gB) of this slab's first and last HW chunks, which the
// neighbouring GPUs need before they can close their separator systems.  One thread per (column, edge chunk); writes go
// straight into the neighbours' buffers (peer memory over NVLink): to_lower = the lower GPU's `ghi`, to_upper = the
// upper GPU's `glo`, both [A|B][BW][HW][n1].
template <int RK, int BW, int M>
__global__ void __launch_bounds__(128)
zslab_edge_kernel(const double* __restrict__ f, long long n1, int n_local, const double* __restrict__ planes,
                  double* __restrict__ to_lower, double* __restrict__ to_upper, int HW, const __grid_constant__ ChunkTables tab,
                  const __grid_constant__ OpParams op) {
    constexpr int HL = Halo<RK>::L, HR = Halo<RK>::R;
    constexpr int HB = HL > HR ? HL : HR;
    const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n1) return;
    const int e = blockIdx.y;                       // 0..HW-1: head chunks, HW..2HW-1: tail chunks
    const int P = n_local / M;
    const int c = e < HW ? e : P - 2 * HW + e;      // chunk index inside the slab
    double v[M + HL + HR];
#pragma unroll
    for (int j = 0; j < M + HL + HR; ++j) {
        const int row = c * M - HL + j;
        double val;
        if (row < 0) val = planes[(long long)(HB + row) * n1 + x];                       // rows -HB..-1
        else if (row >= n_local) val = planes[(long long)(HB + row - n_local) * n1 + x];  // rows n_local..
        else val = f[(long long)row * n1 + x];
        v[j] = val;
    }
    double r[M];
#pragma unroll
    for (int i = 0; i < M; ++i) r[i] = rhs_eval<RK>(&v[i + HL], op);
    double a_[2], b_[2];
    chunk_interior<BW, M>(r, tab, a_, b_);
    double* dst = e < HW ? to_lower : to_upper;
    const int j = e < HW ? e : e - HW;
#pragma unroll
    for (int k = 0; k < BW; ++k) {
        dst[((long long)(0 * BW + k) * HW + j) * n1 + x] = a_[k];
        dst[((long long)(1 * BW + k) * HW + j) * n1 + x] = b_[k];
    }
}

// Chunks per CTA of the cluster + TMA kernel for P chunks per line and XT-column tiles: the largest power of two
// <= pc_max that divides P into a cluster of at most 8 CTAs, keeps a CTA between 64 and 512 threads and 256 rows (one
// tensor-map box), holds the halo reach HW, and fits three tile buffers in shared memory.  0: no such configuration.
inline int ctma_chunks_per_cta(int P, int M, int XT, int HB, int HW, int BWc, bool banded, int pc_max, size_t* smem_out,
                               int* nbuf_out = nullptr) {
    constexpr size_t cap = 227 * 1024 - 256;
    for (int pc = pc_max; pc >= 1; pc >>= 1) {
        if (P % pc != 0) continue;
        const int c = P / pc;
        if (c < 1 || c > 8) continue;
        if (XT * pc > 512 || XT * pc < 64 || pc * M > 256 || (banded && HW > pc)) continue;
        for (int nbuf = kXtBuf; nbuf >= 2; --nbuf) {   // three rotating tile buffers; two when the ping-pong exchange arrays are large
            const size_t need = sizeof(double) * ((size_t)nbuf * (pc * M + 2 * HB) * XT + 4 * BWc * (size_t)(pc + 2 * HW) * XT);
            if (need > cap) continue;
            if (smem_out) *smem_out = need;
            if (nbuf_out) *nbuf_out = nbuf;
            return pc;
        }
    }
    return 0;
}

template <int RK, int BW, int M>
cudaError_t launch_ctma(const BandedOp* h, const double* f, double* out, long long n1, long long n3, long long in_slab,
                        long long out_slab, cudaStream_t st, int XT, int pc_max = 16, const ZSlabHost* z = nullptr) {
    const int n = z ? z->n_local : h->n, P = n / M;   // z-slab mode: rows held here; the tables stay those of the whole line
    constexpr int BWc = BW > 0 ? BW : 1;
    if (n % M != 0 || (n1 & 1) || n1 >= (1LL << 31) || n3 >= (1LL << 31) || n1 < XT / 2) return cudaErrorInvalidConfiguration;
    if (h->op.edge_in || h->op.edge_out || in_slab != n1 * n || out_slab != n1 * n) return cudaErrorInvalidConfiguration;
    if (BW > 0 && (h->tab.dense || 2 * h->tab.W + 2 > P)) return cudaErrorInvalidConfiguration;
    if (z && (BW == 0 || n3 != 1)) return cudaErrorInvalidConfiguration;
    constexpr int HB = Halo<RK>::L > Halo<RK>::R ? Halo<RK>::L : Halo<RK>::R;
    const int HW = (BW > 0 ? h->tab.W : 0) + 1;
    if (((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return cudaErrorInvalidConfiguration;
    constexpr size_t cap = 227 * 1024 - 256;
    size_t smem = 0;
    int nbuf = kXtBuf;
    const int PC = ctma_chunks_per_cta(P, M, XT, HB, HW, BWc, BW > 0, pc_max, &smem, &nbuf);
    if (PC == 0) return cudaErrorInvalidConfiguration;
    const int C = P / PC;
    CUtensorMap tm_in, tm_halo, tm_out;
    if (!encode_field_map(&tm_in, f, n1, n, n3, XT, PC * M) || !encode_field_map(&tm_halo, f, n1, n, n3, XT, HB) ||
        !encode_field_map(&tm_out, out, n1, n, n3, XT, PC * M))
        return cudaErrorInvalidConfiguration;
    auto kern = chunk_strided_ctma_kernel<RK, BW, M>;
    static bool attr_done = false;
    static int max_clusters[9][3] = {};
    ZSlabArgs zs;
    std::memset(&zs, 0, sizeof(zs));
    if (z) {
        if (!encode_field_map(&zs.tm_planes, z->planes, n1, 2 * HB, 1, XT, HB)) return cudaErrorInvalidConfiguration;
        zs.glo = z->glo; zs.ghi = z->ghi; zs.n1 = n1; zs.enabled = 1;
    }
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const int tiles_x = (int)((n1 + XT - 1) / XT);
    const long long ntiles = (long long)tiles_x * n3;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(XT * PC);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int& mc = max_clusters[C][pc_max < 16 ? 2 : (XT == 64 ? 1 : 0)];
    if (mc == 0) {
        cfg.gridDim = dim3((unsigned)(148 / C * C));
        int nc = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
        if (e != cudaSuccess) return e;
        mc = nc > 0 ? nc : 1;
    }
    const long long ncl = ntiles < mc ? ntiles : mc;
    cfg.gridDim = dim3((unsigned)(ncl * C));
    g_last_variant = pc_max < 16 ? kCTma32s : (XT == 64 ? kCTma64 : kCTma32);
    return cudaLaunchKernelEx(&cfg, kern, tm_in, tm_halo, tm_out, n, XT, C, PC, tiles_x, ntiles, h->tab, h->op, zs, nbuf);
}

template <int RK, int BW, int M>
cudaError_t launch_chunk(const BandedOp* h, int axis, const double* f, double* out, long long n1, long long n3,
                         long long in_slab, long long out_slab, cudaStream_t st, int mode, int xth) {
    const int n = h->n, P = n / M;
    if (axis == 0) {
        if (xth == kXTma) {  // TMA pipeline on the operator's own chunk length
            if constexpr (M == 32) { g_last_variant = kXTma; return launch_xtma<RK, BW, 32, 256>(h, h->tab, f, out, n3, st); }
            else if constexpr (M == 16) { g_last_variant = kXTma; return launch_xtma<RK, BW, 16, 512>(h, h->tab, f, out, n3, st); }
            else return cudaErrorInvalidConfiguration;
        }
        if (xth == kXTma16) {  // TMA pipeline on 16-point chunks (twice the threads per tile byte)
            if (M == 32 && h->has_tab16) { g_last_variant = kXTma16; return launch_xtma<RK, BW, 16, 512>(h, h->tab16, f, out, n3, st); }
            return cudaErrorInvalidConfiguration;
        }
        if (xth == 0 && xtma_in_planner() && n3 >= 148 * 8) {  // heuristic path (small problems, captured streams): TMA pipeline if it fits
            cudaError_t e = cudaErrorInvalidConfiguration;
            if constexpr (M == 32) e = launch_xtma<RK, BW, 32, 256>(h, h->tab, f, out, n3, st);
            else if constexpr (M == 16) e = launch_xtma<RK, BW, 16, 512>(h, h->tab, f, out, n3, st);
            if (e == cudaSuccess) { g_last_variant = kXTma; return e; }
            cudaGetLastError();
        }
        if (xth != 256 && P <= 128) { g_last_variant = 128; return launch_x<RK, BW, M, 128>(h, f, out, n3, st); }
        g_last_variant = 256;
        return launch_x<RK, BW, M, 256>(h, f, out, n3, st);
    }
    if (mode == kSTma) return launch_stma<RK, BW, M>(h, f, out, n1, n3, in_slab, out_slab, st);
    if (mode == kCTma64) return launch_ctma<RK, BW, M>(h, f, out, n1, n3, in_slab, out_slab, st, 64);
    if (mode == kCTma32) return launch_ctma<RK, BW, M>(h, f, out, n1, n3, in_slab, out_slab, st, 32);
    if (mode == kCTma32s) return launch_ctma<RK, BW, M>(h, f, out, n1, n3, in_slab, out_slab, st, 32, 4);  // two CTAs per SM
    if constexpr (M == 32) {
        // cluster kernel: P chunks split over C = P/PC CTAs (portable cluster sizes only); mode 3: PC=8, mode 4: PC=4
        const int PC = (mode == 4) ? 4 : 8;
        const int C = P / PC, XTc = kClThreads / PC;
        // explicit stencils (Gaussian) have no compute phase worth pipelining: the plain streaming form of this kernel
        // (no shared staging, two CTAs per SM) runs at the copy rate, so it is their default
        const bool stencil_default = (BW == 0) && (mode == 0) && n1 >= XTc;
        if ((mode == 3 || mode == 4 || stencil_default) && P % PC == 0 && (BW == 0 || C == 1 || C == 2 || C == 4 || C == 8) &&
            n1 >= XTc / 2 && (BW == 0 || h->tab.W + 1 <= kClMaxHW)) {
            const int tiles_x = (int)((n1 + XTc - 1) / XTc);
            const long long nblocks = (long long)tiles_x * n3 * C;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)nblocks);
            cfg.blockDim = dim3(kClThreads);
            cfg.dynamicSmemBytes = 0;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)(BW > 0 ? C : 1);
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            g_last_variant = 3;
            if (PC == 8)
                return cudaLaunchKernelEx(&cfg, chunk_strided_cluster_kernel<RK, BW, M, 8>, f, out, n1, n, in_slab, out_slab, tiles_x,
                                          C, h->tab, h->op);
            return cudaLaunchKernelEx(&cfg, chunk_strided_cluster_kernel<RK, BW, M, 4>, f, out, n1, n, in_slab, out_slab, tiles_x, C,
                                      h->tab, h->op);
        }
    }
    if constexpr (M == 32) {
        // cpipe: persistent clusters of C = P/16 CTAs, 256-byte row segments.  Default for the outermost axis
        // (row stride n1 of megabytes) and for lines the single-CTA pipeline can only tile 8 columns wide.
        const int C = P / kCpPC;
        const bool fits = (P % kCpPC == 0) && (C == 1 || C == 2 || C == 4 || C == 8) && n1 >= kCpXT &&
                          (BW == 0 || (!h->tab.dense && h->tab.W + 1 <= kCpMaxHW));
        const bool want = (mode == 5) || (mode == kCpipeT) || (mode == 0 && (n1 * (long long)sizeof(double) >= (1 << 20) || P > 32));
        if (mode == kCpipeT && !fits) return cudaErrorInvalidConfiguration;
        if (fits && want) {
            constexpr int HLR = Halo<RK>::L + Halo<RK>::R;
            constexpr int BWc = (BW > 0 ? BW : 1);
            const size_t smem = sizeof(double) * ((size_t)(kCpPC * M + HLR) * kCpXT + 4 * BWc * (size_t)(kCpPC + 2 * kCpMaxHW) * kCpXT +
                                                  BWc * (size_t)kCpPC * kCpXT);
            auto kern = chunk_strided_cpipe_kernel<RK, BW, M>;
            static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            const int tiles_x = (int)((n1 + kCpXT - 1) / kCpXT);
            const long long ntiles = (long long)tiles_x * n3;
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(kCpPC * kCpXT);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)C;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (max_clusters[C] == 0) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                cfg.gridDim = dim3((unsigned)(148 / C * C));
                int nc = 0;
                e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
                if (e != cudaSuccess) return e;
                max_clusters[C] = nc > 0 ? nc : 1;
            }
            g_last_variant = 5;
            const long long ncl = ntiles < max_clusters[C] ? ntiles : max_clusters[C];
            cfg.gridDim = dim3((unsigned)(ncl * C));
            const int vec16 = (n1 % 2 == 0) && (in_slab % 2 == 0) && ((reinterpret_cast<uintptr_t>(f) & 15) == 0);
            CUtensorMap tm_in, tm_halo;
            std::memset(&tm_in, 0, sizeof(tm_in));
            std::memset(&tm_halo, 0, sizeof(tm_halo));
            int tma_in = 0;
            if (mode == kCpipeT) {  // tile loads through tensor maps instead of per-thread cp.async
                if (Halo<RK>::L != Halo<RK>::R || h->op.edge_in || !vec16 || n1 >= (1LL << 31) || n3 >= (1LL << 31) ||
                    !encode_field_map(&tm_in, f, n1, in_slab / n1, n3, kCpXT, kCpBoxRows) || (kCpPC * M) % kCpBoxRows != 0 ||
                    !encode_field_map(&tm_halo, f, n1, in_slab / n1, n3, kCpXT, Halo<RK>::L))
                    return cudaErrorInvalidConfiguration;
                tma_in = 1;
                g_last_variant = kCpipeT;
            }
            return cudaLaunchKernelEx(&cfg, kern, f, out, n1, n, in_slab, out_slab, tiles_x, ntiles, C, vec16, tma_in, tm_in, tm_halo,
                                      h->tab, h->op);
        }
    }
    const int threads = (mode == 2) ? 256 : 512;
    int XT = 1;
    while (XT * 2 * P <= threads) XT *= 2;
    while (XT > 1 && XT / 2 >= n1) XT /= 2;
    if (XT * P > threads) return cudaErrorInvalidConfiguration;
    const int tiles_x = (int)((n1 + XT - 1) / XT);
    const long long ntiles = (long long)tiles_x * n3;
    const size_t smem_g = sizeof(double) * 3 * (BW > 0 ? BW : 1) * (size_t)P * XT;
    const int rows_in = n + (h->op.edge_in ? 1 : 0);
    const size_t smem_pipe = sizeof(double) * (((size_t)rows_in * XT + 1) & ~(size_t)1) + smem_g;
    if ((mode == 0 || mode == 6) && smem_pipe <= 200 * 1024 && ntiles >= 148) {
        static bool attr_done = false;
        auto kern = chunk_strided_pipe_kernel<RK, BW, M, 512>;
        if (!attr_done) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return e;
            attr_done = true;
        }
        const int vec16 = (XT % 2 == 0) && (n1 % 2 == 0) && (in_slab % 2 == 0) && ((reinterpret_cast<uintptr_t>(f) & 15) == 0);
        const long long grid = ntiles < 148 ? ntiles : 148;
        g_last_variant = 6;
        kern<<<(unsigned)grid, XT * P, smem_pipe, st>>>(f, out, n1, n, in_slab, out_slab, tiles_x, XT, ntiles, vec16, h->tab, h->op);
    } else if (threads == 256) {
        g_last_variant = 2;
        chunk_strided_kernel<RK, BW, M, 256><<<(unsigned)ntiles, XT * P, smem_g, st>>>(f, out, n1, n, in_slab, out_slab, tiles_x,
                                                                                       XT, h->tab, h->op);
    } else {
        g_last_variant = 1;
        chunk_strided_kernel<RK, BW, M, 512><<<(unsigned)ntiles, XT * P, smem_g, st>>>(f, out, n1, n, in_slab, out_slab, tiles_x,
                                                                                       XT, h->tab, h->op);
    }
    return cudaGetLastError();
}

// Deterministic dispatch: the variant a shape gets when no plan has been made for it.  Measured on B200 at the BASELINE shapes
// (profiles/r02*_vsweep*.jsonl); every entry falls back to the shape heuristic (mode 0) when the variant does not cover the shape.
//   x axis                       TMA bulk-copy pipeline
//   y axis (rows KBs apart)      banded: single-CTA pipelines while a whole line fits one CTA 16 columns wide (pipe1; the TMA one at
//                                <= 16 chunks), cpipe with tensor-map loads beyond; CD06: cluster + TMA; stencils: streaming kernel
//   z axis (rows MBs apart)      256-byte row segments always: cpipe (cpipe_t for lines of 64 chunks), CD06: cluster + TMA
template <int RK, int BW, int M>
int default_variant(const BandedOp* h, int axis, long long n1, long long n3) {
    (void)n3;
    if (axis == 0) return kXTma;
    const int P = h->n / M;
    if (BW == 0) return 0;                                   // explicit stencils: the heuristic's streaming kernel
    const bool far_rows = n1 * (long long)sizeof(double) >= (1 << 20);
    if (M != 32) return 0;
    if (BW == 1 && RK == RK_D1_5) return (n1 % 2 == 0 && P >= 8) ? kCTma32 : 0;
    if (BW == 1) return 0;                                   // staggered operators (edge planes): heuristic
    if (!far_rows) {
        if (P <= 16) return (n1 % 2 == 0) ? kSTma : 6;
        if (P <= 32) return (RK == RK_SYM_9) ? kCpipeT : 6;  // CF90's 9-point stencil + 13-block gather: the cluster split wins
        return kCpipeT;
    }
    if (P <= 16) return (RK == RK_SYM_9) ? kCpipeT : 6;
    if (P <= 32) return (RK == RK_SYM_9) ? kCpipeT : 5;
    return kCpipeT;
}

template <int RK, int BW, int M>
cudaError_t launch_default(const BandedOp* h, int axis, const double* f, double* out, long long n1, long long n3, long long in_slab,
                           long long out_slab, cudaStream_t st) {
    const int v = default_variant<RK, BW, M>(h, axis, n1, n3);
    if (v != 0) {
        const cudaError_t e = launch_chunk<RK, BW, M>(h, axis, f, out, n1, n3, in_slab, out_slab, st, v, v);
        if (e != cudaErrorInvalidConfiguration) return e;
        cudaGetLastError();                                  // the table's pick does not cover this shape (odd n1, tiny extents, ...)
    }
    return launch_chunk<RK, BW, M>(h, axis, f, out, n1, n3, in_slab, out_slab, st, 0, 0);
}

// Timing-based plan (explicit: banded_op_plan, or the first call when PDO_TUNE=1): like FFTW's planner and 2DECOMP's
// best_2d_grid (both timing-based in the reference) it times the candidates on real arrays of the shape and remembers the
// winner in the handle.  The candidates agree to rounding (same per-chunk algebra), so the choice does not change the answer
// beyond 1e-15.
template <int RK, int BW, int M>
cudaError_t plan_by_timing(const BandedOp* h, int axis, const double* f, double* out, long long n1, long long n3, long long in_slab,
                           long long out_slab, cudaStream_t st) {
    const int cand_x[3] = {128, 256, kXTma};
    const int cand_s[8] = {6, 5, 3, 1, kSTma, kCTma32s, kCTma32, kCpipeT};
    const int* cand = axis == 0 ? cand_x : cand_s;
    const int ncand = axis == 0 ? (xtma_in_planner() ? 3 : 2) : (xtma_in_planner() ? 8 : 4);
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
        cudaGetLastError();
        return launch_default<RK, BW, M>(h, axis, f, out, n1, n3, in_slab, out_slab, st);
    }
    int best = 0;
    float best_ms = 1e30f;
    for (int c = 0; c < ncand; ++c) {
        float ms_c = 1e30f;
        bool ok = true;
        for (int rep = 0; rep < 3 && ok; ++rep) {  // rep 0 warms up (function attributes, instruction cache)
            cudaEventRecord(e0, st);
            ok = launch_chunk<RK, BW, M>(h, axis, f, out, n1, n3, in_slab, out_slab, st, cand[c], cand[c]) == cudaSuccess;
            cudaEventRecord(e1, st);
            if (cudaEventSynchronize(e1) != cudaSuccess) ok = false;
            float ms = 0.f;
            if (ok && rep > 0 && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && ms < ms_c) ms_c = ms;
        }
        if (!ok) { cudaGetLastError(); continue; }
        if (ms_c < best_ms) { best_ms = ms_c; best = cand[c]; }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    {
        std::lock_guard<std::mutex> lk(g_plan_mutex);
        int slot = -1;
        for (int i = 0; i < h->nplans; ++i)
            if (h->plans[i].axis == axis && h->plans[i].n1 == n1 && h->plans[i].n3 == n3) slot = i;
        if (slot < 0 && h->nplans < BandedOp::kMaxPlans) slot = h->nplans++;
        if (slot >= 0) { h->plans[slot].axis = axis; h->plans[slot].n1 = n1; h->plans[slot].n3 = n3; h->plans[slot].choice = best; }
    }
    if (std::getenv("PDO_TUNE_VERBOSE"))
        std::fprintf(stderr, "[padeops_b200] plan rk=%d bw=%d n=%d axis=%d n1=%lld n3=%lld -> variant %d (%.3f ms)\n", RK, BW, h->n, axis,
                     n1, n3, best, best_ms);
    g_last_variant = best;
    return cudaSuccess;  // the timed runs already produced `out`
}

template <int RK, int BW, int M>
cudaError_t launch_planned(const BandedOp* h, int axis, const double* f, double* out, long long n1, long long n3,
                           long long in_slab, long long out_slab, cudaStream_t st, bool make_plan = false) {
    const int forced_mode = strided_mode(), forced_x = x_threads();
    const bool forced = (axis == 0) ? forced_x != 0 : forced_mode != 0;
    if (make_plan) return plan_by_timing<RK, BW, M>(h, axis, f, out, n1, n3, in_slab, out_slab, st);
    if (forced) return launch_chunk<RK, BW, M>(h, axis, f, out, n1, n3, in_slab, out_slab, st, forced_mode, forced_x);
    const long long pts = n1 * h->n * n3;
    if (pts < (1LL << 24)) return launch_chunk<RK, BW, M>(h, axis, f, out, n1, n3, in_slab, out_slab, st, 0, 0);
    {
        std::lock_guard<std::mutex> lk(g_plan_mutex);
        for (int i = 0; i < h->nplans; ++i)
            if (h->plans[i].axis == axis && h->plans[i].n1 == n1 && h->plans[i].n3 == n3)
                return launch_chunk<RK, BW, M>(h, axis, f, out, n1, n3, in_slab, out_slab, st, h->plans[i].choice, h->plans[i].choice);
    }
    if (tuning_enabled()) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone)
            return plan_by_timing<RK, BW, M>(h, axis, f, out, n1, n3, in_slab, out_slab, st);
        cudaGetLastError();
    }
    return launch_default<RK, BW, M>(h, axis, f, out, n1, n3, in_slab, out_slab, st);
}

template <int RK, int BW>
cudaError_t launch_any(const BandedOp* h, int axis, const double* f, double* out, long long n1, long long n3,
                       long long in_slab, long long out_slab, cudaStream_t st, int force_generic) {
    const bool whole_line_chunked = h->M > 0 && h->n / h->M <= 128;
    const bool mk = force_generic == 2;     // 2: time the candidates and store the plan (banded_op_plan)
    if (mk) force_generic = 0;
    if (h->M == 32 && whole_line_chunked && !force_generic) return launch_planned<RK, BW, 32>(h, axis, f, out, n1, n3, in_slab, out_slab, st, mk);
    if (h->M == 16 && !force_generic) return launch_planned<RK, BW, 16>(h, axis, f, out, n1, n3, in_slab, out_slab, st, mk);
    if (h->M == 8 && !force_generic) return launch_planned<RK, BW, 8>(h, axis, f, out, n1, n3, in_slab, out_slab, st, mk);
    // generic
    const int n = h->n;
    const long long tot = n1 * n * n3;
    const int thr = 256;
    long long blocks = (tot + thr - 1) / thr;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    rhs_generic_kernel<RK><<<(unsigned)blocks, thr, 0, st>>>(f, out, n1, n, n3, in_slab, out_slab, h->op);
    if constexpr (BW > 0) {
        const long long nlines = n1 * n3;
        line_solve_generic_kernel<BW><<<(unsigned)((nlines + 127) / 128), 128, 0, st>>>(
            out, n1, n, n3, out_slab, h->d_line, h->op.edge_out);
    }
    return cudaGetLastError();
}

}  // namespace

int banded_debug_last_variant() { return g_last_variant; }
void banded_set_tuning(bool on) { g_tuning = on ? 1 : 0; }

// host-only view of the cluster + TMA kernel's launch shape (tests check its invariants without a GPU)
int banded_debug_ctma_config(int P, int XT, int HB, int HW, int BW, int pc_max, long long* smem_bytes) {
    size_t smem = 0;
    const int pc = ctma_chunks_per_cta(P, 32, XT, HB, HW, BW > 0 ? BW : 1, BW > 0, pc_max, &smem);
    if (smem_bytes) *smem_bytes = (long long)smem;
    return pc;
}

void banded_debug_set_variant(int strided_mode, int x_threads) {
    g_strided_mode = strided_mode;
    g_x_threads = x_threads;
}

// ---- z-slab (distributed line) entry points: see banded.cuh ----
namespace {
template <int RK, int BW>
cudaError_t zslab_edges_t(const BandedOp* h, const double* f, long long n1, int n_local, const double* planes, double* to_lower,
                          double* to_upper, cudaStream_t st) {
    const int HW = h->tab.W + 1;
    dim3 grid((unsigned)((n1 + 127) / 128), (unsigned)(2 * HW));
    zslab_edge_kernel<RK, BW, 32><<<grid, 128, 0, st>>>(f, n1, n_local, planes, to_lower, to_upper, HW, h->tab, h->op);
    return cudaGetLastError();
}
template <int RK, int BW>
cudaError_t zslab_apply_t(const BandedOp* h, const double* f, double* out, long long n1, const ZSlabHost& z, cudaStream_t st) {
    const long long slab = n1 * z.n_local;
    cudaError_t e = launch_ctma<RK, BW, 32>(h, f, out, n1, 1, slab, slab, st, 32, 16, &z);  // 8 chunks x 32 columns per CTA
    if (e == cudaErrorInvalidConfiguration) {
        cudaGetLastError();
        e = launch_ctma<RK, BW, 32>(h, f, out, n1, 1, slab, slab, st, 32, 4, &z);
    }
    return e;
}
}  // namespace

int banded_zslab_halo_rows(const BandedOp* h) {
    switch (h->rk) {
        case RK_D1_7: case RK_D2_7: return 3;
        case RK_D1_5: case RK_D2_5: return 2;
        case RK_SYM_9: return 4;
        default: return -1;
    }
}
int banded_zslab_halo_chunks(const BandedOp* h, int n_local) {
    if (h->M != 32 || h->bw == 0 || h->tab.dense || n_local % 32 != 0 || h->op.edge_in || h->op.edge_out) return -1;
    const int HW = h->tab.W + 1, HB = banded_zslab_halo_rows(h);
    if (n_local / 32 < 2 * HW || HB < 0) return -1;
    if (!ctma_chunks_per_cta(n_local / 32, 32, 32, HB, HW, h->bw, true, 4, nullptr) &&
        !ctma_chunks_per_cta(n_local / 32, 32, 32, HB, HW, h->bw, true, 16, nullptr))
        return -1;  // no cluster shape covers this slab
    return HW;
}
cudaError_t banded_zslab_edges(const BandedOp* h, const double* f, long long n1, int n_local, const double* planes, double* to_lower,
                               double* to_upper, cudaStream_t st) {
    if (banded_zslab_halo_chunks(h, n_local) < 0) return cudaErrorInvalidConfiguration;
    switch (h->rk * 10 + h->bw) {
        case RK_D1_7 * 10 + 2: return zslab_edges_t<RK_D1_7, 2>(h, f, n1, n_local, planes, to_lower, to_upper, st);
        case RK_D2_7 * 10 + 2: return zslab_edges_t<RK_D2_7, 2>(h, f, n1, n_local, planes, to_lower, to_upper, st);
        case RK_D1_5 * 10 + 1: return zslab_edges_t<RK_D1_5, 1>(h, f, n1, n_local, planes, to_lower, to_upper, st);
        case RK_SYM_9 * 10 + 2: return zslab_edges_t<RK_SYM_9, 2>(h, f, n1, n_local, planes, to_lower, to_upper, st);
        default: return cudaErrorInvalidValue;
    }
}
cudaError_t banded_zslab_apply(const BandedOp* h, const double* f, double* out, long long n1, int n_local, const double* planes,
                               const double* glo, const double* ghi, cudaStream_t st) {
    if (banded_zslab_halo_chunks(h, n_local) < 0) return cudaErrorInvalidConfiguration;
    ZSlabHost z{n_local, planes, glo, ghi};
    cudaError_t e;
    switch (h->rk * 10 + h->bw) {
        case RK_D1_7 * 10 + 2: e = zslab_apply_t<RK_D1_7, 2>(h, f, out, n1, z, st); break;
        case RK_D2_7 * 10 + 2: e = zslab_apply_t<RK_D2_7, 2>(h, f, out, n1, z, st); break;
        case RK_D1_5 * 10 + 1: e = zslab_apply_t<RK_D1_5, 1>(h, f, out, n1, z, st); break;
        case RK_SYM_9 * 10 + 2: e = zslab_apply_t<RK_SYM_9, 2>(h, f, out, n1, z, st); break;
        default: e = cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) cudaGetLastError();
    return e;
}

cudaError_t banded_dispatch(const BandedOp* h, int axis, const double* f, double* out, long long n1, long long n3,
                            long long in_slab, long long out_slab, cudaStream_t st, int force_generic);

cudaError_t banded_op_apply(const BandedOp* h, int axis, const double* f, double* out, long long na, long long nb,
                            cudaStream_t st, int force_generic) {
    const int n = h->n;
    long long n1, n3;
    if (axis == 0) { n1 = 1; n3 = na * nb; }
    else if (axis == 1) { n1 = na; n3 = nb; }
    else if (axis == 2) { n1 = na * nb; n3 = 1; }
    else return cudaErrorInvalidValue;
    if (n1 * n3 == 0) return cudaSuccess;
    const long long in_slab = n1 * (n + (h->op.edge_in || (h->op.edge_out && h->rk == RK_D2_5) ? 1 : 0));
    const long long out_slab = n1 * (n + (h->op.edge_out ? 1 : 0));
    const cudaError_t e = banded_dispatch(h, axis, f, out, n1, n3, in_slab, out_slab, st, force_generic);
    if (e != cudaSuccess) cudaGetLastError();  // do not leave a stale error for the next launch check to trip over
    return e;
}

// Explicit planning: times the kernel candidates for this operator along `axis` of a pencil with extents (na, nb) on scratch
// arrays of that shape and stores the winner in the handle; later calls on the same shape use it.  Synchronises the device.
cudaError_t banded_op_plan(const BandedOp* h, int axis, long long na, long long nb, int* chosen) {
    const int n = h->n;
    if (chosen) *chosen = 0;
    if (n <= 1 || na * nb <= 0 || axis < 0 || axis > 2) return cudaSuccess;
    const size_t planes_in = (size_t)n + ((h->op.edge_in || (h->op.edge_out && h->rk == RK_D2_5)) ? 1 : 0);
    const size_t planes_out = (size_t)n + (h->op.edge_out ? 1 : 0);
    double *f = nullptr, *o = nullptr;
    cudaError_t e = cudaMalloc(&f, sizeof(double) * planes_in * (size_t)na * (size_t)nb);
    if (e == cudaSuccess) e = cudaMalloc(&o, sizeof(double) * planes_out * (size_t)na * (size_t)nb);
    if (e == cudaSuccess) e = cudaMemset(f, 0, sizeof(double) * planes_in * (size_t)na * (size_t)nb);
    if (e == cudaSuccess) e = banded_op_apply(h, axis, f, o, na, nb, nullptr, 2);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && chosen) *chosen = g_last_variant;
    if (f) cudaFree(f);
    if (o) cudaFree(o);
    if (e != cudaSuccess) cudaGetLastError();
    return e;
}

cudaError_t banded_dispatch(const BandedOp* h, int axis, const double* f, double* out, long long n1, long long n3,
                            long long in_slab, long long out_slab, cudaStream_t st, int force_generic) {
    const int key = h->rk * 10 + h->bw;
    switch (key) {
        case RK_D1_7 * 10 + 2: return launch_any<RK_D1_7, 2>(h, axis, f, out, n1, n3, in_slab, out_slab, st, force_generic);
        case RK_D2_7 * 10 + 2: return launch_any<RK_D2_7, 2>(h, axis, f, out, n1, n3, in_slab, out_slab, st, force_generic);
        case RK_D1_5 * 10 + 1: return launch_any<RK_D1_5, 1>(h, axis, f, out, n1, n3, in_slab, out_slab, st, force_generic);
        case RK_SYM_9 * 10 + 2: return launch_any<RK_SYM_9, 2>(h, axis, f, out, n1, n3, in_slab, out_slab, st, force_generic);
        case RK_SYM_9 * 10 + 0: return launch_any<RK_SYM_9, 0>(h, axis, f, out, n1, n3, in_slab, out_slab, st, force_generic);
        case RK_D2_5 * 10 + 1: return launch_any<RK_D2_5, 1>(h, axis, f, out, n1, n3, in_slab, out_slab, st, force_generic);
        case RK_STAG_E2C * 10 + 1: return launch_any<RK_STAG_E2C, 1>(h, axis, f, out, n1, n3, in_slab, out_slab, st, force_generic);
        case RK_STAG_C2E * 10 + 1: return launch_any<RK_STAG_C2E, 1>(h, axis, f, out, n1, n3, in_slab, out_slab, st, force_generic);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace pdo
