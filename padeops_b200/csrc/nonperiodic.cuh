// nonperiodic.cuh — non-periodic closures of the compact operators: CD10 first / second
// derivative and the CF90 filter with boundary codes (bc1, bcn) in {0, 1, -1}^2.
//
// Reference: derivatives/cd10.F90:29-96 (boundary schemes), 429-707 (ComputePenta1/2), 823-851 (SolveXPenta1),
// 1143-1262 / 1636-1731 (non-periodic RHS); filters/cf90.F90:24-47, 276-418, 532-558, 672-801.
//
// First delivery: correctness path.  One pointwise RHS kernel (coalesced) + one thread per line for the pentadiagonal LU
// sweeps, tables in global memory — the shape of the any-n periodic kernels (≈48 B/pt), not the chunked fast path.
// The per-point RHS and the per-line sweeps are __host__ __device__ functions so that CPU tests can execute exactly the
// code the kernels execute (pdo_debug_np_line_host; not reachable from any public entry point — the API needs a GPU).
//
// bc = 0: one-sided rows at that end; bc = +1 / -1: the interior row applied to the even / odd reflection of f about the
// boundary node, which reproduces the reference's written-out rows operand for operand (f(2)+f(2), -f(n-1)-f(n-5), ...).
#pragma once
#include <cuda_runtime.h>

namespace pdo {

enum NpKind { NP_CD10_D1 = 0, NP_CD10_D2 = 1, NP_CF90 = 2, NP_CD06_D1 = 3, NP_GAUSS = 4, NP_LSTSQ = 5 };   // CD06: one-sided closure only (cd06.F90:264-327); Gaussian / least-squares: explicit, no solve (gaussian.F90:22-46, 215-330; lstsq.F90:14-46, 169-212)

struct NpCoefs {
    double in[5];     // interior stencil: D1/D2: a, b, c (grid spacing folded in); CF90: a, b, c, d, e
    double r1[5];     // one-sided row 1   (D1: p q r s scaled by w1/dx; D2: five coefficients; CF90: {1})
    double r2[2];     // row 2
    double r3[3];     // row 3
    double r4[4];     // row 4 (D1, CF90)
};

#if defined(__CUDACC__)
#define PDO_HD __host__ __device__ __forceinline__
#else
#define PDO_HD inline
#endif

// F(j): value of the line at 1-based index j in [1, n]
template <int KIND, class Acc>
PDO_HD double np_rhs_point(int i, int n, int bc1, int bcn, const NpCoefs& c, Acc F) {
    constexpr int NB = (KIND == NP_CD10_D2 || KIND == NP_CD06_D1) ? 3 : 4;   // rows owned by the one-sided closure at each end
    if (bc1 == 0 && i <= NB) {
        if (KIND == NP_CD10_D1) {
            if (i == 1) return c.r1[0] * F(1) + c.r1[1] * F(2) + c.r1[2] * F(3) + c.r1[3] * F(4);
            if (i == 2) return c.r2[0] * (F(3) - F(1));
            if (i == 3) return c.r3[0] * (F(4) - F(2)) + c.r3[1] * (F(5) - F(1));
            return c.r4[0] * (F(5) - F(3)) + c.r4[1] * (F(6) - F(2)) + c.r4[2] * (F(7) - F(1));
        } else if (KIND == NP_CD10_D2) {
            if (i == 1) return c.r1[0] * F(1) + c.r1[1] * F(2) + c.r1[2] * F(3) + c.r1[3] * F(4) + c.r1[4] * F(5);
            if (i == 2) return c.r2[0] * (F(3) - 2.0 * F(2) + F(1));
            return c.r3[0] * (F(4) - 2.0 * F(3) + F(2)) + c.r3[1] * (F(5) - 2.0 * F(3) + F(1));
        } else if (KIND == NP_CD06_D1) {
            if (i == 1) return c.r1[0] * F(1) + c.r1[1] * F(2) + c.r1[2] * F(3) + c.r1[3] * F(4);
            if (i == 2) return c.r2[0] * (F(3) - F(1));
            return c.r3[0] * (F(4) - F(2)) + c.r3[1] * (F(5) - F(1));
        } else {
            if (i == 1) return (KIND == NP_GAUSS || KIND == NP_LSTSQ) ? c.r1[0] * (F(1)) + c.r1[1] * (F(2)) : c.r1[0] * (F(1));
            if (i == 2) return c.r2[0] * (F(2)) + c.r2[1] * (F(3) + F(1));
            if (i == 3) return c.r3[0] * (F(3)) + c.r3[1] * (F(4) + F(2)) + c.r3[2] * (F(5) + F(1));
            return c.r4[0] * (F(4)) + c.r4[1] * (F(5) + F(3)) + c.r4[2] * (F(6) + F(2)) + c.r4[3] * (F(7) + F(1));
        }
    }
    if (bcn == 0 && i > n - NB) {
        const int m = n - i + 1;   // 1 = last row, 2 = last but one, ...
        if (KIND == NP_CD10_D1) {
            if (m == 1) return -c.r1[0] * F(n) - c.r1[1] * F(n - 1) - c.r1[2] * F(n - 2) - c.r1[3] * F(n - 3);
            if (m == 2) return c.r2[0] * (F(n) - F(n - 2));
            if (m == 3) return c.r3[0] * (F(n - 1) - F(n - 3)) + c.r3[1] * (F(n) - F(n - 4));
            return c.r4[0] * (F(n - 2) - F(n - 4)) + c.r4[1] * (F(n - 1) - F(n - 5)) + c.r4[2] * (F(n) - F(n - 6));
        } else if (KIND == NP_CD10_D2) {
            if (m == 1) return c.r1[0] * F(n) + c.r1[1] * F(n - 1) + c.r1[2] * F(n - 2) + c.r1[3] * F(n - 3) + c.r1[4] * F(n - 4);
            if (m == 2) return c.r2[0] * (F(n) - 2.0 * F(n - 1) + F(n - 2));
            return c.r3[0] * (F(n - 1) - 2.0 * F(n - 2) + F(n - 3)) + c.r3[1] * (F(n) - 2.0 * F(n - 2) + F(n - 4));
        } else if (KIND == NP_CD06_D1) {
            if (m == 1) return -c.r1[0] * F(n) - c.r1[1] * F(n - 1) - c.r1[2] * F(n - 2) - c.r1[3] * F(n - 3);
            if (m == 2) return c.r2[0] * (F(n) - F(n - 2));
            return c.r3[0] * (F(n - 1) - F(n - 3)) + c.r3[1] * (F(n) - F(n - 4));
        } else {
            if (m == 1) return (KIND == NP_GAUSS || KIND == NP_LSTSQ) ? c.r1[0] * (F(n)) + c.r1[1] * (F(n - 1)) : c.r1[0] * (F(n));
            if (m == 2) return c.r2[0] * (F(n - 1)) + c.r2[1] * (F(n) + F(n - 2));
            if (m == 3) return c.r3[0] * (F(n - 2)) + c.r3[1] * (F(n - 1) + F(n - 3)) + c.r3[2] * (F(n) + F(n - 4));
            return c.r4[0] * (F(n - 3)) + c.r4[1] * (F(n - 2) + F(n - 4)) + c.r4[2] * (F(n - 1) + F(n - 5)) + c.r4[3] * (F(n) + F(n - 6));
        }
    }
    // interior row on the reflected line: G(j) = F(j) inside, s1 F(2-j) below 1, sn F(2n-j) above n
    const double s1 = (double)bc1, sn = (double)bcn;
    auto G = [&](int j) -> double { return j < 1 ? s1 * F(2 - j) : (j > n ? sn * F(2 * n - j) : F(j)); };
    if (KIND == NP_CD10_D1) {
        return c.in[0] * (G(i + 1) - G(i - 1)) + c.in[1] * (G(i + 2) - G(i - 2)) + c.in[2] * (G(i + 3) - G(i - 3));
    } else if (KIND == NP_CD06_D1) {   // cd06.F90:574-575: the b06 term comes first
        return c.in[1] * (G(i + 2) - G(i - 2)) + c.in[0] * (G(i + 1) - G(i - 1));
    } else if (KIND == NP_CD10_D2) {
        const double f0 = F(i);
        return c.in[0] * (G(i + 1) - 2.0 * f0 + G(i - 1)) + c.in[1] * (G(i + 2) - 2.0 * f0 + G(i - 2)) +
               c.in[2] * (G(i + 3) - 2.0 * f0 + G(i - 3));
    } else {
        return c.in[0] * (F(i)) + c.in[1] * (G(i + 1) + G(i - 1)) + c.in[2] * (G(i + 2) + G(i - 2)) + c.in[3] * (G(i + 3) + G(i - 3)) +
               c.in[4] * (G(i + 4) + G(i - 4));
    }
}

// SolveXPenta1 on one line, in place: y(i) at y[(i-1)*es].  tab = f[n] g[n] obc[n] at[n] eobc[n] (penta columns 8, 9, 7, 5, 10
// of the reference's table); operand order as in cd10.F90:833-848.
PDO_HD void np_solve_line(double* y, long long es, int n, const double* tab) {
    const double *tf = tab, *tg = tab + n, *obc = tab + 2 * (long long)n, *at = tab + 3 * (long long)n, *eobc = tab + 4 * (long long)n;
    double ym2 = y[0];
    double ym1 = y[es] - tf[1] * ym2;
    y[es] = ym1;
    for (int i = 2; i < n; ++i) {
        const double v = y[i * es] - tg[i] * ym2 - tf[i] * ym1;
        y[i * es] = v;
        ym2 = ym1; ym1 = v;
    }
    double yp1 = y[(n - 1) * es] * obc[n - 1];
    y[(n - 1) * es] = yp1;
    double yp0 = y[(n - 2) * es] * obc[n - 2] - eobc[n - 2] * yp1;
    y[(n - 2) * es] = yp0;
    for (int i = n - 3; i >= 0; --i) {
        const double v = y[i * es] * obc[i] - yp1 * at[i] * obc[i] - yp0 * eobc[i];
        y[i * es] = v;
        yp1 = yp0; yp0 = v;
    }
}

// Host side (tables.cpp-style, plain double like the reference's ComputePenta*): coefficients and the 5n-double table.
// Returns 0; 2 / 7 for lines shorter than the closures allow (cd10: 8, cf90: 10); 324 for bad boundary codes.
int np_build_coefs(int kind, double dx, NpCoefs* out);
int np_build_table(int kind, int n, int bc1, int bcn, double* tab5n);
int np_build_rows(int kind, int n, int bc1, int bcn, double* rows5n);   // bt[n] b[n] d[n] a[n] at[n]

// Chunked fast path of one (bc1, bcn) system (np_chunk.cu): three table sets (first / mid / last chunk; sizeof = 3 NpChunkSet,
// kept as raw doubles so that this header does not pull np_chunk_tables.h in) + the position-dependent separator inverse on the device.
struct NpFast {
    bool ok = false;
    int P = 0, W = 0;
    double sets[3][(5 * 32 + 2 * 64 + 6)];
    double* d_G = nullptr;
};
cudaError_t np_fast_create(NpFast* t, int kind, int n, int bc1, int bcn);
void np_fast_destroy(NpFast* t);

struct NpOp {
    int kind = 0, n = 0;
    NpCoefs co{};
    double* d_tab[9] = {};   // device tables for (bc1, bcn) in {0, 1, -1}^2, index 3*slot(bc1) + slot(bcn), slot: 0 -> 0, 1 -> 1, -1 -> 2
    NpFast fast[9];          // same indexing; ok = false where the line is not chunkable (the sweeps above then do the work)
};
cudaError_t np_fast_apply(const NpFast& t, int kind, const NpCoefs& co, int n, int axis, const double* f, double* out, long long n1,
                          long long n3, int bc1, int bcn, cudaStream_t st);
// PDO_NP_FAST=0 keeps every non-periodic call on the one-thread-per-line sweeps (tests compare the two paths)
void np_set_fast_path(int mode);   // -1 environment / default (on), 0 off, 1 on
cudaError_t np_op_create(NpOp* h, int kind, int n, double dx, int* ierr_out);
void np_op_destroy(NpOp* h);
// axis 0: f(n,na,nb); 1: f(na,n,nb); 2: f(na,nb,n).  Device pointers; f and out must not alias.
cudaError_t np_op_apply(const NpOp* h, int axis, const double* f, double* out, long long na, long long nb, int bc1, int bcn, cudaStream_t st);
// the same arithmetic executed on the host (test hook only)
int np_apply_host(int kind, int n, double dx, int bc1, int bcn, int axis, const double* f, double* out, long long na, long long nb);

}  // namespace pdo
