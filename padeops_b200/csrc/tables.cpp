// tables.cpp — host-side factorisation of the cyclic banded LHS into chunk tables (see tables.h).
// All arithmetic in long double (x87 80-bit here), rounded to double once at the end.
#include "tables.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace pdo {
namespace {

typedef long double ld;

// LU (no pivoting; the matrices are symmetric positive definite) of the N x N non-cyclic Toeplitz
// band matrix with diagonal 1, first off-diagonals b1, second off-diagonals b2 (b2 = 0 and BW = 1
// for tridiagonal).  Doolittle:  l2_i = b2/g_{i-2};  l1_i = (b1 - l2_i u1_{i-2})/g_{i-1};
// g_i = 1 - l2_i b2 - l1_i u1_{i-1};  u1_i = b1 - l1_i b2.
struct BandLU {
    int N;
    ld b2;
    std::vector<ld> l1, l2, g, u1;
    BandLU(int N_, ld b1, ld b2_) : N(N_), b2(b2_), l1(N_, 0), l2(N_, 0), g(N_, 0), u1(N_, 0) {
        for (int i = 0; i < N; ++i) {
            l2[i] = (i >= 2) ? b2 / g[i - 2] : (ld)0;
            l1[i] = (i >= 1) ? (b1 - l2[i] * (i >= 2 ? u1[i - 2] : (ld)0)) / g[i - 1] : (ld)0;
            g[i] = (ld)1 - l2[i] * b2 - l1[i] * (i >= 1 ? u1[i - 1] : (ld)0);
            u1[i] = b1 - l1[i] * b2;
        }
    }
    void solve(ld* x) const {  // in place
        for (int i = 0; i < N; ++i) {
            ld v = x[i];
            if (i >= 1) v -= l1[i] * x[i - 1];
            if (i >= 2) v -= l2[i] * x[i - 2];
            x[i] = v;
        }
        for (int i = N - 1; i >= 0; --i) {
            ld v = x[i];
            if (i + 1 < N) v -= u1[i] * x[i + 1];
            if (i + 2 < N) v -= b2 * x[i + 2];
            x[i] = v / g[i];
        }
    }
};

// Small dense solve with partial pivoting (sizes <= 4 here).
void dense_solve(int n, ld* A, ld* b) {
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (fabsl(A[r * n + c]) > fabsl(A[piv * n + c])) piv = r;
        if (piv != c) {
            for (int k = 0; k < n; ++k) { ld t = A[c * n + k]; A[c * n + k] = A[piv * n + k]; A[piv * n + k] = t; }
            ld t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        for (int r = c + 1; r < n; ++r) {
            ld m = A[r * n + c] / A[c * n + c];
            for (int k = c; k < n; ++k) A[r * n + k] -= m * A[c * n + k];
            b[r] -= m * b[c];
        }
    }
    for (int r = n - 1; r >= 0; --r) {
        ld v = b[r];
        for (int k = r + 1; k < n; ++k) v -= A[r * n + k] * b[k];
        b[r] = v / A[r * n + r];
    }
}

// First column c of the inverse of the n x n CYCLIC matrix circ[b2 b1 1 b1 b2]  (A^{-1}[i][j] = c[(i-j) mod n]).
// Woodbury on the non-cyclic band B:  A = B + E, E = the corner entries.  E c touches only rows
// {0..BW-1, n-BW..n-1} and reads only c at the same indices, so  c = B^{-1} e0 - sum_j t_j B^{-1} E(:,j)
// with t = c restricted to those 2 BW indices, found from a 2BW x 2BW system.
std::vector<ld> cyclic_inverse_first_column(int n, int BW, ld b1, ld b2) {
    BandLU B(n, b1, b2);
    const int K = 2 * BW;
    std::vector<int> idx(K);
    for (int j = 0; j < BW; ++j) { idx[j] = j; idx[BW + j] = n - BW + j; }
    // corner entries: A[i][j] for (i - j) mod n in {±1, ±2} that wrap around
    auto corner = [&](int i, int j) -> ld {
        int d = ((i - j) % n + n) % n;
        int dist = d < n - d ? d : n - d;
        bool wraps = std::abs(i - j) != dist;  // true only for entries that live in a corner
        if (!wraps || dist == 0 || dist > BW) return 0;
        return dist == 1 ? b1 : b2;
    };
    std::vector<std::vector<ld>> Y(K, std::vector<ld>(n, 0));  // Y_j = B^{-1} E(:, idx_j)
    for (int j = 0; j < K; ++j) {
        for (int q = 0; q < K; ++q) Y[j][idx[q]] = corner(idx[q], idx[j]);
        B.solve(Y[j].data());
    }
    std::vector<ld> y0(n, 0);
    y0[0] = 1;
    B.solve(y0.data());
    // t_q = y0[idx_q] - sum_j t_j Y_j[idx_q]   →  (I + Yr) t = y0r
    std::vector<ld> M(K * K), rhs(K);
    for (int q = 0; q < K; ++q) {
        rhs[q] = y0[idx[q]];
        for (int j = 0; j < K; ++j) M[q * K + j] = (q == j ? (ld)1 : (ld)0) + Y[j][idx[q]];
    }
    dense_solve(K, M.data(), rhs.data());
    std::vector<ld> c(n);
    for (int i = 0; i < n; ++i) {
        ld v = y0[i];
        for (int j = 0; j < K; ++j) v -= rhs[j] * Y[j][i];
        c[i] = v;
    }
    return c;
}

}  // namespace

int build_chunk_tables(int n, int M, int BW, double b1d, double b2d, ChunkTables* out) {
    std::memset(out, 0, sizeof(*out));
    if (M > kMaxChunk || M < 2 * BW + 2 || n % M != 0 || BW < 1 || BW > 2) return -1;
    const int P = n / M, mi = M - BW;
    const ld b1 = b1d, b2 = (BW == 2) ? (ld)b2d : (ld)0;
    if (n < 2 * BW + 1) return -1;
    out->n = n; out->M = M; out->P = P; out->BW = BW; out->b1 = b1d; out->b2 = (BW == 2) ? b2d : 0.0;

    BandLU T(mi, b1, b2);
    for (int i = 0; i < mi; ++i) {
        out->l1[i] = (double)T.l1[i];
        out->l2[i] = (double)T.l2[i];
        out->ginv[i] = (double)((ld)1 / T.g[i]);
        out->ug[i] = (double)(T.u1[i] / T.g[i]);
        out->bg[i] = (double)(b2 / T.g[i]);
    }
    // Left spike: interior rows 0..BW-1 see the previous chunk's separators (sa = its local M-2, sb = M-1 for
    // penta; the single separator for tri).  Row 0: b2*sa + b1*sb ; row 1: b2*sb.   V(:,q) = T^{-1} E_L(:,q).
    // Right spike: rows mi-2: b2*ta ; row mi-1: b1*ta + b2*tb.                      U(:,q) = T^{-1} E_R(:,q).
    for (int q = 0; q < BW; ++q) {
        std::vector<ld> v(mi, 0), u(mi, 0);
        if (BW == 2) {
            if (q == 0) { v[0] = b2; u[mi - 2] += b2; u[mi - 1] += b1; }
            else { v[0] = b1; v[1] += b2; u[mi - 1] = b2; }
        } else {
            v[0] = b1; u[mi - 1] = b1;
        }
        T.solve(v.data());
        T.solve(u.data());
        for (int i = 0; i < mi; ++i) { out->V[i][q] = (double)v[i]; out->U[i][q] = (double)u[i]; }
    }
    // Separator inverse = the separator-separator blocks of A^{-1} (inverse of a Schur complement).
    std::vector<ld> c = cyclic_inverse_first_column(n, BW, b1, b2);
    ld cmax = 0;
    for (int i = 0; i < n; ++i) cmax = fabsl(c[i]) > cmax ? fabsl(c[i]) : cmax;
    auto blk = [&](int d, int a, int b) -> ld {  // A^{-1}[sep a of chunk p][sep b of chunk p+d]
        long long k = (long long)(-d) * M + a - b;
        k %= n; if (k < 0) k += n;
        return c[(int)k];
    };
    int W = 0;
    for (int d = 0; d <= P / 2; ++d)
        for (int a = 0; a < BW; ++a)
            for (int b = 0; b < BW; ++b)
                if (fabsl(blk(d, a, b)) > (ld)1e-19 * cmax || fabsl(blk(-d, a, b)) > (ld)1e-19 * cmax) W = d;
    int dense = 0;
    if (2 * W + 1 >= P) {
        dense = 1;
        W = P / 2;  // offsets d = -(P-1)/2 .. P/2 cover every chunk exactly once
    }
    if (W > kMaxW) return -1;
    out->W = W; out->dense = dense;
    for (int d = -W; d <= W; ++d)
        for (int a = 0; a < BW; ++a)
            for (int b = 0; b < BW; ++b) out->G[d + W][a * BW + b] = (double)blk(d, a, b);
    if (dense && P % 2 == 0) {  // d = -P/2 and +P/2 are the same chunk: keep only +P/2
        for (int q = 0; q < 4; ++q) out->G[0][q] = 0.0;
    }
    return 0;
}

int build_line_tables(int n, int BW, double b1d, double b2d, LineTablesHost* out) {
    const int mi = n - BW;
    if (BW < 1 || BW > 2 || mi < 2 * BW) return -1;
    const ld b1 = b1d, b2 = (BW == 2) ? (ld)b2d : (ld)0;
    out->n = n; out->BW = BW; out->b1 = b1d; out->b2 = (BW == 2) ? b2d : 0.0;
    double* D = (double*)std::calloc((size_t)6 * n + 4, sizeof(double));
    if (!D) return -2;
    out->data = D;
    BandLU T(mi, b1, b2);
    for (int i = 0; i < mi; ++i) {
        D[i] = (double)T.l1[i];
        D[n + i] = (double)T.l2[i];
        D[2 * n + i] = (double)((ld)1 / T.g[i]);
        D[3 * n + i] = (double)T.u1[i];
    }
    for (int q = 0; q < BW; ++q) {
        std::vector<ld> v(mi, 0), u(mi, 0);
        if (BW == 2) {
            if (q == 0) { v[0] = b2; u[mi - 2] += b2; u[mi - 1] += b1; }
            else { v[0] = b1; v[1] += b2; u[mi - 1] = b2; }
        } else {
            v[0] = b1; u[mi - 1] = b1;
        }
        T.solve(v.data());
        T.solve(u.data());
        for (int i = 0; i < mi; ++i) D[4 * n + 2 * i + q] = (double)(v[i] + u[i]);  // one chunk: s_prev == s_own
    }
    std::vector<ld> c = cyclic_inverse_first_column(n, BW, b1, b2);
    for (int a = 0; a < BW; ++a)
        for (int b = 0; b < BW; ++b) {
            int k = ((a - b) % n + n) % n;
            D[6 * n + a * BW + b] = (double)c[k];
        }
    return 0;
}

}  // namespace pdo
