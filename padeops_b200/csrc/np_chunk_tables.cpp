// np_chunk_tables.cpp — host-side factorisation of a NON-CYCLIC pentadiagonal system with position-dependent boundary rows
// into chunk tables: the non-periodic counterpart of tables.cpp, feeding the chunked (fast-path) kernels of the
// non-periodic closures.  Kernels for it are the next step; this file and its CPU test
// (tests/test_np_chunk_tables_cpu.py) fix the algebra and the table layout first.
//
// Same cut as the cyclic case: chunks of M rows, the last 2 rows of every chunk are separators.  What changes:
//   * the interior block T_p of the FIRST and LAST chunk contains boundary rows, so three table sets (first / mid / last);
//   * the couplings between a chunk's interior and the separators around it are row entries, not (b1, b2) constants;
//   * the separator system is block tridiagonal, not block circulant: its inverse depends on the position, so the
//     truncated inverse G is stored per chunk, G[p][d] = Sinv[p][p - W + d] (zero blocks outside the matrix).
// All arithmetic in long double, rounded to double once.
#include <cmath>
#include <cstring>
#include <vector>

#include "np_chunk_tables.h"

namespace pdo {
namespace {

typedef long double ld;

struct Band {   // rows of the n x n pentadiagonal matrix, 0-based: A[i][i-2..i+2] = bt, b, d, a, at
    int n;
    std::vector<ld> bt, b, d, a, at;
    ld at_(int i, int j) const {
        const int k = j - i;
        if (j < 0 || j >= n) return 0;
        switch (k) { case -2: return bt[i]; case -1: return b[i]; case 0: return d[i]; case 1: return a[i]; case 2: return at[i]; default: return 0; }
    }
};

// LU without pivoting of the band block rows/cols [r0, r0 + m): returns per-row factors
struct BlockLU {
    int m;
    std::vector<ld> l1, l2, g, u1, u2;
    BlockLU(const Band& A, int r0, int m_) : m(m_), l1(m_, 0), l2(m_, 0), g(m_, 0), u1(m_, 0), u2(m_, 0) {
        for (int i = 0; i < m; ++i) {
            const int r = r0 + i;
            const ld sub2 = i >= 2 ? A.bt[r] : (ld)0, sub1 = i >= 1 ? A.b[r] : (ld)0;
            l2[i] = i >= 2 ? sub2 / g[i - 2] : (ld)0;
            l1[i] = i >= 1 ? (sub1 - l2[i] * (i >= 2 ? u1[i - 2] : (ld)0)) / g[i - 1] : (ld)0;
            g[i] = A.d[r] - l2[i] * (i >= 2 ? u2[i - 2] : (ld)0) - l1[i] * (i >= 1 ? u1[i - 1] : (ld)0);
            u1[i] = (i + 1 < m ? A.a[r] : (ld)0) - l1[i] * (i >= 1 ? u2[i - 1] : (ld)0);
            u2[i] = i + 2 < m ? A.at[r] : (ld)0;
        }
    }
    void solve(ld* x) const {
        for (int i = 0; i < m; ++i) {
            ld v = x[i];
            if (i >= 1) v -= l1[i] * x[i - 1];
            if (i >= 2) v -= l2[i] * x[i - 2];
            x[i] = v;
        }
        for (int i = m - 1; i >= 0; --i) {
            ld v = x[i];
            if (i + 1 < m) v -= u1[i] * x[i + 1];
            if (i + 2 < m) v -= u2[i] * x[i + 2];
            x[i] = v / g[i];
        }
    }
};

void fill_set(const Band& A, int p, int M, NpChunkSet* s) {
    std::memset(s, 0, sizeof(*s));
    const int mi = M - 2, r0 = p * M, P = A.n / M;
    BlockLU T(A, r0, mi);
    for (int i = 0; i < mi; ++i) {
        s->l1[i] = (double)T.l1[i];
        s->l2[i] = (double)T.l2[i];
        s->ginv[i] = (double)((ld)1 / T.g[i]);
        s->ug[i] = (double)(T.u1[i] / T.g[i]);
        s->bg[i] = (double)(T.u2[i] / T.g[i]);
    }
    // spikes: V = T^{-1} A[I_p, S_{p-1}],  U = T^{-1} A[I_p, S_p]
    for (int q = 0; q < 2; ++q) {
        std::vector<ld> v(mi, 0), u(mi, 0);
        for (int i = 0; i < mi; ++i) {
            if (p > 0) v[i] = A.at_(r0 + i, r0 - 2 + q);
            u[i] = A.at_(r0 + i, r0 + mi + q);
        }
        T.solve(v.data());
        T.solve(u.data());
        for (int i = 0; i < mi; ++i) { s->V[i][q] = (double)v[i]; s->U[i][q] = (double)u[i]; }
    }
    // reduced right-hand side pieces: gA = r_sep - cA . z_tail (own separator rows), gB = - cB . z_head (the PREVIOUS chunk's
    // separator rows acting on my head)
    const int s0 = r0 + mi, s1 = s0 + 1;
    s->cA[0] = (double)A.at_(s0, s0 - 2); s->cA[1] = (double)A.at_(s0, s0 - 1); s->cA[2] = (double)A.at_(s1, s1 - 2);
    if (p > 0) {
        const int q0 = r0 - 2, q1 = r0 - 1;
        s->cB[0] = (double)A.at_(q0, r0); s->cB[1] = (double)A.at_(q1, r0); s->cB[2] = (double)A.at_(q1, r0 + 1);
    }
    (void)P;
}

}  // namespace

int build_np_chunk_tables(int n, int M, const double* rows5n, NpChunkTables* out) {
    if (M != 32 && M != 16 && M != 8) return -1;
    if (n % M != 0 || n / M < 2) return -1;
    const int P = n / M, mi = M - 2;
    Band A;
    A.n = n;
    A.bt.assign(rows5n, rows5n + n); A.b.assign(rows5n + n, rows5n + 2 * (size_t)n); A.d.assign(rows5n + 2 * (size_t)n, rows5n + 3 * (size_t)n);
    A.a.assign(rows5n + 3 * (size_t)n, rows5n + 4 * (size_t)n); A.at.assign(rows5n + 4 * (size_t)n, rows5n + 5 * (size_t)n);
    out->n = n; out->M = M; out->P = P;
    fill_set(A, 0, M, &out->first);
    fill_set(A, P > 2 ? 1 : 0, M, &out->mid);
    fill_set(A, P - 1, M, &out->last);
    // separator-separator blocks of A^{-1}: one banded solve per separator column
    BlockLU F(A, 0, n);
    std::vector<std::vector<ld>> col(2 * (size_t)P, std::vector<ld>(n, 0));
    for (int q = 0; q < P; ++q)
        for (int b = 0; b < 2; ++b) {
            std::vector<ld>& y = col[2 * (size_t)q + b];
            y[q * M + mi + b] = 1;
            F.solve(y.data());
        }
    auto blk = [&](int p, int a, int q, int b) -> ld { return col[2 * (size_t)q + b][p * M + mi + a]; };
    ld smax = 0;
    for (int p = 0; p < P; ++p)
        for (int q = 0; q < P; ++q)
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) smax = fabsl(blk(p, a, q, b)) > smax ? fabsl(blk(p, a, q, b)) : smax;
    int W = 0;
    for (int p = 0; p < P; ++p)
        for (int q = 0; q < P; ++q) {
            ld m = 0;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) m = fabsl(blk(p, a, q, b)) > m ? fabsl(blk(p, a, q, b)) : m;
            const int d = q > p ? q - p : p - q;
            if (m > (ld)1e-19 * smax && d > W) W = d;
        }
    if (W > kNpMaxW) return -1;
    out->W = W;
    out->G.assign((size_t)P * (2 * W + 1) * 4, 0.0);
    for (int p = 0; p < P; ++p)
        for (int d = 0; d <= 2 * W; ++d) {
            const int q = p - W + d;
            if (q < 0 || q >= P) continue;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) out->G[((size_t)p * (2 * W + 1) + d) * 4 + a * 2 + b] = (double)blk(p, a, q, b);
        }
    return 0;
}

}  // namespace pdo
