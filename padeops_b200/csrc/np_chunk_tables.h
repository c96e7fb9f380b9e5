// np_chunk_tables.h — chunk tables of a non-cyclic pentadiagonal system with boundary rows (see np_chunk_tables.cpp).
#pragma once
#include <vector>

namespace pdo {

constexpr int kNpMaxW = 16;

// One chunk position class (first / mid / last).  Per chunk, with r = the M right-hand-side values of the chunk:
//   y_i = r_i - l1_i y_{i-1} - l2_i y_{i-2};   z_i = y_i ginv_i - ug_i z_{i+1} - bg_i z_{i+2}        (i < M-2)
//   gA = (r[M-2] - cA0 z[M-4] - cA1 z[M-3],  r[M-1] - cA2 z[M-3]);   gB = (-cB0 z[0],  -cB1 z[0] - cB2 z[1])
//   h_p = gA_p + gB_{p+1}  (gB_P = 0);   s_p = sum_d G[p][d] h_{p-W+d};   x = z - V s_{p-1} - U s_p,  x_sep = s_p
struct NpChunkSet {
    double l1[32], l2[32], ginv[32], ug[32], bg[32];
    double V[32][2], U[32][2];
    double cA[3], cB[3];
};
struct NpChunkTables {
    int n, M, P, W;
    NpChunkSet first, mid, last;
    std::vector<double> G;   // [P][2W+1][4], row-major 2x2 blocks
};

// rows5n = bt[n] b[n] d[n] a[n] at[n].  Returns 0, or -1 if (n, M) is not chunkable (n % M, fewer than 2 chunks, reach > kNpMaxW).
int build_np_chunk_tables(int n, int M, const double* rows5n, NpChunkTables* out);

}  // namespace pdo
