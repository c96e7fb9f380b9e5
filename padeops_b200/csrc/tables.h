// tables.h — precomputed factor tables for the chunked cyclic banded solves (host + device).
//
// The reference factors the cyclic tri-/pentadiagonal LHS once per operator (ComputeLU,
// derivatives/cd10.F90:351-427, cd06.F90:221-262) and then sweeps every line sequentially with
// running corner sums.  A length-n dependent sweep cannot stay on chip on a GPU, so this library
// factors the same matrix differently — still once, on the host, in extended precision:
//
//   the line is cut into P = n/M chunks of M points; the last BW points of every chunk are
//   separators (BW = half bandwidth: 1 tri, 2 penta).  Removing the separators leaves P identical
//   non-cyclic Toeplitz blocks T (size M-BW), so one small LU serves every chunk.  The separator
//   unknowns obey a block-circulant Schur system whose inverse is read off the first column of
//   A^{-1}; its blocks decay like rho^(M d), so only |d| <= W neighbour chunks are kept (entries
//   below 1e-19 relative are dropped; W covers all chunks when the decay is slow).
//
// Per chunk, in registers:   z = T^{-1} r ;  g = r_sep - (coupling) z ;  s = sum_d G[d] g_{p+d} ;
//                            x = z - V s_{p-1} - U s_p ;  x_sep = s_p.
// Flop count per point equals the reference's sweep (which drags 4 corner columns through every row).
#pragma once
#include <cstdint>

namespace pdo {

constexpr int kMaxChunk = 32;  // M
constexpr int kMaxW = 16;      // |d| <= W neighbour chunks in the separator solve

// Everything the kernels need for one cyclic banded matrix [b2 b1 1 b1 b2] at line length n, chunk M.
// Passed BY VALUE as a __grid_constant__ kernel parameter: it lives in the constant bank, and with the
// chunk loops fully unrolled every factor becomes a c[0x0][imm] operand of the DFMA that uses it.
struct ChunkTables {
    double l1[kMaxChunk];    // forward:  y_i = r_i - l1_i y_{i-1} - l2_i y_{i-2}
    double l2[kMaxChunk];
    double ginv[kMaxChunk];  // backward: z_i = y_i ginv_i - ug_i z_{i+1} - bg_i z_{i+2}   (ug = u1 ginv, bg = b2 ginv:
    double ug[kMaxChunk];    //           the 1/g scaling is folded into the factors so the dependent chain is
    double bg[kMaxChunk];    //           one FMA per row)
    double V[kMaxChunk][2];  // x_i -= V[i][0]*sprev[0] + V[i][1]*sprev[1]   (left spike)
    double U[kMaxChunk][2];  // x_i -= U[i][0]*sown[0]  + U[i][1]*sown[1]    (right spike)
    double G[2 * kMaxW + 1][4];  // separator inverse blocks, G[d+W] row-major BWxBW
    double b1, b2;           // off-diagonals of the LHS (alpha, beta)
    int W;                   // neighbour reach actually needed
    int dense;               // 1: loop covers every chunk exactly once (2W+1 >= P)
    int n, M, P, BW;
};

// Builds the tables in long double.  Returns 0, or -1 if (n, M) is not chunkable for this matrix
// (n % M != 0, M-BW < BW, or the separator inverse needs more reach than kMaxW without being dense).
int build_chunk_tables(int n, int M, int BW, double b1, double b2, ChunkTables* out);

// Any-n variant (one chunk = the whole line, P = 1): same factorisation with table length n, kept in
// device global memory and used by the generic one-thread-per-line kernels for line lengths that are
// not a multiple of 8.  Layout of `data` (doubles): l1[n] l2[n] ginv[n] u1[n] VU[n][2] G0[4].
struct LineTablesHost {
    int n, BW;
    double b1, b2;
    double* data;  // malloc'ed, 6*n + 4 doubles; caller frees
};
int build_line_tables(int n, int BW, double b1, double b2, LineTablesHost* out);

}  // namespace pdo
