// stagg_np.cuh — NON-PERIODIC staggered 6th-order compact operators in z (second half): the eight
// z-operators of cd06stagg%init_nonperiodic with even / odd symmetry or one-sided closures at each wall — what a wall-bounded
// igrid differentiates and interpolates with.
//
// Reference: derivatives/cd06stagg.F90:17-56 (constants), 197-231 (init_nonperiodic), 820-1059 (the operators) and the
// included derivatives/STAGG_CD06_files/{ComputeTri_allRoutines, TridiagSolver_allRoutines, D1RHS_{E2C,C2E,C2C,E2E}_common,
// InterpRHS_{E2C,C2E}_common, D2RHS_{C2C,E2E}_common}.F90.
//
// Correctness path (like nonperiodic.cuh for the collocated schemes): one thread per z-line, the right-hand side is formed
// inside the forward Thomas sweep (no RHS pass through memory), tables (ddn*den, den, cp) in global memory.  Complex fields use
// the real tables: a complex line is two interleaved real lines.  The per-point RHS and the sweeps are __host__ __device__ so
// that CPU tests execute exactly the code the kernel executes (pdo_debug_stagg_np_host).
//
// Cells sit at (k + 1/2) dz, edges at k dz, k = 0 ..; n cells, nE = n + 1 edges.  Indices below are 0-based.
#pragma once
#include <cuda_runtime.h>

#include "nonperiodic.cuh"   // PDO_HD

namespace pdo {

enum StaggNpOp { SNP_D1_E2C = 0, SNP_D1_C2E = 1, SNP_D1_C2C = 2, SNP_D1_E2E = 3, SNP_INTERP_E2C = 4, SNP_INTERP_C2E = 5,
                 SNP_D2_C2C = 6, SNP_D2_E2E = 7, SNP_COUNT = 8 };

struct StaggNpFlags { int botEven, topEven, botSided, topSided; };

// rows the operator reads / writes for n cells
PDO_HD int snp_rows_in(int op, int n) { return (op == SNP_D1_E2C || op == SNP_D1_E2E || op == SNP_INTERP_E2C || op == SNP_D2_E2E) ? n + 1 : n; }
PDO_HD int snp_rows_out(int op, int n) { return (op == SNP_D1_C2E || op == SNP_D1_E2E || op == SNP_INTERP_C2E || op == SNP_D2_E2E) ? n + 1 : n; }

// cd06stagg.F90:17-56
struct StaggNpConst {
    double o, o2;                                  // 1/dx, 1/dx^2
    double w1, w2, w3, q_p, q_pp, r_pp, p, q, r, s; // one-sided collocated closure (rows 1-3 of C2C)
};

// right-hand side of row k (0-based) of operator OP; F(j): input value at 0-based index j of the line
template <int OP, class Acc>
PDO_HD double snp_rhs(int k, int n, const StaggNpFlags& fl, const StaggNpConst& c, Acc F) {
    const int nE = n + 1;
    const double w0s = 223.0 / 186.0, w1s = 61.0 / 62.0;
    if (OP == SNP_D1_E2C) {                        // D1RHS_E2C_common.F90
        const double a06 = (63.0 / 62.0) * c.o, b06 = ((17.0 / 62.0) / 3.0) * c.o;
        const double al1 = 37.0 / 183.0, al0 = -1.0;
        const double a0 = (1.0 / 24.0) * (al0 - 23.0), b0 = (1.0 / 8.0) * (-9.0 * al0 + 7.0);
        const double c0 = (1.0 / 8.0) * (9.0 * al0 + 1.0), d0 = -(1.0 / 24.0) * (al0 + 1.0);
        const double a1 = (3.0 / 8.0) * (3.0 - 2.0 * al1), b1 = (1.0 / 8.0) * (-1.0 + 22.0 * al1);
        if (k == n - 1) {
            if (fl.topSided) return -w0s * c.o * (a0 * F(nE - 1) + b0 * F(nE - 2) + c0 * F(nE - 3) + d0 * F(nE - 4));
            if (fl.topEven) return b06 * (F(nE - 2) - F(nE - 3)) + a06 * (F(nE - 1) - F(nE - 2));
            return -b06 * (F(nE - 2) + F(nE - 3)) + a06 * (F(nE - 1) - F(nE - 2));
        }
        if (k == n - 2 && fl.topSided) return w1s * c.o * ((-b1 / 3.0) * F(nE - 4) + (-a1) * F(nE - 3) + (a1) * F(nE - 2) + (b1 / 3.0) * F(nE - 1));
        if (k == 0) {
            if (fl.botSided) return w0s * c.o * (a0 * F(0) + b0 * F(1) + c0 * F(2) + d0 * F(3));
            if (fl.botEven) return b06 * (F(2) - F(1)) + a06 * (F(1) - F(0));
            return b06 * (F(2) + F(1)) + a06 * (F(1) - F(0));
        }
        if (k == 1 && fl.botSided) return w1s * c.o * ((-b1 / 3.0) * F(0) + (-a1) * F(1) + (a1) * F(2) + (b1 / 3.0) * F(3));
        return b06 * (F(k + 2) - F(k - 1)) + a06 * (F(k + 1) - F(k));
    } else if (OP == SNP_D1_C2E) {                 // D1RHS_C2E_common.F90
        const double a06 = (63.0 / 62.0) * c.o, b06 = ((17.0 / 62.0) / 3.0) * c.o;
        const double a0 = -71.0 / 24.0, b0 = 47.0 / 8.0, c0 = -31.0 / 8.0, d0 = 23.0 / 24.0, a1 = 12.0 / 11.0;
        if (k == nE - 1) {
            if (fl.topSided) return (a0 * F(n - 1) + b0 * F(n - 2) + c0 * F(n - 3) + d0 * F(n - 4)) * (-c.o);
            if (fl.topEven) return 0.0;
            return -2.0 * b06 * (F(n - 2)) - 2.0 * a06 * (F(n - 1));
        }
        if (k == 0) {
            if (fl.botSided) return (a0 * F(0) + b0 * F(1) + c0 * F(2) + d0 * F(3)) * c.o;
            if (fl.botEven) return 0.0;
            return 2.0 * b06 * F(1) + 2.0 * a06 * F(0);
        }
        if (k == nE - 2) {
            if (fl.topSided) return (F(n - 1) - F(n - 2)) * (a1 * c.o);
            const double base = a06 * (F(k) - F(k - 1));
            return fl.topEven ? base + b06 * (F(n - 1) - F(n - 3)) : base - b06 * (F(n - 1) + F(n - 3));
        }
        if (k == 1) {
            if (fl.botSided) return (F(1) - F(0)) * (a1 * c.o);
            const double base = a06 * (F(1) - F(0));
            return fl.botEven ? base + b06 * (F(2) - F(0)) : base + b06 * (F(2) + F(0));
        }
        return a06 * (F(k) - F(k - 1)) + b06 * (F(k + 1) - F(k - 2));
    } else if (OP == SNP_D1_C2C) {                 // D1RHS_C2C_common.F90 (top rows are written last: they win on the shortest lines)
        const double a06 = ((14.0 / 9.0) / 2.0) * c.o, b06 = ((1.0 / 9.0) / 4.0) * c.o;
        const double a_np_3 = c.w3 * c.q_pp * c.o, b_np_3 = c.w3 * c.r_pp * c.o, a_np_2 = c.w2 * c.q_p * c.o;
        const double a_np_1 = c.w1 * (-c.p * c.o), b_np_1 = c.w1 * (c.q * c.o), c_np_1 = c.w1 * (c.r * c.o), d_np_1 = c.w1 * (c.s * c.o);
        if (k >= n - 3) {
            const int m = n - 1 - k;   // 0 = last row
            if (fl.topSided) {
                if (m == 2) return a_np_3 * (F(n - 2) - F(n - 4)) + b_np_3 * (F(n - 1) - F(n - 5));
                if (m == 1) return a_np_2 * (F(n - 1) - F(n - 3));
                return -a_np_1 * F(n - 1) - b_np_1 * F(n - 2) - c_np_1 * F(n - 3) - d_np_1 * F(n - 4);
            }
            if (m == 2) return b06 * (F(n - 1) - F(n - 5)) + a06 * (F(n - 2) - F(n - 4));
            if (fl.topEven) {
                if (m == 1) return b06 * (F(n - 1) - F(n - 4)) + a06 * (F(n - 1) - F(n - 3));
                return b06 * (F(n - 2) - F(n - 3)) + a06 * (F(n - 1) - F(n - 2));
            }
            if (m == 1) return -b06 * (F(n - 1) + F(n - 4)) + a06 * (F(n - 1) - F(n - 3));
            return -b06 * (F(n - 2) + F(n - 3)) - a06 * (F(n - 1) + F(n - 2));
        }
        if (k <= 2) {
            if (fl.botSided) {
                if (k == 0) return a_np_1 * F(0) + b_np_1 * F(1) + c_np_1 * F(2) + d_np_1 * F(3);
                if (k == 1) return a_np_2 * (F(2) - F(0));
                return a_np_3 * (F(3) - F(1)) + b_np_3 * (F(4) - F(0));
            }
            if (k == 2) return b06 * (F(4) - F(0)) + a06 * (F(3) - F(1));
            if (fl.botEven) {
                if (k == 0) return b06 * (F(2) - F(1)) + a06 * (F(1) - F(0));
                return b06 * (F(3) - F(0)) + a06 * (F(2) - F(0));
            }
            if (k == 0) return b06 * (F(2) + F(1)) + a06 * (F(1) + F(0));
            return b06 * (F(3) + F(0)) + a06 * (F(2) - F(0));
        }
        return b06 * (F(k + 2) - F(k - 2)) + a06 * (F(k + 1) - F(k - 1));
    } else if (OP == SNP_D1_E2E) {                 // D1RHS_E2E_common.F90 (no one-sided variant in the reference)
        const double a06 = ((14.0 / 9.0) / 2.0) * c.o, b06 = ((1.0 / 9.0) / 4.0) * c.o;
        if (k == nE - 1) return fl.topEven ? 0.0 : -b06 * (F(nE - 3) + F(nE - 3)) - a06 * (F(nE - 2) + F(nE - 2));
        if (k == 0) return fl.botEven ? 0.0 : b06 * (F(2) + F(2)) + a06 * (F(1) + F(1));
        const double base = a06 * (F(k + 1) - F(k - 1));
        if (k == nE - 2) return fl.topEven ? base + b06 * (F(nE - 2) - F(nE - 4)) : base - b06 * (F(nE - 2) + F(nE - 4));
        if (k == 1) return fl.botEven ? base + b06 * (F(3) - F(1)) : base + b06 * (F(3) + F(1));
        return base + b06 * (F(k + 2) - F(k - 2));
    } else if (OP == SNP_INTERP_E2C) {             // InterpRHS_E2C_common.F90
        const double b = (1.0 / 10.0) / 2.0, a = (3.0 / 2.0) / 2.0, al0 = 1.0;
        const double a0 = (1.0 / 16.0) * (5.0 - al0), b0 = (1.0 / 16.0) * (9.0 * al0 + 15.0);
        const double c0 = (1.0 / 16.0) * (9.0 * al0 - 5.0), d0 = (1.0 / 16.0) * (1.0 - al0);
        if (k == n - 1) {
            if (fl.topSided) return a0 * F(nE - 1) + b0 * F(nE - 2) + c0 * F(nE - 3) + d0 * F(nE - 4);
            if (fl.topEven) return (b) * (F(nE - 2) + F(nE - 3)) + (a) * (F(nE - 1) + F(nE - 2));
            return (b) * (-F(nE - 2) + F(nE - 3)) + (a) * (F(nE - 1) + F(nE - 2));
        }
        if (k == 0) {
            if (fl.botSided) return a0 * F(0) + b0 * F(1) + c0 * F(2) + d0 * F(3);
            if (fl.botEven) return (b) * (F(2) + F(1)) + (a) * (F(1) + F(0));
            return (b) * (F(2) - F(1)) + (a) * (F(1) + F(0));
        }
        return (b) * (F(k + 2) + F(k - 1)) + (a) * (F(k + 1) + F(k));
    } else if (OP == SNP_INTERP_C2E) {             // InterpRHS_C2E_common.F90
        const double b = (1.0 / 10.0) / 2.0, a = (3.0 / 2.0) / 2.0;
        const double a0 = 15.0 / 8.0, b0 = -5.0 / 4.0, c0 = 3.0 / 8.0, al1 = 1.0 / 6.0;
        const double a1 = (1.0 / 8.0) * (9.0 + 10.0 * al1);
        if (k == nE - 1) {
            if (fl.topSided) return a0 * F(n - 1) + b0 * F(n - 2) + c0 * F(n - 3);
            if (fl.topEven) return 2.0 * b * F(n - 2) + 2.0 * a * F(n - 1);
            return 0.0;
        }
        if (k == 0) {
            if (fl.botSided) return a0 * F(0) + b0 * F(1) + c0 * F(2);
            if (fl.botEven) return 2.0 * b * F(1) + 2.0 * a * F(0);
            return 0.0;
        }
        if (k == nE - 2) {
            if (fl.topSided) return (a1 / 2.0) * (F(n - 1) + F(n - 2));
            const double base = (a) * (F(k) + F(k - 1));
            return fl.topEven ? base + (b) * (F(n - 1) + F(n - 3)) : base + (b) * (-F(n - 1) + F(n - 3));
        }
        if (k == 1) {
            if (fl.botSided) return (a1 / 2.0) * (F(1) + F(0));
            const double base = (a) * (F(1) + F(0));
            return fl.botEven ? base + (b) * (F(2) + F(0)) : base + (b) * (F(2) - F(0));
        }
        return (a) * (F(k) + F(k - 1)) + (b) * (F(k + 1) + F(k - 2));
    } else if (OP == SNP_D2_C2C) {                 // D2RHS_C2C_common.F90
        const double a06 = (12.0 / 11.0) * c.o2, b06 = ((3.0 / 11.0) / 4.0) * c.o2;
        const double diag = 2.0 * (b06 + a06) * F(k);
        if (k == n - 1)
            return (fl.topEven ? b06 * (F(n - 2) + F(n - 3)) + a06 * (F(n - 1) + F(n - 2)) : b06 * (-F(n - 2) + F(n - 3)) + a06 * (-F(n - 1) + F(n - 2))) - diag;
        if (k == 0)
            return (fl.botEven ? b06 * (F(2) + F(1)) + a06 * (F(1) + F(0)) : b06 * (F(2) - F(1)) + a06 * (F(1) - F(0))) - diag;
        const double base = a06 * (F(k + 1) + F(k - 1));
        if (k == n - 2) return (fl.topEven ? base + b06 * (F(n - 1) + F(n - 4)) : base + b06 * (-F(n - 1) + F(n - 4))) - diag;
        if (k == 1) return (fl.botEven ? base + b06 * (F(3) + F(0)) : base + b06 * (F(3) - F(0))) - diag;
        return (base + b06 * (F(k + 2) + F(k - 2))) - diag;
    } else {                                       // D2RHS_E2E_common.F90
        const double a06 = (12.0 / 11.0) * c.o2, b06 = ((3.0 / 11.0) / 4.0) * c.o2;
        const double diag = 2.0 * (b06 + a06) * F(k);
        if (k == nE - 1) return (fl.topEven ? 2.0 * b06 * F(nE - 3) + 2.0 * a06 * F(nE - 2) : 0.0) - diag;
        if (k == 0) return (fl.botEven ? b06 * (F(2) + F(2)) + a06 * (F(1) + F(1)) : 0.0) - diag;
        const double base = a06 * (F(k + 1) + F(k - 1));
        if (k == nE - 2) return (fl.topEven ? base + b06 * (F(nE - 2) + F(nE - 4)) : base + b06 * (-F(nE - 2) + F(nE - 4))) - diag;
        if (k == 1) return (fl.botEven ? base + b06 * (F(3) + F(1)) : base + b06 * (F(3) - F(1))) - diag;
        return (base + b06 * (F(k + 2) + F(k - 2))) - diag;
    }
}

// One line: RHS formed inside the forward sweep of SolveZTriREAL (TridiagSolver_allRoutines.F90:1-17), then the backward sweep.
// in / out: element j of the line at in[j * es] / out[j * es]; tab = t1[nout] (ddn*den), t2[nout] (den), t3[nout] (cp).
template <int OP>
PDO_HD void snp_line(const double* in, double* out, long long es, int n, const StaggNpFlags& fl, const StaggNpConst& c, const double* tab) {
    const int nout = snp_rows_out(OP, n);
    const double *t1 = tab, *t2 = tab + nout, *t3 = tab + 2 * (long long)nout;
    auto F = [&](int j) -> double { return in[(long long)j * es]; };
    double prev = snp_rhs<OP>(0, n, fl, c, F) * t2[0];
    out[0] = prev;
    for (int k = 1; k < nout; ++k) {
        const double v = snp_rhs<OP>(k, n, fl, c, F) * t2[k] - prev * t1[k];
        out[(long long)k * es] = v;
        prev = v;
    }
    for (int k = nout - 2; k >= 0; --k) {
        const double v = out[(long long)k * es] - t3[k] * prev;
        out[(long long)k * es] = v;
        prev = v;
    }
}

// ---- host side ----
// the eight tridiagonal systems (ComputeTri_allRoutines.F90): rows ddn[nout] dg[nout] dup[nout]; returns 0 or 21 (n <= 4)
int snp_build_rows(int op, int n, const StaggNpFlags& fl, double* rows3n);
// Thomas factors t1 = ddn*den, t2 = den, t3 = cp (same recurrences as the reference's cp / den loops)
int snp_build_table(int op, int n, const StaggNpFlags& fl, double* tab3n);
StaggNpConst snp_constants(double dx);

struct StaggNp {
    int n = 0;
    StaggNpFlags fl{};
    StaggNpConst co{};
    double* d_tab[SNP_COUNT] = {};
};
cudaError_t snp_create(StaggNp* h, int n, double dx, const StaggNpFlags& fl, int* ierr_out);
void snp_destroy(StaggNp* h);
// in(ncols, rows_in) -> out(ncols, rows_out): ncols = n1 * n2 (x 2 for complex data); device pointers, no aliasing
cudaError_t snp_apply(const StaggNp* h, int op, const double* in, double* out, long long ncols, cudaStream_t st);
// the same arithmetic on the host (test hook only)
int snp_apply_host(int op, int n, double dx, const StaggNpFlags& fl, const double* in, double* out, long long ncols);

}  // namespace pdo
