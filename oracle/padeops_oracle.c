/*
 * padeops_oracle.c — CPU restatement of the PadeOps operator hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker / the timed CPU arm.  The product path (padeops_b200/) never links or
 * calls it and fails loudly when its CUDA library is missing.
 *
 * Each routine follows the Fortran loop nest it cites (paths relative to /root/reference/src,
 * "2D»" = dependencies/2decomp_fft-1.5.847.tar.gz » 2decomp_fft/src).  Same sweeps, same operand
 * order inside each expression, column-major arrays (first index fastest) so the same buffers can
 * be handed to the CUDA library.  The reference itself is Fortran 2003 + MPI and cannot be compiled
 * in the build image (no Fortran compiler, no MPI), and it ships no golden vectors: the oracle is
 * pinned against the reference's own known-answer tests (analytic fields, transfer functions,
 * modified wavenumbers, manufactured Poisson solution) and against independent dense cyclic
 * solves — see tests/test_oracle_*.py and DESIGN.md §3.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX2(i, c, n) ((size_t)(c) * (size_t)(n) + (size_t)(i)) /* LU(i,c), 0-based, column-major */

/* ------------------------------------------------------------------------------------------ */
/* Coefficients (derivatives/cd10.F90:16-27, cd06.F90:14-21, cd06stagg.F90:174-176,307,411,532;  */
/* filters/cf90.F90:16-22, gaussian.F90:16-20)                                                  */
/* ------------------------------------------------------------------------------------------ */
static const double alpha10d1 = 1.0 / 2.0, beta10d1 = 1.0 / 20.0;
static const double a10d1 = (17.0 / 12.0) / 2.0, b10d1 = (101.0 / 150.0) / 4.0, c10d1 = (1.0 / 100.0) / 6.0;
static const double alpha10d2 = 334.0 / 899.0, beta10d2 = 43.0 / 1798.0;
static const double a10d2 = (1065.0 / 1798.0) / 1.0, b10d2 = (1038.0 / 899.0) / 4.0, c10d2 = (79.0 / 1798.0) / 9.0;
static const double alpha06d1 = 1.0 / 3.0, a06d1 = (14.0 / 9.0) / 2.0, b06d1 = (1.0 / 9.0) / 4.0;
static const double alpha90 = 6.6624e-1, beta90 = 1.6688e-1;
static const double a90 = 9.9965e-1, b90 = 6.6652e-1, c90 = 1.6674e-1, d90 = 4.0e-5, e90 = -5.0e-6;
static const double agf = 3565.0 / 10368.0, bgf = 3091.0 / 12960.0, cgf = 1997.0 / 25920.0,
                    dgf = 149.0 / 12960.0, egf = 107.0 / 103680.0;

/* ------------------------------------------------------------------------------------------ */
/* Cyclic pentadiagonal LU — derivatives/cd10.F90:351-427 (same routine filters/cf90.F90:198-274) */
/* LU(n,9) columns: 0 b, 1 eg, 2 k, 3 l, 4 1/g, 5 h, 6 ff, 7 v, 8 w                              */
/* ------------------------------------------------------------------------------------------ */
void pdo_oracle_penta_lu(int n, double e, double a, double d, double c, double f, double *LU)
{
    memset(LU, 0, sizeof(double) * 9 * (size_t)n);
    double *b = LU + IDX2(0, 0, n), *eg = LU + IDX2(0, 1, n), *k = LU + IDX2(0, 2, n);
    double *l = LU + IDX2(0, 3, n), *g = LU + IDX2(0, 4, n), *h = LU + IDX2(0, 5, n);
    double *ff = LU + IDX2(0, 6, n), *v = LU + IDX2(0, 7, n), *w = LU + IDX2(0, 8, n);
    int i;
    double s;
#define F1(x) ((x) - 1) /* Fortran 1-based index → C */
    /* Step 1 */
    g[F1(1)] = d;
    b[F1(2)] = a / g[F1(1)];
    h[F1(1)] = c;
    k[F1(1)] = f / g[F1(1)];
    w[F1(1)] = a;
    v[F1(1)] = e;
    l[F1(1)] = c / g[F1(1)];
    g[F1(2)] = d - b[F1(2)] * h[F1(1)];
    k[F1(2)] = -k[F1(1)] * h[F1(1)] / g[F1(2)];
    w[F1(2)] = e - b[F1(2)] * w[F1(1)];
    v[F1(2)] = -b[F1(2)] * v[F1(1)];
    l[F1(2)] = (f - l[F1(1)] * h[F1(1)]) / g[F1(2)];
    h[F1(2)] = c - b[F1(2)] * f;
    /* Step 2 */
    for (i = 3; i <= n - 3; ++i) {
        b[F1(i)] = (a - (e / g[F1(i - 2)]) * h[F1(i - 2)]) / g[F1(i - 1)];
        h[F1(i)] = c - b[F1(i)] * f;
        g[F1(i)] = d - (e / g[F1(i - 2)]) * f - b[F1(i)] * h[F1(i - 1)];
    }
    /* Step 3 */
    b[F1(n - 2)] = (a - (e / g[F1(n - 4)]) * h[F1(n - 4)]) / g[F1(n - 3)];
    g[F1(n - 2)] = d - (e / g[F1(n - 4)]) * f - b[F1(n - 2)] * h[F1(n - 3)];
    /* Step 4 */
    for (i = 3; i <= n - 4; ++i) {
        k[F1(i)] = -(k[F1(i - 2)] * f + k[F1(i - 1)] * h[F1(i - 1)]) / g[F1(i)];
        v[F1(i)] = -(e / g[F1(i - 2)]) * v[F1(i - 2)] - b[F1(i)] * v[F1(i - 1)];
    }
    /* Step 5 */
    k[F1(n - 3)] = (e - k[F1(n - 5)] * f - k[F1(n - 4)] * h[F1(n - 4)]) / g[F1(n - 3)];
    k[F1(n - 2)] = (a - k[F1(n - 4)] * f - k[F1(n - 3)] * h[F1(n - 3)]) / g[F1(n - 2)];
    v[F1(n - 3)] = f - (e / g[F1(n - 5)]) * v[F1(n - 5)] - b[F1(n - 3)] * v[F1(n - 4)];
    v[F1(n - 2)] = c - (e / g[F1(n - 4)]) * v[F1(n - 4)] - b[F1(n - 2)] * v[F1(n - 3)];
    s = 0.0;
    for (i = 1; i <= n - 2; ++i) s += k[F1(i)] * v[F1(i)];
    g[F1(n - 1)] = d - s;
    /* Step 6 */
    for (i = 3; i <= n - 3; ++i) {
        w[F1(i)] = -(e / g[F1(i - 2)]) * w[F1(i - 2)] - b[F1(i)] * w[F1(i - 1)];
        l[F1(i)] = -(l[F1(i - 2)] * f + l[F1(i - 1)] * h[F1(i - 1)]) / g[F1(i)];
    }
    /* Step 7 */
    w[F1(n - 2)] = f - (e / g[F1(n - 4)]) * w[F1(n - 4)] - b[F1(n - 2)] * w[F1(n - 3)];
    s = 0.0;
    for (i = 1; i <= n - 2; ++i) s += k[F1(i)] * w[F1(i)];
    w[F1(n - 1)] = c - s;
    l[F1(n - 2)] = (e - l[F1(n - 4)] * f - l[F1(n - 3)] * h[F1(n - 3)]) / g[F1(n - 2)];
    s = 0.0;
    for (i = 1; i <= n - 2; ++i) s += l[F1(i)] * v[F1(i)];
    l[F1(n - 1)] = (a - s) / g[F1(n - 1)];
    s = 0.0;
    for (i = 1; i <= n - 1; ++i) s += l[F1(i)] * w[F1(i)];
    g[F1(n)] = d - s;
    /* eg(3:n-2) = e/g(1:n-4); ff(1:n-4) = f; g = 1/g */
    for (i = 3; i <= n - 2; ++i) eg[F1(i)] = e / g[F1(i - 2)];
    for (i = 1; i <= n - 4; ++i) ff[F1(i)] = f;
    for (i = 1; i <= n; ++i) g[F1(i)] = 1.0 / g[F1(i)];
}

/* ------------------------------------------------------------------------------------------ */
/* Cyclic tridiagonal LU — derivatives/cd06.F90:221-262 (= cd06stagg.F90:631-672)                */
/* LU(n,5) columns: 0 b/c(i-1), 1 h, 2 1/c, 3 aa/c, 4 v/c                                        */
/* ------------------------------------------------------------------------------------------ */
void pdo_oracle_tri_lu(int n, double b, double d, double a, double *LU)
{
    memset(LU, 0, sizeof(double) * 5 * (size_t)n);
    double *bc = LU + IDX2(0, 0, n), *h = LU + IDX2(0, 1, n), *c = LU + IDX2(0, 2, n);
    double *aa = LU + IDX2(0, 3, n), *v = LU + IDX2(0, 4, n);
    int i;
    double s;
    c[F1(1)] = d;
    v[F1(1)] = b;
    h[F1(1)] = a / c[F1(1)];
    for (i = 2; i <= n - 1; ++i) c[F1(i)] = d - (b / c[F1(i - 1)]) * a;
    for (i = 2; i <= n - 2; ++i) {
        v[F1(i)] = -(b / c[F1(i - 1)]) * v[F1(i - 1)];
        h[F1(i)] = -(a / c[F1(i)]) * h[F1(i - 1)];
    }
    v[F1(n - 1)] = a - (b / c[F1(n - 2)]) * v[F1(n - 2)];
    h[F1(n - 1)] = (b - h[F1(n - 2)] * a) / c[F1(n - 1)];
    s = 0.0;
    for (i = 1; i <= n - 1; ++i) s += h[F1(i)] * v[F1(i)];
    c[F1(n)] = d - s;
    for (i = 2; i <= n - 1; ++i) bc[F1(i)] = b / c[F1(i - 1)];
    for (i = 1; i <= n - 2; ++i) aa[F1(i)] = a;
    for (i = 1; i <= n; ++i) c[F1(i)] = 1.0 / c[F1(i)];
    for (i = 1; i <= n; ++i) aa[F1(i)] = aa[F1(i)] * c[F1(i)];
    for (i = 1; i <= n; ++i) v[F1(i)] = v[F1(i)] * c[F1(i)];
}

/* ------------------------------------------------------------------------------------------ */
/* Penta solves.  X: cd10.F90:712-749 (scalar recurrence per line).  Y/Z: cd10.F90:751-821        */
/* (array statements over the contiguous index; Z = Y with n1 := n1*n2, n3 := 1, including the    */
/* plane-sized sum1/sum2 temporaries).  CF90's SolveX/Y/ZLU (cf90.F90:421-530) are identical.     */
/* ------------------------------------------------------------------------------------------ */
#define LUc(i, c) LU[IDX2((i) - 1, (c) - 1, n)] /* LU(i,c) with Fortran indices */

void pdo_oracle_penta_solve_x(const double *LU, int n, double *y, int64_t nlines)
{
    for (int64_t ln = 0; ln < nlines; ++ln) {
        double *Y = y + ln * n - 1; /* Y[i], i = 1..n */
        double sum1, sum2;
        int i;
        Y[2] = Y[2] - LUc(2, 1) * Y[1];
        sum1 = LUc(1, 3) * Y[1] + LUc(2, 3) * Y[2];
        sum2 = LUc(1, 4) * Y[1] + LUc(2, 4) * Y[2];
        for (i = 3; i <= n - 2; ++i) {
            Y[i] = Y[i] - LUc(i, 1) * Y[i - 1] - LUc(i, 2) * Y[i - 2];
            sum1 = sum1 + LUc(i, 3) * Y[i];
            sum2 = sum2 + LUc(i, 4) * Y[i];
        }
        Y[n - 1] = Y[n - 1] - sum1;
        Y[n] = (Y[n] - sum2 - LUc(n - 1, 4) * Y[n - 1]) * LUc(n, 5);
        Y[n - 1] = (Y[n - 1] - LUc(n - 1, 9) * Y[n]) * LUc(n - 1, 5);
        Y[n - 2] = (Y[n - 2] - LUc(n - 2, 8) * Y[n - 1] - LUc(n - 2, 9) * Y[n]) * LUc(n - 2, 5);
        Y[n - 3] = (Y[n - 3] - LUc(n - 3, 6) * Y[n - 2] - LUc(n - 3, 8) * Y[n - 1] - LUc(n - 3, 9) * Y[n]) * LUc(n - 3, 5);
        for (i = n - 4; i >= 1; --i)
            Y[i] = (Y[i] - LUc(i, 6) * Y[i + 1] - LUc(i, 7) * Y[i + 2] - LUc(i, 8) * Y[n - 1] - LUc(i, 9) * Y[n]) * LUc(i, 5);
    }
}

/* y(n1, n, n3): solve along the middle index, vectorised over the first */
void pdo_oracle_penta_solve_y(const double *LU, int n, double *y, int64_t n1, int64_t n3)
{
    double *sum1 = (double *)malloc(sizeof(double) * (size_t)n1);
    double *sum2 = (double *)malloc(sizeof(double) * (size_t)n1);
    for (int64_t k = 0; k < n3; ++k) {
        double *P = y + k * n1 * n;
#define ROW(j) (P + ((int64_t)(j) - 1) * n1)
        int64_t i;
        int j;
        {
            double *y1 = ROW(1), *y2 = ROW(2);
            const double b2 = LUc(2, 1), k1 = LUc(1, 3), k2 = LUc(2, 3), l1 = LUc(1, 4), l2 = LUc(2, 4);
            for (i = 0; i < n1; ++i) y2[i] = y2[i] - b2 * y1[i];
            for (i = 0; i < n1; ++i) sum1[i] = k1 * y1[i] + k2 * y2[i];
            for (i = 0; i < n1; ++i) sum2[i] = l1 * y1[i] + l2 * y2[i];
        }
        for (j = 3; j <= n - 2; ++j) {
            double *yj = ROW(j), *ym1 = ROW(j - 1), *ym2 = ROW(j - 2);
            const double bj = LUc(j, 1), ej = LUc(j, 2), kj = LUc(j, 3), lj = LUc(j, 4);
            for (i = 0; i < n1; ++i) yj[i] = yj[i] - bj * ym1[i] - ej * ym2[i];
            for (i = 0; i < n1; ++i) sum1[i] = sum1[i] + kj * yj[i];
            for (i = 0; i < n1; ++i) sum2[i] = sum2[i] + lj * yj[i];
        }
        {
            double *yn = ROW(n), *yn1 = ROW(n - 1), *yn2 = ROW(n - 2), *yn3 = ROW(n - 3);
            const double ln1 = LUc(n - 1, 4), gn = LUc(n, 5), wn1 = LUc(n - 1, 9), gn1 = LUc(n - 1, 5);
            const double vn2 = LUc(n - 2, 8), wn2 = LUc(n - 2, 9), gn2 = LUc(n - 2, 5);
            const double hn3 = LUc(n - 3, 6), vn3 = LUc(n - 3, 8), wn3 = LUc(n - 3, 9), gn3 = LUc(n - 3, 5);
            for (i = 0; i < n1; ++i) yn1[i] = yn1[i] - sum1[i];
            for (i = 0; i < n1; ++i) yn[i] = (yn[i] - sum2[i] - ln1 * yn1[i]) * gn;
            for (i = 0; i < n1; ++i) yn1[i] = (yn1[i] - wn1 * yn[i]) * gn1;
            for (i = 0; i < n1; ++i) yn2[i] = (yn2[i] - vn2 * yn1[i] - wn2 * yn[i]) * gn2;
            for (i = 0; i < n1; ++i) yn3[i] = (yn3[i] - hn3 * yn2[i] - vn3 * yn1[i] - wn3 * yn[i]) * gn3;
            for (j = n - 4; j >= 1; --j) {
                double *yj = ROW(j), *yp1 = ROW(j + 1), *yp2 = ROW(j + 2);
                const double hj = LUc(j, 6), fj = LUc(j, 7), vj = LUc(j, 8), wj = LUc(j, 9), gj = LUc(j, 5);
                for (i = 0; i < n1; ++i)
                    yj[i] = (yj[i] - hj * yp1[i] - fj * yp2[i] - vj * yn1[i] - wj * yn[i]) * gj;
            }
        }
#undef ROW
    }
    free(sum1);
    free(sum2);
}

/* ------------------------------------------------------------------------------------------ */
/* Tri solves.  X: cd06.F90:345-373.  Y/Z: cd06.F90:375-429, cd06stagg.F90:248-299.              */
/* ------------------------------------------------------------------------------------------ */
void pdo_oracle_tri_solve_x(const double *LU, int n, double *y, int64_t nlines)
{
    for (int64_t ln = 0; ln < nlines; ++ln) {
        double *Y = y + ln * n - 1;
        double sum1;
        int i;
        sum1 = LUc(1, 2) * Y[1];
        for (i = 2; i <= n - 1; ++i) {
            Y[i] = Y[i] - LUc(i, 1) * Y[i - 1];
            sum1 = sum1 + LUc(i, 2) * Y[i];
        }
        Y[n] = Y[n] - sum1;
        Y[n] = Y[n] * LUc(n, 3);
        Y[n - 1] = Y[n - 1] * LUc(n - 1, 3) - Y[n] * LUc(n - 1, 5);
        for (i = n - 2; i >= 1; --i) Y[i] = Y[i] * LUc(i, 3) - Y[i + 1] * LUc(i, 4) - Y[n] * LUc(i, 5);
    }
}

/* y(n1, n, n3) with row stride rs (= n1 normally) and plane stride ps: lets the staggered ops  */
/* solve the first n planes of an (n1,n2,n+1) array exactly like the reference's slices do.     */
void pdo_oracle_tri_solve_y(const double *LU, int n, double *y, int64_t n1, int64_t n3)
{
    double *sum1 = (double *)malloc(sizeof(double) * (size_t)n1);
    for (int64_t k = 0; k < n3; ++k) {
        double *P = y + k * n1 * n;
#define ROW(j) (P + ((int64_t)(j) - 1) * n1)
        int64_t i;
        int j;
        {
            double *y1 = ROW(1);
            const double h1 = LUc(1, 2);
            for (i = 0; i < n1; ++i) sum1[i] = h1 * y1[i];
        }
        for (j = 2; j <= n - 1; ++j) {
            double *yj = ROW(j), *ym1 = ROW(j - 1);
            const double bj = LUc(j, 1), hj = LUc(j, 2);
            for (i = 0; i < n1; ++i) yj[i] = yj[i] - bj * ym1[i];
            for (i = 0; i < n1; ++i) sum1[i] = sum1[i] + hj * yj[i];
        }
        {
            double *yn = ROW(n), *yn1 = ROW(n - 1);
            const double cn = LUc(n, 3), cn1 = LUc(n - 1, 3), vn1 = LUc(n - 1, 5);
            for (i = 0; i < n1; ++i) yn[i] = yn[i] - sum1[i];
            for (i = 0; i < n1; ++i) yn[i] = yn[i] * cn;
            for (i = 0; i < n1; ++i) yn1[i] = yn1[i] * cn1 - yn[i] * vn1;
            for (j = n - 2; j >= 1; --j) {
                double *yj = ROW(j), *yp1 = ROW(j + 1);
                const double cj = LUc(j, 3), aj = LUc(j, 4), vj = LUc(j, 5);
                for (i = 0; i < n1; ++i) yj[i] = yj[i] * cj - yp1[i] * aj - yn[i] * vj;
            }
        }
#undef ROW
    }
    free(sum1);
}

/* ------------------------------------------------------------------------------------------ */
/* Periodic RHS stencils.  The Fortran writes the wrapped rows out by hand (cd10.F90:1118-1142,  */
/* 1265-1308, 1428-1468, 1601-1635; cd06.F90:515-550, 596-630, 674-704; cf90.F90:610-669;       */
/* gaussian.F90:137-187); every row is the same expression with indices taken mod n, which is   */
/* what W() does.  Operand order inside the expression is the Fortran's.                         */
/* Layout: f(n1, n, n3); x-direction = (n1=1, n, n3=nlines).                                     */
/* ------------------------------------------------------------------------------------------ */
enum { RHS_D1_7 = 0, RHS_D2_7 = 1, RHS_D1_5 = 2, RHS_SYM_9 = 3, RHS_D2_5 = 4 };

static inline int wrapi(int j, int n) { return j < 0 ? j + n : (j >= n ? j - n : j); }

void pdo_oracle_rhs(int kind, const double *co, const double *f, double *r, int64_t n1, int n, int64_t n3)
{
    for (int64_t k = 0; k < n3; ++k) {
        const double *F = f + k * n1 * n;
        double *R = r + k * n1 * n;
        for (int j = 0; j < n; ++j) {
#define W(o) (F + (int64_t)wrapi(j + (o), n) * n1)
            double *rj = R + (int64_t)j * n1;
            int64_t i;
            if (kind == RHS_D1_7) {
                const double a = co[0], b = co[1], c = co[2];
                const double *p1 = W(1), *m1 = W(-1), *p2 = W(2), *m2 = W(-2), *p3 = W(3), *m3 = W(-3);
                for (i = 0; i < n1; ++i) rj[i] = a * (p1[i] - m1[i]) + b * (p2[i] - m2[i]) + c * (p3[i] - m3[i]);
            } else if (kind == RHS_D2_7) {
                const double a = co[0], b = co[1], c = co[2];
                const double *p0 = W(0), *p1 = W(1), *m1 = W(-1), *p2 = W(2), *m2 = W(-2), *p3 = W(3), *m3 = W(-3);
                for (i = 0; i < n1; ++i)
                    rj[i] = a * (p1[i] - 2.0 * p0[i] + m1[i]) + b * (p2[i] - 2.0 * p0[i] + m2[i]) + c * (p3[i] - 2.0 * p0[i] + m3[i]);
            } else if (kind == RHS_D1_5) {
                const double a = co[0], b = co[1];
                const double *p1 = W(1), *m1 = W(-1), *p2 = W(2), *m2 = W(-2);
                for (i = 0; i < n1; ++i) rj[i] = a * (p1[i] - m1[i]) + b * (p2[i] - m2[i]);
            } else if (kind == RHS_D2_5) {
                const double a = co[0], b = co[1];
                const double *p0 = W(0), *p1 = W(1), *m1 = W(-1), *p2 = W(2), *m2 = W(-2);
                for (i = 0; i < n1; ++i)
                    rj[i] = a * (p1[i] - 2.0 * p0[i] + m1[i]) + b * (p2[i] - 2.0 * p0[i] + m2[i]);
            } else { /* RHS_SYM_9 */
                const double a = co[0], b = co[1], c = co[2], d = co[3], e = co[4];
                const double *p0 = W(0), *p1 = W(1), *m1 = W(-1), *p2 = W(2), *m2 = W(-2), *p3 = W(3), *m3 = W(-3),
                             *p4 = W(4), *m4 = W(-4);
                for (i = 0; i < n1; ++i)
                    rj[i] = a * (p0[i]) + b * (p1[i] + m1[i]) + c * (p2[i] + m2[i]) + d * (p3[i] + m3[i]) + e * (p4[i] + m4[i]);
            }
#undef W
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Operator entry points (type-bound dd1/dd2/dd3, d2d1..3, filter1..3)                           */
/* axis: 0 = x (f(n,na,nb)), 1 = y (f(na,n,nb)), 2 = z (f(na,nb,n))                              */
/* cd10.F90:2029-2447, cd06.F90:775-839, cf90.F90:1020-1228, gaussian.F90:104-187,336,564        */
/* ------------------------------------------------------------------------------------------ */
static void shape_for_axis(int axis, int n, int64_t na, int64_t nb, int64_t *n1, int64_t *n3)
{
    (void)n;
    if (axis == 0) { *n1 = 1; *n3 = na * nb; }
    else if (axis == 1) { *n1 = na; *n3 = nb; }
    else { *n1 = na * nb; *n3 = 1; }
}

/* which: 1 = first derivative (dd*), 2 = second derivative (d2d*).  LU is the matching table. */
int pdo_oracle_cd10(const double *LU, int n, double dx, int which, int axis, const double *f, double *df, int64_t na, int64_t nb)
{
    int64_t n1, n3, tot = (int64_t)n * na * nb;
    if (n == 1) { memset(df, 0, sizeof(double) * (size_t)tot); return 0; }
    const double onebydx = 1.0 / dx, onebydx2 = onebydx / dx;
    double co[3];
    if (which == 1) { co[0] = a10d1 * onebydx; co[1] = b10d1 * onebydx; co[2] = c10d1 * onebydx; }
    else { co[0] = a10d2 * onebydx2; co[1] = b10d2 * onebydx2; co[2] = c10d2 * onebydx2; }
    shape_for_axis(axis, n, na, nb, &n1, &n3);
    pdo_oracle_rhs(which == 1 ? RHS_D1_7 : RHS_D2_7, co, f, df, n1, n, n3);
    if (axis == 0) pdo_oracle_penta_solve_x(LU, n, df, n3);
    else pdo_oracle_penta_solve_y(LU, n, df, n1, n3);
    return 0;
}

int pdo_oracle_cd10_lu(int n, int which, double *LU)
{
    if (n >= 8) {
        if (which == 1) pdo_oracle_penta_lu(n, beta10d1, alpha10d1, 1.0, alpha10d1, beta10d1, LU);
        else pdo_oracle_penta_lu(n, beta10d2, alpha10d2, 1.0, alpha10d2, beta10d2, LU);
        return 0;
    }
    if (n == 1) { for (int i = 0; i < 9; ++i) LU[i] = 1.0; return 0; }
    return 2; /* cd10.F90:224 */
}

int pdo_oracle_cd06_lu(int n, double *LU)
{
    if (n >= 6) { pdo_oracle_tri_lu(n, alpha06d1, 1.0, alpha06d1, LU); return 0; }
    if (n == 1) { for (int i = 0; i < 5; ++i) LU[i] = 1.0; return 0; }
    return 3; /* cd06.F90:158 */
}

int pdo_oracle_cd06(const double *LU, int n, double dx, int axis, const double *f, double *df, int64_t na, int64_t nb)
{
    int64_t n1, n3, tot = (int64_t)n * na * nb;
    if (n == 1) { memset(df, 0, sizeof(double) * (size_t)tot); return 0; }
    const double onebydx = 1.0 / dx;
    double co[2] = { a06d1 * onebydx, b06d1 * onebydx };
    shape_for_axis(axis, n, na, nb, &n1, &n3);
    pdo_oracle_rhs(RHS_D1_5, co, f, df, n1, n, n3);
    if (axis == 0) pdo_oracle_tri_solve_x(LU, n, df, n3);
    else pdo_oracle_tri_solve_y(LU, n, df, n1, n3);
    return 0;
}

int pdo_oracle_cf90_lu(int n, double *LU)
{
    if (n >= 10) { pdo_oracle_penta_lu(n, beta90, alpha90, 1.0, alpha90, beta90, LU); return 0; }
    if (n == 1) { for (int i = 0; i < 9; ++i) LU[i] = 1.0; return 0; }
    return 7; /* cf90.F90:128 */
}

int pdo_oracle_cf90(const double *LU, int n, int axis, const double *f, double *out, int64_t na, int64_t nb)
{
    int64_t n1, n3, tot = (int64_t)n * na * nb;
    if (n == 1) { memcpy(out, f, sizeof(double) * (size_t)tot); return 0; }
    const double co[5] = { a90, b90, c90, d90, e90 };
    shape_for_axis(axis, n, na, nb, &n1, &n3);
    pdo_oracle_rhs(RHS_SYM_9, co, f, out, n1, n, n3);
    if (axis == 0) pdo_oracle_penta_solve_x(LU, n, out, n3);
    else pdo_oracle_penta_solve_y(LU, n, out, n1, n3);
    return 0;
}

int pdo_oracle_gaussian(int n, int axis, const double *f, double *out, int64_t na, int64_t nb)
{
    int64_t n1, n3, tot = (int64_t)n * na * nb;
    if (n == 1) { memcpy(out, f, sizeof(double) * (size_t)tot); return 0; }
    const double co[5] = { agf, bgf, cgf, dgf, egf };
    shape_for_axis(axis, n, na, nb, &n1, &n3);
    pdo_oracle_rhs(RHS_SYM_9, co, f, out, n1, n, n3);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Staggered CD06 in z, periodic — derivatives/cd06stagg.F90:170-195 (init), 301-629 (RHS),      */
/* 248-299 (solve), 820-1059 (entry points).  `m` = n1*n2 doubles per plane (complex data: pass  */
/* m = 2*n1*n2; the LU is real so re/im are independent lines).  Cells n planes, edges n+1.      */
/* op: 0 ddz_E2C, 1 ddz_C2E, 2 interp_E2C, 3 interp_C2E, 4 d2dz2_C2C, 5 d2dz2_E2E                */
/* ------------------------------------------------------------------------------------------ */
int pdo_oracle_stagg_lu(int n, int which /*0 D1, 1 D2, 2 interp*/, double *LU)
{
    static const double al[3] = { 9.0 / 62.0, 2.0 / 11.0, 3.0 / 10.0 };
    if (n <= 4) return 21; /* cd06stagg.F90:182-184 */
    pdo_oracle_tri_lu(n, al[which], 1.0, al[which], LU);
    return 0;
}

/* plane pointer helpers; planes are 1-based like the Fortran */
#define PL(p, k) ((p) + ((int64_t)(k) - 1) * m)

int pdo_oracle_stagg(const double *LU, int n, double dx, int op, const double *f, double *out, int64_t m)
{
    const double onebydx = 1.0 / dx;
    int64_t i;
    int k;
    double a06, b06;
    if (op == 0 || op == 1) { a06 = (63.0 / 62.0) * onebydx; b06 = ((17.0 / 62.0) / 3.0) * onebydx; }
    else if (op == 2 || op == 3) { a06 = (3.0 / 2.0) * (1.0 / 2.0); b06 = (1.0 / 10.0) * (1.0 / 2.0); }
    else { a06 = (12.0 / 11.0) * (onebydx * onebydx); b06 = ((3.0 / 11.0) / 4.0) * (onebydx * onebydx); }

    if (op == 0 || op == 2) {
        /* E2C (cd06stagg.F90:301-323, 526-548): reads edge plane n+1 of the caller's array
           (sequence association through the fE(:,:,1:n) slice); wraps with planes 2 / n. */
        const double sg = (op == 0) ? -1.0 : 1.0;
        for (k = 1; k <= n; ++k) {
            const double *p1 = PL(f, k + 1), *p0 = PL(f, k);
            const double *p2 = (k == n) ? PL(f, 2) : PL(f, k + 2);
            const double *m1 = (k == 1) ? PL(f, n) : PL(f, k - 1);
            double *r = PL(out, k);
            for (i = 0; i < m; ++i) r[i] = a06 * (p1[i] + sg * p0[i]) + b06 * (p2[i] + sg * m1[i]);
        }
        pdo_oracle_tri_solve_y(LU, n, out, m, 1);
    } else if (op == 1 || op == 3) {
        /* C2E (cd06stagg.F90:349-372, 578-604): edge k sits between cells k-1 and k */
        const double sg = (op == 1) ? -1.0 : 1.0;
        for (k = 1; k <= n; ++k) {
            const double *p0 = PL(f, k);
            const double *m1 = PL(f, wrapi(k - 2, n) + 1);
            const double *p1 = PL(f, wrapi(k, n) + 1);
            const double *m2 = PL(f, wrapi(k - 3, n) + 1);
            double *r = PL(out, k);
            for (i = 0; i < m; ++i) r[i] = a06 * (p0[i] + sg * m1[i]) + b06 * (p1[i] + sg * m2[i]);
        }
        pdo_oracle_tri_solve_y(LU, n, out, m, 1);
        memcpy(PL(out, n + 1), PL(out, 1), sizeof(double) * (size_t)m); /* :859, 969 */
    } else {
        /* collocated second derivative C2C / E2E (cd06stagg.F90:405-524) */
        for (k = 1; k <= n; ++k) {
            const double *p0 = PL(f, k);
            const double *p1 = PL(f, wrapi(k, n) + 1), *m1 = PL(f, wrapi(k - 2, n) + 1);
            const double *p2 = PL(f, wrapi(k + 1, n) + 1), *m2 = PL(f, wrapi(k - 3, n) + 1);
            double *r = PL(out, k);
            for (i = 0; i < m; ++i)
                r[i] = a06 * (p1[i] - 2.0 * p0[i] + m1[i]) + b06 * (p2[i] - 2.0 * p0[i] + m2[i]);
        }
        pdo_oracle_tri_solve_y(LU, n, out, m, 1);
        if (op == 5) memcpy(PL(out, n + 1), PL(out, 1), sizeof(double) * (size_t)m); /* :1036 */
    }
    return 0;
}
