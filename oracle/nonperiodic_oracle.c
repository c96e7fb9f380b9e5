/* nonperiodic_oracle.c — CPU restatement of the NON-PERIODIC CD10, CF90 and CD06 closures (groundwork: the
 * CUDA library does not implement them yet and returns PDO_E_UNSUPPORTED for periodic = .false.).
 *
 * TEST INFRASTRUCTURE ONLY (see padeops_oracle.c).  Follows derivatives/cd10.F90 statement by statement:
 *   boundary-scheme constants            cd10.F90:29-96
 *   ComputePenta1 / ComputePenta2        cd10.F90:429-575, 577-707   (LHS rows per (bc1, bcn) in {0, 1, -1}^2 and its LU)
 *   SolveXPenta1 (same sweeps for 2)     cd10.F90:823-851
 *   ComputeXD1RHS, periodic = .false.    cd10.F90:1143-1262
 *   ComputeXD2RHS, periodic = .false.    cd10.F90:1636-1731
 * bc = 0: one-sided closure; bc = 1: f symmetric about the boundary node; bc = -1: f antisymmetric.  In dd1 the caller's
 * (bc1, bcn) select the RHS rows and the matching LHS; in d2d1 likewise (cd10.F90:2029-2097, 2239-2305).
 * Lines are handled one at a time with an element stride, so every axis of f(n1,n2,n3) uses the same code. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- interior scheme (cd10.F90:16-27) ---- */
static const double alpha10d1 = 1.0 / 2.0, beta10d1 = 1.0 / 20.0;
static const double a10d1 = (17.0 / 12.0) / 2.0, b10d1 = (101.0 / 150.0) / 4.0, c10d1 = (1.0 / 100.0) / 6.0;
static const double alpha10d2 = 334.0 / 899.0, beta10d2 = 43.0 / 1798.0;
static const double a10d2 = (1065.0 / 1798.0) / 1.0, b10d2 = (1038.0 / 899.0) / 4.0, c10d2 = (79.0 / 1798.0) / 9.0;

/* ---- first-derivative boundary schemes and weights (cd10.F90:33-77); evaluated once, in the reference's order ---- */
typedef struct {
    double alpha, p, q, r, s;
    double q_hat, r_hat, s_hat, alpha_hat, beta_hat;
    double q_p, alpha_p;
    double alpha_ppp, beta_ppp, q_ppp, r_ppp, s_ppp;
    double alpha_pp, beta_pp, q_pp, r_pp, s_pp;
    double w1, w2, w3, w4;
} D1Consts;

static D1Consts d1_consts(void)
{
    D1Consts c;
    c.alpha = 3.0; c.p = -17.0 / 6.0; c.q = 3.0 / 2.0; c.r = 3.0 / 2.0; c.s = -1.0 / 6.0;
    c.q_hat = a10d1; c.r_hat = b10d1; c.s_hat = c10d1; c.alpha_hat = alpha10d1; c.beta_hat = beta10d1;
    c.q_p = 3.0 / 4.0; c.alpha_p = 1.0 / 4.0;
    c.alpha_ppp = (8 * c.r_hat - 175 * c.s_hat) / (18 * c.r_hat - 550 * c.s_hat);
    c.beta_ppp = (1.0 / 20.0) * (-3 + 8 * c.alpha_ppp);
    c.q_ppp = (1.0 / 12.0) * (12 - 7 * c.alpha_ppp);
    c.r_ppp = (1.0 / 600.0) * (568 * c.alpha_ppp - 183);
    c.s_ppp = (1.0 / 300.0) * (9 * c.alpha_ppp - 4);
    {   /* cd10.F90:60-65; the literal `1._rkind/3_rkind` there is real/integer = 1/3 */
        const double t = (c.s * (c.r_hat + 2 * c.s_hat) - c.q * (c.q_hat + c.r_hat + c.s_hat));
        const double u = (c.q + c.s) * (c.q_hat + c.r_hat - c.s_hat * (c.q_ppp / c.s_ppp - 1));
        c.alpha_pp = ((17 * t) / (72 * u) - 8.0 / 9.0) / ((19 * t) / (24 * u) - 1.0 / 3.0);
    }
    c.beta_pp = (1.0 / 12.0) * (-1 + 3 * c.alpha_pp);
    c.q_pp = (2.0 / 18.0) * (8 - 3 * c.alpha_pp);
    c.r_pp = (1.0 / 72.0) * (-17 + 57 * c.alpha_pp);
    c.s_pp = 0.0;
    c.w1 = (c.q_hat + 2 * c.r_hat + 3 * c.s_hat) / (c.q + c.s);
    c.w2 = (1 / c.q_p) * (c.r_hat + c.s_hat * (1 + c.q_ppp / c.s_ppp) - c.r * (c.q_hat + 2 * c.r_hat + 3 * c.s_hat) / (c.q + c.s));
    c.w3 = (c.q_hat + c.r_hat + c.s_hat * (1 - c.q_ppp / c.s_ppp)) / (c.r_pp);
    c.w4 = c.s_hat / c.s_ppp;
    return c;
}

/* ---- second-derivative boundary schemes (cd10.F90:83-99) ---- */
static const double b1_alpha10d2 = 11.0;
#define b1_a10d2 ((11 * b1_alpha10d2 + 35) / 12)
#define b1_b10d2 (-(5 * b1_alpha10d2 + 26) / 3)
#define b1_c10d2 ((b1_alpha10d2 + 19) / 2)
#define b1_d10d2 ((b1_alpha10d2 - 14) / 3)
#define b1_e10d2 ((11 - b1_alpha10d2) / 12)
static const double b2_alpha10d2 = 1.0 / 10.0;
#define b2_a10d2 ((4 * (1 - b2_alpha10d2) / 3) / 1.0)
static const double b3_alpha10d2 = 344.0 / 1179.0;
#define b3_beta10d2 ((38.0 * b3_alpha10d2 - 9.0) / 214.0)
#define b3_a10d2 (((696 - 1191 * b3_alpha10d2) / 428) / 1.0)
#define b3_b10d2 (((2454 * b3_alpha10d2 - 294) / 535) / 4.0)

/* penta(n, 11) column-major like the Fortran: column c (1-based) of row i (1-based) = P[(c-1)*n + (i-1)] */
#define COL(P, c) ((P) + (size_t)((c) - 1) * (size_t)n - 1) /* 1-based row index */

static void penta_factor(int n, double *P)  /* Steps 1-3 of ComputePenta1/2 (cd10.F90:556-572, 688-704) */
{
    double *bt = COL(P, 1), *b = COL(P, 2), *d = COL(P, 3), *a = COL(P, 4), *at = COL(P, 5);
    double *e = COL(P, 6), *obc = COL(P, 7), *f = COL(P, 8), *g = COL(P, 9), *eobc = COL(P, 10);
    obc[1] = 1.0 / d[1];
    obc[2] = 1.0 / (d[2] - b[2] * a[1] * obc[1]);
    e[1] = a[1];
    f[2] = b[2] * obc[1];
    for (int i = 3; i <= n; ++i) {
        g[i] = bt[i] * obc[i - 2];
        e[i - 1] = a[i - 1] - f[i - 1] * at[i - 2];
        f[i] = (b[i] - g[i] * e[i - 2]) * obc[i - 1];
        obc[i] = 1.0 / (d[i] - f[i] * e[i - 1] - g[i] * at[i - 2]);
    }
    for (int i = 1; i <= n; ++i) eobc[i] = e[i] * obc[i];
}

/* ComputePenta1 (which = 1, cd10.F90:429-575) / ComputePenta2 (which = 2, :577-707).  P: n*11 doubles, zeroed here.
 * Returns 0, or 2 if n is too short for the closures (the reference does not guard; 8 points keep the stencils apart). */
int pdo_oracle_cd10_np_penta(int n, int which, int bc1, int bcn, double *P)
{
    if (n < 8) return 2;
    memset(P, 0, sizeof(double) * 11 * (size_t)n);
    double *bt = COL(P, 1), *b = COL(P, 2), *d = COL(P, 3), *a = COL(P, 4), *at = COL(P, 5);
    if (which == 1) {
        const D1Consts c = d1_consts();
        for (int i = 1; i <= n; ++i) { at[i] = c.beta_hat; bt[i] = c.beta_hat; a[i] = c.alpha_hat; b[i] = c.alpha_hat; d[i] = 1.0; }
        switch (bc1) {
        case 0:
            bt[1] = c.w1 * 0; b[1] = c.w1 * 0; d[1] = c.w1 * 1; a[1] = c.w1 * c.alpha; at[1] = c.w1 * 0;
            bt[2] = c.w2 * 0; b[2] = c.w2 * c.alpha_p; d[2] = c.w2 * 1; a[2] = c.w2 * c.alpha_p; at[2] = c.w2 * 0;
            bt[3] = c.w3 * c.beta_pp; b[3] = c.w3 * c.alpha_pp; d[3] = c.w3 * 1; a[3] = c.w3 * c.alpha_pp; at[3] = c.w3 * c.beta_pp;
            bt[4] = c.w4 * c.beta_ppp; b[4] = c.w4 * c.alpha_ppp; d[4] = c.w4 * 1; a[4] = c.w4 * c.alpha_ppp; at[4] = c.w4 * c.beta_ppp;
            break;
        case 1:
            bt[1] = 0; b[1] = 0; d[1] = 1; a[1] = 0; at[1] = 0;
            bt[2] = 0; b[2] = c.alpha_hat; d[2] = 1 - c.beta_hat; a[2] = c.alpha_hat; at[2] = c.beta_hat;
            break;
        case -1:
            bt[1] = 0; b[1] = 0; d[1] = 1; a[1] = 2 * c.alpha_hat; at[1] = 2 * c.beta_hat;
            bt[2] = 0; b[2] = c.alpha_hat; d[2] = 1 + c.beta_hat; a[2] = c.alpha_hat; at[2] = c.beta_hat;
            break;
        default: return 324;
        }
        switch (bcn) {
        case 0:
            bt[n] = c.w1 * 0; b[n] = c.w1 * c.alpha; d[n] = c.w1 * 1; a[n] = c.w1 * 0; at[n] = c.w1 * 0;
            bt[n - 1] = c.w2 * 0; b[n - 1] = c.w2 * c.alpha_p; d[n - 1] = c.w2 * 1; a[n - 1] = c.w2 * c.alpha_p; at[n - 1] = c.w2 * 0;
            bt[n - 2] = c.w3 * c.beta_pp; b[n - 2] = c.w3 * c.alpha_pp; d[n - 2] = c.w3 * 1; a[n - 2] = c.w3 * c.alpha_pp; at[n - 2] = c.w3 * c.beta_pp;
            bt[n - 3] = c.w4 * c.beta_ppp; b[n - 3] = c.w4 * c.alpha_ppp; d[n - 3] = c.w4 * 1; a[n - 3] = c.w4 * c.alpha_ppp; at[n - 3] = c.w4 * c.beta_ppp;
            break;
        case 1:
            bt[n] = 0; b[n] = 0; d[n] = 1; a[n] = 0; at[n] = 0;
            bt[n - 1] = c.beta_hat; b[n - 1] = c.alpha_hat; d[n - 1] = 1 - c.beta_hat; a[n - 1] = c.alpha_hat; at[n - 1] = 0;
            break;
        case -1:
            bt[n] = 2 * c.beta_hat; b[n] = 2 * c.alpha_hat; d[n] = 1; a[n] = 0; at[n] = 0;
            bt[n - 1] = c.beta_hat; b[n - 1] = c.alpha_hat; d[n - 1] = 1 + c.beta_hat; a[n - 1] = c.alpha_hat; at[n - 1] = 0;
            break;
        default: return 324;
        }
    } else {
        for (int i = 1; i <= n; ++i) { at[i] = beta10d2; bt[i] = beta10d2; d[i] = 1.0; a[i] = alpha10d2; b[i] = alpha10d2; }
        switch (bc1) {
        case 0:
            bt[1] = 0; b[1] = 0; d[1] = 1; a[1] = b1_alpha10d2; at[1] = 0;
            bt[2] = 0; b[2] = b2_alpha10d2; d[2] = 1; a[2] = b2_alpha10d2; at[2] = 0;
            bt[3] = b3_beta10d2; b[3] = b3_alpha10d2; d[3] = 1; a[3] = b3_alpha10d2; at[3] = b3_beta10d2;
            break;
        case 1:
            bt[1] = 0; b[1] = 0; d[1] = 1; a[1] = 2 * alpha10d2; at[1] = 2 * beta10d2;
            bt[2] = 0; b[2] = alpha10d2; d[2] = 1 + beta10d2; a[2] = alpha10d2; at[2] = beta10d2;
            break;
        case -1:
            bt[1] = 0; b[1] = 0; d[1] = 1; a[1] = 0; at[1] = 0;
            bt[2] = 0; b[2] = alpha10d2; d[2] = 1 - beta10d2; a[2] = alpha10d2; at[2] = beta10d2;
            break;
        default: return 324;
        }
        switch (bcn) {
        case 0:
            bt[n - 2] = b3_beta10d2; b[n - 2] = b3_alpha10d2; d[n - 2] = 1; a[n - 2] = b3_alpha10d2; at[n - 2] = b3_beta10d2;
            bt[n - 1] = 0; b[n - 1] = b2_alpha10d2; d[n - 1] = 1; a[n - 1] = b2_alpha10d2; at[n - 1] = 0;
            bt[n] = 0; b[n] = b1_alpha10d2; d[n] = 1; a[n] = 0; at[n] = 0;
            break;
        case 1:
            bt[n - 1] = beta10d2; b[n - 1] = alpha10d2; d[n - 1] = 1 + beta10d2; a[n - 1] = alpha10d2; at[n - 1] = 0;
            bt[n] = 2 * beta10d2; b[n] = 2 * alpha10d2; d[n] = 1; a[n] = 0; at[n] = 0;
            break;
        case -1:
            bt[n - 1] = beta10d2; b[n - 1] = alpha10d2; d[n - 1] = 1 - beta10d2; a[n - 1] = alpha10d2; at[n - 1] = 0;
            bt[n] = 0; b[n] = 0; d[n] = 1; a[n] = 0; at[n] = 0;
            break;
        default: return 324;
        }
    }
    penta_factor(n, P);
    return 0;
}

/* SolveXPenta1 on one line y[1..n] (1-based view), cd10.F90:833-848 */
static void penta_solve_line(int n, const double *P, double *y)
{
    const double *at = COL(P, 5), *obc = COL(P, 7), *f = COL(P, 8), *g = COL(P, 9), *eobc = COL(P, 10);
    y[2] = y[2] - f[2] * y[1];
    for (int i = 3; i <= n; ++i) y[i] = y[i] - g[i] * y[i - 2] - f[i] * y[i - 1];
    y[n] = y[n] * obc[n];
    y[n - 1] = y[n - 1] * obc[n - 1] - eobc[n - 1] * y[n];
    for (int i = n - 2; i >= 1; --i) y[i] = y[i] * obc[i] - y[i + 2] * at[i] * obc[i] - y[i + 1] * eobc[i];
}

/* ComputeXD1RHS, periodic = .false. (cd10.F90:1143-1262) on one line; f, R are 1-based views */
static void d1_rhs_line(int n, double onebydx, int bc1, int bcn, const double *f, double *R)
{
    const D1Consts c = d1_consts();
    const double a10 = c.q_hat * onebydx, b10 = c.r_hat * onebydx, c10 = c.s_hat * onebydx;
    const double a_np_4 = c.w4 * c.q_ppp * onebydx, b_np_4 = c.w4 * c.r_ppp * onebydx, c_np_4 = c.w4 * c.s_ppp * onebydx;
    const double a_np_3 = c.w3 * c.q_pp * onebydx, b_np_3 = c.w3 * c.r_pp * onebydx;
    const double a_np_2 = c.w2 * c.q_p * onebydx;
    const double a_np_1 = c.w1 * (c.p * onebydx), b_np_1 = c.w1 * (c.q * onebydx), c_np_1 = c.w1 * (c.r * onebydx), d_np_1 = c.w1 * (c.s * onebydx);
    switch (bc1) {
    case 0:
        R[1] = a_np_1 * f[1] + b_np_1 * f[2] + c_np_1 * f[3] + d_np_1 * f[4];
        R[2] = a_np_2 * (f[3] - f[1]);
        R[3] = a_np_3 * (f[4] - f[2]) + b_np_3 * (f[5] - f[1]);
        R[4] = a_np_4 * (f[5] - f[3]) + b_np_4 * (f[6] - f[2]) + c_np_4 * (f[7] - f[1]);
        break;
    case 1:
        R[1] = 0.0;
        R[2] = a10 * (f[3] - f[1]) + b10 * (f[4] - f[2]) + c10 * (f[5] - f[3]);
        R[3] = a10 * (f[4] - f[2]) + b10 * (f[5] - f[1]) + c10 * (f[6] - f[2]);
        R[4] = a10 * (f[5] - f[3]) + b10 * (f[6] - f[2]) + c10 * (f[7] - f[1]);
        break;
    default: /* -1 */
        R[1] = a10 * (f[2] + f[2]) + b10 * (f[3] + f[3]) + c10 * (f[4] + f[4]);
        R[2] = a10 * (f[3] - f[1]) + b10 * (f[4] + f[2]) + c10 * (f[5] + f[3]);
        R[3] = a10 * (f[4] - f[2]) + b10 * (f[5] - f[1]) + c10 * (f[6] + f[2]);
        R[4] = a10 * (f[5] - f[3]) + b10 * (f[6] - f[2]) + c10 * (f[7] - f[1]);
        break;
    }
    for (int i = 5; i <= n - 4; ++i) R[i] = a10 * (f[i + 1] - f[i - 1]) + b10 * (f[i + 2] - f[i - 2]) + c10 * (f[i + 3] - f[i - 3]);
    switch (bcn) {
    case 0:
        R[n - 3] = a_np_4 * (f[n - 2] - f[n - 4]) + b_np_4 * (f[n - 1] - f[n - 5]) + c_np_4 * (f[n] - f[n - 6]);
        R[n - 2] = a_np_3 * (f[n - 1] - f[n - 3]) + b_np_3 * (f[n] - f[n - 4]);
        R[n - 1] = a_np_2 * (f[n] - f[n - 2]);
        R[n] = -a_np_1 * f[n] - b_np_1 * f[n - 1] - c_np_1 * f[n - 2] - d_np_1 * f[n - 3];
        break;
    case 1:
        R[n - 3] = a10 * (f[n - 2] - f[n - 4]) + b10 * (f[n - 1] - f[n - 5]) + c10 * (f[n] - f[n - 6]);
        R[n - 2] = a10 * (f[n - 1] - f[n - 3]) + b10 * (f[n] - f[n - 4]) + c10 * (f[n - 1] - f[n - 5]);
        R[n - 1] = a10 * (f[n] - f[n - 2]) + b10 * (f[n - 1] - f[n - 3]) + c10 * (f[n - 2] - f[n - 4]);
        R[n] = 0.0;
        break;
    default: /* -1 */
        R[n - 3] = a10 * (f[n - 2] - f[n - 4]) + b10 * (f[n - 1] - f[n - 5]) + c10 * (f[n] - f[n - 6]);
        R[n - 2] = a10 * (f[n - 1] - f[n - 3]) + b10 * (f[n] - f[n - 4]) + c10 * (-f[n - 1] - f[n - 5]);
        R[n - 1] = a10 * (f[n] - f[n - 2]) + b10 * (-f[n - 1] - f[n - 3]) + c10 * (-f[n - 2] - f[n - 4]);
        R[n] = a10 * (-f[n - 1] - f[n - 1]) + b10 * (-f[n - 2] - f[n - 2]) + c10 * (-f[n - 3] - f[n - 3]);
        break;
    }
}

/* ComputeXD2RHS, periodic = .false. (cd10.F90:1636-1731) on one line */
static void d2_rhs_line(int n, double onebydx2, int bc1, int bcn, const double *f, double *R)
{
    const double two = 2.0;
    const double a10 = a10d2 * onebydx2, b10 = b10d2 * onebydx2, c10 = c10d2 * onebydx2;
    const double a_np_3 = b3_a10d2 * onebydx2, b_np_3 = b3_b10d2 * onebydx2;
    const double a_np_2 = b2_a10d2 * onebydx2;
    const double a_np_1 = b1_a10d2 * onebydx2, b_np_1 = b1_b10d2 * onebydx2, c_np_1 = b1_c10d2 * onebydx2, d_np_1 = b1_d10d2 * onebydx2,
                 e_np_1 = b1_e10d2 * onebydx2;
    switch (bc1) {
    case 0:
        R[1] = a_np_1 * f[1] + b_np_1 * f[2] + c_np_1 * f[3] + d_np_1 * f[4] + e_np_1 * f[5];
        R[2] = a_np_2 * (f[3] - two * f[2] + f[1]);
        R[3] = a_np_3 * (f[4] - two * f[3] + f[2]) + b_np_3 * (f[5] - two * f[3] + f[1]);
        break;
    case 1:
        R[1] = a10 * (f[2] - two * f[1] + f[2]) + b10 * (f[3] - two * f[1] + f[3]) + c10 * (f[4] - two * f[1] + f[4]);
        R[2] = a10 * (f[3] - two * f[2] + f[1]) + b10 * (f[4] - two * f[2] + f[2]) + c10 * (f[5] - two * f[2] + f[3]);
        R[3] = a10 * (f[4] - two * f[3] + f[2]) + b10 * (f[5] - two * f[3] + f[1]) + c10 * (f[6] - two * f[3] + f[2]);
        break;
    default: /* -1 */
        R[1] = a10 * (f[2] - two * f[1] - f[2]) + b10 * (f[3] - two * f[1] - f[3]) + c10 * (f[4] - two * f[1] - f[4]);
        R[2] = a10 * (f[3] - two * f[2] + f[1]) + b10 * (f[4] - two * f[2] - f[2]) + c10 * (f[5] - two * f[2] - f[3]);
        R[3] = a10 * (f[4] - two * f[3] + f[2]) + b10 * (f[5] - two * f[3] + f[1]) + c10 * (f[6] - two * f[3] - f[2]);
        break;
    }
    for (int i = 4; i <= n - 3; ++i)
        R[i] = a10 * (f[i + 1] - two * f[i] + f[i - 1]) + b10 * (f[i + 2] - two * f[i] + f[i - 2]) + c10 * (f[i + 3] - two * f[i] + f[i - 3]);
    switch (bcn) {
    case 0:
        R[n - 2] = a_np_3 * (f[n - 1] - two * f[n - 2] + f[n - 3]) + b_np_3 * (f[n] - two * f[n - 2] + f[n - 4]);
        R[n - 1] = a_np_2 * (f[n] - two * f[n - 1] + f[n - 2]);
        R[n] = a_np_1 * f[n] + b_np_1 * f[n - 1] + c_np_1 * f[n - 2] + d_np_1 * f[n - 3] + e_np_1 * f[n - 4];
        break;
    case 1:
        R[n - 2] = a10 * (f[n - 1] - two * f[n - 2] + f[n - 3]) + b10 * (f[n] - two * f[n - 2] + f[n - 4]) + c10 * (f[n - 1] - two * f[n - 2] + f[n - 5]);
        R[n - 1] = a10 * (f[n] - two * f[n - 1] + f[n - 2]) + b10 * (f[n - 1] - two * f[n - 1] + f[n - 3]) + c10 * (f[n - 2] - two * f[n - 1] + f[n - 4]);
        R[n] = a10 * (f[n - 1] - two * f[n] + f[n - 1]) + b10 * (f[n - 2] - two * f[n] + f[n - 2]) + c10 * (f[n - 3] - two * f[n] + f[n - 3]);
        break;
    default: /* -1 */
        R[n - 2] = a10 * (f[n - 1] - two * f[n - 2] + f[n - 3]) + b10 * (f[n] - two * f[n - 2] + f[n - 4]) + c10 * (-f[n - 1] - two * f[n - 2] + f[n - 5]);
        R[n - 1] = a10 * (f[n] - two * f[n - 1] + f[n - 2]) + b10 * (-f[n - 1] - two * f[n - 1] + f[n - 3]) + c10 * (-f[n - 2] - two * f[n - 1] + f[n - 4]);
        R[n] = a10 * (-f[n - 1] - two * f[n] + f[n - 1]) + b10 * (-f[n - 2] - two * f[n] + f[n - 2]) + c10 * (-f[n - 3] - two * f[n] + f[n - 3]);
        break;
    }
}

/* cd10%dd1/dd2/dd3 (which = 1) or d2d1/2/3 (which = 2) with periodic = .false. along `axis` of f(n1,n2,n3) given as
 * (n, na, nb) like pdo_oracle_cd10: axis 0: f(n,na,nb); 1: f(na,n,nb); 2: f(na,nb,n). */
int pdo_oracle_cd10_np(int n, double dx, int which, int bc1, int bcn, int axis, const double *f, double *df, int64_t na, int64_t nb)
{
    double *P = (double *)malloc(sizeof(double) * 11 * (size_t)n);
    double *lf = (double *)malloc(sizeof(double) * (size_t)n), *lr = (double *)malloc(sizeof(double) * (size_t)n);
    if (!P || !lf || !lr) { free(P); free(lf); free(lr); return -1; }
    int rc = pdo_oracle_cd10_np_penta(n, which, bc1, bcn, P);
    if (rc) { free(P); free(lf); free(lr); return rc; }
    const double onebydx = 1.0 / dx, onebydx2 = onebydx / dx;
    int64_t stride, nlines_in, nlines_out, in_step, out_step;
    /* a line starts at  base = io*out_step + ii*in_step  and advances by `stride` */
    if (axis == 0) { stride = 1; nlines_in = na * nb; nlines_out = 1; in_step = n; out_step = 0; }
    else if (axis == 1) { stride = na; nlines_in = na; nlines_out = nb; in_step = 1; out_step = na * (int64_t)n; }
    else { stride = na * nb; nlines_in = na * nb; nlines_out = 1; in_step = 1; out_step = 0; }
    for (int64_t io = 0; io < nlines_out; ++io)
        for (int64_t ii = 0; ii < nlines_in; ++ii) {
            const int64_t base = io * out_step + ii * in_step;
            for (int i = 0; i < n; ++i) lf[i] = f[base + (int64_t)i * stride];
            if (which == 1) d1_rhs_line(n, onebydx, bc1, bcn, lf - 1, lr - 1);
            else d2_rhs_line(n, onebydx2, bc1, bcn, lf - 1, lr - 1);
            penta_solve_line(n, P, lr - 1);
            for (int i = 0; i < n; ++i) df[base + (int64_t)i * stride] = lr[i];
        }
    free(P); free(lf); free(lr);
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * CF90 non-periodic filter: filters/cf90.F90:24-47 (boundary constants), 276-418 (ComputePenta), 532-558 (SolveXPenta,
 * the same sweeps as cd10's), 672-801 (ComputeXRHS, periodic = .false.).  First / last point are the identity for bc = 0.
 * --------------------------------------------------------------------------------------------------------------- */
static const double alpha90 = 6.6624e-1, beta90 = 1.6688e-1;
static const double a90 = 9.9965e-1, b90 = 6.6652e-1, c90 = 1.6674e-1, d90 = 4.0e-5, e90 = -5.0e-6;
static const double b2_alpha90 = 4.997e-1, b2_a90 = 9.997e-1, b2_b90 = 4.9985e-1;
static const double b3_alpha90 = 6.6624e-1, b3_beta90 = 1.6688e-1, b3_a90 = 9.9952e-1, b3_b90 = 6.6656e-1, b3_c90 = 1.668e-1;
static const double b4_alpha90 = 6.6624e-1, b4_beta90 = 1.6688e-1, b4_a90 = 9.9968e-1, b4_b90 = 6.6652e-1, b4_c90 = 1.6672e-1,
                    b4_d90 = 4.0e-5;

int pdo_oracle_cf90_np_penta(int n, int bc1, int bcn, double *P)
{
    if (n < 10) return 7; /* cf90.F90:128: the periodic guard; 10 points also keep the two closures apart */
    memset(P, 0, sizeof(double) * 11 * (size_t)n);
    double *bt = COL(P, 1), *b = COL(P, 2), *d = COL(P, 3), *a = COL(P, 4), *at = COL(P, 5);
    for (int i = 1; i <= n; ++i) { at[i] = beta90; bt[i] = beta90; a[i] = alpha90; b[i] = alpha90; d[i] = 1.0; }
    switch (bc1) {
    case 0:
        bt[1] = 0; b[1] = 0; d[1] = 1; a[1] = 0; at[1] = 0;
        bt[2] = 0; b[2] = b2_alpha90; d[2] = 1; a[2] = b2_alpha90; at[2] = 0;
        bt[3] = b3_beta90; b[3] = b3_alpha90; d[3] = 1; a[3] = b3_alpha90; at[3] = b3_beta90;
        bt[4] = b4_beta90; b[4] = b4_alpha90; d[4] = 1; a[4] = b4_alpha90; at[4] = b4_beta90;
        break;
    case 1:
        bt[1] = 0; b[1] = 0; d[1] = 1; a[1] = 2 * alpha90; at[1] = 2 * beta90;
        bt[2] = 0; b[2] = alpha90; d[2] = 1 + beta90; a[2] = alpha90; at[2] = beta90;
        break;
    case -1:
        bt[1] = 0; b[1] = 0; d[1] = 1; a[1] = 0; at[1] = 0;
        bt[2] = 0; b[2] = alpha90; d[2] = 1 - beta90; a[2] = alpha90; at[2] = beta90;
        break;
    default: return 324;
    }
    switch (bcn) {
    case 0:
        bt[n] = 0; b[n] = 0; d[n] = 1; a[n] = 0; at[n] = 0;
        bt[n - 1] = 0; b[n - 1] = b2_alpha90; d[n - 1] = 1; a[n - 1] = b2_alpha90; at[n - 1] = 0;
        bt[n - 2] = b3_beta90; b[n - 2] = b3_alpha90; d[n - 2] = 1; a[n - 2] = b3_alpha90; at[n - 2] = b3_beta90;
        bt[n - 3] = b4_beta90; b[n - 3] = b4_alpha90; d[n - 3] = 1; a[n - 3] = b4_alpha90; at[n - 3] = b4_beta90;
        break;
    case 1:
        bt[n] = 2 * beta90; b[n] = 2 * alpha90; d[n] = 1; a[n] = 0; at[n] = 0;
        bt[n - 1] = beta90; b[n - 1] = alpha90; d[n - 1] = 1 + beta90; a[n - 1] = alpha90; at[n - 1] = 0;
        break;
    case -1:
        bt[n] = 0; b[n] = 0; d[n] = 1; a[n] = 0; at[n] = 0;
        bt[n - 1] = beta90; b[n - 1] = alpha90; d[n - 1] = 1 - beta90; a[n - 1] = alpha90; at[n - 1] = 0;
        break;
    default: return 324;
    }
    penta_factor(n, P);
    return 0;
}

/* ComputeXRHS, periodic = .false. (cf90.F90:672-801) on one line; 1-based views */
static void cf90_rhs_line(int n, int bc1, int bcn, const double *f, double *R)
{
    switch (bc1) {
    case 0:
        R[1] = 1.0 * (f[1]);
        R[2] = b2_a90 * (f[2]) + b2_b90 * (f[3] + f[1]);
        R[3] = b3_a90 * (f[3]) + b3_b90 * (f[4] + f[2]) + b3_c90 * (f[5] + f[1]);
        R[4] = b4_a90 * (f[4]) + b4_b90 * (f[5] + f[3]) + b4_c90 * (f[6] + f[2]) + b4_d90 * (f[7] + f[1]);
        break;
    case 1:
        R[1] = a90 * (f[1]) + b90 * (f[2] + f[2]) + c90 * (f[3] + f[3]) + d90 * (f[4] + f[4]) + e90 * (f[5] + f[5]);
        R[2] = a90 * (f[2]) + b90 * (f[3] + f[1]) + c90 * (f[4] + f[2]) + d90 * (f[5] + f[3]) + e90 * (f[6] + f[4]);
        R[3] = a90 * (f[3]) + b90 * (f[4] + f[2]) + c90 * (f[5] + f[1]) + d90 * (f[6] + f[2]) + e90 * (f[7] + f[3]);
        R[4] = a90 * (f[4]) + b90 * (f[5] + f[3]) + c90 * (f[6] + f[2]) + d90 * (f[7] + f[1]) + e90 * (f[8] + f[2]);
        break;
    default: /* -1 */
        R[1] = a90 * (f[1]) + b90 * (f[2] - f[2]) + c90 * (f[3] - f[3]) + d90 * (f[4] - f[4]) + e90 * (f[5] - f[5]);
        R[2] = a90 * (f[2]) + b90 * (f[3] + f[1]) + c90 * (f[4] - f[2]) + d90 * (f[5] - f[3]) + e90 * (f[6] - f[4]);
        R[3] = a90 * (f[3]) + b90 * (f[4] + f[2]) + c90 * (f[5] + f[1]) + d90 * (f[6] - f[2]) + e90 * (f[7] - f[3]);
        R[4] = a90 * (f[4]) + b90 * (f[5] + f[3]) + c90 * (f[6] + f[2]) + d90 * (f[7] + f[1]) + e90 * (f[8] - f[2]);
        break;
    }
    for (int i = 5; i <= n - 4; ++i)
        R[i] = a90 * (f[i]) + b90 * (f[i + 1] + f[i - 1]) + c90 * (f[i + 2] + f[i - 2]) + d90 * (f[i + 3] + f[i - 3]) + e90 * (f[i + 4] + f[i - 4]);
    switch (bcn) {
    case 0:
        R[n - 3] = b4_a90 * (f[n - 3]) + b4_b90 * (f[n - 2] + f[n - 4]) + b4_c90 * (f[n - 1] + f[n - 5]) + b4_d90 * (f[n] + f[n - 6]);
        R[n - 2] = b3_a90 * (f[n - 2]) + b3_b90 * (f[n - 1] + f[n - 3]) + b3_c90 * (f[n] + f[n - 4]);
        R[n - 1] = b2_a90 * (f[n - 1]) + b2_b90 * (f[n] + f[n - 2]);
        R[n] = 1.0 * (f[n]);
        break;
    case 1:
        R[n - 3] = a90 * (f[n - 3]) + b90 * (f[n - 2] + f[n - 4]) + c90 * (f[n - 1] + f[n - 5]) + d90 * (f[n] + f[n - 6]) + e90 * (f[n - 1] + f[n - 7]);
        R[n - 2] = a90 * (f[n - 2]) + b90 * (f[n - 1] + f[n - 3]) + c90 * (f[n] + f[n - 4]) + d90 * (f[n - 1] + f[n - 5]) + e90 * (f[n - 2] + f[n - 6]);
        R[n - 1] = a90 * (f[n - 1]) + b90 * (f[n] + f[n - 2]) + c90 * (f[n - 1] + f[n - 3]) + d90 * (f[n - 2] + f[n - 4]) + e90 * (f[n - 3] + f[n - 5]);
        R[n] = a90 * (f[n]) + b90 * (f[n - 1] + f[n - 1]) + c90 * (f[n - 2] + f[n - 2]) + d90 * (f[n - 3] + f[n - 3]) + e90 * (f[n - 4] + f[n - 4]);
        break;
    default: /* -1 */
        R[n - 3] = a90 * (f[n - 3]) + b90 * (f[n - 2] + f[n - 4]) + c90 * (f[n - 1] + f[n - 5]) + d90 * (f[n] + f[n - 6]) + e90 * (-f[n - 1] + f[n - 7]);
        R[n - 2] = a90 * (f[n - 2]) + b90 * (f[n - 1] + f[n - 3]) + c90 * (f[n] + f[n - 4]) + d90 * (-f[n - 1] + f[n - 5]) + e90 * (-f[n - 2] + f[n - 6]);
        R[n - 1] = a90 * (f[n - 1]) + b90 * (f[n] + f[n - 2]) + c90 * (-f[n - 1] + f[n - 3]) + d90 * (-f[n - 2] + f[n - 4]) + e90 * (-f[n - 3] + f[n - 5]);
        R[n] = a90 * (f[n]) + b90 * (-f[n - 1] + f[n - 1]) + c90 * (-f[n - 2] + f[n - 2]) + d90 * (-f[n - 3] + f[n - 3]) + e90 * (-f[n - 4] + f[n - 4]);
        break;
    }
}

/* cf90%filter1/2/3 with periodic = .false. along `axis`; (n, na, nb) as in pdo_oracle_cd10_np */
int pdo_oracle_cf90_np(int n, int bc1, int bcn, int axis, const double *f, double *out, int64_t na, int64_t nb)
{
    double *P = (double *)malloc(sizeof(double) * 11 * (size_t)n);
    double *lf = (double *)malloc(sizeof(double) * (size_t)n), *lr = (double *)malloc(sizeof(double) * (size_t)n);
    if (!P || !lf || !lr) { free(P); free(lf); free(lr); return -1; }
    int rc = pdo_oracle_cf90_np_penta(n, bc1, bcn, P);
    if (rc) { free(P); free(lf); free(lr); return rc; }
    int64_t stride, nlines_in, nlines_out, in_step, out_step;
    if (axis == 0) { stride = 1; nlines_in = na * nb; nlines_out = 1; in_step = n; out_step = 0; }
    else if (axis == 1) { stride = na; nlines_in = na; nlines_out = nb; in_step = 1; out_step = na * (int64_t)n; }
    else { stride = na * nb; nlines_in = na * nb; nlines_out = 1; in_step = 1; out_step = 0; }
    for (int64_t io = 0; io < nlines_out; ++io)
        for (int64_t ii = 0; ii < nlines_in; ++ii) {
            const int64_t base = io * out_step + ii * in_step;
            for (int i = 0; i < n; ++i) lf[i] = f[base + (int64_t)i * stride];
            cf90_rhs_line(n, bc1, bcn, lf - 1, lr - 1);
            penta_solve_line(n, P, lr - 1);
            for (int i = 0; i < n; ++i) out[base + (int64_t)i * stride] = lr[i];
        }
    free(P); free(lf); free(lr);
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * CD06 non-periodic first derivative: derivatives/cd06.F90:27-58 (boundary scheme and weights), 264-327 (ComputeTri1; only
 * the one-sided closure bc = 0 is complete in the reference, the symmetric cases are marked "Incomplete"), 432-449
 * (SolveXTri1), 551-590 (ComputeXD1RHS, periodic = .false.).  dd1/dd2/dd3 take no boundary codes (cd06.F90:775-839).
 * --------------------------------------------------------------------------------------------------------------- */
typedef struct { double alpha, p, q, r, s, qhat, rhat, alpha_hat, q_p, alpha_p, alpha_pp, q_pp, r_pp, w1, w2, w3; } C6Consts;
static C6Consts c6_consts(void)
{
    C6Consts c;
    c.alpha = 3.0; c.p = 17.0 / 6.0; c.q = 3.0 / 2.0; c.r = 3.0 / 2.0; c.s = -1.0 / 6.0;
    c.qhat = (14.0 / 9.0) / 2.0; c.rhat = (1.0 / 9.0) / 4.0; c.alpha_hat = 1.0 / 3.0;
    c.q_p = 3.0 / 4.0; c.alpha_p = 1.0 / 4.0;
    c.alpha_pp = ((40 * c.alpha_hat - 1) * c.q + 7 * (4 * c.alpha_hat - 1) * c.s) / (16 * (c.alpha_hat + 2) * c.q + 8 * (1 - 4 * c.alpha_hat) * c.s);
    c.q_pp = (1.0 / 3.0) * (c.alpha_pp + 2);
    c.r_pp = (1.0 / 12.0) * (4 * c.alpha_pp - 1);
    c.w1 = (2 * c.alpha_hat + 1) / (2 * (c.q + c.s));
    c.w2 = ((8 * c.alpha_hat + 7) * c.q - 6 * (2 * c.alpha_hat + 1) * c.r + (8 * c.alpha_hat + 7) * c.s) / (9 * (c.q + c.s));
    c.w3 = (4 * (c.alpha_hat + 2) * c.q + 2 * (1 - 4 * c.alpha_hat) * c.s) / (9 * (c.q + c.s));
    return c;
}

/* ComputeTri1 with bc1 = bcn = 0 (cd06.F90:264-327): T(n,3) column-major = a*den, den, cp */
int pdo_oracle_cd06_np_tri(int n, double *T)
{
    if (n < 6) return 3;
    const C6Consts k = c6_consts();
    double *a = (double *)malloc(sizeof(double) * 5 * (size_t)n);
    if (!a) return -1;
    double *b = a + n, *c = b + n, *cp = c + n, *den = cp + n;
    for (int i = 0; i < n; ++i) { a[i] = k.alpha_hat; b[i] = 1.0; c[i] = k.alpha_hat; }
    a[0] = k.w1 * 0; a[1] = k.w2 * k.alpha_p; a[2] = k.w3 * k.alpha_pp;
    b[0] = k.w1 * 1; b[1] = k.w2 * 1; b[2] = k.w3 * 1;
    c[0] = k.w1 * k.alpha; c[1] = k.w2 * k.alpha_p; c[2] = k.w3 * k.alpha_pp;
    c[n - 1] = k.w1 * 0; c[n - 2] = k.w2 * k.alpha_p; c[n - 3] = k.w3 * k.alpha_pp;
    b[n - 1] = k.w1 * 1; b[n - 2] = k.w2 * 1; b[n - 3] = k.w3 * 1;
    a[n - 1] = k.w1 * k.alpha; a[n - 2] = k.w2 * k.alpha_p; a[n - 3] = k.w3 * k.alpha_pp;
    cp[0] = c[0] / b[0];
    for (int i = 1; i < n - 1; ++i) cp[i] = c[i] / (b[i] - a[i] * cp[i - 1]);
    cp[n - 1] = 0.0;   /* never read by the solve */
    den[0] = 1.0 / b[0];
    for (int i = 1; i < n; ++i) den[i] = 1.0 / (b[i] - a[i] * cp[i - 1]);
    for (int i = 0; i < n; ++i) { T[i] = a[i] * den[i]; T[n + i] = den[i]; T[2 * (size_t)n + i] = cp[i]; }
    /* also hand back the raw rows for tests: T[3n..6n) = a, b, c */
    for (int i = 0; i < n; ++i) { T[3 * (size_t)n + i] = a[i]; T[4 * (size_t)n + i] = b[i]; T[5 * (size_t)n + i] = c[i]; }
    free(a);
    return 0;
}

int pdo_oracle_cd06_np(int n, double dx, int axis, const double *f, double *df, int64_t na, int64_t nb)
{
    double *T = (double *)malloc(sizeof(double) * 6 * (size_t)n);
    double *lf = (double *)malloc(sizeof(double) * (size_t)n), *ly = (double *)malloc(sizeof(double) * (size_t)n);
    if (!T || !lf || !ly) { free(T); free(lf); free(ly); return -1; }
    int rc = pdo_oracle_cd06_np_tri(n, T);
    if (rc) { free(T); free(lf); free(ly); return rc; }
    const C6Consts k = c6_consts();
    const double onebydx = 1.0 / dx;
    const double a06 = k.qhat * onebydx, b06 = k.rhat * onebydx;
    const double a_np_3 = k.w3 * k.q_pp * onebydx, b_np_3 = k.w3 * k.r_pp * onebydx, a_np_2 = k.w2 * k.q_p * onebydx;
    const double a_np_1 = k.w1 * (-k.p * onebydx), b_np_1 = k.w1 * (k.q * onebydx), c_np_1 = k.w1 * (k.r * onebydx), d_np_1 = k.w1 * (k.s * onebydx);
    int64_t stride, nlines_in, nlines_out, in_step, out_step;
    if (axis == 0) { stride = 1; nlines_in = na * nb; nlines_out = 1; in_step = n; out_step = 0; }
    else if (axis == 1) { stride = na; nlines_in = na; nlines_out = nb; in_step = 1; out_step = na * (int64_t)n; }
    else { stride = na * nb; nlines_in = na * nb; nlines_out = 1; in_step = 1; out_step = 0; }
    for (int64_t io = 0; io < nlines_out; ++io)
        for (int64_t ii = 0; ii < nlines_in; ++ii) {
            const int64_t base = io * out_step + ii * in_step;
            for (int i = 0; i < n; ++i) lf[i] = f[base + (int64_t)i * stride];
            const double *F = lf - 1;   /* 1-based */
            double *R = ly - 1;
            R[1] = a_np_1 * F[1] + b_np_1 * F[2] + c_np_1 * F[3] + d_np_1 * F[4];
            R[2] = a_np_2 * (F[3] - F[1]);
            R[3] = a_np_3 * (F[4] - F[2]) + b_np_3 * (F[5] - F[1]);
            for (int i = 4; i <= n - 3; ++i) R[i] = b06 * (F[i + 2] - F[i - 2]) + a06 * (F[i + 1] - F[i - 1]);
            R[n - 2] = a_np_3 * (F[n - 1] - F[n - 3]) + b_np_3 * (F[n] - F[n - 4]);
            R[n - 1] = a_np_2 * (F[n] - F[n - 2]);
            R[n] = -a_np_1 * F[n] - b_np_1 * F[n - 1] - c_np_1 * F[n - 2] - d_np_1 * F[n - 3];
            /* SolveXTri1 (cd06.F90:439-447): Tri1(:,1) = a*den, (:,2) = den, (:,3) = cp */
            const double *t1 = T - 1, *t2 = T + n - 1, *t3 = T + 2 * (size_t)n - 1;
            R[1] = R[1] * t2[1];
            for (int i = 2; i <= n; ++i) R[i] = R[i] * t2[i] - R[i - 1] * t1[i];
            for (int i = n - 1; i >= 1; --i) R[i] = R[i] - t3[i] * R[i + 1];
            for (int i = 0; i < n; ++i) df[base + (int64_t)i * stride] = ly[i];
        }
    free(T); free(lf); free(ly);
    return 0;
}

/* SolveXPenta1 on one line (in place), exported so that tests can check the LU against a dense solve of the assembled rows */
void pdo_oracle_cd10_np_solve_line(int n, const double *P, double *y) { penta_solve_line(n, P, y - 1); }
