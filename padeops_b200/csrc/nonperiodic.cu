// nonperiodic.cu — see nonperiodic.cuh.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nonperiodic.cuh"

namespace pdo {

namespace {

// ---- interior schemes (cd10.F90:16-27, cf90.F90:16-22) ----
constexpr double alpha10d1 = 1.0 / 2.0, beta10d1 = 1.0 / 20.0;
constexpr double a10d1 = (17.0 / 12.0) / 2.0, b10d1 = (101.0 / 150.0) / 4.0, c10d1 = (1.0 / 100.0) / 6.0;
constexpr double alpha10d2 = 334.0 / 899.0, beta10d2 = 43.0 / 1798.0;
constexpr double a10d2 = (1065.0 / 1798.0) / 1.0, b10d2 = (1038.0 / 899.0) / 4.0, c10d2 = (79.0 / 1798.0) / 9.0;
constexpr double alpha90 = 6.6624e-1, beta90 = 1.6688e-1, a90 = 9.9965e-1, b90 = 6.6652e-1, c90 = 1.6674e-1, d90 = 4.0e-5, e90 = -5.0e-6;

// ---- first-derivative boundary schemes and weights, evaluated in the reference's order (cd10.F90:33-77) ----
struct D1B {
    double alpha = 3.0, p = -17.0 / 6.0, q = 3.0 / 2.0, r = 3.0 / 2.0, s = -1.0 / 6.0;
    double q_hat = a10d1, r_hat = b10d1, s_hat = c10d1;
    double q_p = 3.0 / 4.0, alpha_p = 1.0 / 4.0;
    double alpha_ppp, beta_ppp, q_ppp, r_ppp, s_ppp, alpha_pp, beta_pp, q_pp, r_pp, w1, w2, w3, w4;
    D1B() {
        alpha_ppp = (8 * r_hat - 175 * s_hat) / (18 * r_hat - 550 * s_hat);
        beta_ppp = (1.0 / 20.0) * (-3 + 8 * alpha_ppp);
        q_ppp = (1.0 / 12.0) * (12 - 7 * alpha_ppp);
        r_ppp = (1.0 / 600.0) * (568 * alpha_ppp - 183);
        s_ppp = (1.0 / 300.0) * (9 * alpha_ppp - 4);
        const double t = s * (r_hat + 2 * s_hat) - q * (q_hat + r_hat + s_hat);
        const double u = (q + s) * (q_hat + r_hat - s_hat * (q_ppp / s_ppp - 1));
        alpha_pp = ((17 * t) / (72 * u) - 8.0 / 9.0) / ((19 * t) / (24 * u) - 1.0 / 3.0);
        beta_pp = (1.0 / 12.0) * (-1 + 3 * alpha_pp);
        q_pp = (2.0 / 18.0) * (8 - 3 * alpha_pp);
        r_pp = (1.0 / 72.0) * (-17 + 57 * alpha_pp);
        w1 = (q_hat + 2 * r_hat + 3 * s_hat) / (q + s);
        w2 = (1 / q_p) * (r_hat + s_hat * (1 + q_ppp / s_ppp) - r * (q_hat + 2 * r_hat + 3 * s_hat) / (q + s));
        w3 = (q_hat + r_hat + s_hat * (1 - q_ppp / s_ppp)) / (r_pp);
        w4 = s_hat / s_ppp;
    }
};
// ---- second-derivative boundary schemes (cd10.F90:83-99) ----
constexpr double b1_alpha10d2 = 11.0;
constexpr double b1_a10d2 = (11 * b1_alpha10d2 + 35) / 12, b1_b10d2 = -(5 * b1_alpha10d2 + 26) / 3, b1_c10d2 = (b1_alpha10d2 + 19) / 2,
                 b1_d10d2 = (b1_alpha10d2 - 14) / 3, b1_e10d2 = (11 - b1_alpha10d2) / 12;
constexpr double b2_alpha10d2 = 1.0 / 10.0, b2_a10d2 = (4 * (1 - b2_alpha10d2) / 3) / 1.0;
constexpr double b3_alpha10d2 = 344.0 / 1179.0, b3_beta10d2 = (38.0 * b3_alpha10d2 - 9.0) / 214.0;
constexpr double b3_a10d2 = ((696 - 1191 * b3_alpha10d2) / 428) / 1.0, b3_b10d2 = ((2454 * b3_alpha10d2 - 294) / 535) / 4.0;
// ---- CF90 boundary rows (cf90.F90:24-47) ----
constexpr double b2_alpha90 = 4.997e-1, b2_a90 = 9.997e-1, b2_b90 = 4.9985e-1;
constexpr double b3_alpha90 = 6.6624e-1, b3_beta90 = 1.6688e-1, b3_a90 = 9.9952e-1, b3_b90 = 6.6656e-1, b3_c90 = 1.668e-1;
constexpr double b4_alpha90 = 6.6624e-1, b4_beta90 = 1.6688e-1, b4_a90 = 9.9968e-1, b4_b90 = 6.6652e-1, b4_c90 = 1.6672e-1, b4_d90 = 4.0e-5;

// ---- CD06 first-derivative boundary scheme and weights (cd06.F90:27-58) ----
struct C6B {
    double alpha = 3.0, p = 17.0 / 6.0, q = 3.0 / 2.0, r = 3.0 / 2.0, s = -1.0 / 6.0;
    double qhat = (14.0 / 9.0) / 2.0, rhat = (1.0 / 9.0) / 4.0, alpha_hat = 1.0 / 3.0, q_p = 3.0 / 4.0, alpha_p = 1.0 / 4.0;
    double alpha_pp, q_pp, r_pp, w1, w2, w3;
    C6B() {
        alpha_pp = ((40 * alpha_hat - 1) * q + 7 * (4 * alpha_hat - 1) * s) / (16 * (alpha_hat + 2) * q + 8 * (1 - 4 * alpha_hat) * s);
        q_pp = (1.0 / 3.0) * (alpha_pp + 2);
        r_pp = (1.0 / 12.0) * (4 * alpha_pp - 1);
        w1 = (2 * alpha_hat + 1) / (2 * (q + s));
        w2 = ((8 * alpha_hat + 7) * q - 6 * (2 * alpha_hat + 1) * r + (8 * alpha_hat + 7) * s) / (9 * (q + s));
        w3 = (4 * (alpha_hat + 2) * q + 2 * (1 - 4 * alpha_hat) * s) / (9 * (q + s));
    }
};

struct Row { double bt, b, d, a, at; };

inline int slot(int bc) { return bc == 0 ? 0 : (bc == 1 ? 1 : 2); }

}  // namespace

int np_build_coefs(int kind, double dx, NpCoefs* c) {
    std::memset(c, 0, sizeof(*c));
    const double onebydx = 1.0 / dx, onebydx2 = onebydx / dx;
    if (kind == NP_CD10_D1) {          // cd10.F90:1143-1160
        const D1B k;
        c->in[0] = k.q_hat * onebydx; c->in[1] = k.r_hat * onebydx; c->in[2] = k.s_hat * onebydx;
        c->r4[0] = k.w4 * k.q_ppp * onebydx; c->r4[1] = k.w4 * k.r_ppp * onebydx; c->r4[2] = k.w4 * k.s_ppp * onebydx;
        c->r3[0] = k.w3 * k.q_pp * onebydx; c->r3[1] = k.w3 * k.r_pp * onebydx;
        c->r2[0] = k.w2 * k.q_p * onebydx;
        c->r1[0] = k.w1 * (k.p * onebydx); c->r1[1] = k.w1 * (k.q * onebydx); c->r1[2] = k.w1 * (k.r * onebydx); c->r1[3] = k.w1 * (k.s * onebydx);
    } else if (kind == NP_CD10_D2) {   // cd10.F90:1637-1651
        c->in[0] = a10d2 * onebydx2; c->in[1] = b10d2 * onebydx2; c->in[2] = c10d2 * onebydx2;
        c->r3[0] = b3_a10d2 * onebydx2; c->r3[1] = b3_b10d2 * onebydx2;
        c->r2[0] = b2_a10d2 * onebydx2;
        c->r1[0] = b1_a10d2 * onebydx2; c->r1[1] = b1_b10d2 * onebydx2; c->r1[2] = b1_c10d2 * onebydx2; c->r1[3] = b1_d10d2 * onebydx2;
        c->r1[4] = b1_e10d2 * onebydx2;
    } else if (kind == NP_CD06_D1) {   // cd06.F90:551-563
        const C6B k;
        c->in[0] = k.qhat * onebydx; c->in[1] = k.rhat * onebydx;
        c->r3[0] = k.w3 * k.q_pp * onebydx; c->r3[1] = k.w3 * k.r_pp * onebydx;
        c->r2[0] = k.w2 * k.q_p * onebydx;
        c->r1[0] = k.w1 * (-k.p * onebydx); c->r1[1] = k.w1 * (k.q * onebydx); c->r1[2] = k.w1 * (k.r * onebydx); c->r1[3] = k.w1 * (k.s * onebydx);
    } else if (kind == NP_CF90) {      // cf90.F90:672-801
        c->in[0] = a90; c->in[1] = b90; c->in[2] = c90; c->in[3] = d90; c->in[4] = e90;
        c->r1[0] = 1.0;
        c->r2[0] = b2_a90; c->r2[1] = b2_b90;
        c->r3[0] = b3_a90; c->r3[1] = b3_b90; c->r3[2] = b3_c90;
        c->r4[0] = b4_a90; c->r4[1] = b4_b90; c->r4[2] = b4_c90; c->r4[3] = b4_d90;
    } else if (kind == NP_GAUSS || kind == NP_LSTSQ) {     // gaussian.F90:15-46, lstsq.F90:14-46 (the same boundary rows)
        if (kind == NP_GAUSS) {
            c->in[0] = 3565.0 / 10368.0; c->in[1] = 3091.0 / 12960.0; c->in[2] = 1997.0 / 25920.0; c->in[3] = 149.0 / 12960.0; c->in[4] = 107.0 / 103680.0;
        } else {
            c->in[0] = 0.5; c->in[1] = (double)0.6744132f / 2.0; c->in[2] = 0.0 / 2.0; c->in[3] = (double)(-0.1744132f) / 2.0; c->in[4] = 0.0 / 2.0;   // real(0.6744132, rkind): a default-real literal widened to double
        }
        c->r1[0] = 5.0 / 6.0; c->r1[1] = 1.0 / 6.0;
        c->r2[0] = 2.0 / 3.0; c->r2[1] = 1.0 / 6.0;
        c->r3[0] = 31.0 / 64.0; c->r3[1] = 7.0 / 32.0; c->r3[2] = 5.0 / 128.0;
        c->r4[0] = 17.0 / 48.0; c->r4[1] = 15.0 / 64.0; c->r4[2] = 7.0 / 96.0; c->r4[3] = 1.0 / 64.0;
    } else {
        return -1;
    }
    return 0;
}

// The rows of ComputePenta1 / ComputePenta2 (cd10.F90:429-554, 577-686) and cf90's ComputePenta (cf90.F90:276-398):
// rows5n = bt[n] b[n] d[n] a[n] at[n] (sub-sub, sub, diagonal, super, super-super entries of row i).
static int np_rows(int kind, int n, int bc1, int bcn, std::vector<Row>& R);

int np_build_rows(int kind, int n, int bc1, int bcn, double* rows5n) {
    std::vector<Row> R;
    if (int rc = np_rows(kind, n, bc1, bcn, R)) return rc;
    for (int i = 1; i <= n; ++i) {
        rows5n[i - 1] = R[i].bt; rows5n[(size_t)n + i - 1] = R[i].b; rows5n[2 * (size_t)n + i - 1] = R[i].d;
        rows5n[3 * (size_t)n + i - 1] = R[i].a; rows5n[4 * (size_t)n + i - 1] = R[i].at;
    }
    return 0;
}

static int np_rows(int kind, int n, int bc1, int bcn, std::vector<Row>& R) {
    if (kind == NP_CD06_D1) {   // ComputeTri1 (cd06.F90:264-327): a tridiagonal system carried as a pentadiagonal one with empty outer bands
        if (n < 6) return 3;
        if (bc1 != 0 || bcn != 0) return 1002;   // the reference's symmetric cases are marked incomplete
        const C6B k;
        R.assign((size_t)n + 1, Row{0, k.alpha_hat, 1.0, k.alpha_hat, 0});
        const Row one_sided[3] = {{0, k.w1 * 0, k.w1 * 1, k.w1 * k.alpha, 0}, {0, k.w2 * k.alpha_p, k.w2 * 1, k.w2 * k.alpha_p, 0},
                                  {0, k.w3 * k.alpha_pp, k.w3 * 1, k.w3 * k.alpha_pp, 0}};
        for (int j = 0; j < 3; ++j) { R[1 + j] = one_sided[j]; const Row& s = one_sided[j]; R[n - j] = Row{0, s.a, s.d, s.b, 0}; }
        return 0;
    }
    if (kind == NP_CF90 ? n < 10 : n < 8) return kind == NP_CF90 ? 7 : 2;
    if ((bc1 != 0 && bc1 != 1 && bc1 != -1) || (bcn != 0 && bcn != 1 && bcn != -1)) return 324;
    R.assign((size_t)n + 1, Row{0, 0, 0, 0, 0});   // 1-based
    double al, be;
    if (kind == NP_CD10_D1) { al = alpha10d1; be = beta10d1; }
    else if (kind == NP_CD10_D2) { al = alpha10d2; be = beta10d2; }
    else { al = alpha90; be = beta90; }
    for (int i = 1; i <= n; ++i) R[i] = Row{be, al, 1.0, al, be};
    if (kind == NP_CD10_D1) {
        const D1B k;
        const Row one_sided[4] = {{k.w1 * 0, k.w1 * 0, k.w1 * 1, k.w1 * k.alpha, k.w1 * 0},
                                  {k.w2 * 0, k.w2 * k.alpha_p, k.w2 * 1, k.w2 * k.alpha_p, k.w2 * 0},
                                  {k.w3 * k.beta_pp, k.w3 * k.alpha_pp, k.w3 * 1, k.w3 * k.alpha_pp, k.w3 * k.beta_pp},
                                  {k.w4 * k.beta_ppp, k.w4 * k.alpha_ppp, k.w4 * 1, k.w4 * k.alpha_ppp, k.w4 * k.beta_ppp}};
        if (bc1 == 0) for (int j = 0; j < 4; ++j) R[1 + j] = one_sided[j];
        if (bc1 == 1) { R[1] = Row{0, 0, 1, 0, 0}; R[2] = Row{0, al, 1 - be, al, be}; }
        if (bc1 == -1) { R[1] = Row{0, 0, 1, 2 * al, 2 * be}; R[2] = Row{0, al, 1 + be, al, be}; }
        if (bcn == 0) for (int j = 0; j < 4; ++j) { const Row& s = one_sided[j]; R[n - j] = Row{s.at, s.a, s.d, s.b, s.bt}; }   // mirrored
        if (bcn == 1) { R[n] = Row{0, 0, 1, 0, 0}; R[n - 1] = Row{be, al, 1 - be, al, 0}; }
        if (bcn == -1) { R[n] = Row{2 * be, 2 * al, 1, 0, 0}; R[n - 1] = Row{be, al, 1 + be, al, 0}; }
    } else if (kind == NP_CD10_D2) {
        const Row one_sided[3] = {{0, 0, 1, b1_alpha10d2, 0}, {0, b2_alpha10d2, 1, b2_alpha10d2, 0},
                                  {b3_beta10d2, b3_alpha10d2, 1, b3_alpha10d2, b3_beta10d2}};
        if (bc1 == 0) for (int j = 0; j < 3; ++j) R[1 + j] = one_sided[j];
        if (bc1 == 1) { R[1] = Row{0, 0, 1, 2 * al, 2 * be}; R[2] = Row{0, al, 1 + be, al, be}; }
        if (bc1 == -1) { R[1] = Row{0, 0, 1, 0, 0}; R[2] = Row{0, al, 1 - be, al, be}; }
        if (bcn == 0) for (int j = 0; j < 3; ++j) { const Row& s = one_sided[j]; R[n - j] = Row{s.at, s.a, s.d, s.b, s.bt}; }
        if (bcn == 1) { R[n] = Row{2 * be, 2 * al, 1, 0, 0}; R[n - 1] = Row{be, al, 1 + be, al, 0}; }
        if (bcn == -1) { R[n] = Row{0, 0, 1, 0, 0}; R[n - 1] = Row{be, al, 1 - be, al, 0}; }
    } else {
        const Row one_sided[4] = {{0, 0, 1, 0, 0}, {0, b2_alpha90, 1, b2_alpha90, 0}, {b3_beta90, b3_alpha90, 1, b3_alpha90, b3_beta90},
                                  {b4_beta90, b4_alpha90, 1, b4_alpha90, b4_beta90}};
        if (bc1 == 0) for (int j = 0; j < 4; ++j) R[1 + j] = one_sided[j];
        if (bc1 == 1) { R[1] = Row{0, 0, 1, 2 * al, 2 * be}; R[2] = Row{0, al, 1 + be, al, be}; }
        if (bc1 == -1) { R[1] = Row{0, 0, 1, 0, 0}; R[2] = Row{0, al, 1 - be, al, be}; }
        if (bcn == 0) for (int j = 0; j < 4; ++j) { const Row& s = one_sided[j]; R[n - j] = Row{s.at, s.a, s.d, s.b, s.bt}; }
        if (bcn == 1) { R[n] = Row{2 * be, 2 * al, 1, 0, 0}; R[n - 1] = Row{be, al, 1 + be, al, 0}; }
        if (bcn == -1) { R[n] = Row{0, 0, 1, 0, 0}; R[n - 1] = Row{be, al, 1 - be, al, 0}; }
    }
    return 0;
}

// rows, then the LU recurrences of ComputePenta* (cd10.F90:556-572)
int np_build_table(int kind, int n, int bc1, int bcn, double* tab) {
    std::vector<Row> R;
    if (int rc = np_rows(kind, n, bc1, bcn, R)) return rc;
    // Steps 1-3: obc = 1/pivot, e = modified super-diagonal, f / g = multipliers
    std::vector<double> e((size_t)n + 1, 0.0), obc((size_t)n + 1, 0.0), f((size_t)n + 1, 0.0), g((size_t)n + 1, 0.0);
    obc[1] = 1.0 / R[1].d;
    obc[2] = 1.0 / (R[2].d - R[2].b * R[1].a * obc[1]);
    e[1] = R[1].a;
    f[2] = R[2].b * obc[1];
    for (int i = 3; i <= n; ++i) {
        g[i] = R[i].bt * obc[i - 2];
        e[i - 1] = R[i - 1].a - f[i - 1] * R[i - 2].at;
        f[i] = (R[i].b - g[i] * e[i - 2]) * obc[i - 1];
        obc[i] = 1.0 / (R[i].d - f[i] * e[i - 1] - g[i] * R[i - 2].at);
    }
    for (int i = 1; i <= n; ++i) {
        tab[i - 1] = f[i];
        tab[(size_t)n + i - 1] = g[i];
        tab[2 * (size_t)n + i - 1] = obc[i];
        tab[3 * (size_t)n + i - 1] = R[i].at;
        tab[4 * (size_t)n + i - 1] = e[i] * obc[i];
    }
    return 0;
}

namespace {

struct LineAcc {
    const double* p;   // element (1) of the line
    long long es;
    __host__ __device__ double operator()(int j) const { return p[(long long)(j - 1) * es]; }
};

// f(n1, n, n3): one thread per point, x fastest (coalesced); writes the RHS into out
template <int KIND>
__global__ void __launch_bounds__(256) np_rhs_kernel(const double* __restrict__ f, double* __restrict__ out, long long n1, int n, long long n3,
                                                     int bc1, int bcn, const __grid_constant__ NpCoefs co) {
    const long long tot = n1 * n * n3;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
        const long long x = idx % n1;
        const long long r = idx / n1;
        const int i = (int)(r % n);
        const long long k = r / n;
        const LineAcc F{f + k * n1 * n + x, n1};
        out[idx] = np_rhs_point<KIND>(i + 1, n, bc1, bcn, co, F);
    }
}

__global__ void __launch_bounds__(128) np_solve_kernel(double* __restrict__ y, long long n1, int n, long long n3, const double* __restrict__ tab) {
    const long long nlines = n1 * n3;
    for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (long long)gridDim.x * blockDim.x) {
        const long long k = l / n1, x = l - k * n1;
        np_solve_line(y + k * n1 * n + x, n1, n, tab);
    }
}

void shape(int axis, int n, long long na, long long nb, long long* n1, long long* n3) {
    if (axis == 0) { *n1 = 1; *n3 = na * nb; }
    else if (axis == 1) { *n1 = na; *n3 = nb; }
    else { *n1 = na * nb; *n3 = 1; }
    (void)n;
}

unsigned blocks_for(long long work, int threads) {
    long long b = (work + threads - 1) / threads;
    const long long cap = 148LL * 32;
    return (unsigned)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace

cudaError_t np_op_create(NpOp* h, int kind, int n, double dx, int* ierr_out) {
    *ierr_out = 0;
    h->kind = kind; h->n = n;
    for (auto& p : h->d_tab) p = nullptr;
    if (np_build_coefs(kind, dx, &h->co) != 0) return cudaErrorInvalidValue;
    if (n == 1) return cudaSuccess;   // degenerate: handled by the callers (derivative 0 / filter identity)
    if (kind == NP_GAUSS || kind == NP_LSTSQ) {           // explicit filters: no system, no tables
        if (n < 8) *ierr_out = 1001;
        return cudaSuccess;
    }
    std::vector<double> tab(5 * (size_t)n);
    const int codes[3] = {0, 1, -1};
    const int ncodes = kind == NP_CD06_D1 ? 1 : 3;   // cd06: the one-sided closure only
    for (int a = 0; a < ncodes; ++a)
        for (int b = 0; b < ncodes; ++b) {
            const int rc = np_build_table(kind, n, codes[a], codes[b], tab.data());
            if (rc) { *ierr_out = rc; np_op_destroy(h); return cudaSuccess; }
            double* d = nullptr;
            cudaError_t e = cudaMalloc(&d, sizeof(double) * tab.size());
            if (e == cudaSuccess) e = cudaMemcpy(d, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { if (d) cudaFree(d); np_op_destroy(h); return e; }
            h->d_tab[3 * a + b] = d;
            e = np_fast_create(&h->fast[3 * a + b], kind, n, codes[a], codes[b]);
            if (e != cudaSuccess) { np_op_destroy(h); return e; }
        }
    return cudaSuccess;
}

void np_op_destroy(NpOp* h) {
    for (auto& p : h->d_tab) { if (p) cudaFree(p); p = nullptr; }
    for (auto& t : h->fast) np_fast_destroy(&t);
}

namespace {
int g_np_fast = -1;
bool np_fast_enabled() {
    if (g_np_fast < 0) {
        const char* e = std::getenv("PDO_NP_FAST");
        g_np_fast = (e && std::atoi(e) == 0) ? 0 : 1;
    }
    return g_np_fast != 0;
}
}  // namespace
void np_set_fast_path(int mode) { g_np_fast = mode; }

cudaError_t np_op_apply(const NpOp* h, int axis, const double* f, double* out, long long na, long long nb, int bc1, int bcn, cudaStream_t st) {
    if (axis < 0 || axis > 2 || (bc1 != 0 && bc1 != 1 && bc1 != -1) || (bcn != 0 && bcn != 1 && bcn != -1)) return cudaErrorInvalidValue;
    long long n1, n3;
    shape(axis, h->n, na, nb, &n1, &n3);
    const long long tot = n1 * h->n * n3;
    if (tot == 0) return cudaSuccess;
    const unsigned gb = blocks_for(tot, 256);
    if (h->kind == NP_GAUSS || h->kind == NP_LSTSQ) {   // the two kinds differ in their coefficients only
        np_rhs_kernel<NP_GAUSS><<<gb, 256, 0, st>>>(f, out, n1, h->n, n3, bc1, bcn, h->co);
        return cudaGetLastError();
    }
    const double* tab = h->d_tab[3 * slot(bc1) + slot(bcn)];
    if (!tab) return cudaErrorInvalidValue;
    // chunked fast path: one fused pass (np_chunk.cu); lines it does not cover keep the sweeps below
    const NpFast& fast = h->fast[3 * slot(bc1) + slot(bcn)];
    if (fast.ok && np_fast_enabled()) {
        const cudaError_t e = np_fast_apply(fast, h->kind, h->co, h->n, axis, f, out, n1, n3, bc1, bcn, st);
        if (e != cudaErrorInvalidConfiguration) return e;
        cudaGetLastError();
    }
    if (h->kind == NP_CD06_D1) np_rhs_kernel<NP_CD06_D1><<<gb, 256, 0, st>>>(f, out, n1, h->n, n3, bc1, bcn, h->co);
    else if (h->kind == NP_CD10_D1) np_rhs_kernel<NP_CD10_D1><<<gb, 256, 0, st>>>(f, out, n1, h->n, n3, bc1, bcn, h->co);
    else if (h->kind == NP_CD10_D2) np_rhs_kernel<NP_CD10_D2><<<gb, 256, 0, st>>>(f, out, n1, h->n, n3, bc1, bcn, h->co);
    else np_rhs_kernel<NP_CF90><<<gb, 256, 0, st>>>(f, out, n1, h->n, n3, bc1, bcn, h->co);
    np_solve_kernel<<<blocks_for(n1 * n3, 128), 128, 0, st>>>(out, n1, h->n, n3, tab);
    return cudaGetLastError();
}

int np_apply_host(int kind, int n, double dx, int bc1, int bcn, int axis, const double* f, double* out, long long na, long long nb) {
    NpCoefs co;
    if (np_build_coefs(kind, dx, &co) != 0) return -1;
    std::vector<double> tab(5 * (size_t)n);
    if (kind == NP_GAUSS || kind == NP_LSTSQ) { if (n < 8) return 1001; }
    else if (int rc = np_build_table(kind, n, bc1, bcn, tab.data())) return rc;
    long long n1, n3;
    shape(axis, n, na, nb, &n1, &n3);
    for (long long k = 0; k < n3; ++k)
        for (long long x = 0; x < n1; ++x) {
            const LineAcc F{f + k * n1 * n + x, n1};
            double* y = out + k * n1 * n + x;
            for (int i = 0; i < n; ++i) {
                double v;
                if (kind == NP_CD06_D1) v = np_rhs_point<NP_CD06_D1>(i + 1, n, bc1, bcn, co, F);
                else if (kind == NP_CD10_D1) v = np_rhs_point<NP_CD10_D1>(i + 1, n, bc1, bcn, co, F);
                else if (kind == NP_CD10_D2) v = np_rhs_point<NP_CD10_D2>(i + 1, n, bc1, bcn, co, F);
                else if (kind == NP_GAUSS || kind == NP_LSTSQ) v = np_rhs_point<NP_GAUSS>(i + 1, n, bc1, bcn, co, F);
                else v = np_rhs_point<NP_CF90>(i + 1, n, bc1, bcn, co, F);
                y[(long long)i * n1] = v;
            }
            if (kind != NP_GAUSS && kind != NP_LSTSQ) np_solve_line(y, n1, n, tab.data());
        }
    return 0;
}

}  // namespace pdo
