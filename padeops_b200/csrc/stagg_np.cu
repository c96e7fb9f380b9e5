// stagg_np.cu — host tables, kernel and host twin of the non-periodic staggered compact operators (stagg_np.cuh).
#include <cstring>
#include <vector>

#include "stagg_np.cuh"

namespace pdo {

StaggNpConst snp_constants(double dx) {
    // cd06stagg.F90:17-56
    StaggNpConst c{};
    c.o = 1.0 / dx;
    c.o2 = c.o / dx;
    c.p = 17.0 / 6.0; c.q = 3.0 / 2.0; c.r = 3.0 / 2.0; c.s = -1.0 / 6.0;
    const double alpha_hat = 1.0 / 3.0;
    c.q_p = 3.0 / 4.0;
    const double alpha_pp = ((40 * alpha_hat - 1) * c.q + 7 * (4 * alpha_hat - 1) * c.s) / (16 * (alpha_hat + 2) * c.q + 8 * (1 - 4 * alpha_hat) * c.s);
    c.q_pp = (1.0 / 3.0) * (alpha_pp + 2);
    c.r_pp = (1.0 / 12.0) * (4 * alpha_pp - 1);
    c.w1 = (2 * alpha_hat + 1) / (2 * (c.q + c.s));
    c.w2 = ((8 * alpha_hat + 7) * c.q - 6 * (2 * alpha_hat + 1) * c.r + (8 * alpha_hat + 7) * c.s) / (9 * (c.q + c.s));
    c.w3 = (4 * (alpha_hat + 2) * c.q + 2 * (1 - 4 * alpha_hat) * c.s) / (9 * (c.q + c.s));
    return c;
}

// ComputeTri_allRoutines.F90: the eight systems, row by row (0-based: Fortran row i -> i - 1)
int snp_build_rows(int op, int n, const StaggNpFlags& fl, double* rows) {
    if (n <= 4) return 21;   // cd06stagg.F90:216-218
    if (op < 0 || op >= SNP_COUNT) return 1001;
    const int m = snp_rows_out(op, n);
    double *ddn = rows, *dg = rows + m, *dup = rows + 2 * (size_t)m;
    const StaggNpConst c = snp_constants(1.0);
    const double alpha_p = 1.0 / 4.0, alpha_hat = 1.0 / 3.0;
    const double alpha_pp = ((40 * alpha_hat - 1) * c.q + 7 * (4 * alpha_hat - 1) * c.s) / (16 * (alpha_hat + 2) * c.q + 8 * (1 - 4 * alpha_hat) * c.s);
    const double w0s = 223.0 / 186.0, w1s = 61.0 / 62.0;
    auto fill = [&](double al) { for (int i = 0; i < m; ++i) { ddn[i] = al; dg[i] = 1.0; dup[i] = al; } };
    switch (op) {
        case SNP_D1_E2C: {          // :1-47
            const double al = 9.0 / 62.0, al1 = 37.0 / 183.0, al0 = -1.0;
            fill(al);
            if (fl.topSided) { ddn[m - 1] = w0s * al0; dg[m - 1] = w0s; ddn[m - 2] = w1s * al1; dup[m - 2] = w1s * al1; dg[m - 2] = w1s; }
            else if (fl.topEven) { dup[m - 1] = 0.0; dg[m - 1] = 1.0 - al; }
            else { dg[m - 1] = 1.0 + al; dup[m - 1] = 0.0; }
            if (fl.botSided) { dup[0] = w0s * al0; dg[0] = w0s; ddn[1] = w1s * al1; dup[1] = w1s * al1; dg[1] = w1s; }
            else if (fl.botEven) { ddn[0] = 0.0; dg[0] = 1.0 - al; }
            else { ddn[0] = 0.0; dg[0] = 1.0 + al; }
            break;
        }
        case SNP_D1_C2E: {          // :49-92
            const double al = 9.0 / 62.0;
            fill(al);
            if (fl.topSided) { ddn[m - 1] = 0.0; dup[m - 1] = 0.0; ddn[m - 2] = 1.0 / 22.0; dup[m - 2] = 1.0 / 22.0; }
            else if (fl.topEven) { dup[m - 1] = 0.0; ddn[m - 1] = 0.0; }
            else { ddn[m - 1] = 2.0 * al; dup[m - 1] = 0.0; }
            if (fl.botSided) { dup[0] = 0.0; ddn[0] = 0.0; dup[1] = 1.0 / 22.0; ddn[1] = 1.0 / 22.0; }
            else if (fl.botEven) { ddn[0] = 0.0; dup[0] = 0.0; }
            else { ddn[0] = 0.0; dup[0] = 2.0 * al; }
            break;
        }
        case SNP_D1_C2C: {          // :94-160
            const double al = 1.0 / 3.0, alLOW = 3.0;
            fill(al);
            if (fl.topSided) {
                dup[m - 1] = c.w1 * 0.0; dup[m - 2] = c.w2 * alpha_p; dup[m - 3] = c.w3 * alpha_pp;
                dg[m - 1] = c.w1 * 1.0; dg[m - 2] = c.w2 * 1.0; dg[m - 3] = c.w3 * 1.0;
                ddn[m - 1] = c.w1 * alLOW; ddn[m - 2] = c.w2 * alpha_p; ddn[m - 3] = c.w3 * alpha_pp;
            } else if (fl.topEven) dg[m - 1] = 1.0 - al;
            else dg[m - 1] = 1.0 + al;
            if (fl.botSided) {
                ddn[0] = c.w1 * 0.0; ddn[1] = c.w2 * alpha_p; ddn[2] = c.w3 * alpha_pp;
                dg[0] = c.w1 * 1.0; dg[1] = c.w2 * 1.0; dg[2] = c.w3 * 1.0;
                dup[0] = c.w1 * alLOW; dup[1] = c.w2 * alpha_p; dup[2] = c.w3 * alpha_pp;
            } else if (fl.botEven) dg[0] = 1.0 - al;
            else dg[0] = 1.0 + al;
            break;
        }
        case SNP_D1_E2E: {          // :162-195
            const double al = 1.0 / 3.0;
            fill(al);
            if (fl.topEven) { dg[m - 1] = 1.0; ddn[m - 1] = 0.0; } else { dg[m - 1] = 1.0; ddn[m - 1] = 2.0 * al; }
            if (fl.botEven) { dg[0] = 1.0; dup[0] = 0.0; } else { dg[0] = 1.0; dup[0] = 2.0 * al; }
            break;
        }
        case SNP_INTERP_C2E: {      // :197-241
            const double al = 3.0 / 10.0, al1 = 1.0 / 6.0;
            fill(al);
            if (fl.topSided) { dup[m - 1] = 0.0; ddn[m - 1] = 0.0; dup[m - 2] = al1; ddn[m - 2] = al1; }
            else if (fl.topEven) { ddn[m - 1] = 2.0 * al; dg[m - 1] = 1.0; }
            else { ddn[m - 1] = 0.0; dg[m - 1] = 1.0; }
            if (fl.botSided) { dup[0] = 0.0; ddn[0] = 0.0; dg[0] = 1.0; dup[1] = al1; ddn[1] = al1; dg[1] = 1.0; }
            else if (fl.botEven) { dup[0] = 2.0 * al; dg[0] = 1.0; }
            else { dup[0] = 0.0; dg[0] = 1.0; }
            break;
        }
        case SNP_INTERP_E2C: {      // :244-287
            const double al = 3.0 / 10.0, al0 = 1.0;
            fill(al);
            if (fl.topSided) { ddn[m - 1] = al0; dup[m - 1] = 0.0; }
            else if (fl.topEven) dg[m - 1] = 1.0 + al;
            else dg[m - 1] = 1.0 - al;
            if (fl.botSided) { dup[0] = al0; ddn[0] = 0.0; }
            else if (fl.botEven) dg[0] = 1.0 + al;
            else dg[0] = 1.0 - al;
            break;
        }
        case SNP_D2_E2E: {          // :290-325
            const double al = 2.0 / 11.0;
            fill(al);
            if (fl.topEven) { dg[m - 1] = 1.0; ddn[m - 1] = 2.0 * al; } else { dg[m - 1] = 1.0; ddn[m - 1] = 0.0; }
            if (fl.botEven) { dg[0] = 1.0; dup[0] = 2.0 * al; } else { dg[0] = 1.0; dup[0] = 0.0; }
            break;
        }
        default: {                  // SNP_D2_C2C :327-366
            const double al = 2.0 / 11.0;
            fill(al);
            dg[m - 1] = fl.topEven ? 1.0 + al : 1.0 - al;
            dg[0] = fl.botEven ? 1.0 + al : 1.0 - al;
            break;
        }
    }
    return 0;
}

int snp_build_table(int op, int n, const StaggNpFlags& fl, double* tab) {
    const int m = snp_rows_out(op, n);
    std::vector<double> rows(3 * (size_t)m);
    if (int rc = snp_build_rows(op, n, fl, rows.data())) return rc;
    const double *ddn = rows.data(), *dg = ddn + m, *dup = ddn + 2 * (size_t)m;
    double *t1 = tab, *t2 = tab + m, *t3 = tab + 2 * (size_t)m;   // ddn*den, den, cp
    std::vector<double> cp(m, 0.0), den(m, 0.0);
    cp[0] = dup[0] / dg[0];
    for (int i = 1; i < m - 1; ++i) cp[i] = dup[i] / (dg[i] - ddn[i] * cp[i - 1]);
    den[0] = 1.0 / dg[0];
    for (int i = 1; i < m; ++i) den[i] = 1.0 / (dg[i] - ddn[i] * cp[i - 1]);
    for (int i = 0; i < m; ++i) { t1[i] = ddn[i] * den[i]; t2[i] = den[i]; t3[i] = cp[i]; }
    return 0;
}

namespace {

template <int OP>
__global__ void __launch_bounds__(128) snp_kernel(const double* __restrict__ in, double* __restrict__ out, long long ncols, int n,
                                                  StaggNpFlags fl, StaggNpConst co, const double* __restrict__ tab) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    snp_line<OP>(in + c, out + c, ncols, n, fl, co, tab);
}

template <int OP>
void snp_host(const double* in, double* out, long long ncols, int n, const StaggNpFlags& fl, const StaggNpConst& co, const double* tab) {
    for (long long c = 0; c < ncols; ++c) snp_line<OP>(in + c, out + c, ncols, n, fl, co, tab);
}

}  // namespace

cudaError_t snp_create(StaggNp* h, int n, double dx, const StaggNpFlags& fl, int* ierr_out) {
    *ierr_out = 0;
    h->n = n; h->fl = fl; h->co = snp_constants(dx);
    for (int op = 0; op < SNP_COUNT; ++op) {
        const int m = snp_rows_out(op, n);
        std::vector<double> tab(3 * (size_t)(m > 0 ? m : 1));
        if (int rc = snp_build_table(op, n, fl, tab.data())) { *ierr_out = rc; snp_destroy(h); return cudaSuccess; }
        cudaError_t e = cudaMalloc(&h->d_tab[op], sizeof(double) * tab.size());
        if (e == cudaSuccess) e = cudaMemcpy(h->d_tab[op], tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { snp_destroy(h); return e; }
    }
    return cudaSuccess;
}
void snp_destroy(StaggNp* h) {
    for (int op = 0; op < SNP_COUNT; ++op) { if (h->d_tab[op]) cudaFree(h->d_tab[op]); h->d_tab[op] = nullptr; }
}

cudaError_t snp_apply(const StaggNp* h, int op, const double* in, double* out, long long ncols, cudaStream_t st) {
    if (ncols <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((ncols + 127) / 128);
#define SNP_CASE(OP) case OP: snp_kernel<OP><<<blocks, 128, 0, st>>>(in, out, ncols, h->n, h->fl, h->co, h->d_tab[OP]); break;
    switch (op) {
        SNP_CASE(SNP_D1_E2C) SNP_CASE(SNP_D1_C2E) SNP_CASE(SNP_D1_C2C) SNP_CASE(SNP_D1_E2E)
        SNP_CASE(SNP_INTERP_E2C) SNP_CASE(SNP_INTERP_C2E) SNP_CASE(SNP_D2_C2C) SNP_CASE(SNP_D2_E2E)
        default: return cudaErrorInvalidValue;
    }
#undef SNP_CASE
    return cudaGetLastError();
}

int snp_apply_host(int op, int n, double dx, const StaggNpFlags& fl, const double* in, double* out, long long ncols) {
    if (op < 0 || op >= SNP_COUNT) return 1001;
    const int m = snp_rows_out(op, n);
    std::vector<double> tab(3 * (size_t)(m > 0 ? m : 1));
    if (int rc = snp_build_table(op, n, fl, tab.data())) return rc;
    const StaggNpConst co = snp_constants(dx);
#define SNP_CASE(OP) case OP: snp_host<OP>(in, out, ncols, n, fl, co, tab.data()); break;
    switch (op) {
        SNP_CASE(SNP_D1_E2C) SNP_CASE(SNP_D1_C2E) SNP_CASE(SNP_D1_C2C) SNP_CASE(SNP_D1_E2E)
        SNP_CASE(SNP_INTERP_E2C) SNP_CASE(SNP_INTERP_C2E) SNP_CASE(SNP_D2_C2C) SNP_CASE(SNP_D2_E2E)
    }
#undef SNP_CASE
    return 0;
}

}  // namespace pdo
