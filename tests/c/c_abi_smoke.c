/* c_abi_smoke.c — plain C99 caller of libpadeops_b200.so through include/padeops_b200.h, exactly the way the Fortran shim
 * modules (fortran/ *.F90, bind(C) interfaces) call it: by-value ints and doubles, void* fields, host and device pointers,
 * int status codes.  tests/test_c_abi_smoke.py compiles it with gcc (no CUDA headers, no Python in the loop) and runs it on
 * the GPU box.  Mirrors tests/test_cd10.F90:51-81 (sin x + sin y + sin z on a periodic box, d1 and d2 on every axis),
 * tests/test_cf90.F90 (single-mode transfer function) and tests/test_PoissonPeriodic.F90:109-118 (manufactured solution).
 * Exit code 0 = every check passed; otherwise the number of failed checks (messages on stderr). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "padeops_b200.h"

static int nfail = 0;
#define CHECK(cond, ...)                                                  \
    do {                                                                  \
        if (!(cond)) {                                                    \
            ++nfail;                                                      \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__);          \
            fprintf(stderr, __VA_ARGS__);                                 \
            fprintf(stderr, " [%s]\n", pdo_last_error());                 \
        }                                                                 \
    } while (0)

static double maxabs_diff(const double* a, const double* b, size_t n) {
    double m = 0.0;
    for (size_t i = 0; i < n; ++i) { const double d = fabs(a[i] - b[i]); if (d > m) m = d; }
    return m;
}

int main(void) {
    const double pi = 3.14159265358979323846;
    const int n = 64;                      /* 64^3 periodic box */
    const double dx = 2.0 * pi / n;
    const size_t N = (size_t)n * n * n, bytes = N * sizeof(double);
    double *f = malloc(bytes), *df = malloc(bytes), *ref = malloc(bytes);
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) f[((size_t)k * n + j) * n + i] = sin(i * dx) + sin(j * dx) + sin(k * dx);

    /* ---- error conventions of the constructors (cd10.F90:219-226, cd06.F90:158, cf90.F90:128, cd06stagg n <= 4) ---- */
    pdo_cd10_t c10 = NULL; pdo_cd06_t c06 = NULL; pdo_cf90_t cf = NULL; pdo_cd06stagg_t st = NULL;
    CHECK(pdo_cd10_init(&c10, 5, dx, 1, 0, 0) == 2, "cd10 n = 5 must return 2");
    CHECK(pdo_cd06_init(&c06, 4, dx, 1, 0, 0) == 3, "cd06 n = 4 must return 3");
    CHECK(pdo_cf90_init(&cf, 9, 1) == 7, "cf90 n = 9 must return 7");
    CHECK(pdo_cd06stagg_init_periodic(&st, 4, dx) == 21, "cd06stagg n = 4 must return 21");

    /* ---- cd10 on HOST arrays (the unmodified-caller path) ---- */
    CHECK(pdo_cd10_init(&c10, n, dx, 1, 0, 0) == 0, "cd10 init");
    CHECK(pdo_cd10_getsize(c10) == n, "GetSize");
    typedef int (*dd_fn)(pdo_cd10_t, const double*, double*, int, int, int, int, void*);
    const dd_fn d1[3] = {pdo_cd10_dd1, pdo_cd10_dd2, pdo_cd10_dd3}, d2[3] = {pdo_cd10_d2d1, pdo_cd10_d2d2, pdo_cd10_d2d3};
    for (int ax = 0; ax < 3; ++ax) {
        CHECK(d1[ax](c10, f, df, n, n, 0, 0, NULL) == 0, "dd%d (host)", ax + 1);
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) ref[((size_t)k * n + j) * n + i] = cos((ax == 0 ? i : ax == 1 ? j : k) * dx);
        CHECK(maxabs_diff(df, ref, N) < 1e-12, "dd%d: max error %.3e (10th order on one mode: round-off only)", ax + 1, maxabs_diff(df, ref, N));
        CHECK(d2[ax](c10, f, df, n, n, 0, 0, NULL) == 0, "d2d%d (host)", ax + 1);
        for (size_t q = 0; q < N; ++q) ref[q] = -ref[q];   /* placeholder, overwritten below */
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) ref[((size_t)k * n + j) * n + i] = -sin((ax == 0 ? i : ax == 1 ? j : k) * dx);
        CHECK(maxabs_diff(df, ref, N) < 1e-11, "d2d%d: max error %.3e", ax + 1, maxabs_diff(df, ref, N));
    }
    CHECK(pdo_cd10_dd1(c10, f, df, n, n, 2, 0, NULL) == 324, "bad boundary code must return 324 (cd10.F90:2044-2046)");

    /* ---- the same on DEVICE-resident fields (type(c_ptr) fields held by the Fortran caller) ---- */
    void *d_f = NULL, *d_df = NULL;
    CHECK(pdo_malloc(&d_f, bytes) == 0 && pdo_malloc(&d_df, bytes) == 0, "pdo_malloc");
    CHECK(pdo_h2d(d_f, f, bytes, NULL) == 0, "pdo_h2d");
    const long long l0 = (long long)pdo_launch_count();
    CHECK(pdo_cd10_dd2(c10, (const double*)d_f, (double*)d_df, n, n, 0, 0, NULL) == 0, "dd2 (device)");
    CHECK((long long)pdo_launch_count() > l0, "a device call must launch a kernel of this library");
    CHECK(pdo_d2h(df, d_df, bytes, NULL) == 0 && pdo_stream_sync(NULL) == 0, "pdo_d2h");
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) ref[((size_t)k * n + j) * n + i] = cos(j * dx);
    CHECK(maxabs_diff(df, ref, N) < 1e-12, "dd2 (device): max error %.3e", maxabs_diff(df, ref, N));
    int v1 = 0, v2 = 0;
    CHECK(pdo_cd10_plan(c10, 1, n, n, &v1, &v2) == 0, "pdo_cd10_plan");

    /* ---- cf90: a single mode comes back scaled by the transfer function (tests/test_cf90.F90:108-116) ---- */
    CHECK(pdo_cf90_init(&cf, n, 1) == 0, "cf90 init");
    {
        const int kw = 5;
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) f[((size_t)k * n + j) * n + i] = cos(kw * i * dx);
        CHECK(pdo_cf90_filter1(cf, f, df, n, n, 0, 0, NULL) == 0, "filter1");
        const double w = kw * dx;
        const double T = (9.9965e-1 + 2 * 6.6652e-1 * cos(w) + 2 * 1.6674e-1 * cos(2 * w) + 2 * 4.0e-5 * cos(3 * w) + 2 * -5.0e-6 * cos(4 * w)) /
                         (1.0 + 2 * 6.6624e-1 * cos(w) + 2 * 1.6688e-1 * cos(2 * w));
        for (size_t q = 0; q < N; ++q) ref[q] = T * f[q];
        CHECK(maxabs_diff(df, ref, N) < 1e-13, "cf90 transfer function: max error %.3e", maxabs_diff(df, ref, N));
    }

    /* ---- decomposition arithmetic + single-rank transposes + PoissonPeriodic (test_PoissonPeriodic.F90:109-118) ---- */
    CHECK(pdo_comm_init(0, 1, NULL) == 0, "pdo_comm_init (one rank)");
    {
        pdo_decomp_info inf;
        CHECK(pdo_decomp_info_for(17, 9, 5, 2, 2, 3, &inf) == 0, "decomp_info_for");
        CHECK(inf.xsz[0] == 17 && inf.ysz[1] == 9 && inf.zsz[2] == 5, "a pencil holds whole lines along its axis");
        CHECK(pdo_decomp_info_for(4, 4, 4, 8, 1, 0, &inf) == 6, "bad 2D grid must return 6 (decomp_2d.f90:507-514)");
        pdo_decomp_t dc = NULL;
        CHECK(pdo_decomp_init(&dc, n, n, n, 1, 1) == 0, "decomp init");
        CHECK(pdo_transpose_x_to_y(dc, f, df, 1, NULL) == 0, "transpose_x_to_y (host arrays)");
        CHECK(memcmp(f, df, bytes) == 0, "one rank: the transpose is the identity, bit for bit");
        pdo_decomp_destroy(dc);
        const int nx = 64, ny = 32, nz = 16;
        const double hx = 2 * pi / nx, hy = 2 * pi / ny, hz = 2 * pi / nz;
        const size_t M = (size_t)nx * ny * nz;
        double *rhs = malloc(M * sizeof(double)), *sol = malloc(M * sizeof(double));
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    const double ft = sin(6 * i * hx) * cos(3 * j * hy) * sin(1 * k * hz);
                    sol[((size_t)k * ny + j) * nx + i] = ft;
                    rhs[((size_t)k * ny + j) * nx + i] = -(36.0 + 9.0 + 1.0) * ft;
                }
        pdo_poisson_t po = NULL;
        CHECK(pdo_poisson_init(&po, nx, ny, nz, hx, hy, hz, 1, 1, 1, NULL, NULL, NULL) == 0, "PoissonPeriodic init");
        CHECK(pdo_poisson_solve(po, rhs, rhs, NULL) == 0, "poisson_solve (in place, host array)");
        CHECK(maxabs_diff(rhs, sol, M) < 1e-13, "Poisson manufactured solution: max error %.3e", maxabs_diff(rhs, sol, M));
        pdo_poisson_destroy(po);
        free(rhs); free(sol);
    }

    pdo_free(d_f); pdo_free(d_df);
    pdo_cd10_destroy(c10); pdo_cf90_destroy(cf);
    pdo_comm_finalize();
    free(f); free(df); free(ref);
    if (nfail == 0) printf("C_ABI_SMOKE PASS (launches = %lld, planned variants %d / %d)\n", (long long)pdo_launch_count(), v1, v2);
    return nfail;
}
