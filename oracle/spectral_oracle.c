/*
 * spectral_oracle.c — CPU restatement of the pointwise spectral pieces of the hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see padeops_oracle.c header).
 *
 * Follows (paths relative to /root/reference/src):
 *   utilities/PoissonPeriodic.F90:226-261  GetWaveNums + ifftshift   (= utilities/fft_3d.F90:899-934)
 *   utilities/PoissonPeriodic.F90:89-111   poisson3D_multiply
 * The FFT passes themselves are FFTW 3.3.5 in the reference (dependencies/fftw-3.3.5.tar.gz); tests
 * compose them from numpy's pocketfft (oracle/oracle.py) or from oracle/_ref/lib/libfftw3.so, which
 * `make -C oracle ref` builds from that very tarball.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* k(i) = (-pi + (i-1)*two*pi/real(n - mod(n,2)))/dx ; k = ifftshift(k) */
void pdo_oracle_wavenums(int n, double dx, double *k)
{
    const double pi = 3.141592653589793238462643383279502884197; /* constants.F90 pi */
    const double two = 2.0;
    const int dummy = n - (n % 2);
    double *t = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 1; i <= n; ++i) t[i - 1] = (-pi + (double)(i - 1) * two * pi / (double)dummy) / dx;
    if (n % 2 == 0) {
        memcpy(k, t + n / 2, sizeof(double) * (size_t)(n / 2));
        memcpy(k + n / 2, t, sizeof(double) * (size_t)(n / 2));
    } else {
        memcpy(k, t + (n + 1) / 2 - 1, sizeof(double) * (size_t)((n + 1) / 2));
        memcpy(k + (n + 1) / 2, t, sizeof(double) * (size_t)((n - 1) / 2));
    }
    free(t);
}

/* RHS_hat(nxh,nyh,nzh) complex interleaved; kx,ky,kz local slices; zero mode cleared when owned */
void pdo_oracle_poisson_multiply(double *rhs_hat, int64_t nxh, int64_t nyh, int64_t nzh, const double *kx,
                                 const double *ky, const double *kz, int have_zero)
{
    for (int64_t k = 0; k < nzh; ++k) {
        const double kz_sq = kz[k] * kz[k];
        for (int64_t j = 0; j < nyh; ++j) {
            const double ky_sq = ky[j] * ky[j];
            double *row = rhs_hat + 2 * ((k * nyh + j) * nxh);
            for (int64_t i = 0; i < nxh; ++i) {
                const double m = -1.0 / (kx[i] * kx[i] + ky_sq + kz_sq + 1.e-20);
                row[2 * i] = row[2 * i] * m;
                row[2 * i + 1] = row[2 * i + 1] * m;
            }
        }
    }
    if (have_zero) { rhs_hat[0] = 0.0; rhs_hat[1] = 0.0; }
}
