// ig_padepoisson.inc.cuh — part of igrid.cu: textually included there, ONE translation unit (the sections share file-local helpers).
// PadePoissonMod::padepoisson: periodic and wall-bounded projection, pressure, divergence check, C ABI.
// Not a stand-alone header: do not include it anywhere else.

// ================================================================================================
// padepoisson (periodic in z)
// ================================================================================================
struct pdo_padepoisson_s {
    pdo_spectral_t sp = nullptr, spE = nullptr;
    pdo_pade6stagg_t derivZ = nullptr;
    pdo_decomp_t dC = nullptr, dE = nullptr;  // spectral decompositions of the cell / edge grids (borrowed from sp / spE)
    pdo_decomp_info sC, sE;
    double *k1sq = nullptr, *k2sq = nullptr, *k3sq = nullptr;  // z-pencil slices of GetWaveNums(nx,dx)^2, (ny,dy)^2; k3mod^2
    double mfact = 1.0;
    double2 *f2d = nullptr, *f2dy = nullptr, *w2 = nullptr, *uhatInZ = nullptr, *dwdz = nullptr;
    double* div_tmp = nullptr;  // real x-pencil, used when the caller passes no divergence array
    bool alias = false;         // one rank in the column communicator: y- and z-pencil layouts coincide, transposes are skipped
    const double2* phat_y = nullptr;  // where the last projection left the pressure (y-pencil layout)
    // PeriodicInZ = .false. (walls; PadePoisson.F90:180-230, 459-623): even / odd extensions to 2 nz planes and their tables
    bool periodic_in_z = true;
    double2 *fext = nullptr, *wext = nullptr, *k3modcm = nullptr, *k3modcp = nullptr;
    double* k3sq_ext = nullptr;
    ZColsPlan ext_plan;
    // computeStokesPressure (:232-296, 320-384): signed z-pencil slices of GetWaveNums, 1 / (lambda sinh(lambda Lz)) per column
    bool stokes = false;
    double Lz = 0.0;
    double *k1z = nullptr, *k2z = nullptr, *denfact = nullptr;
    double2* vhatInZ = nullptr;
    // GetStokesPressure (:641-714) keeps the two harmonic pressure pieces for getPressure; getPressureAndUpdateRHS adds whatever
    // the LAST getPressure left there (:1146-1156 — it runs ProjectStokesPressure, which does not refresh them); allocated zeroed
    double2 *phat_z1 = nullptr, *phat_z2 = nullptr;
    // PeriodicInZ: the z-operators of the projection are circulant, i.e. diagonal in kz.  symE2C / symC2E are the symbols of
    // derivZ%ddz_E2C / ddz_C2E (nz complex numbers each), measured at init as the transform of the operators' response to a unit
    // pulse — whatever the scheme (cd06 or Fourier collocation) — so that dealiasing + projection can run between ONE forward and
    // ONE backward z transform per field (poiss_dealias_project_kz)
    double2 *symE2C = nullptr, *symC2E = nullptr;
};

namespace {

// f2dy = i (k1 u + k2 v)   (PadePoisson.F90:392-401)
int poiss_div_xy(pdo_padepoisson_s* p, const double2* u, const double2* v, double2* out, cudaStream_t st) {
    const pdo_spectral_s* s = p->sp;
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const double *k1 = s->k1y, *k2 = s->k2;
    return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
        const double a = k1[(int)(i % n1)], b = k2[(int)((i / n1) % n2)];
        const double2 uu = u[i], vv = v[i];
        const double re = a * uu.x + b * vv.x, im = a * uu.y + b * vv.y;
        out[i] = make_double2(-im, re);
    });
}

// steps shared by PeriodicProjection / Periodic_getPressure*: leaves phat in f2d (z-pencil) and what in w2 (z-pencil)
int poiss_solve(pdo_padepoisson_s* p, const double2* uhat, const double2* vhat, const double2* what, cudaStream_t st) {
    if (p->alias && fft3d_own_z(p->sp->ft)) {
        // one GPU column: y- and z-pencils coincide and the hand-written z pass takes the pointwise work on its first load —
        // "+ i (k1 u + k2 v)" on the way into the forward transform, "-kradsq_inv mfact" on the way into the backward one
        if (int rc = pdo_pade6stagg_ddz_E2C(p->derivZ, (const double*)what, (double*)p->f2d, 1, 0, 0, st)) return rc;
        FftPro div;
        div.U = uhat; div.V = vhat; div.KU = p->sp->k1y; div.KV = p->sp->k2;
        if (int rc = fft3d_z_fused(p->sp->ft, p->f2d, p->f2d, -1, div, st)) return rc;
        FftPro inv;
        inv.poisson = 1; inv.A = p->k1sq; inv.B = p->k2sq; inv.C = p->k3sq; inv.scale = p->mfact;
        return fft3d_z_fused(p->sp->ft, p->f2d, p->f2d, +1, inv, st);
    }
    if (int rc = poiss_div_xy(p, uhat, vhat, p->f2dy, st)) return rc;
    const double2 *uz = p->f2dy, *wz = what;
    if (!p->alias) {
        if (int rc = decomp_transpose_device(p->dC, 2, (const double*)p->f2dy, (double*)p->uhatInZ, 2, st)) return rc;
        if (int rc = decomp_transpose_device(p->dE, 2, (const double*)what, (double*)p->w2, 2, st)) return rc;
        uz = p->uhatInZ; wz = p->w2;
    }
    if (int rc = pdo_pade6stagg_ddz_E2C(p->derivZ, (const double*)wz, (double*)p->f2d, 1, 0, 0, st)) return rc;
    const long long n = vol(p->sC.zsz);
    double2* f2d = p->f2d;
    if (int rc = launch_ew(n, st, [=] __device__(long long i) { double2 a = f2d[i]; const double2 b = uz[i]; a.x += b.x; a.y += b.y; f2d[i] = a; })) return rc;
    if (int rc = fft3d_z_inplace(p->sp->ft, f2d, -1, st)) return rc;
    const int n1 = p->sC.zsz[0], n2 = p->sC.zsz[1];
    const double *k1sq = p->k1sq, *k2sq = p->k2sq, *k3sq = p->k3sq;
    const double mfact = p->mfact;
    if (int rc = launch_ew(n, st, [=] __device__(long long i) {  // f2d = -kradsq_inv f2d, mfact folded in (:413-415, 103-108)
            const int ii = (int)(i % n1);
            const long long t = i / n1;
            const int jj = (int)(t % n2), kk = (int)(t / n2);
            const double kradsq = k1sq[ii] + k2sq[jj] + k3sq[kk];
            const double m = (kradsq <= 1.e-14) ? 0.0 : -(1.0 / kradsq) * mfact;
            double2 a = f2d[i];
            a.x *= m; a.y *= m;
            f2d[i] = a;
        })) return rc;
    return fft3d_z_inplace(p->sp->ft, f2d, +1, st);
}

// w2 -= ddz_C2E(f2d); what <- w2; f2dy <- f2d; u -= i k1 p, v -= i k2 p   (:417-431)
int poiss_correct(pdo_padepoisson_s* p, double2* uhat, double2* vhat, double2* what, cudaStream_t st) {
    if (int rc = pdo_pade6stagg_ddz_C2E(p->derivZ, (const double*)p->f2d, (double*)p->dwdz, 1, 0, 0, st)) return rc;
    double2* w2 = p->alias ? what : p->w2;
    const double2* dw = p->dwdz;
    if (int rc = launch_ew(vol(p->sE.zsz), st, [=] __device__(long long i) { double2 a = w2[i]; const double2 b = dw[i]; a.x -= b.x; a.y -= b.y; w2[i] = a; })) return rc;
    const double2* ph = p->f2d;
    if (!p->alias) {
        if (int rc = decomp_transpose_device(p->dE, 3, (const double*)p->w2, (double*)what, 2, st)) return rc;
        if (int rc = decomp_transpose_device(p->dC, 3, (const double*)p->f2d, (double*)p->f2dy, 2, st)) return rc;
        ph = p->f2dy;
    }
    p->phat_y = ph;
    const pdo_spectral_s* s = p->sp;
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const double *k1 = s->k1y, *k2 = s->k2;
    return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
        const double a = k1[(int)(i % n1)], b = k2[(int)((i / n1) % n2)];
        const double2 q = ph[i];
        double2 uu = uhat[i], vv = vhat[i];
        uu.x += a * q.y; uu.y -= a * q.x;  // u - i k1 p
        vv.x += b * q.y; vv.y -= b * q.x;
        uhat[i] = uu; vhat[i] = vv;
    });
}

// symbols of the periodic z-derivatives: response to a unit pulse in plane 0 of every column, column 0 transformed on the host
int poiss_measure_symbols(pdo_padepoisson_s* p, cudaStream_t st) {
    const int nz = p->sp->nz;
    const long long cols = (long long)p->sC.zsz[0] * p->sC.zsz[1];
    if (cols < 1) return 0;
    std::vector<double2> resp((size_t)nz), sym((size_t)nz);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    auto transform = [&](double2** dev) -> int {
        for (int k = 0; k < nz; ++k) {
            long double re = 0.0L, im = 0.0L;
            for (int z = 0; z < nz; ++z) {
                const long long e = ((long long)k * z) % nz;
                const long double c = cosl(two_pi * (long double)e / (long double)nz), sn = -sinl(two_pi * (long double)e / (long double)nz);
                re += (long double)resp[z].x * c - (long double)resp[z].y * sn;
                im += (long double)resp[z].x * sn + (long double)resp[z].y * c;
            }
            sym[(size_t)k] = make_double2((double)re, (double)im);
        }
        PDO_CUDA(cudaMalloc(dev, sizeof(double2) * (size_t)nz));
        PDO_CUDA(cudaMemcpy(*dev, sym.data(), sizeof(double2) * (size_t)nz, cudaMemcpyHostToDevice));
        return 0;
    };
    double2 *e = p->dwdz, *c = p->f2d;
    // ddz_E2C: pulse on edge plane 0 (= plane nz, the periodic image the operator may read)
    PDO_CUDA(cudaMemsetAsync(e, 0, sizeof(double2) * (size_t)cols * (size_t)(nz + 1), st));
    if (int rc = launch_ew(cols, st, [=] __device__(long long i) { e[i] = make_double2(1.0, 0.0); e[i + cols * nz] = make_double2(1.0, 0.0); })) return rc;
    if (int rc = pdo_pade6stagg_ddz_E2C(p->derivZ, (const double*)e, (double*)c, 1, 0, 0, st)) return rc;
    PDO_CUDA(cudaMemcpy2DAsync(resp.data(), sizeof(double2), c, sizeof(double2) * (size_t)cols, sizeof(double2), (size_t)nz, cudaMemcpyDeviceToHost, st));
    PDO_CUDA(cudaStreamSynchronize(st));
    if (int rc = transform(&p->symE2C)) return rc;
    // ddz_C2E: pulse on cell plane 0
    PDO_CUDA(cudaMemsetAsync(c, 0, sizeof(double2) * (size_t)cols * (size_t)nz, st));
    if (int rc = launch_ew(cols, st, [=] __device__(long long i) { c[i] = make_double2(1.0, 0.0); })) return rc;
    if (int rc = pdo_pade6stagg_ddz_C2E(p->derivZ, (const double*)c, (double*)e, 1, 0, 0, st)) return rc;
    PDO_CUDA(cudaMemcpy2DAsync(resp.data(), sizeof(double2), e, sizeof(double2) * (size_t)cols, sizeof(double2), (size_t)nz, cudaMemcpyDeviceToHost, st));
    PDO_CUDA(cudaStreamSynchronize(st));
    return transform(&p->symC2E);
}

// dealias (spectral.F90:343-363) + PeriodicProjection (PadePoisson.F90:386-432) of z-pencil (u, v, w) between one forward and one
// backward z transform per field.  In kz-space every step is pointwise: the mask, the divergence D_E2C w + i k1 u + i k2 v, the solve
// -1 / kradsq, the corrections w - D_C2E p, u - i k1 p, v - i k2 p.  Same mathematics as the reference's sequence (its compact z-solves
// are exactly circulant); the operations differ, the results agree to rounding (2e-15 measured on the CPU restatement of both sequences).  Replaces six z
// transforms + two compact solves + five pointwise passes by six transforms + one pass.
int poiss_dealias_project_kz(pdo_padepoisson_s* p, double2* zu, double2* zv, double2* zw, cudaStream_t st) {
    pdo_spectral_s* s = p->sp;
    const int nz = s->nz, n1 = p->sC.zsz[0], n2 = p->sC.zsz[1];
    const long long cols = (long long)n1 * n2, n = cols * nz;
    if (int rc = fft3d_z_inplace(s->ft, zu, -1, st)) return rc;
    if (int rc = fft3d_z_inplace(s->ft, zv, -1, st)) return rc;
    if (int rc = fft3d_z_inplace(s->ft, zw, -1, st)) return rc;
    const double *gx = s->gx, *gy = s->gyz, *gz = s->gz, *k1 = s->k1y, *k2 = s->k2z;
    const double *k1sq = p->k1sq, *k2sq = p->k2sq, *k3sq = p->k3sq;
    const double2 *dE = p->symE2C, *dC = p->symC2E;
    const double inv_nz = s->normfactz;
    if (int rc = launch_ew(n, st, [=] __device__(long long i) {
            const int ii = (int)(i % n1);
            const long long t = i / n1;
            const int jj = (int)(t % n2), kk = (int)(t / n2);
            const double m = gx[ii] * gy[jj] * gz[kk];
            double2 U = zu[i], V = zv[i], W = zw[i];
            U.x *= m; U.y *= m; V.x *= m; V.y *= m; W.x *= m; W.y *= m;
            const double a = k1[ii], b = k2[jj];
            const double2 de = dE[kk], dc = dC[kk];
            // f = D_E2C W + i (k1 U + k2 V)
            double2 f = make_double2(de.x * W.x - de.y * W.y, de.x * W.y + de.y * W.x);
            f.x += -(a * U.y + b * V.y);
            f.y += a * U.x + b * V.x;
            const double kradsq = k1sq[ii] + k2sq[jj] + k3sq[kk];
            const double q = (kradsq <= 1.e-14) ? 0.0 : -(1.0 / kradsq);
            const double2 ph = make_double2(f.x * q, f.y * q);
            W.x -= dc.x * ph.x - dc.y * ph.y; W.y -= dc.x * ph.y + dc.y * ph.x;
            U.x += a * ph.y; U.y -= a * ph.x;      // u - i k1 p
            V.x += b * ph.y; V.y -= b * ph.x;
            zu[i] = make_double2(U.x * inv_nz, U.y * inv_nz);
            zv[i] = make_double2(V.x * inv_nz, V.y * inv_nz);
            zw[i] = make_double2(W.x * inv_nz, W.y * inv_nz);
        })) return rc;
    if (int rc = fft3d_z_inplace(s->ft, zu, +1, st)) return rc;
    if (int rc = fft3d_z_inplace(s->ft, zv, +1, st)) return rc;
    if (int rc = fft3d_z_inplace(s->ft, zw, +1, st)) return rc;
    PDO_CUDA(cudaMemcpyAsync(zw + cols * nz, zw, sizeof(double2) * (size_t)cols, cudaMemcpyDeviceToDevice, st));   // plane nz+1 := plane 1
    p->phat_y = nullptr;
    return 0;
}

// PeriodicProjection (PadePoisson.F90:386-432) on z-PENCIL copies of (uhat, vhat, what), in place: the pointwise steps do not care
// which pencil they run in, so a caller that already holds the three fields in the z-pencil (igrid, after dealiasing there) skips
// the four transposes of the y-pencil entry point.  Same operations per element, same results.  phat stays in f2d (z-pencil).
int poiss_projection_z(pdo_padepoisson_s* p, double2* zu, double2* zv, double2* zw, cudaStream_t st) {
    const pdo_spectral_s* s = p->sp;
    const int n1 = p->sC.zsz[0], n2 = p->sC.zsz[1];
    const long long n = vol(p->sC.zsz);
    const double *k1 = s->k1y, *k2 = s->k2z;
    double2* f2d = p->f2d;
    if (int rc = pdo_pade6stagg_ddz_E2C(p->derivZ, (const double*)zw, (double*)f2d, 1, 0, 0, st)) return rc;
    const double *k1sq = p->k1sq, *k2sq = p->k2sq, *k3sq = p->k3sq;
    const double mfact = p->mfact;
    if (fft3d_own_z(s->ft)) {
        FftPro div;
        div.U = zu; div.V = zv; div.KU = k1; div.KV = k2;
        if (int rc = fft3d_z_fused(s->ft, f2d, f2d, -1, div, st)) return rc;
        FftPro inv;
        inv.poisson = 1; inv.A = k1sq; inv.B = k2sq; inv.C = k3sq; inv.scale = mfact;
        if (int rc = fft3d_z_fused(s->ft, f2d, f2d, +1, inv, st)) return rc;
    } else {
        if (int rc = launch_ew(n, st, [=] __device__(long long i) {   // f2d += i (k1 u + k2 v)
                const double a = k1[(int)(i % n1)], b = k2[(int)((i / n1) % n2)];
                const double2 uu = zu[i], vv = zv[i];
                const double re = a * uu.x + b * vv.x, im = a * uu.y + b * vv.y;
                double2 q = f2d[i];
                q.x += -im; q.y += re;
                f2d[i] = q;
            })) return rc;
        if (int rc = fft3d_z_inplace(s->ft, f2d, -1, st)) return rc;
        if (int rc = launch_ew(n, st, [=] __device__(long long i) {
                const int ii = (int)(i % n1);
                const long long t = i / n1;
                const int jj = (int)(t % n2), kk = (int)(t / n2);
                const double kradsq = k1sq[ii] + k2sq[jj] + k3sq[kk];
                const double m = (kradsq <= 1.e-14) ? 0.0 : -(1.0 / kradsq) * mfact;
                double2 a = f2d[i];
                a.x *= m; a.y *= m;
                f2d[i] = a;
            })) return rc;
        if (int rc = fft3d_z_inplace(s->ft, f2d, +1, st)) return rc;
    }
    if (int rc = pdo_pade6stagg_ddz_C2E(p->derivZ, (const double*)f2d, (double*)p->dwdz, 1, 0, 0, st)) return rc;
    const double2* dw = p->dwdz;
    if (int rc = launch_ew(vol(p->sE.zsz), st, [=] __device__(long long i) { double2 a = zw[i]; const double2 b = dw[i]; a.x -= b.x; a.y -= b.y; zw[i] = a; })) return rc;
    p->phat_y = nullptr;   // the pressure was not taken to the y-pencil
    return launch_ew(n, st, [=] __device__(long long i) {
        const double a = k1[(int)(i % n1)], b = k2[(int)((i / n1) % n2)];
        const double2 q = f2d[i];
        double2 uu = zu[i], vv = zv[i];
        uu.x += a * q.y; uu.y -= a * q.x;  // u - i k1 p
        vv.x += b * q.y; vv.y -= b * q.x;
        zu[i] = uu; zv[i] = vv;
    });
}

int poiss_divergence(pdo_padepoisson_s* p, const double2* uhat, const double2* vhat, const double2* what, double* div, cudaStream_t st) {
    const double2* wz = what;
    if (!p->alias) {
        if (int rc = decomp_transpose_device(p->dE, 2, (const double*)what, (double*)p->w2, 2, st)) return rc;
        wz = p->w2;
    }
    if (int rc = pdo_pade6stagg_ddz_E2C(p->derivZ, (const double*)wz, (double*)p->f2d, 1, -1, -1, st)) return rc;
    double2* f = p->f2d;
    if (!p->alias) {
        if (int rc = decomp_transpose_device(p->dC, 3, (const double*)p->f2d, (double*)p->f2dy, 2, st)) return rc;
        f = p->f2dy;
    }
    const pdo_spectral_s* s = p->sp;
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const double *k1 = s->k1y, *k2 = s->k2;
    if (int rc = launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {  // + i k1 u + i k2 v  (:1191-1200)
            const double a = k1[(int)(i % n1)], b = k2[(int)((i / n1) % n2)];
            const double2 uu = uhat[i], vv = vhat[i];
            double2 q = f[i];
            q.x += -a * uu.y - b * vv.y;
            q.y += a * uu.x + b * vv.x;
            f[i] = q;
        })) return rc;
    return fft3d_backward_yx(s->ft, f, div, false, st);
}

// p_maxval(maxval(a)) (use_abs = 0, as DivergenceCheck does) or of |a|
int global_max(pdo_spectral_s* s, const double* a, long long n, int use_abs, double* out, cudaStream_t st) {
    const int blocks = 1024;
    max_kernel<<<blocks, 256, 0, st>>>(a, n, use_abs, s->partial);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    double hpart[1024];
    PDO_CUDA(cudaMemcpyAsync(hpart, s->partial, sizeof(double) * blocks, cudaMemcpyDeviceToHost, st));
    PDO_CUDA(cudaStreamSynchronize(st));
    double m = -1.0e300;
    for (int i = 0; i < blocks; ++i) m = hpart[i] > m ? hpart[i] : m;
    return pdo_p_maxval(m, out);
}

// ProjectStokesPressure (PadePoisson.F90:320-384), in place on the z-pencil arrays: one thread per (kx, ky) column.  The harmonic
// pressure chat cosh(lambda (Lz - z)) (bottom wall) and chat cosh(lambda z) (top wall, computed from the ALREADY corrected top
// plane) cancels w on the walls; cosh / sinh are evaluated on the fly with the reference's clipping (arguments >= 32 -> 4e13 /
// 1e13 / -4e13 / 4e13) instead of being read from four 3-D tables.
__global__ void __launch_bounds__(128) stokes_kernel(double2* __restrict__ u, double2* __restrict__ v, double2* __restrict__ w2, long long cols,
                                                     int n1, int nz, double Lz, const double* __restrict__ k1z, const double* __restrict__ k2z,
                                                     const double* __restrict__ denfact, double2* __restrict__ pz1, double2* __restrict__ pz2) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const double k1 = k1z[(int)(c % n1)], k2 = k2z[(int)(c / n1)];
    const double lam = sqrt(k1 * k1 + k2 * k2), den = denfact[c], dzl = Lz / (double)nz;
    // bottom BC
    const double2 w0 = w2[c];
    double2 ch = make_double2(-w0.x * den, -w0.y * den);
    for (int k = 0; k < nz; ++k) {
        const double zc = 0.5 * ((double)k * dzl + (double)(k + 1) * dzl);
        const double t = lam * (Lz - zc);
        const double cb = t < 32.0 ? cosh(t) : 4.0e13;
        const double2 ph = make_double2(-ch.y * cb, ch.x * cb);      // imi * chat * cosh_bot
        if (pz1) pz1[c + cols * k] = ph;
        double2 a = u[c + cols * k]; a.x -= k1 * ph.x; a.y -= k1 * ph.y; u[c + cols * k] = a;
        double2 b = v[c + cols * k]; b.x -= k2 * ph.x; b.y -= k2 * ph.y; v[c + cols * k] = b;
    }
    w2[c] = make_double2(0.0, 0.0);
    for (int k = 1; k <= nz; ++k) {
        const double t = lam * (Lz - (double)k * dzl);
        const double sb = t < 32.0 ? -lam * sinh(t) : -4.0e13;
        double2 a = w2[c + cols * k]; a.x -= ch.x * sb; a.y -= ch.y * sb; w2[c + cols * k] = a;
    }
    // top BC
    const double2 wn = w2[c + cols * nz];
    ch = make_double2(wn.x * den, wn.y * den);
    for (int k = 0; k < nz; ++k) {
        const double zc = 0.5 * ((double)k * dzl + (double)(k + 1) * dzl);
        double t = lam * zc;
        const double ct = t < 32.0 ? cosh(t) : 1.0e13;
        const double2 ph = make_double2(-ch.y * ct, ch.x * ct);
        if (pz2) pz2[c + cols * k] = ph;
        double2 a = u[c + cols * k]; a.x -= k1 * ph.x; a.y -= k1 * ph.y; u[c + cols * k] = a;
        double2 b = v[c + cols * k]; b.x -= k2 * ph.x; b.y -= k2 * ph.y; v[c + cols * k] = b;
        t = lam * ((double)k * dzl);
        const double stp = t < 32.0 ? lam * sinh(t) : 4.0e13;
        double2 e = w2[c + cols * k]; e.x -= ch.x * stp; e.y -= ch.y * stp; w2[c + cols * k] = e;
    }
    w2[c + cols * nz] = make_double2(0.0, 0.0);
}

// PressureProjection with walls (PadePoisson.F90:444-623): the horizontal divergence is extended
// evenly and w oddly about both walls to 2 nz planes, one c2c-z pair solves and projects (the half-cell shifts ride on
// k3modcm / k3modcp), the upper halves come back and w is zero on both walls.
int poiss_wall_projection(pdo_padepoisson_s* p, double2* uhat, double2* vhat, double2* what, cudaStream_t st, bool store_stokes = false) {
    const int nz = p->sp->nz, n1 = p->sC.zsz[0], n2 = p->sC.zsz[1];
    const long long cols = (long long)n1 * n2, next = cols * 2 * nz;
    const double2 *uz = nullptr, *wz = what;
    double2 *uZ = uhat, *vZ = vhat;          // computeStokesPressure: u and v in the z-pencil, corrected in place
    if (p->stokes) {                          // Step 0 (:444-458)
        double2* w2 = what;
        if (!p->alias) {
            if (int rc = decomp_transpose_device(p->dC, 2, (const double*)uhat, (double*)p->uhatInZ, 2, st)) return rc;
            if (int rc = decomp_transpose_device(p->dC, 2, (const double*)vhat, (double*)p->vhatInZ, 2, st)) return rc;
            if (int rc = decomp_transpose_device(p->dE, 2, (const double*)what, (double*)p->w2, 2, st)) return rc;
            uZ = p->uhatInZ; vZ = p->vhatInZ; w2 = p->w2;
        }
        stokes_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, st>>>(uZ, vZ, w2, cols, n1, nz, p->Lz, p->k1z, p->k2z, p->denfact,
                                                                      store_stokes ? p->phat_z1 : nullptr, store_stokes ? p->phat_z2 : nullptr);
        PDO_CUDA(cudaGetLastError());
        g_launches += 1;
        const double *k1z = p->k1z, *k2z = p->k2z;
        double2* f2d = p->f2d;
        const double2 *cu = uZ, *cv = vZ;
        if (int rc = launch_ew(cols * nz, st, [=] __device__(long long i) {   // f2d = i (k1inZ u + k2inZ v)
                const double a = k1z[(int)(i % n1)], b = k2z[(int)((i / n1) % n2)];
                const double2 uu = cu[i], vv = cv[i];
                f2d[i] = make_double2(-(a * uu.y + b * vv.y), a * uu.x + b * vv.x);
            })) return rc;
        uz = p->f2d; wz = w2;
    } else {
        if (int rc = poiss_div_xy(p, uhat, vhat, p->f2dy, st)) return rc;                       // Step 1
        uz = p->f2dy;
        if (!p->alias) {                                                                          // Step 2
            if (int rc = decomp_transpose_device(p->dC, 2, (const double*)p->f2dy, (double*)p->uhatInZ, 2, st)) return rc;
            if (int rc = decomp_transpose_device(p->dE, 2, (const double*)what, (double*)p->w2, 2, st)) return rc;
            uz = p->uhatInZ; wz = p->w2;
        }
    }
    double2 *fe = p->fext, *we = p->wext;
    if (int rc = launch_ew(next, st, [=] __device__(long long i) {                            // Step 3
            const long long c = i % cols;
            const int kk = (int)(i / cols);
            fe[i] = kk < nz ? uz[c + cols * (nz - 1 - kk)] : uz[c + cols * (kk - nz)];
            if (kk < nz - 1) { const double2 a = wz[c + cols * (nz - 1 - kk)]; we[i] = make_double2(-a.x, -a.y); }
            else we[i] = wz[c + cols * (kk - (nz - 1))];
        })) return rc;
    if (int rc = zcols_exec(&p->ext_plan, 2 * nz, cols, fe, -1, st)) return rc;               // Step 4
    if (int rc = zcols_exec(&p->ext_plan, 2 * nz, cols, we, -1, st)) return rc;
    const double *k1sq = p->k1sq, *k2sq = p->k2sq, *k3sq = p->k3sq_ext;
    const double2 *cm = p->k3modcm, *cp = p->k3modcp;
    const double mfact = p->mfact;
    if (int rc = launch_ew(next, st, [=] __device__(long long i) {                            // Steps 5-6 (+ mfact of Step 7)
            const int ii = (int)(i % n1);
            const long long t = i / n1;
            const int jj = (int)(t % n2), kk = (int)(t / n2);
            const double kradsq = k1sq[ii] + k2sq[jj] + k3sq[kk];
            const double kinv = (kradsq <= 1.e-14) ? 0.0 : 1.0 / kradsq;
            double2 f = fe[i], w = we[i];
            const double2 a = cm[kk], b = cp[kk];
            // f = f + i cm w;  f = -f kinv
            f.x += -(a.x * w.y + a.y * w.x);
            f.y += a.x * w.x - a.y * w.y;
            f.x = -f.x * kinv; f.y = -f.y * kinv;
            // w = w - i cp f
            w.x -= -(b.x * f.y + b.y * f.x);
            w.y -= b.x * f.x - b.y * f.y;
            fe[i] = make_double2(f.x * mfact, f.y * mfact);
            we[i] = make_double2(w.x * mfact, w.y * mfact);
        })) return rc;
    if (int rc = zcols_exec(&p->ext_plan, 2 * nz, cols, fe, +1, st)) return rc;               // Step 7
    if (int rc = zcols_exec(&p->ext_plan, 2 * nz, cols, we, +1, st)) return rc;
    double2* f2d = p->f2d;
    double2* w2 = p->alias ? what : p->w2;
    if (int rc = launch_ew(cols * (nz + 1), st, [=] __device__(long long i) {
            const int kk = (int)(i / cols);
            if (kk < nz) f2d[i] = fe[i + cols * nz];
            w2[i] = (kk == 0 || kk == nz) ? make_double2(0.0, 0.0) : we[i + cols * (nz - 1)];
        })) return rc;
    if (p->stokes) {                                                                          // :597-609
        const double *k1z = p->k1z, *k2z = p->k2z;
        const double2* pr = p->f2d;
        if (int rc = launch_ew(cols * nz, st, [=] __device__(long long i) {   // u -= i k1inZ p, v -= i k2inZ p, in the z-pencil
                const double a = k1z[(int)(i % n1)], b = k2z[(int)((i / n1) % n2)];
                const double2 q = pr[i];
                double2 uu = uZ[i], vv = vZ[i];
                uu.x += a * q.y; uu.y -= a * q.x;
                vv.x += b * q.y; vv.y -= b * q.x;
                uZ[i] = uu; vZ[i] = vv;
            })) return rc;
        p->phat_y = nullptr;
        if (!p->alias) {
            if (int rc = decomp_transpose_device(p->dE, 3, (const double*)p->w2, (double*)what, 2, st)) return rc;
            if (int rc = decomp_transpose_device(p->dC, 3, (const double*)uZ, (double*)uhat, 2, st)) return rc;
            if (int rc = decomp_transpose_device(p->dC, 3, (const double*)vZ, (double*)vhat, 2, st)) return rc;
        }
        return 0;
    }
    const double2* ph = p->f2d;
    if (!p->alias) {                                                                          // Step 8
        if (int rc = decomp_transpose_device(p->dE, 3, (const double*)p->w2, (double*)what, 2, st)) return rc;
        if (int rc = decomp_transpose_device(p->dC, 3, (const double*)p->f2d, (double*)p->f2dy, 2, st)) return rc;
        ph = p->f2dy;
    }
    p->phat_y = ph;
    const pdo_spectral_s* s = p->sp;
    const int m1 = s->si.ysz[0], m2 = s->si.ysz[1];
    const double *k1 = s->k1y, *k2 = s->k2;
    return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {                        // Step 9
        const double a = k1[(int)(i % m1)], b = k2[(int)((i / m1) % m2)];
        const double2 q = ph[i];
        double2 uu = uhat[i], vv = vhat[i];
        uu.x += a * q.y; uu.y -= a * q.x;
        vv.x += b * q.y; vv.y -= b * q.x;
        uhat[i] = uu; vhat[i] = vv;
    });
}

// the pressure of the wall-bounded solver after poiss_wall_projection (getPressure :880-896, getPressureAndUpdateRHS :1144-1160):
// phat = f2d (+ phat_z1 + phat_z2 with computeStokesPressure), taken to the y-pencil and to physical space
int poiss_wall_pressure_out(pdo_padepoisson_s* p, double* pressure, cudaStream_t st) {
    const double2* ph = p->phat_y;
    if (p->stokes) {
        double2* f2d = p->f2d;
        const double2 *z1 = p->phat_z1, *z2 = p->phat_z2;
        if (z1) {
            if (int rc = launch_ew(vol(p->sC.zsz), st, [=] __device__(long long i) {
                    double2 a = f2d[i];
                    const double2 b = z1[i], c = z2[i];
                    a.x += b.x; a.y += b.y;
                    a.x += c.x; a.y += c.y;
                    f2d[i] = a;
                })) return rc;
        }
        ph = p->f2d;
        if (!p->alias) {
            if (int rc = decomp_transpose_device(p->dC, 3, (const double*)p->f2d, (double*)p->f2dy, 2, st)) return rc;
            ph = p->f2dy;
        }
    }
    const size_t bytes = sizeof(double) * (size_t)vol(p->sp->pi.xsz);
    return with_device_views(pressure, 0, pressure, bytes, st, [&](const void*, void* d_o) {
        return fft3d_backward_yx(p->sp->ft, ph, (double*)d_o, false, st);
    });
}
int poiss_alloc_stokes_pieces(pdo_padepoisson_s* p) {
    if (p->phat_z1) return 0;
    const size_t bytes = sizeof(double2) * (size_t)vol(p->sC.zsz);
    PDO_CUDA(cudaMalloc(&p->phat_z1, bytes));
    PDO_CUDA(cudaMalloc(&p->phat_z2, bytes));
    PDO_CUDA(cudaMemset(p->phat_z1, 0, bytes));
    PDO_CUDA(cudaMemset(p->phat_z2, 0, bytes));
    return 0;
}

int poiss_projection(pdo_padepoisson_s* p, double2* u, double2* v, double2* w, cudaStream_t st) {
    if (!p->periodic_in_z) return poiss_wall_projection(p, u, v, w, st);
    if (int rc = poiss_solve(p, u, v, w, st)) return rc;
    return poiss_correct(p, u, v, w, st);
}

int poiss_divergence_check(pdo_padepoisson_s* p, double2* u, double2* v, double2* w, double* div, bool fix, double* max_div, cudaStream_t st) {
    if (!div) div = p->div_tmp;
    const long long n = vol(p->sp->pi.xsz);
    if (int rc = poiss_divergence(p, u, v, w, div, st)) return rc;
    double md = 0.0;
    if (fix || max_div) { if (int rc = global_max(p->sp, div, n, 0, &md, st)) return rc; }
    if (fix && md > 1.e-13) {  // PadePoisson.F90:1209-1241
        if (int rc = poiss_projection(p, u, v, w, st)) return rc;
        if (int rc = poiss_divergence(p, u, v, w, div, st)) return rc;
        if (int rc = global_max(p->sp, div, n, 0, &md, st)) return rc;
        if (md > 1.e-10) { if (int rc = poiss_projection(p, u, v, w, st)) return rc; }
    }
    if (max_div) *max_div = md;
    return 0;
}

}  // namespace

extern "C" {

int pdo_padepoisson_init(pdo_padepoisson_t* h, double dx, double dy, double dz, pdo_spectral_t sp, pdo_spectral_t spE,
                         pdo_pade6stagg_t derivZ) {
    return pdo_padepoisson_init2(h, dx, dy, dz, sp, spE, derivZ, 1);
}
int pdo_padepoisson_init2(pdo_padepoisson_t* h, double dx, double dy, double dz, pdo_spectral_t sp, pdo_spectral_t spE,
                          pdo_pade6stagg_t derivZ, int periodic_in_z) {
    return pdo_padepoisson_init3(h, dx, dy, dz, sp, spE, derivZ, periodic_in_z, 0, 0.0);
}
int pdo_padepoisson_init3(pdo_padepoisson_t* h, double dx, double dy, double dz, pdo_spectral_t sp, pdo_spectral_t spE,
                          pdo_pade6stagg_t derivZ, int periodic_in_z, int compute_stokes_pressure, double Lz) {
    if (!h || !sp || !spE || !derivZ) return fail(PDO_E_BADARG, "null argument");
    *h = nullptr;
    if (spE->nz != sp->nz + 1 || spE->nx != sp->nx || spE->ny != sp->ny) return fail(PDO_E_BADARG, "spE must be the (nx, ny, nz+1) edge type of sp");
    if (periodic_in_z && !derivZ->periodic)
        return fail(PDO_E_BADARG, "padepoisson: PeriodicInZ = .true. needs a derivZ initialised with isPeriodic = .true.");
    if (!periodic_in_z && derivZ->periodic)
        return fail(PDO_E_BADARG, "padepoisson: PeriodicInZ = .false. needs a derivZ initialised with isPeriodic = .false.");
    // PadePoisson.F90:215-218 — the two decompositions must split x and y identically in the z-pencil
    if (sp->si.zst[0] != spE->si.zst[0] || sp->si.zst[1] != spE->si.zst[1])
        return fail(423, "Failed at initializing Padepoisson. sp_gp and sp_gpE have different x and y starts in z-decomp");
    pdo_padepoisson_s* p = new (std::nothrow) pdo_padepoisson_s();
    if (!p) return fail(PDO_E_BADARG, "out of memory");
    p->sp = sp; p->spE = spE; p->derivZ = derivZ;
    p->dC = fft3d_spec_decomp(sp->ft); p->dE = fft3d_spec_decomp(spE->ft);
    p->sC = sp->si; p->sE = spE->si;
    const int nz = sp->nz;
    // InitPeriodicPoissonSolver (:76-128): k1, k2 straight from GetWaveNums (no oddball flip), k3 through the z scheme's symbol
    std::vector<double> k1 = wavenums(sp->nx, dx), k2 = wavenums(sp->ny, dy), k3 = wavenums(nz, dz), k3m(nz);
    pdo_pade6stagg_get_modified_wavenumbers(derivZ, k3.data(), k3m.data(), nz);
    for (auto& v : k1) v = v * v;
    for (auto& v : k2) v = v * v;
    for (auto& v : k3m) v = v * v;
    int rc = upload(&p->k1sq, k1, p->sC.zst[0] - 1, p->sC.zsz[0]);
    if (!rc) rc = upload(&p->k2sq, k2, p->sC.zst[1] - 1, p->sC.zsz[1]);
    if (!rc) rc = upload(&p->k3sq, k3m, 0, nz);
    p->mfact = 1.0 / (double)nz;
    p->alias = (sp->p_col == 1);
    p->periodic_in_z = periodic_in_z != 0;
    cudaError_t e = cudaSuccess;
    if (!rc && !p->periodic_in_z) {
        // :183-210: k3 = GetWaveNums(2 nz, dz) through the z scheme's symbol; tfm / tfp = exp(-+ i dz/2 k3); mfact = 1 / (2 nz)
        const int nze = 2 * nz;
        std::vector<double> k3e = wavenums(nze, dz), k3me(nze), k3sq(nze);
        pdo_pade6stagg_get_modified_wavenumbers(derivZ, k3e.data(), k3me.data(), nze);
        std::vector<double2> cm(nze), cp(nze);
        for (int k = 0; k < nze; ++k) {
            const double ph = (dz / 2.0) * k3e[k];
            cm[k] = make_double2(k3me[k] * std::cos(ph), -k3me[k] * std::sin(ph));   // k3mod exp(-i dz/2 k3)
            cp[k] = make_double2(k3me[k] * std::cos(ph), k3me[k] * std::sin(ph));    // k3mod exp(+i dz/2 k3)
            k3sq[k] = k3me[k] * k3me[k];
        }
        p->mfact = 1.0 / (double)nze;
        rc = upload(&p->k3sq_ext, k3sq, 0, nze);
        const size_t ext = sizeof(double2) * (size_t)p->sC.zsz[0] * p->sC.zsz[1] * (size_t)nze;
        if (!rc) {
            e = cudaMalloc(&p->k3modcm, sizeof(double2) * nze);
            if (e == cudaSuccess) e = cudaMalloc(&p->k3modcp, sizeof(double2) * nze);
            if (e == cudaSuccess) e = cudaMemcpy(p->k3modcm, cm.data(), sizeof(double2) * nze, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(p->k3modcp, cp.data(), sizeof(double2) * nze, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMalloc(&p->fext, ext);
            if (e == cudaSuccess) e = cudaMalloc(&p->wext, ext);
            if (e != cudaSuccess) rc = fail(PDO_E_CUDA, "padepoisson wall buffers: %s", cudaGetErrorString(e));
        }
        if (!rc && compute_stokes_pressure) {   // :232-296
            p->stokes = true;
            p->Lz = Lz > 0.0 ? Lz : (double)nz * dz;
            std::vector<double> w1 = wavenums(sp->nx, dx), w2v = wavenums(sp->ny, dy);   // GetWaveNums, no oddball flip (:248-249)
            const int c1 = p->sC.zsz[0], c2 = p->sC.zsz[1], o1 = p->sC.zst[0] - 1, o2 = p->sC.zst[1] - 1;
            std::vector<double> den((size_t)c1 * c2);
            for (int j = 0; j < c2; ++j)
                for (int i = 0; i < c1; ++i) {
                    const double lam = std::sqrt(w1[o1 + i] * w1[o1 + i] + w2v[o2 + j] * w2v[o2 + j]);
                    double d = (lam * p->Lz < 500.0) ? 1.0 / (lam * std::sinh(lam * p->Lz) + 1.0e-13) : 0.0;
                    if (d < 1.0e-16) d = 0.0;
                    den[(size_t)j * c1 + i] = d;
                }
            if (o1 == 0 && o2 == 0) den[0] = 0.0;   // "if (nrank == 0) this%denFact(1,1) = 0": the owner of the mean mode
            rc = upload(&p->k1z, w1, o1, c1);
            if (!rc) rc = upload(&p->k2z, w2v, o2, c2);
            if (!rc) rc = upload(&p->denfact, den, 0, den.size());
            if (!rc) rc = comm_shared_malloc((void**)&p->vhatInZ, sizeof(double2) * (size_t)vol(p->sC.zsz));
        }
    }
    if (!rc) {
        e = cudaMalloc(&p->f2d, sizeof(double2) * (size_t)vol(p->sC.zsz));
        if (e == cudaSuccess) e = cudaMalloc(&p->dwdz, sizeof(double2) * (size_t)vol(p->sE.zsz));
        if (e == cudaSuccess) e = cudaMalloc(&p->div_tmp, sizeof(double) * (size_t)vol(sp->pi.xsz));
        if (e != cudaSuccess) rc = fail(PDO_E_CUDA, "padepoisson buffers: %s", cudaGetErrorString(e));
        // transpose destinations: peer-writable (collective, same order on every rank)
        if (!rc) rc = comm_shared_malloc((void**)&p->uhatInZ, sizeof(double2) * (size_t)vol(p->sC.zsz));
        if (!rc) rc = comm_shared_malloc((void**)&p->w2, sizeof(double2) * (size_t)vol(p->sE.zsz));
        if (!rc) rc = comm_shared_malloc((void**)&p->f2dy, sizeof(double2) * (size_t)vol(p->sC.ysz));
    }
    if (!rc && p->periodic_in_z && sp->periodicInZ) rc = poiss_measure_symbols(p, nullptr);
    if (rc) { pdo_padepoisson_destroy(p); return rc; }
    *h = p;
    return 0;
}

int pdo_padepoisson_destroy(pdo_padepoisson_t p) {
    if (!p) return 0;
    void* ptrs[] = {p->k1sq, p->k2sq, p->k3sq, p->f2d, p->dwdz, p->div_tmp, p->symE2C, p->symC2E, p->phat_z1, p->phat_z2};
    for (void* q : ptrs) if (q) cudaFree(q);
    void* wall[] = {p->fext, p->wext, p->k3modcm, p->k3modcp, p->k3sq_ext, p->k1z, p->k2z, p->denfact};
    for (void* q : wall) if (q) cudaFree(q);
    comm_shared_free(p->vhatInZ);       // same order as the allocations on every rank
    comm_shared_free(p->uhatInZ);
    comm_shared_free(p->w2);
    comm_shared_free(p->f2dy);
    zcols_destroy(&p->ext_plan);
    delete p;
    return 0;
}

}  // extern "C"

namespace {
// Runs body(u, v, w) on device views of the three spectral arrays; host arrays are staged in and (when writable) out.
template <class Body>
int with_uvw(pdo_padepoisson_s* p, const double* u, const double* v, const double* w, bool writeback, cudaStream_t st, Body body) {
    const size_t bC = sizeof(double2) * (size_t)vol(p->sC.ysz), bE = sizeof(double2) * (size_t)vol(p->sE.ysz);
    const double* in[3] = {u, v, w};
    const size_t bytes[3] = {bC, bC, bE};
    double2* dev[3] = {nullptr, nullptr, nullptr};
    bool staged[3] = {false, false, false};
    int rc = 0;
    for (int i = 0; i < 3 && !rc; ++i) {       // no early return: the staged copies made so far are released on every exit path
        if (is_device_ptr(in[i])) { dev[i] = (double2*)in[i]; continue; }
        cudaError_t e = cudaMalloc(&dev[i], bytes[i]);
        if (e == cudaSuccess) {
            staged[i] = true;
            e = cudaMemcpyAsync(dev[i], in[i], bytes[i], cudaMemcpyHostToDevice, st);
        }
        if (e != cudaSuccess) rc = fail(PDO_E_CUDA, "padepoisson staging: %s", cudaGetErrorString(e));
    }
    if (!rc) rc = body(dev[0], dev[1], dev[2]);
    for (int i = 0; i < 3; ++i) {
        if (!staged[i]) continue;
        if (!rc && writeback) {
            if (cudaMemcpyAsync((void*)in[i], dev[i], bytes[i], cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = fail(PDO_E_CUDA, "D2H failed");
        }
        cudaStreamSynchronize(st);
        cudaFree(dev[i]);
    }
    return rc;
}
}  // namespace

extern "C" {

int pdo_padepoisson_pressure_projection(pdo_padepoisson_t p, double* uhat, double* vhat, double* what, void* stream) {
    if (!p || !uhat || !vhat || !what) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    return with_uvw(p, uhat, vhat, what, true, st, [&](double2* u, double2* v, double2* w) { return poiss_projection(p, u, v, w, st); });
}

int pdo_padepoisson_get_pressure(pdo_padepoisson_t p, const double* uhat, const double* vhat, const double* what, double* pressure,
                                 void* stream) {
    if (!p || !uhat || !vhat || !what || !pressure) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (!p->periodic_in_z) {
        // walls (:762-896): the projection's Steps 0-7 on COPIES of the intent(in) right-hand sides, GetStokesPressure's two
        // pressure pieces kept, phat = f2d + phat_z1 + phat_z2
        if (p->stokes) { if (int rc = poiss_alloc_stokes_pieces(p)) return rc; }
        return with_uvw(p, uhat, vhat, what, false, st, [&](double2* u, double2* v, double2* w) -> int {
            const size_t bC = sizeof(double2) * (size_t)vol(p->sp->si.ysz), bE = sizeof(double2) * (size_t)vol(p->spE->si.ysz);
            double2* cp[3] = {nullptr, nullptr, nullptr};
            const double2* src[3] = {u, v, w};
            const size_t nb[3] = {bC, bC, bE};
            int rc = 0;
            for (int i = 0; i < 3 && !rc; ++i) {
                if (cudaMalloc(&cp[i], nb[i]) != cudaSuccess) { cudaGetLastError(); rc = fail(PDO_E_CUDA, "getPressure: out of device memory"); break; }
                if (cudaMemcpyAsync(cp[i], src[i], nb[i], cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = fail(PDO_E_CUDA, "getPressure: copy failed");
            }
            if (!rc) rc = poiss_wall_projection(p, cp[0], cp[1], cp[2], st, true);
            if (!rc) rc = poiss_wall_pressure_out(p, pressure, st);
            cudaStreamSynchronize(st);
            for (int i = 0; i < 3; ++i) if (cp[i]) cudaFree(cp[i]);
            return rc;
        });
    }
    return with_uvw(p, uhat, vhat, what, false, st, [&](double2* u, double2* v, double2* w) -> int {
        if (int rc = poiss_solve(p, u, v, w, st)) return rc;
        const double2* ph = p->f2d;
        if (!p->alias) {
            if (int rc = decomp_transpose_device(p->dC, 3, (const double*)p->f2d, (double*)p->f2dy, 2, st)) return rc;
            ph = p->f2dy;
        }
        const size_t bytes = sizeof(double) * (size_t)vol(p->sp->pi.xsz);
        return with_device_views(pressure, 0, pressure, bytes, st, [&](const void*, void* d_o) {
            return fft3d_backward_yx(p->sp->ft, ph, (double*)d_o, false, st);
        });
    });
}

int pdo_padepoisson_get_pressure_and_update_rhs(pdo_padepoisson_t p, double* uhat, double* vhat, double* what, double* pressure,
                                                void* stream) {
    if (!p || !uhat || !vhat || !what || !pressure) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (!p->periodic_in_z)   // walls (:963-1160): the projection in place, then its pressure (with the Stokes pieces of the last getPressure)
        return with_uvw(p, uhat, vhat, what, true, st, [&](double2* u, double2* v, double2* w) -> int {
            if (int rc = poiss_wall_projection(p, u, v, w, st, false)) return rc;
            return poiss_wall_pressure_out(p, pressure, st);
        });
    return with_uvw(p, uhat, vhat, what, true, st, [&](double2* u, double2* v, double2* w) -> int {
        if (int rc = poiss_projection(p, u, v, w, st)) return rc;  // leaves phat at phat_y
        const size_t bytes = sizeof(double) * (size_t)vol(p->sp->pi.xsz);
        return with_device_views(pressure, 0, pressure, bytes, st, [&](const void*, void* d_o) {
            return fft3d_backward_yx(p->sp->ft, p->phat_y, (double*)d_o, false, st);
        });
    });
}

int pdo_padepoisson_divergence_check(pdo_padepoisson_t p, double* uhat, double* vhat, double* what, double* divergence, int fix_div,
                                     double* max_div, void* stream) {
    if (!p || !uhat || !vhat || !what) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    return with_uvw(p, uhat, vhat, what, fix_div != 0, st, [&](double2* u, double2* v, double2* w) -> int {
        if (!divergence) return poiss_divergence_check(p, u, v, w, nullptr, fix_div != 0, max_div, st);
        const size_t bytes = sizeof(double) * (size_t)vol(p->sp->pi.xsz);
        return with_device_views(divergence, 0, divergence, bytes, st, [&](const void*, void* d_o) {
            return poiss_divergence_check(p, u, v, w, (double*)d_o, fix_div != 0, max_div, st);
        });
    });
}

}  // extern "C"
