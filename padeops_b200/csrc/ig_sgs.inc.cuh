// ig_sgs.inc.cuh — part of igrid.cu: textually included there, ONE translation unit (the sections share file-local helpers).
// sgs_igrid: eddy-viscosity SGS term of the right-hand side (inside igrid.cu's anonymous namespace).
// Not a stand-alone header: do not include it anywhere else.

// ---- Step 6 of populate_rhs: the SGS term (sgsmod_igrid.F90:156-268), eddy-viscosity models with a global constant ----
struct Grad9 { const double* p[9]; };

// nu = cmodel_global * kernel(duidxj)   (get_SGS_kernel + multiply_by_model_constant, eddyViscosity.F90:38-95)
int sgs_nu(const SgsConst& c, const Grad9& G, double* nu, long long n, cudaStream_t st) {
    return launch_ew(n, st, [=] __device__(long long i) {
        double d[9], S[6];
#pragma unroll
        for (int k = 0; k < 9; ++k) d[k] = G.p[k][i];
        sgs_sij(d, S);
        nu[i] = c.cmodel * sgs_kernel_point(c, d, S);
    });
}
// tau = -2 nu S: S = a (diagonal components) or 0.5 (a + b)
int sgs_tau(double* tau, const double* nu, const double* a, const double* b, long long n, cudaStream_t st) {
    if (b) return launch_ew(n, st, [=] __device__(long long i) { tau[i] = -2.0 * nu[i] * (0.5 * (a[i] + b[i])); });
    return launch_ew(n, st, [=] __device__(long long i) { tau[i] = -2.0 * nu[i] * a[i]; });
}
// dst -= src (complex arrays viewed as doubles)
int csub(double2* dst, const double2* src, long long n, cudaStream_t st) {
    double* d = (double*)dst;
    const double* sp = (const double*)src;
    return launch_ew(2 * n, st, [=] __device__(long long i) { d[i] -= sp[i]; });
}
// dst -= i k f  (mTimes_ik*_ip / _oop followed by "rhs = rhs - cbuffy")
int csub_ik(pdo_spectral_s* s, int which, double2* dst, const double2* f, cudaStream_t st) {
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const double* k = which == 1 ? s->k1y : s->k2;
    const int w = which;
    return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
        const double kv = (w == 1) ? k[(int)(i % n1)] : k[(int)((i / n1) % n2)];
        const double2 q = f[i];
        double2 a = dst[i];
        a.x -= -kv * q.y; a.y -= kv * q.x;
        dst[i] = a;
    });
}
// interpolate_eddy_viscosity(.true.) (eddyViscosity.F90:97-113): x -> y -> z on gpC, interpz_C2E on the REAL array, z -> y -> x on
// gpE, negative values clipped; a transpose inside a 1-rank group is the identity and is skipped
int sgs_interp_nu(pdo_igrid_s* g, cudaStream_t st) {
    pdo_decomp_t pC = fft3d_phys_decomp(g->spC->ft), pE = fft3d_phys_decomp(g->spE->ft);
    const bool tx = g->spC->p_row > 1, tz = g->spC->p_col > 1;
    const double* a = g->sgs_nuC;
    // (tx && !tz): the y-pencil IS the z-pencil; it lands in rzC so that the edge result can take ry
    double* ybuf = (tx && !tz) ? g->sgs_rzC : g->sgs_ry;
    if (tx) { IG(decomp_transpose_device(pC, 0, a, ybuf, 1, st)); a = ybuf; }
    if (tz) { IG(decomp_transpose_device(pC, 2, a, g->sgs_rzC, 1, st)); a = g->sgs_rzC; }
    double* zout = tz ? g->sgs_rzE : (tx ? g->sgs_ry : g->sgs_nuE);
    IG(pdo_pade6stagg_interpz_C2E(g->ops, a, zout, 0, 0, 0, st));
    const double* b = zout;
    if (tz) { double* yd = tx ? g->sgs_ry : g->sgs_nuE; IG(decomp_transpose_device(pE, 3, b, yd, 1, st)); b = yd; }
    if (tx) IG(decomp_transpose_device(pE, 1, b, g->sgs_nuE, 1, st));
    double* nuE = g->sgs_nuE;
    return launch_ew(g->nRE, st, [=] __device__(long long i) { if (nuE[i] < 0.0) nuE[i] = 0.0; });
}

int ig_sgs_rhs(pdo_igrid_s* g, double2* ru, double2* rv, double2* rw, cudaStream_t st) {
    pdo_spectral_s *C = g->spC, *E = g->spE;
    Grad9 GC, GE;
    for (int k = 0; k < 9; ++k) { GC.p[k] = g->gradC[k]; GE.p[k] = g->gradE[k]; }
    double **dC = g->gradC, **dE = g->gradE;
    // getTauSGS :156-203
    IG(sgs_nu(g->sgs, GC, g->sgs_nuC, g->nRC, st));
    if (g->sgs_explicit_edge) IG(sgs_nu(g->sgs, GE, g->sgs_nuE, g->nRE, st));
    else IG(sgs_interp_nu(g, st));
    double *TC = g->rbC[0], *TE = g->rbE[0];
    double2 *fC = g->yC[0], *fE = g->yE[0], *gC2 = g->yC[1], *gE2 = g->yE[1];
    const double2* z = nullptr;
    double2* t = nullptr;
    // ddx(tau11) -> urhs
    IG(sgs_tau(TC, g->sgs_nuC, dC[0], nullptr, g->nRC, st));
    IG(fftC(g, TC, fC, st));
    IG(csub_ik(C, 1, ru, fC, st));
    // ddy(tau22) -> vrhs
    IG(sgs_tau(TC, g->sgs_nuC, dC[4], nullptr, g->nRC, st));
    IG(fftC(g, TC, fC, st));
    IG(csub_ik(C, 2, rv, fC, st));
    // ddz(tau33) -> wrhs
    IG(sgs_tau(TC, g->sgs_nuC, dC[8], nullptr, g->nRC, st));
    IG(fftC(g, TC, fC, st));
    IG(zviewC(g, fC, g->zC[0], &z, st));
    t = ztarget(g, gE2, g->zE[0]);
    ZOP(pdo_pade6stagg_ddz_C2E, z, t);
    IG(zcommitE(g, t, gE2, st));
    IG(csub(rw, gE2, g->nYE, st));
    // tau12: ddx -> vrhs, ddy -> urhs
    IG(sgs_tau(TC, g->sgs_nuC, dC[1], dC[3], g->nRC, st));
    IG(fftC(g, TC, fC, st));
    IG(csub_ik(C, 1, rv, fC, st));
    IG(csub_ik(C, 2, ru, fC, st));
    // tau13 (edges): ddz -> urhs, ddx -> wrhs;  tau23: ddz -> vrhs, ddy -> wrhs
    for (int c = 0; c < 2; ++c) {
        IG(sgs_tau(TE, g->sgs_nuE, c == 0 ? dE[2] : dE[5], c == 0 ? dE[6] : dE[7], g->nRE, st));
        IG(fftE(g, TE, fE, st));
        IG(zviewE(g, fE, g->zE[0], &z, st));
        t = ztarget(g, gC2, g->zC[0]);
        ZOP(pdo_pade6stagg_ddz_E2C, z, t);
        IG(zcommitC(g, t, gC2, st));
        IG(csub(c == 0 ? ru : rv, gC2, g->nYC, st));
        IG(csub_ik(E, c == 0 ? 1 : 2, rw, fE, st));
    }
    return 0;
}
