// ig_pade6stagg.inc.cuh — part of igrid.cu: textually included there, ONE translation unit (the sections share file-local helpers).
// PadeDerOps::Pade6stagg: cd06 / fourierColl / wall dispatch, C ABI.
// Not a stand-alone header: do not include it anywhere else.

// ================================================================================================
// Pade6stagg (periodic)
// ================================================================================================
struct pdo_pade6stagg_s {
    int gp_zsz[3], sp_zsz[3];
    double dz;
    int scheme;
    pdo_cd06stagg_t der = nullptr;
    pdo_spectral_t spectC = nullptr;   // scheme = fourierColl: the spectral type whose z transforms and tables are used (borrowed)
    // isPeriodic = .false., cd06: the nine wall handles derOO .. derSS (PadeDerOps.F90:92-110) at index 3 (bot + 1) + (top + 1),
    // bot / top = -1 odd, 0 one-sided, +1 even
    bool periodic = true;
    pdo_cd06stagg_t wall[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {
typedef int (*stagg_fn)(pdo_cd06stagg_t, const double*, double*, int, int, int, void*);

// Fourier collocation in z (spectral.F90:365-702): c2c-z forward on the first nz planes, multiply plane k by table(k), c2c-z
// backward, x 1/nz; edge outputs get plane nz+1 := plane 1.  Complex arrays live on the spectral z-pencil, real ones on the
// physical z-pencil (r2c / c2r in the reference, oddball mode untouched: zfourier_real).
// which: 0 ddz_E2C, 1 ddz_C2E, 2 interp_E2C, 3 interp_C2E, 4 d2dz2_C2C, 5 d2dz2_E2E
int pade_fourier(pdo_pade6stagg_s* p, int which, const double* in, double* out, int is_complex, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    pdo_spectral_s* s = p->spectC;
    const int nz = s->nz;
    const int* zs = is_complex ? s->si.zsz : s->pi.zsz;
    const size_t esz = is_complex ? sizeof(double2) : sizeof(double);
    const size_t plane = (size_t)zs[0] * zs[1];
    const bool edge_in = (which == 0 || which == 2 || which == 5), edge_out = (which == 1 || which == 3 || which == 5);
    const size_t bin = esz * plane * (size_t)(nz + (edge_in ? 1 : 0)), bout = esz * plane * (size_t)(nz + (edge_out ? 1 : 0));
    static const int tab_of[6] = {ZT_K3_E2C, ZT_K3_C2E, ZT_E2C, ZT_C2E, ZT_MK3SQ, ZT_MK3SQ};
    return with_device_views(in, bin, out, bout, st, [&](const void* di, void* d_o) -> int {
        if (is_complex) {
            double2* w = (double2*)d_o;
            if (di != d_o) PDO_CUDA(cudaMemcpyAsync(w, di, esz * plane * (size_t)nz, cudaMemcpyDeviceToDevice, st));
            if (int rc = zfourier_complex(s, w, tab_of[which], st)) return rc;
        } else {
            if (int rc = zfourier_real(s, (const double*)di, (double*)d_o, tab_of[which], st)) return rc;
        }
        if (edge_out) PDO_CUDA(cudaMemcpyAsync((char*)d_o + esz * plane * (size_t)nz, d_o, esz * plane, cudaMemcpyDeviceToDevice, st));
        return 0;
    });
}

int pade_apply(pdo_pade6stagg_s* p, stagg_fn fn, int which, const double* in, double* out, int is_complex, int bot, int top, void* st) {
    if (!p) return fail(PDO_E_BADARG, "null handle");
    const int* z = is_complex ? p->sp_zsz : p->gp_zsz;
    if (!p->periodic) {
        // PadeDerOps.F90:185-205, 449-482, ...: the first-order operators take bot, top in {-1, 0, +1}, the second derivatives
        // {-1, +1}; any other code gives output = 0
        const bool second = which >= 4;
        const bool ok = bot >= -1 && bot <= 1 && top >= -1 && top <= 1 && !(second && (bot == 0 || top == 0));
        if (!ok) {
            if (!out) return fail(PDO_E_BADARG, "null field pointer");
            const bool edge_out = (which == 1 || which == 3 || which == 5);
            const size_t bytes = sizeof(double) * (is_complex ? 2 : 1) * (size_t)z[0] * z[1] * (size_t)(z[2] + (edge_out ? 1 : 0));
            if (is_device_ptr(out)) PDO_CUDA(cudaMemsetAsync(out, 0, bytes, (cudaStream_t)st));
            else std::memset(out, 0, bytes);
            return 0;
        }
        return fn(p->wall[3 * (bot + 1) + (top + 1)], in, out, z[0], z[1], is_complex, st);
    }
    if (p->scheme == PDO_SCHEME_FOURIER) return pade_fourier(p, which, in, out, is_complex, st);
    return fn(p->der, in, out, z[0], z[1], is_complex, st);
}
}  // namespace

extern "C" {

int pdo_pade6stagg_init2(pdo_pade6stagg_t* h, const int gp_zsz[3], const int sp_zsz[3], double dz, int scheme, int is_periodic,
                         pdo_spectral_t spectC) {
    if (!h || !gp_zsz || !sp_zsz) return fail(PDO_E_BADARG, "null argument");
    *h = nullptr;
    if (scheme == PDO_SCHEME_FD02) return fail(PDO_E_UNSUPPORTED, "Pade6stagg: scheme fd02 is not built (cd06 and fourierColl are)");
    if (!is_periodic && scheme != PDO_SCHEME_CD06) return fail(323, "Invalid choice for numerical scheme in vertical direction");  // PadeDerOps.F90:121
    if (scheme != PDO_SCHEME_CD06 && scheme != PDO_SCHEME_FOURIER) return fail(434, "Invalid choice of numerical scheme in vertical");  // PadeDerOps.F90:84
    if (scheme == PDO_SCHEME_FOURIER) {
        if (!spectC) return fail(43, "You need to pass in a spectral derived type if you want to use Fourier differentiation in z");  // :77
        if (!spectC->periodicInZ) return fail(PDO_E_BADARG, "fourierColl needs a spectral type initialised with init_periodicInZ");
        if (spectC->si.zsz[0] != sp_zsz[0] || spectC->si.zsz[1] != sp_zsz[1] || spectC->si.zsz[2] != sp_zsz[2])
            return fail(PDO_E_BADARG, "spectral type and sp_gpC disagree on the z-pencil");
        if (spectC->pi.zsz[0] != gp_zsz[0] || spectC->pi.zsz[1] != gp_zsz[1] || spectC->pi.zsz[2] != gp_zsz[2])
            return fail(PDO_E_BADARG, "spectral type and gpC disagree on the z-pencil");
    }
    pdo_pade6stagg_s* p = new (std::nothrow) pdo_pade6stagg_s();
    if (!p) return fail(PDO_E_BADARG, "out of memory");
    std::memcpy(p->gp_zsz, gp_zsz, sizeof(int) * 3);
    std::memcpy(p->sp_zsz, sp_zsz, sizeof(int) * 3);
    p->dz = dz; p->scheme = scheme;
    p->periodic = is_periodic != 0;
    if (!p->periodic) {
        // derOO .. derSS (:92-110): the Even flag of a one-sided wall never reaches a row (the sided branch comes first)
        for (int bot = -1; bot <= 1; ++bot)
            for (int top = -1; top <= 1; ++top) {
                int rc = pdo_cd06stagg_init_nonperiodic(&p->wall[3 * (bot + 1) + (top + 1)], gp_zsz[2], dz, top == 1, bot == 1, top == 0, bot == 0);
                if (rc) { pdo_pade6stagg_destroy(p); return rc; }
            }
    } else if (scheme == PDO_SCHEME_CD06) {
        int rc = pdo_cd06stagg_init_periodic(&p->der, gp_zsz[2], dz);  // derPeriodic%init(gp%zsz(3), dz)  :79-80
        if (rc) { delete p; return rc; }
    } else {
        p->spectC = spectC;   // the tables are the spectral type's own (spectral.F90:843-856), built on first use
        if (int rc = spectral_ztables(spectC)) { delete p; return rc; }
    }
    *h = p;
    return 0;
}
int pdo_pade6stagg_init(pdo_pade6stagg_t* h, const int gp_zsz[3], const int sp_zsz[3], double dz, int scheme, int is_periodic) {
    return pdo_pade6stagg_init2(h, gp_zsz, sp_zsz, dz, scheme, is_periodic, nullptr);
}
int pdo_pade6stagg_destroy(pdo_pade6stagg_t p) {
    if (!p) return 0;
    pdo_cd06stagg_destroy(p->der);
    for (int i = 0; i < 9; ++i) pdo_cd06stagg_destroy(p->wall[i]);
    delete p;
    return 0;
}
#define PDO_PADE_FN(name, target, which)                                                                                   \
    int name(pdo_pade6stagg_t p, const double* in, double* out, int is_complex, int bot, int top, void* st) {              \
        return pade_apply(p, target, which, in, out, is_complex, bot, top, st);                                            \
    }
PDO_PADE_FN(pdo_pade6stagg_ddz_C2E, pdo_cd06stagg_ddz_C2E, 1)
PDO_PADE_FN(pdo_pade6stagg_ddz_E2C, pdo_cd06stagg_ddz_E2C, 0)
PDO_PADE_FN(pdo_pade6stagg_interpz_C2E, pdo_cd06stagg_interpz_C2E, 3)
PDO_PADE_FN(pdo_pade6stagg_interpz_E2C, pdo_cd06stagg_interpz_E2C, 2)
PDO_PADE_FN(pdo_pade6stagg_d2dz2_C2C, pdo_cd06stagg_d2dz2_C2C, 4)
PDO_PADE_FN(pdo_pade6stagg_d2dz2_E2E, pdo_cd06stagg_d2dz2_E2E, 5)

// getmodCD06stagg (PadeDerOps.F90:1034-1053)
int pdo_pade6stagg_get_modified_wavenumbers(pdo_pade6stagg_t p, const double* k, double* kp, int n) {
    if (!p || !k || !kp) return fail(PDO_E_BADARG, "null argument");
    if (p->scheme == PDO_SCHEME_FOURIER) {   // PadeDerOps.F90:1003-1004
        for (int i = 0; i < n; ++i) kp[i] = k[i];
        return 0;
    }
    const double alpha = 9.0 / 62.0, beta = 0.0, a = 63.0 / 62.0, b = 17.0 / 62.0, c = 0.0;
    for (int i = 0; i < n; ++i) {
        const double omega = k[i] * p->dz;
        double v = (2.0 * a * std::sin(omega / 2.0) + (2.0 / 3.0) * b * std::sin(3.0 * omega / 2.0) + (2.0 / 5.0) * c * std::sin(5.0 * omega / 2.0)) /
                   (1.0 + 2.0 * alpha * std::cos(omega) + 2.0 * beta * std::cos(2.0 * omega));
        kp[i] = v / p->dz;
    }
    return 0;
}

}  // extern "C"
