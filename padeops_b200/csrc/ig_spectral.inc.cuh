// ig_spectral.inc.cuh — part of igrid.cu: textually included there, ONE translation unit (the sections share file-local helpers).
// spectralMod::spectral: tables, pointwise passes, z-Fourier operators, C ABI.
// Not a stand-alone header: do not include it anywhere else.

// ================================================================================================
// spectral
// ================================================================================================
struct pdo_spectral_s {
    int nx, ny, nz, nxh, p_row, p_col;
    double dx, dy, dz;
    pdo_fft3d_t ft = nullptr;
    pdo_decomp_info pi, si;
    bool periodicInZ = false;
    double normfactz = 1.0;
    std::vector<double> h_k1, h_k2, h_gx, h_gy, h_gz;  // global 1-D tables (nxh, ny, nxh, ny, nz)
    double *k1y = nullptr, *k2 = nullptr;               // local slice of k1 (ysz0 == zsz0 entries), full k2
    double* k2z = nullptr;                              // y slice of k2 owned by the z-pencil
    double *gx = nullptr, *gy = nullptr, *gyz = nullptr, *gz = nullptr;  // dealias masks: x slice, y full, y slice of the z-pencil, z
    double2* ctmpz = nullptr;
    double* partial = nullptr;  // reduction scratch
    // z-Fourier tables of init_periodic_inZ_procedures (spectral.F90:843-856), built on first use: 6 tables of nz complex numbers
    // (k3_E2Cshift, k3_C2Eshift, E2Cshift, C2Eshift, mk3sq, k3_C2Cder), then the same six with the oddball entry set to 1 for the
    // REAL procedures, which leave that mode untouched
    double2* ztab = nullptr;
    ZColsPlan rz_plan;          // c2c-z over pairs of real columns
    double2* rz_work = nullptr;
    size_t rz_cap = 0;
};
enum { ZT_K3_E2C = 0, ZT_K3_C2E = 1, ZT_E2C = 2, ZT_C2E = 3, ZT_MK3SQ = 4, ZT_K3_C2C = 5, ZT_COUNT = 6 };

namespace {

// fout = i k fin (* scale: lets a caller fold the inverse transform's 1/(nx ny) into this pass)
int spectral_mtimes(pdo_spectral_s* s, int which, const double2* fin, double2* fout, cudaStream_t st, double scale = 1.0) {
    const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
    const long long n = vol(s->si.ysz);
    const double* k = which == 1 ? s->k1y : s->k2;
    if (which == 1)
        return launch_ew(n, st, [=] __device__(long long i) {
            const double kv = k[(int)(i % n1)] * scale;
            const double2 v = fin[i];
            fout[i] = make_double2(-kv * v.y, kv * v.x);
        });
    return launch_ew(n, st, [=] __device__(long long i) {
        const double kv = k[(int)((i / n1) % n2)] * scale;
        const double2 v = fin[i];
        fout[i] = make_double2(-kv * v.y, kv * v.x);
    });
}

// ---- z-Fourier operators: c2c-z forward, x table(k), c2c-z backward, x 1/nz (spectral.F90:365-702) ----
// host side of the tables: [2][ZT_COUNT][nz], complex procedures first, then the REAL procedures' twins (oddball entry = 1)
std::vector<double2> build_ztables_host(int nz, double dz) {
    std::vector<double> k3 = wavenums(nz, dz);   // GetWaveNums(nz, dz), no sign flip of the oddball (spectral.F90:845)
    std::vector<double2> t(2 * ZT_COUNT * (size_t)nz);
    for (int k = 0; k < nz; ++k) {
        const double kk = k3[k], ph = kk * dz / 2.0, c = std::cos(ph), sn = std::sin(ph);
        t[(size_t)ZT_K3_E2C * nz + k] = make_double2(-kk * sn, kk * c);    // i k e^{+i k dz/2}
        t[(size_t)ZT_K3_C2E * nz + k] = make_double2(kk * sn, kk * c);     // i k e^{-i k dz/2}
        t[(size_t)ZT_E2C * nz + k] = make_double2(c, sn);                  // e^{+i k dz/2}
        t[(size_t)ZT_C2E * nz + k] = make_double2(c, -sn);                 // e^{-i k dz/2}
        t[(size_t)ZT_MK3SQ * nz + k] = make_double2(-(kk * kk), 0.0);      // -k^2
        t[(size_t)ZT_K3_C2C * nz + k] = make_double2(0.0, kk);             // i k
    }
    for (int i = 0; i < ZT_COUNT; ++i)
        for (int k = 0; k < nz; ++k)
            t[(size_t)(ZT_COUNT + i) * nz + k] = k == nz / 2 ? make_double2(1.0, 0.0) : t[(size_t)i * nz + k];
    return t;
}
int spectral_ztables(pdo_spectral_s* s) {
    if (s->ztab) return 0;
    if (!s->periodicInZ) return fail(PDO_E_BADARG, "spectral type was not initialised with init_periodicInZ");
    std::vector<double2> t = build_ztables_host(s->nz, s->dz);
    PDO_CUDA(cudaMalloc(&s->ztab, sizeof(double2) * t.size()));
    PDO_CUDA(cudaMemcpy(s->ztab, t.data(), sizeof(double2) * t.size(), cudaMemcpyHostToDevice));
    return 0;
}
inline const double2* ztable(const pdo_spectral_s* s, int which, bool real_variant) {
    return s->ztab + (size_t)((real_variant ? ZT_COUNT : 0) + which) * s->nz;
}
// w(cols, nz) *= tab(k) * scale
int ztable_multiply(double2* w, long long cols, int nz, const double2* tab, double scale, cudaStream_t st) {
    return launch_ew(cols * nz, st, [=] __device__(long long i) {
        double2 t = tab[(int)(i / cols)];
        const double2 v = w[i];
        t.x *= scale; t.y *= scale;
        w[i] = make_double2(v.x * t.x - v.y * t.y, v.x * t.y + v.y * t.x);
    });
}
// complex z-pencil array of the spectral decomposition, in place on the first nz planes of w
int zfourier_complex(pdo_spectral_s* s, double2* w, int which, cudaStream_t st) {
    if (int rc = spectral_ztables(s)) return rc;
    const long long cols = (long long)s->si.zsz[0] * s->si.zsz[1];
    if (int rc = fft3d_z_inplace(s->ft, w, -1, st)) return rc;
    if (int rc = ztable_multiply(w, cols, s->nz, ztable(s, which, false), 1.0 / (double)s->nz, st)) return rc;
    return fft3d_z_inplace(s->ft, w, +1, st);
}
// REAL z-pencil array of the physical decomposition: in(P, nz [+1]) -> out(P, nz).  The table with the oddball entry = 1 is
// conjugate-symmetric in k, so the operator maps real columns to real columns and is linear over C: two real columns a, b
// are transformed as ONE complex column a + i b and come back as a' + i b' (half the transform work of a zero-padded c2c,
// the same arithmetic as the reference's r2c / c2r pair up to rounding).  P odd: the last column is paired with zeros.
int zfourier_real(pdo_spectral_s* s, const double* in, double* out, int which, cudaStream_t st) {
    if (int rc = spectral_ztables(s)) return rc;
    const int nz = s->nz;
    const long long P = (long long)s->pi.zsz[0] * s->pi.zsz[1], Pc = (P + 1) / 2;
    const size_t need = sizeof(double2) * (size_t)Pc * nz;
    if (s->rz_cap < need) {
        if (s->rz_work) cudaFree(s->rz_work);
        s->rz_work = nullptr; s->rz_cap = 0;
        PDO_CUDA(cudaMalloc(&s->rz_work, need));
        s->rz_cap = need;
    }
    double2* w = s->rz_work;
    if (P & 1) PDO_CUDA(cudaMemsetAsync(w, 0, need, st));
    PDO_CUDA(cudaMemcpy2DAsync(w, sizeof(double2) * Pc, in, sizeof(double) * P, sizeof(double) * P, nz, cudaMemcpyDeviceToDevice, st));
    if (int rc = zcols_exec(&s->rz_plan, nz, Pc, w, -1, st)) return rc;
    if (int rc = ztable_multiply(w, Pc, nz, ztable(s, which, true), 1.0 / (double)nz, st)) return rc;
    if (int rc = zcols_exec(&s->rz_plan, nz, Pc, w, +1, st)) return rc;
    PDO_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * P, w, sizeof(double2) * Pc, sizeof(double) * P, nz, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// z-pencil array a(zsz0, zsz1, nz) *= gx(i) gy(j) gz(k) * scale
int spectral_mask_z(pdo_spectral_s* s, double2* a, double scale, cudaStream_t st) {
    const int n1 = s->si.zsz[0], n2 = s->si.zsz[1];
    const long long n = (long long)n1 * n2 * s->nz;
    const double *gx = s->gx, *gy = s->gyz, *gz = s->gz;
    return launch_ew(n, st, [=] __device__(long long i) {
        const int ii = (int)(i % n1);
        const long long t = i / n1;
        const int jj = (int)(t % n2);
        const int kk = (int)(t / n2);
        const double m = gx[ii] * gy[jj] * gz[kk] * scale;
        double2 v = a[i];
        v.x *= m; v.y *= m;
        a[i] = v;
    });
}

// the z-pencil part of dealias (spectral.F90:343-363) on a z-pencil array of the spectral decomposition, in place
int spectral_dealias_zwork(pdo_spectral_s* s, double2* work, cudaStream_t st) {
    if (int rc = fft3d_z_inplace(s->ft, work, -1, st)) return rc;
    if (fft3d_own_z(s->ft)) return fft3d_z_pro(s->ft, work, work, +1, s->gx, s->gyz, s->gz, s->normfactz, st);   // mask and 1/nz on the first load
    if (int rc = spectral_mask_z(s, work, s->normfactz, st)) return rc;
    return fft3d_z_inplace(s->ft, work, +1, st);  // take_ifftz
}

int spectral_dealias(pdo_spectral_s* s, double2* fhat, cudaStream_t st) {
    if (!s->periodicInZ) {  // 2-D mask (spectral.F90:329-338 with the table of :1147-1159)
        const int n1 = s->si.ysz[0], n2 = s->si.ysz[1];
        const double *gx = s->gx, *gy = s->gy;
        return launch_ew(vol(s->si.ysz), st, [=] __device__(long long i) {
            const double m = gx[(int)(i % n1)] * gy[(int)((i / n1) % n2)];
            double2 v = fhat[i];
            v.x *= m; v.y *= m;
            fhat[i] = v;
        });
    }
    pdo_decomp_t spec = fft3d_spec_decomp(s->ft);
    double2* work = fhat;  // one rank in the column communicator: y- and z-pencil layouts coincide
    if (s->p_col > 1) {
        work = s->ctmpz;
        if (int rc = decomp_transpose_device(spec, 2, (const double*)fhat, (double*)work, 2, st)) return rc;  // take_fftz
    }
    if (int rc = spectral_dealias_zwork(s, work, st)) return rc;
    if (s->p_col > 1) return decomp_transpose_device(spec, 3, (const double*)work, (double*)fhat, 2, st);
    return 0;
}

int spectral_dealias_edge(pdo_spectral_s* s, double2* fE, cudaStream_t st) {
    if (!s->periodicInZ) return 0;  // the reference does nothing on this branch (spectral.F90:348)
    if (int rc = spectral_dealias_zwork(s, fE, st)) return rc;
    const size_t plane = (size_t)s->si.zsz[0] * s->si.zsz[1];
    PDO_CUDA(cudaMemcpyAsync(fE + plane * s->nz, fE, sizeof(double2) * plane, cudaMemcpyDeviceToDevice, st));  // :361
    return 0;
}

}  // namespace

extern "C" {

int pdo_spectral_init(pdo_spectral_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col,
                      int fix_oddball, int init_periodic_in_z, double dealias_fact) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    if (nx < 2 || ny < 2 || nz < 1) return fail(PDO_E_BADARG, "bad sizes");
    if (init_periodic_in_z && (nz % 2) != 0)
        return fail(104, "You cannot initialize a periodic_inZ spectral type with an odd values nz");  // spectral.F90:773-775
    if (p_row == 0 && p_col == 0) { p_row = 1; p_col = pdo_comm_size(); }
    pdo_spectral_s* s = new (std::nothrow) pdo_spectral_s();
    if (!s) return fail(PDO_E_BADARG, "out of memory");
    s->nx = nx; s->ny = ny; s->nz = nz; s->nxh = nx / 2 + 1; s->dx = dx; s->dy = dy; s->dz = dz;
    s->p_row = p_row; s->p_col = p_col;
    s->periodicInZ = init_periodic_in_z != 0;
    int rc = pdo_fft3d_init(&s->ft, nx, ny, nz, dx, dy, dz, p_row, p_col);
    if (rc) { delete s; return rc; }
    pdo_fft3d_get_physical_info(s->ft, &s->pi);
    pdo_fft3d_get_spectral_info(s->ft, &s->si);
    // 1-D wavenumbers with the oddball sign flip (spectral.F90:1024-1032) and the optional fixOddball (:1189-1199)
    std::vector<double> k1 = wavenums(nx, dx), k2 = wavenums(ny, dy), k3 = wavenums(nz, dz);
    k1[nx / 2] = -k1[nx / 2];
    k2[ny / 2] = -k2[ny / 2];
    k3[nz / 2] = -k3[nz / 2];
    if (fix_oddball) { k1[nx / 2] = 0.0; k2[ny / 2] = 0.0; }
    s->h_k1.assign(k1.begin(), k1.begin() + s->nxh);
    s->h_k2 = k2;
    s->h_gx.assign(s->nxh, 1.0); s->h_gy.assign(ny, 1.0); s->h_gz.assign(nz, 1.0);
    if (s->periodicInZ) {  // zero where |k| >= f pi/d  (spectral.F90:785-814)
        const double kdx = dealias_fact * kPi / dx, kdy = dealias_fact * kPi / dy, kdz = dealias_fact * kPi / dz;
        for (int i = 0; i < s->nxh; ++i) if (std::fabs(s->h_k1[i]) >= kdx) s->h_gx[i] = 0.0;
        for (int j = 0; j < ny; ++j) if (std::fabs(k2[j]) >= kdy) s->h_gy[j] = 0.0;
        for (int k = 0; k < nz; ++k) if (std::fabs(k3[k]) >= kdz) s->h_gz[k] = 0.0;
        s->normfactz = 1.0 / (double)nz;
    } else {               // pass band |k| < (2/3) pi/d, factor hard-wired (spectral.F90:1147-1159)
        const double kdx = (2.0 / 3.0) * kPi / dx, kdy = (2.0 / 3.0) * kPi / dy;
        for (int i = 0; i < s->nxh; ++i) s->h_gx[i] = (std::fabs(s->h_k1[i]) < kdx) ? 1.0 : 0.0;
        for (int j = 0; j < ny; ++j) s->h_gy[j] = (std::fabs(k2[j]) < kdy) ? 1.0 : 0.0;
    }
    const int i0 = s->si.yst[0] - 1, ni = s->si.ysz[0];
    rc = upload(&s->k1y, s->h_k1, i0, ni);
    if (!rc) rc = upload(&s->k2, s->h_k2, 0, ny);
    if (!rc) rc = upload(&s->gx, s->h_gx, i0, ni);
    if (!rc) rc = upload(&s->gy, s->h_gy, 0, ny);
    if (!rc) rc = upload(&s->gyz, s->h_gy, s->si.zst[1] - 1, s->si.zsz[1]);
    if (!rc) rc = upload(&s->k2z, s->h_k2, s->si.zst[1] - 1, s->si.zsz[1]);
    if (!rc) rc = upload(&s->gz, s->h_gz, 0, nz);
    if (!rc && s->periodicInZ && p_col > 1) {
        rc = comm_shared_malloc((void**)&s->ctmpz, sizeof(double2) * (size_t)vol(s->si.zsz));
    }
    if (!rc) {
        cudaError_t e = cudaMalloc(&s->partial, sizeof(double) * 2048);
        if (e != cudaSuccess) rc = fail(PDO_E_CUDA, "spectral scratch: %s", cudaGetErrorString(e));
    }
    if (rc) { pdo_spectral_destroy(s); return rc; }
    *h = s;
    return 0;
}

int pdo_spectral_destroy(pdo_spectral_t s) {
    if (!s) return 0;
    double* ptrs[] = {s->k1y, s->k2, s->k2z, s->gx, s->gy, s->gyz, s->gz, s->partial};
    for (double* p : ptrs) if (p) cudaFree(p);
    comm_shared_free(s->ctmpz);
    if (s->ztab) cudaFree(s->ztab);
    if (s->rz_work) cudaFree(s->rz_work);
    zcols_destroy(&s->rz_plan);
    pdo_fft3d_destroy(s->ft);
    delete s;
    return 0;
}

int pdo_spectral_get_physical_info(pdo_spectral_t s, pdo_decomp_info* info) {
    if (!s || !info) return fail(PDO_E_BADARG, "null argument");
    *info = s->pi;
    return 0;
}
int pdo_spectral_get_spectral_info(pdo_spectral_t s, pdo_decomp_info* info) {
    if (!s || !info) return fail(PDO_E_BADARG, "null argument");
    *info = s->si;
    return 0;
}
int pdo_spectral_get_tables(pdo_spectral_t s, double* k1, double* k2, double* gx, double* gy, double* gz) {
    if (!s) return fail(PDO_E_BADARG, "null handle");
    if (k1) std::memcpy(k1, s->h_k1.data(), sizeof(double) * s->nxh);
    if (k2) std::memcpy(k2, s->h_k2.data(), sizeof(double) * s->ny);
    if (gx) std::memcpy(gx, s->h_gx.data(), sizeof(double) * s->nxh);
    if (gy) std::memcpy(gy, s->h_gy.data(), sizeof(double) * s->ny);
    if (gz) std::memcpy(gz, s->h_gz.data(), sizeof(double) * s->nz);
    return 0;
}

int pdo_spectral_fft(pdo_spectral_t s, const double* in, double* out, void* stream) {
    if (!s) return fail(PDO_E_BADARG, "null handle");
    return pdo_fft3d_fft2_x2y(s->ft, in, out, stream);
}
int pdo_spectral_ifft(pdo_spectral_t s, const double* in, double* out, int set_oddball, void* stream) {
    if (!s) return fail(PDO_E_BADARG, "null handle");
    return pdo_fft3d_ifft2_y2x(s->ft, in, out, set_oddball, stream);
}

static int spectral_ywise(pdo_spectral_t s, const double* in, double* out, void* stream, int op) {
    if (!s || !in || !out) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = sizeof(double2) * (size_t)vol(s->si.ysz);
    return with_device_views(in, bytes, out, bytes, st, [&](const void* di, void* d_o) -> int {
        if (op == 1 || op == 2) return spectral_mtimes(s, op, (const double2*)di, (double2*)d_o, st);
        if (di != d_o) PDO_CUDA(cudaMemcpyAsync(d_o, di, bytes, cudaMemcpyDeviceToDevice, st));
        return spectral_dealias(s, (double2*)d_o, st);
    });
}
int pdo_spectral_mtimes_ik1_oop(pdo_spectral_t s, const double* fin, double* fout, void* st) { return spectral_ywise(s, fin, fout, st, 1); }
int pdo_spectral_mtimes_ik2_oop(pdo_spectral_t s, const double* fin, double* fout, void* st) { return spectral_ywise(s, fin, fout, st, 2); }
int pdo_spectral_mtimes_ik1_ip(pdo_spectral_t s, double* f, void* st) { return spectral_ywise(s, f, f, st, 1); }
int pdo_spectral_mtimes_ik2_ip(pdo_spectral_t s, double* f, void* st) { return spectral_ywise(s, f, f, st, 2); }
int pdo_spectral_dealias(pdo_spectral_t s, double* fhat, void* st) { return spectral_ywise(s, fhat, fhat, st, 3); }

static int spectral_zwise(pdo_spectral_t s, double* a, void* stream, int op) {
    if (!s || !a) return fail(PDO_E_BADARG, "null argument");
    if (!s->periodicInZ) return fail(PDO_E_BADARG, "spectral type was not initialised with init_periodicInZ");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane = (size_t)s->si.zsz[0] * s->si.zsz[1];
    const size_t bytes = sizeof(double2) * plane * (size_t)(s->nz + (op == 0 ? 1 : 0));
    return with_device_views(a, bytes, a, bytes, st, [&](const void* di, void* d_o) -> int {
        if (di != d_o) PDO_CUDA(cudaMemcpyAsync(d_o, di, bytes, cudaMemcpyDeviceToDevice, st));
        double2* w = (double2*)d_o;
        if (op == 0) return spectral_dealias_edge(s, w, st);
        if (op == 1) return fft3d_z_inplace(s->ft, w, -1, st);
        if (int rc = fft3d_z_inplace(s->ft, w, +1, st)) return rc;
        const double nf = s->normfactz;
        return launch_ew((long long)plane * s->nz, st, [=] __device__(long long i) { double2 v = w[i]; v.x *= nf; v.y *= nf; w[i] = v; });
    });
}
int pdo_spectral_dealias_edgefield(pdo_spectral_t s, double* fE, void* st) { return spectral_zwise(s, fE, st, 0); }
int pdo_spectral_take_fft1d_z2z_ip(pdo_spectral_t s, double* a, void* st) { return spectral_zwise(s, a, st, 1); }
int pdo_spectral_take_ifft1d_z2z_ip(pdo_spectral_t s, double* a, void* st) { return spectral_zwise(s, a, st, 2); }

// ddz_C2C_complex_inplace (:528-547), shiftz_E2C / shiftz_C2E (:409-437): complex z-pencil arrays of the spectral decomposition
static int spectral_zcomplex(pdo_spectral_t s, double* a, void* stream, int op) {
    if (!s || !a) return fail(PDO_E_BADARG, "null argument");
    if (!s->periodicInZ) return fail(PDO_E_BADARG, "spectral type was not initialised with init_periodicInZ");
    cudaStream_t st = (cudaStream_t)stream;
    const long long cols = (long long)s->si.zsz[0] * s->si.zsz[1];
    const size_t bytes = sizeof(double2) * (size_t)cols * (size_t)s->nz;
    return with_device_views(a, bytes, a, bytes, st, [&](const void* di, void* d_o) -> int {
        if (di != d_o) PDO_CUDA(cudaMemcpyAsync(d_o, di, bytes, cudaMemcpyDeviceToDevice, st));
        double2* w = (double2*)d_o;
        if (op == 0) return zfourier_complex(s, w, ZT_K3_C2C, st);
        if (int rc = spectral_ztables(s)) return rc;
        return ztable_multiply(w, cols, s->nz, ztable(s, op == 1 ? ZT_E2C : ZT_C2E, false), 1.0, st);
    });
}
// test hook (host only, not in the public header): the z-Fourier tables as the kernels get them, out[2][6][nz] complex
}  // extern "C"
namespace pdo { namespace hooks {
int ztables(int nz, double dz, double* out) {
    if (nz < 2 || (nz & 1) || !out) return fail(PDO_E_BADARG, "bad argument");
    std::vector<double2> t = build_ztables_host(nz, dz);
    std::memcpy(out, t.data(), sizeof(double2) * t.size());
    return 0;
}
}}  // namespace pdo::hooks
extern "C" {
int pdo_spectral_ddz_c2c_complex_ip(pdo_spectral_t s, double* a, void* st) { return spectral_zcomplex(s, a, st, 0); }
int pdo_spectral_shiftz_e2c(pdo_spectral_t s, double* a, void* st) { return spectral_zcomplex(s, a, st, 1); }
int pdo_spectral_shiftz_c2e(pdo_spectral_t s, double* a, void* st) { return spectral_zcomplex(s, a, st, 2); }
// ddz_C2C_real_inplace (:507-526): real z-pencil array of the physical decomposition; the oddball mode passes through
int pdo_spectral_ddz_c2c_real_ip(pdo_spectral_t s, double* a, void* stream) {
    if (!s || !a) return fail(PDO_E_BADARG, "null argument");
    if (!s->periodicInZ) return fail(PDO_E_BADARG, "spectral type was not initialised with init_periodicInZ");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = sizeof(double) * (size_t)vol(s->pi.zsz);
    return with_device_views(a, bytes, a, bytes, st, [&](const void* di, void* d_o) -> int {
        return zfourier_real(s, (const double*)di, (double*)d_o, ZT_K3_C2C, st);
    });
}

}  // extern "C"
