// ig_ops_periodic.inc.cuh — part of igrid.cu: textually included there, ONE translation unit (the sections share file-local helpers).
// igrid_Operators_Periodic::Ops_Periodic, C ABI.
// Not a stand-alone header: do not include it anywhere else.

// ================================================================================================
// igrid_Operators_Periodic::Ops_Periodic (igrid_operators_periodic.F90:13-161): Fourier operators on x-pencil fields of a
// triply periodic box — compositions of the spectral type's transforms, its pointwise passes and PoissonPeriodic
// ================================================================================================
struct pdo_ops_periodic_s {
    pdo_spectral_t spect = nullptr;
    pdo_poisson_t poiss = nullptr;
    double2* cbuffy1 = nullptr;                     // spectral y-pencil
    double *rbuffy = nullptr, *rbuffz1 = nullptr;   // physical y- / z-pencils (allocated only where they differ from the x- / y-pencil)
};

extern "C" {

/* init(nx, ny, nz, dx, dy, dz, gp, InputDir, OutputDir) :86-109; gp enters as its process grid (0, 0 = 1 x nproc) */
int pdo_ops_periodic_init(pdo_ops_periodic_t* h, int nx, int ny, int nz, double dx, double dy, double dz, int p_row, int p_col) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    pdo_ops_periodic_s* o = new (std::nothrow) pdo_ops_periodic_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    // spect%init("x", nx, ny, nz, dx, dy, dz, "four", "2/3rd", 2, fixOddball=.false., init_periodicInZ=.TRUE., dealiasF=2/3)  :94-95
    int rc = pdo_spectral_init(&o->spect, nx, ny, nz, dx, dy, dz, p_row, p_col, 0, 1, 2.0 / 3.0);
    // poiss%init(dx, dy, dz, gp, 1, .true., GetKmod_Fourier x 3): the spectral wavenumbers themselves  :107-108
    if (!rc) rc = pdo_poisson_init(&o->poiss, nx, ny, nz, dx, dy, dz, o->spect->p_row, o->spect->p_col, 1, nullptr, nullptr, nullptr);
    if (!rc) {
        pdo_spectral_s* s = o->spect;
        cudaError_t e = cudaMalloc(&o->cbuffy1, sizeof(double2) * (size_t)vol(s->si.ysz));
        if (e == cudaSuccess && s->p_row > 1) e = cudaMalloc(&o->rbuffy, sizeof(double) * (size_t)vol(s->pi.ysz));
        if (e == cudaSuccess && s->p_col > 1) e = cudaMalloc(&o->rbuffz1, sizeof(double) * (size_t)vol(s->pi.zsz));
        if (e != cudaSuccess) rc = fail(PDO_E_CUDA, "Ops_Periodic buffers: %s", cudaGetErrorString(e));
    }
    if (rc) { pdo_ops_periodic_destroy(o); return rc; }
    *h = o;
    return 0;
}
int pdo_ops_periodic_destroy(pdo_ops_periodic_t o) {
    if (!o) return 0;
    if (o->cbuffy1) cudaFree(o->cbuffy1);
    if (o->rbuffy) cudaFree(o->rbuffy);
    if (o->rbuffz1) cudaFree(o->rbuffz1);
    pdo_poisson_destroy(o->poiss);
    pdo_spectral_destroy(o->spect);
    delete o;
    return 0;
}
/* link_spect :46-52 */
pdo_spectral_t pdo_ops_periodic_spect(pdo_ops_periodic_t o) { return o ? o->spect : nullptr; }

// ddx :117-125, ddy :127-135 (which = 1, 2), dealiasField :56-62 (which = 0): fft, one pointwise pass, ifft
static int ops_periodic_xy(pdo_ops_periodic_t o, int which, const double* f, double* out, void* stream) {
    if (!o || !f || !out) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    pdo_spectral_s* s = o->spect;
    const size_t bytes = sizeof(double) * (size_t)vol(s->pi.xsz);
    return with_device_views(f, bytes, out, bytes, st, [&](const void* di, void* d_o) -> int {
        if (int rc = fft3d_forward_xy(s->ft, (const double*)di, o->cbuffy1, st)) return rc;
        if (which == 0) { if (int rc = spectral_dealias(s, o->cbuffy1, st)) return rc; }
        else if (int rc = spectral_mtimes(s, which, o->cbuffy1, o->cbuffy1, st)) return rc;
        return fft3d_backward_yx(s->ft, o->cbuffy1, (double*)d_o, false, st);
    });
}
int pdo_ops_periodic_ddx(pdo_ops_periodic_t o, const double* f, double* dfdx, void* st) { return ops_periodic_xy(o, 1, f, dfdx, st); }
int pdo_ops_periodic_ddy(pdo_ops_periodic_t o, const double* f, double* dfdy, void* st) { return ops_periodic_xy(o, 2, f, dfdy, st); }
int pdo_ops_periodic_dealias_field(pdo_ops_periodic_t o, double* f, void* st) { return ops_periodic_xy(o, 0, f, f, st); }

/* ddz :149-160: x -> y -> z, spect%ddz_C2C_real_inplace, z -> y -> x (a transpose inside a 1-rank group is the identity and is skipped) */
int pdo_ops_periodic_ddz(pdo_ops_periodic_t o, const double* f, double* dfdz, void* stream) {
    if (!o || !f || !dfdz) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    pdo_spectral_s* s = o->spect;
    pdo_decomp_t gp = fft3d_phys_decomp(s->ft);
    const size_t bytes = sizeof(double) * (size_t)vol(s->pi.xsz);
    return with_device_views(f, bytes, dfdz, bytes, st, [&](const void* di, void* d_o) -> int {
        const double* a = (const double*)di;
        double* out = (double*)d_o;
        const bool tx = s->p_row > 1, tz = s->p_col > 1;
        double* ydst = tx ? o->rbuffy : out;          // where the y-pencil result lives
        if (tx) { if (int rc = decomp_transpose_device(gp, 0, a, o->rbuffy, 1, st)) return rc; a = o->rbuffy; }
        if (tz) {
            if (int rc = decomp_transpose_device(gp, 2, a, o->rbuffz1, 1, st)) return rc;
            if (int rc = zfourier_real(s, o->rbuffz1, o->rbuffz1, ZT_K3_C2C, st)) return rc;
            if (int rc = decomp_transpose_device(gp, 3, o->rbuffz1, ydst, 1, st)) return rc;
        } else if (int rc = zfourier_real(s, a, ydst, ZT_K3_C2C, st)) return rc;
        if (tx) return decomp_transpose_device(gp, 1, o->rbuffy, out, 1, st);
        return 0;
    });
}
/* ddz_cmplx2cmplx :137-145: complex y-pencil of the spectral decomposition, in place */
int pdo_ops_periodic_ddz_cmplx2cmplx(pdo_ops_periodic_t o, double* fhat, void* stream) {
    if (!o || !fhat) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    pdo_spectral_s* s = o->spect;
    pdo_decomp_t spec = fft3d_spec_decomp(s->ft);
    const size_t bytes = sizeof(double2) * (size_t)vol(s->si.ysz);
    return with_device_views(fhat, bytes, fhat, bytes, st, [&](const void* di, void* d_o) -> int {
        if (di != d_o) PDO_CUDA(cudaMemcpyAsync(d_o, di, bytes, cudaMemcpyDeviceToDevice, st));
        double2* w = (double2*)d_o;
        if (s->p_col == 1) return zfourier_complex(s, w, ZT_K3_C2C, st);
        if (int rc = decomp_transpose_device(spec, 2, (const double*)w, (double*)s->ctmpz, 2, st)) return rc;
        if (int rc = zfourier_complex(s, s->ctmpz, ZT_K3_C2C, st)) return rc;
        return decomp_transpose_device(spec, 3, (const double*)s->ctmpz, (double*)w, 2, st);
    });
}
/* ReadField3D :162-187 / WriteField3D :189-205: "<dir>/Run<runID>_<label>_t<tidx>.out", x-pencil of gp through decomp_2d_io */
static std::string ops_periodic_fname(const char* dir, const char* label4, int tidx, int run_id) {
    char name[64];
    std::snprintf(name, sizeof(name), "Run%02d_%.4s_t%06d.out", run_id, label4, tidx);   // "(A3,I2.2,A1,A4,A2,I6.6,A4)"
    return std::string(dir ? dir : ".") + "/" + name;
}
int pdo_ops_periodic_write_field3d(pdo_ops_periodic_t o, const double* field, const char* label4, int tidx, int run_id, const char* outputdir) {
    if (!o || !field || !label4) return fail(PDO_E_BADARG, "null argument");
    return pdo_decomp_write_one(fft3d_phys_decomp(o->spect->ft), 1, field, 1, ops_periodic_fname(outputdir, label4, tidx, run_id).c_str());
}
int pdo_ops_periodic_read_field3d(pdo_ops_periodic_t o, double* field, const char* label4, int tidx, int run_id, const char* inputdir) {
    if (!o || !field || !label4) return fail(PDO_E_BADARG, "null argument");
    return pdo_decomp_read_one(fft3d_phys_decomp(o->spect->ft), 1, field, 1, ops_periodic_fname(inputdir, label4, tidx, run_id).c_str());   // missing file -> 321
}
/* SolvePoisson_oop :70-76 (p != rhs), SolvePoisson_ip :78-84 (p == rhs) */
int pdo_ops_periodic_solve_poisson(pdo_ops_periodic_t o, const double* rhs, double* p, void* stream) {
    if (!o) return fail(PDO_E_BADARG, "null handle");
    return pdo_poisson_solve(o->poiss, rhs, p, stream);
}

}  // extern "C"
