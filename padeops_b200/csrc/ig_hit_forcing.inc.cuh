// ig_hit_forcing.inc.cuh — part of igrid.cu: textually included there, ONE translation unit (the sections share file-local helpers).
// forcingmod::HIT_shell_forcing: sparse-DFT shell forcing, C ABI.
// Not a stand-alone header: do not include it anywhere else.

// ================================================================================================
// forcingmod::HIT_shell_forcing (incompressible/forcingIsotropic.F90:45-314)
//
// The reference z-transforms three whole fields, touches Nwaves modes, and inverse-transforms three whole (almost empty)
// fields: 6 transposes + 6 FFT passes per right-hand side for O(Nwaves) numbers.  Here the forcing is evaluated where it
// lives: a direct DFT of the Nwaves columns (x, y) = (kx, ky) at the single wavenumber kz each (O(Nwaves nz) work, one small
// reduction kernel + one allreduce of 3 Nwaves complex numbers when z is distributed), and the inverse transform of a
// single-mode spectrum is the plane wave itself, added to the right-hand side in place by a second small kernel.
// Same arithmetic per mode (den, fac, conjg, the E2C / C2E shifts of w); sums over z instead of an FFT: rounding-level
// differences.
// ================================================================================================
struct pdo_hit_forcing_s {
    pdo_spectral_t spC = nullptr, spE = nullptr;   // borrowed
    double kmin = 2.0, kmax = 10.0, eps = 0.1, normfact = 1.0;
    int nwaves = 0;
    long long seed0 = 0, seed1 = 0, seed2 = 0, seed3 = 0;
    std::vector<int> waves;      // wave_x[n], wave_y[n], wave_z[n]
    bool have_waves = false, waves_dirty = false;   // dirty: the host copy is newer than d_waves
    bool injected = false;       // set_wavenumbers: the next new time step keeps these instead of drawing (A/B runs against the reference's RNG)
    int* d_waves = nullptr;
    double2* d_part = nullptr;   // (U, V, Wraw) per wave
};

namespace {

void hit_update_seeds(pdo_hit_forcing_s* f) {   // :122-128
    auto ab = [](long long v) { return v < 0 ? -v : v; };
    f->seed0 = ab(f->seed0 + 2223345);
    f->seed1 = ab(f->seed0 + 1423246);
    f->seed2 = ab(f->seed0 + 8723446);
    f->seed3 = ab(f->seed0 + 3423444);
}
// `count` doubles in [0, 1): SplitMix64.  Fortran's random_seed(put) / random_number (utilities/random.F90:154-174) is
// compiler-specific, so the stream is this documented generator; the reference's own draw can be
// injected with pdo_hit_forcing_set_wavenumbers.
void hit_uniform(double* out, int count, double left, double right, long long seed) {
    unsigned long long state = (unsigned long long)seed;
    const double diff = right - left;
    for (int i = 0; i < count; ++i) {
        state += 0x9E3779B97F4A7C15ULL;
        unsigned long long z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        double a = (double)(z >> 11) * (1.0 / 9007199254740992.0);
        a = diff * a;          // "array = diff*array; array = array + left"
        out[i] = a + left;
    }
}
void hit_waves_from_samples(pdo_hit_forcing_s* f, const double* kabs, const double* zeta, const double* theta) {   // :137-147
    const int n = f->nwaves;
    for (int i = 0; i < n; ++i) {
        double t = kabs[i] * std::sqrt(1 - zeta[i] * zeta[i]) * std::cos(theta[i]);
        f->waves[i] = (int)std::ceil(std::fabs(t));
        t = kabs[i] * std::sqrt(1 - zeta[i] * zeta[i]) * std::sin(theta[i]);
        f->waves[n + i] = (int)std::ceil(std::fabs(t));
        t = kabs[i] * zeta[i];
        f->waves[2 * n + i] = (int)std::ceil(std::fabs(t));
    }
    f->have_waves = true;
    f->waves_dirty = true;
}

struct HitGeom {
    int nxh, ny, nz;
    int x0, nxl;          // first global x index (0-based) and count of the spectral y-pencil
    int zc0, nzc;         // cell planes of this rank
    int ze0, nze;         // edge planes of this rank (global indices 0 .. nz)
};

__device__ __forceinline__ double2 hit_phase(int kz, int zg, int nz, double sign) {
    const int m = (int)(((long long)kz * zg) % nz);
    double sn, cs;
    sincospi(sign * 2.0 * (double)m / (double)nz, &sn, &cs);
    return make_double2(cs, sn);
}
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// block w: (U, V, Wraw)(w) = sum over the local planes of (u, v, w)_hat(kx, ky, z) e^{-2 pi i kz z / nz}
__global__ void __launch_bounds__(128) hit_reduce_kernel(const double2* __restrict__ uh, const double2* __restrict__ vh, const double2* __restrict__ wh,
                                                         const int* __restrict__ waves, int nwaves, HitGeom g, double2* __restrict__ part) {
    const int w = blockIdx.x;
    const int kx = waves[w], ky = waves[nwaves + w], kz = waves[2 * nwaves + w];
    double2 a[3] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    const int il = kx - g.x0;
    if (il >= 0 && il < g.nxl && ky >= 0 && ky < g.ny && kz >= 0 && kz < g.nz) {
        for (int zl = threadIdx.x; zl < g.nzc; zl += blockDim.x) {
            const double2 ph = hit_phase(kz, g.zc0 + zl, g.nz, -1.0);
            const size_t idx = (size_t)il + (size_t)g.nxl * ((size_t)ky + (size_t)g.ny * zl);
            const double2 pu = cmul(uh[idx], ph), pv = cmul(vh[idx], ph);
            a[0].x += pu.x; a[0].y += pu.y; a[1].x += pv.x; a[1].y += pv.y;
        }
        for (int zl = threadIdx.x; zl < g.nze; zl += blockDim.x) {
            const int zg = g.ze0 + zl;
            if (zg >= g.nz) continue;      // "this%what = this%cbuffzE(:,:,1:nz)"
            const double2 pw = cmul(wh[(size_t)il + (size_t)g.nxl * ((size_t)ky + (size_t)g.ny * zl)], hit_phase(kz, zg, g.nz, -1.0));
            a[2].x += pw.x; a[2].y += pw.y;
        }
    }
    __shared__ double sm[128][6];
    for (int c = 0; c < 3; ++c) { sm[threadIdx.x][2 * c] = a[c].x; sm[threadIdx.x][2 * c + 1] = a[c].y; }
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if (threadIdx.x < s) for (int c = 0; c < 6; ++c) sm[threadIdx.x][c] += sm[threadIdx.x + s][c];
        __syncthreads();
    }
    if (threadIdx.x < 3) part[3 * w + threadIdx.x] = make_double2(sm[0][2 * threadIdx.x], sm[0][2 * threadIdx.x + 1]);
}

// thread t: cell plane t and edge plane t of this rank; the waves are applied one after another (embed_forcing_mode :215-251,
// then the inverse z transform of each single-mode spectrum: normfactz x the plane wave)
__global__ void __launch_bounds__(128) hit_apply_kernel(double2* __restrict__ ur, double2* __restrict__ vr, double2* __restrict__ wr,
                                                        const int* __restrict__ waves, int nwaves, HitGeom g, const double2* __restrict__ part,
                                                        const double2* __restrict__ e2c, const double2* __restrict__ c2e, double normfact,
                                                        double eps) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool doC = t < g.nzc, doE = t < g.nze;
    if (!doC && !doE) return;
    const double nwr = (double)nwaves, nfz = 1.0 / (double)g.nz;
    for (int w = 0; w < nwaves; ++w) {
        const int kx = waves[w], ky = waves[nwaves + w], kz = waves[2 * nwaves + w];
        const int il = kx - g.x0;
        if (!(il >= 0 && il < g.nxl && ky >= 0 && ky < g.ny && kz >= 0 && kz < g.nz)) continue;
        const double2 U = part[3 * w], V = part[3 * w + 1], W = cmul(part[3 * w + 2], e2c[kz]);   // shiftz_E2C(what)
        const double den = (U.x * U.x + U.y * U.y) + (V.x * V.x + V.y * V.y) + (W.x * W.x + W.y * W.y) + 1.0e-14;
        const double fac = normfact * eps / den / nwr;
        if (doC) {
            const double2 ph = hit_phase(kz, g.zc0 + t, g.nz, +1.0);
            const size_t idx = (size_t)il + (size_t)g.nxl * ((size_t)ky + (size_t)g.ny * t);
            const double2 fu = cmul(make_double2(fac * U.x * nfz, -fac * U.y * nfz), ph);
            const double2 fv = cmul(make_double2(fac * V.x * nfz, -fac * V.y * nfz), ph);
            double2 a = ur[idx]; a.x += fu.x; a.y += fu.y; ur[idx] = a;
            double2 b = vr[idx]; b.x += fv.x; b.y += fv.y; vr[idx] = b;
        }
        if (doE) {
            const int zg = g.ze0 + t;
            const double2 ph = hit_phase(kz, zg >= g.nz ? 0 : zg, g.nz, +1.0);     // plane nz+1 := plane 1
            const double2 fz = cmul(make_double2(fac * W.x, -fac * W.y), c2e[kz]);  // shiftz_C2E(fzhat)
            const double2 fw = cmul(make_double2(fz.x * nfz, fz.y * nfz), ph);
            const size_t idx = (size_t)il + (size_t)g.nxl * ((size_t)ky + (size_t)g.ny * t);
            double2 a = wr[idx]; a.x += fw.x; a.y += fw.y; wr[idx] = a;
        }
    }
}

// device pointers, y-pencils of the spectral decompositions
int hit_get_rhs_dev(pdo_hit_forcing_s* f, double2* ur, double2* vr, double2* wr, const double2* uh, const double2* vh, const double2* wh,
                    bool new_timestep, cudaStream_t st) {
    const int n = f->nwaves;
    if (new_timestep) {     // :265-268
        if (!f->injected) {
            std::vector<double> a(3 * (size_t)n);
            hit_uniform(a.data(), n, f->kmin, f->kmax, f->seed1);
            hit_uniform(a.data() + n, n, -1.0, 1.0, f->seed2);
            hit_uniform(a.data() + 2 * n, n, 0.0, 2.0 * kPi, f->seed3);
            hit_waves_from_samples(f, a.data(), a.data() + n, a.data() + 2 * n);
        }
        f->injected = false;
        hit_update_seeds(f);
    }
    if (!f->have_waves) return fail(PDO_E_BADARG, "HIT forcing: no wavenumbers yet (newTimestep was never true and none were set)");
    if (f->waves_dirty) {   // once per time step: stream-ordered copy from pageable memory (staged before the call returns)
        PDO_CUDA(cudaMemcpyAsync(f->d_waves, f->waves.data(), sizeof(int) * 3 * n, cudaMemcpyHostToDevice, st));
        f->waves_dirty = false;
    }
    if (int rc = spectral_ztables(f->spC)) return rc;
    pdo_spectral_s *C = f->spC, *E = f->spE;
    HitGeom g;
    g.nxh = C->nxh; g.ny = C->ny; g.nz = C->nz;
    g.x0 = C->si.yst[0] - 1; g.nxl = C->si.ysz[0];
    g.zc0 = C->si.yst[2] - 1; g.nzc = C->si.ysz[2];
    g.ze0 = E->si.yst[2] - 1; g.nze = E->si.ysz[2];
    hit_reduce_kernel<<<n, 128, 0, st>>>(uh, vh, wh, f->d_waves, n, g, f->d_part);
    PDO_CUDA(cudaGetLastError());
    if (int rc = comm_allreduce_sum((double*)f->d_part, 6 * n, st)) return rc;
    const int planes = g.nzc > g.nze ? g.nzc : g.nze;
    hit_apply_kernel<<<(planes + 127) / 128, 128, 0, st>>>(ur, vr, wr, f->d_waves, n, g, f->d_part, ztable(C, ZT_E2C, false), ztable(C, ZT_C2E, false),
                                                            f->normfact, f->eps);
    PDO_CUDA(cudaGetLastError());
    g_launches += 2;
    return 0;
}

}  // namespace

extern "C" {

/* hitforce%init(inputfile, sp_gpC, sp_gpE, spectC, ...) :45-120; the &HIT_Forcing namelist enters as arguments */
int pdo_hit_forcing_init(pdo_hit_forcing_t* h, pdo_spectral_t spectC, pdo_spectral_t spectE, double kmin, double kmax, int nwaves,
                         double eps_amplitude, int tid_start, int rand_seed_to_add) {
    if (!h || !spectC || !spectE) return fail(PDO_E_BADARG, "null argument");
    *h = nullptr;
    if (nwaves < 1) return fail(PDO_E_BADARG, "HIT forcing: Nwaves must be positive");
    if (!spectC->periodicInZ) return fail(PDO_E_BADARG, "HIT forcing needs a spectral type initialised with init_periodicInZ");
    if (spectE->nz != spectC->nz + 1 || spectE->nx != spectC->nx || spectE->ny != spectC->ny || spectE->si.yst[0] != spectC->si.yst[0] ||
        spectE->si.ysz[0] != spectC->si.ysz[0])
        return fail(PDO_E_BADARG, "spectE must be the (nx, ny, nz+1) edge type of spectC on the same process grid");
    pdo_hit_forcing_s* f = new (std::nothrow) pdo_hit_forcing_s();
    if (!f) return fail(PDO_E_BADARG, "out of memory");
    f->spC = spectC; f->spE = spectE;
    f->kmin = kmin; f->kmax = kmax; f->eps = eps_amplitude; f->nwaves = nwaves;
    f->seed0 = (long long)tid_start + (long long)rand_seed_to_add;   // :86
    hit_update_seeds(f);
    const double n3 = (double)spectC->nx * (double)spectC->ny * (double)spectC->nz;
    f->normfact = n3 * n3;                                             // :89
    f->waves.assign(3 * (size_t)nwaves, 0);
    cudaError_t e = cudaMalloc(&f->d_waves, sizeof(int) * 3 * nwaves);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_part, sizeof(double2) * 3 * nwaves);
    if (e != cudaSuccess) { pdo_hit_forcing_destroy(f); return fail(PDO_E_CUDA, "HIT forcing init: %s", cudaGetErrorString(e)); }
    *h = f;
    return 0;
}
int pdo_hit_forcing_destroy(pdo_hit_forcing_t f) {
    if (!f) return 0;
    if (f->d_waves) cudaFree(f->d_waves);
    if (f->d_part) cudaFree(f->d_part);
    delete f;
    return 0;
}
/* the draw of the CURRENT (or the next new) time step, e.g. the reference RNG's wave_x / wave_y / wave_z for an A/B run: the next
   call with newTimestep keeps these instead of drawing (its seeds still advance), later steps draw again */
int pdo_hit_forcing_set_wavenumbers(pdo_hit_forcing_t f, const int* wave_x, const int* wave_y, const int* wave_z) {
    if (!f || !wave_x || !wave_y || !wave_z) return fail(PDO_E_BADARG, "null argument");
    const int n = f->nwaves;
    for (int i = 0; i < n; ++i) { f->waves[i] = wave_x[i]; f->waves[n + i] = wave_y[i]; f->waves[2 * n + i] = wave_z[i]; }
    f->have_waves = true;
    f->waves_dirty = true;
    f->injected = true;
    return 0;
}
int pdo_hit_forcing_get_wavenumbers(pdo_hit_forcing_t f, int* wave_x, int* wave_y, int* wave_z) {
    if (!f || !wave_x || !wave_y || !wave_z) return fail(PDO_E_BADARG, "null argument");
    const int n = f->nwaves;
    for (int i = 0; i < n; ++i) { wave_x[i] = f->waves[i]; wave_y[i] = f->waves[n + i]; wave_z[i] = f->waves[2 * n + i]; }
    return 0;
}
/* getRHS_HITforcing(urhs_xy, vrhs_xy, wrhs_xy, uhat_xy, vhat_xy, what_xy, newTimestep) :254-311; DEVICE pointers (y-pencils) */
int pdo_hit_forcing_get_rhs(pdo_hit_forcing_t f, double* urhs, double* vrhs, double* wrhs, const double* uhat, const double* vhat,
                            const double* what, int new_timestep, void* stream) {
    if (!f || !urhs || !vrhs || !wrhs || !uhat || !vhat || !what) return fail(PDO_E_BADARG, "null argument");
    for (const void* p : {(const void*)urhs, (const void*)vrhs, (const void*)wrhs, (const void*)uhat, (const void*)vhat, (const void*)what})
        if (!is_device_ptr(p)) return fail(PDO_E_BADARG, "HIT forcing works on device-resident right-hand sides");
    return hit_get_rhs_dev(f, (double2*)urhs, (double2*)vrhs, (double2*)wrhs, (const double2*)uhat, (const double2*)vhat, (const double2*)what,
                           new_timestep != 0, (cudaStream_t)stream);
}
/* test hook (host only, not in the public header): seeds after init + `updates` further update_seeds, the draw for the current seeds */
}  // extern "C"
namespace pdo { namespace hooks {
int hit_draw(double kmin, double kmax, int nwaves, int tid_start, int rand_seed_to_add, int updates, long long seeds[4], int* wx,
                       int* wy, int* wz) {
    pdo_hit_forcing_s f;
    f.kmin = kmin; f.kmax = kmax; f.nwaves = nwaves;
    f.seed0 = (long long)tid_start + rand_seed_to_add;
    hit_update_seeds(&f);
    for (int i = 0; i < updates; ++i) hit_update_seeds(&f);
    f.waves.assign(3 * (size_t)nwaves, 0);
    std::vector<double> a(3 * (size_t)nwaves);
    hit_uniform(a.data(), nwaves, kmin, kmax, f.seed1);
    hit_uniform(a.data() + nwaves, nwaves, -1.0, 1.0, f.seed2);
    hit_uniform(a.data() + 2 * nwaves, nwaves, 0.0, 2.0 * kPi, f.seed3);
    hit_waves_from_samples(&f, a.data(), a.data() + nwaves, a.data() + 2 * nwaves);
    seeds[0] = f.seed0; seeds[1] = f.seed1; seeds[2] = f.seed2; seeds[3] = f.seed3;
    for (int i = 0; i < nwaves; ++i) { wx[i] = f.waves[i]; wy[i] = f.waves[nwaves + i]; wz[i] = f.waves[2 * nwaves + i]; }
    return 0;
}
}}  // namespace pdo::hooks
extern "C" {

}  // extern "C"
