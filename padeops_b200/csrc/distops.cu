// distops.cu — vector calculus on decomposed fields (the reference's utilities/operators.F90:17-151: gradient, curl,
// divergence of y-pencil fields) and the z-slab distributed compact solve that makes them cheap on NVLink.
//
// The reference differentiates along an axis that is split across ranks by transposing the whole field to the pencil where
// that axis is local, sweeping, and transposing back (operators.F90:43-51): 2 x the field over the network per derivative.
// Here, when the z axis is split and the slabs are even multiples of the chunk length, the compact solve itself is
// distributed: neighbouring GPUs exchange 2 x halo_rows planes of f and the reduced-system pieces of halo_chunks chunks
// per line (banded.cuh, z-slab mode) — for CD10 on 1024 planes per GPU that is about 5 % of the field — straight into
// each other's memory over NVLink (CUDA IPC), with in-stream epoch flags instead of host synchronisation.  Everything
// else (x split across ranks, uneven slabs, no peer access) takes the reference's transpose choreography on
// decomp.cu's transposes.  Results of the two paths agree to rounding (same per-chunk arithmetic, same tables).
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "banded.cuh"
#include "common.cuh"
#include "handles.h"
#include "spectral_internal.cuh"

using namespace pdo;

namespace {

__device__ __forceinline__ void flag_store_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long flag_load_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// my rows [0, HB) -> the lower GPU's "rows above the slab"; my rows [n_local-HB, n_local) -> the upper GPU's "rows below"
__global__ void __launch_bounds__(256) zslab_push_planes_kernel(const double* __restrict__ f, long long n1, int n_local, int HB,
                                                                double* __restrict__ lower_planes, double* __restrict__ upper_planes) {
    const long long tot = (long long)HB * n1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        lower_planes[tot + i] = f[i];                                  // upper half of the lower GPU's [2*HB][n1]
        upper_planes[i] = f[(long long)(n_local - HB) * n1 + i];       // lower half of the upper GPU's
    }
}
// runs after the producing kernel in the same stream: publishes epoch e to two peers' flag slots
__global__ void zslab_signal_kernel(unsigned long long* a, unsigned long long* b, unsigned long long e) {
    __threadfence_system();
    if (threadIdx.x == 0) flag_store_sys(a, e);
    if (threadIdx.x == 1) flag_store_sys(b, e);
}
// holds the stream until both of my flag slots have reached epoch e; a peer that never arrives traps instead of hanging
__global__ void zslab_wait_kernel(const unsigned long long* a, const unsigned long long* b, unsigned long long e) {
    const unsigned long long* p = threadIdx.x == 0 ? a : b;
    const unsigned long long t0 = global_ns();
    while (flag_load_sys(p) < e) {
        if (global_ns() - t0 > 20000000000ull) __trap();  // 20 s
    }
}

__global__ void __launch_bounds__(256) add_kernel(double* __restrict__ a, const double* __restrict__ b, long long n, double sign) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = a[i] + sign * b[i];
}

unsigned grid_for(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = 148LL * 16;
    return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

// Layout of one rank's symmetric z-slab buffer (identical on all ranks of a z-group):
//   flags[8] (u64: 0 planes<-lower, 1 planes<-upper, 2 g<-lower, 3 g<-upper), then per parity b in {0,1}:
//   planes[2*HB][n1], glo[2*BW*HW][n1], ghi[2*BW*HW][n1]
struct ZSlab {
    bool on = false;
    int HB = 0, HW = 0, BW = 0, n_local = 0;
    long long n1 = 0;
    char* base = nullptr;
    char* lower = nullptr;  // the lower / upper GPU's copy of the same buffer (peer memory)
    char* upper = nullptr;
    unsigned long long epoch = 0;
    size_t bytes = 0;
    size_t off_planes(int b) const { return 64 + (size_t)b * per_parity(); }
    size_t off_glo(int b) const { return off_planes(b) + sizeof(double) * 2 * HB * (size_t)n1; }
    size_t off_ghi(int b) const { return off_glo(b) + sizeof(double) * 2 * BW * HW * (size_t)n1; }
    size_t per_parity() const { return sizeof(double) * (size_t)n1 * (2 * HB + 4 * BW * HW); }
    size_t total() const { return 64 + 2 * per_parity(); }
};

// Phase 1 (exchange): push my edge planes, wait for the neighbours', compute my edge pieces into the neighbours' buffers.
// Phase 2 (solve): wait for the neighbours' edge pieces, run the fused solve.  The phases may sit on different streams
// (phase 2 ordered after phase 1 by the caller) so that the exchange overlaps other derivative kernels.
int zslab_phase1(ZSlab& z, const BandedOp* op, const double* f, cudaStream_t st) {
    const unsigned long long e = ++z.epoch;
    const int b = (int)(e & 1ull);
    auto U64 = [](char* p, int i) { return reinterpret_cast<unsigned long long*>(p) + i; };
    auto D = [](char* p, size_t off) { return reinterpret_cast<double*>(p + off); };
    zslab_push_planes_kernel<<<grid_for((long long)z.HB * z.n1), 256, 0, st>>>(f, z.n1, z.n_local, z.HB, D(z.lower, z.off_planes(b)),
                                                                             D(z.upper, z.off_planes(b)));
    zslab_signal_kernel<<<1, 32, 0, st>>>(U64(z.lower, 1), U64(z.upper, 0), e);   // I am the lower GPU's upper neighbour
    zslab_wait_kernel<<<1, 2, 0, st>>>(U64(z.base, 0), U64(z.base, 1), e);
    PDO_CUDA(cudaGetLastError());
    PDO_CUDA(banded_zslab_edges(op, f, z.n1, z.n_local, D(z.base, z.off_planes(b)), D(z.lower, z.off_ghi(b)), D(z.upper, z.off_glo(b)), st));
    zslab_signal_kernel<<<1, 32, 0, st>>>(U64(z.lower, 3), U64(z.upper, 2), e);
    PDO_CUDA(cudaGetLastError());
    g_launches += 5;
    return 0;
}
int zslab_phase2(ZSlab& z, const BandedOp* op, const double* f, double* out, cudaStream_t st) {
    const unsigned long long e = z.epoch;
    const int b = (int)(e & 1ull);
    auto U64 = [](char* p, int i) { return reinterpret_cast<unsigned long long*>(p) + i; };
    auto D = [](char* p, size_t off) { return reinterpret_cast<double*>(p + off); };
    zslab_wait_kernel<<<1, 2, 0, st>>>(U64(z.base, 2), U64(z.base, 3), e);
    PDO_CUDA(cudaGetLastError());
    PDO_CUDA(banded_zslab_apply(op, f, out, z.n1, z.n_local, D(z.base, z.off_planes(b)), D(z.base, z.off_glo(b)), D(z.base, z.off_ghi(b)), st));
    g_launches += 2;
    return 0;
}
int zslab_run(ZSlab& z, const BandedOp* op, const double* f, double* out, cudaStream_t st) {
    if (int rc = zslab_phase1(z, op, f, st)) return rc;
    return zslab_phase2(z, op, f, out, st);
}

}  // namespace

struct pdo_operators_s {
    pdo_decomp_t gp = nullptr;
    pdo_decomp_info info{};
    int p_row = 1, p_col = 1, c1 = 0, c2 = 0;
    int method = 0;  // 0 cd10, 1 cd06
    pdo_cd10_t c10[3] = {nullptr, nullptr, nullptr};
    pdo_cd06_t c06[3] = {nullptr, nullptr, nullptr};
    ZSlab zs;
    double *xtmp = nullptr, *xdum = nullptr, *ztmp = nullptr, *zdum = nullptr, *ytmp = nullptr, *ytmp2 = nullptr;  // lazily allocated work pencils
    void* hbuf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // device copies of HOST arguments (grow-only)
    size_t hcap[6] = {0, 0, 0, 0, 0, 0};
    cudaStream_t side = nullptr;             // z-slab exchange runs here while the caller's stream differentiates along x / y
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    const BandedOp* op(int axis) const { return method == 0 ? &c10[axis]->d1 : &c06[axis]->d1; }
};

namespace {

inline long long vol3(const int* s) { return (long long)s[0] * s[1] * s[2]; }

int ensure(double** p, long long doubles) {
    if (*p) return 0;
    PDO_CUDA(cudaMalloc(p, sizeof(double) * (size_t)doubles));
    return 0;
}

// derivative of a y-pencil field along `axis`, result in the y-pencil
int ops_dd(pdo_operators_s* o, int axis, const double* f, double* out, cudaStream_t st) {
    const int* ys = o->info.ysz;
    if (axis == 1) {
        PDO_CUDA(banded_op_apply(o->op(1), 1, f, out, ys[0], ys[2], st));
        g_launches += 1;
        return 0;
    }
    if (axis == 0) {
        if (o->p_row == 1) {  // x- and y-pencils coincide
            PDO_CUDA(banded_op_apply(o->op(0), 0, f, out, ys[1], ys[2], st));
            g_launches += 1;
            return 0;
        }
        const int* xs = o->info.xsz;
        if (int rc = ensure(&o->xtmp, vol3(xs))) return rc;
        if (int rc = ensure(&o->xdum, vol3(xs))) return rc;
        if (int rc = decomp_transpose_device(o->gp, 1, f, o->xtmp, 1, st)) return rc;       // operators.F90:43-46
        PDO_CUDA(banded_op_apply(o->op(0), 0, o->xtmp, o->xdum, xs[1], xs[2], st));
        g_launches += 1;
        return decomp_transpose_device(o->gp, 0, o->xdum, out, 1, st);
    }
    if (o->p_col == 1) {  // y- and z-pencils coincide
        PDO_CUDA(banded_op_apply(o->op(2), 2, f, out, ys[0], ys[1], st));
        g_launches += 1;
        return 0;
    }
    if (o->zs.on) return zslab_run(o->zs, o->op(2), f, out, st);
    const int* zs = o->info.zsz;
    if (int rc = ensure(&o->ztmp, vol3(zs))) return rc;
    if (int rc = ensure(&o->zdum, vol3(zs))) return rc;
    if (int rc = decomp_transpose_device(o->gp, 2, f, o->ztmp, 1, st)) return rc;           // operators.F90:48-51
    PDO_CUDA(banded_op_apply(o->op(2), 2, o->ztmp, o->zdum, zs[0], zs[1], st));
    g_launches += 1;
    return decomp_transpose_device(o->gp, 3, o->zdum, out, 1, st);
}

// z-derivative split around other work: zbegin starts the exchange for `f` on the side stream (ordered after everything
// already in `st`), zend finishes on `st`.  Without the z-slab mode zbegin does nothing and zend is the plain derivative.
int ops_zbegin(pdo_operators_s* o, const double* f, cudaStream_t st) {
    if (!(o->p_col > 1 && o->zs.on)) return 0;
    if (!o->side) {
        PDO_CUDA(cudaStreamCreateWithFlags(&o->side, cudaStreamNonBlocking));
        PDO_CUDA(cudaEventCreateWithFlags(&o->ev_fork, cudaEventDisableTiming));
        PDO_CUDA(cudaEventCreateWithFlags(&o->ev_join, cudaEventDisableTiming));
    }
    PDO_CUDA(cudaEventRecord(o->ev_fork, st));
    PDO_CUDA(cudaStreamWaitEvent(o->side, o->ev_fork, 0));
    if (int rc = zslab_phase1(o->zs, o->op(2), f, o->side)) return rc;
    PDO_CUDA(cudaEventRecord(o->ev_join, o->side));
    return 0;
}
int ops_zend(pdo_operators_s* o, const double* f, double* out, cudaStream_t st) {
    if (!(o->p_col > 1 && o->zs.on)) return ops_dd(o, 2, f, out, st);
    PDO_CUDA(cudaStreamWaitEvent(st, o->ev_join, 0));
    return zslab_phase2(o->zs, o->op(2), f, out, st);
}

int ops_add(double* a, const double* b, long long n, double sign, cudaStream_t st) {
    add_kernel<<<grid_for(n), 256, 0, st>>>(a, b, n, sign);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return 0;
}

// Arguments that arrive as HOST arrays (the unmodified Fortran caller) are staged through per-handle device
// buffers: inputs copied in before `body`, outputs copied back after it, then one stream synchronise.  Device (or managed)
// pointers pass straight through, stream-ordered, nothing synchronises.
struct OpsArg { const void* ptr; size_t bytes; bool out; bool inout; void* dev; };
template <class Body>
int with_ops_args(pdo_operators_s* o, OpsArg* a, int na, cudaStream_t st, Body body) {
    bool any_host = false;
    for (int i = 0; i < na; ++i) {
        if (!a[i].ptr) return fail(PDO_E_BADARG, "null argument");
        if (is_device_ptr(a[i].ptr)) { a[i].dev = const_cast<void*>(a[i].ptr); continue; }
        any_host = true;
        if (o->hcap[i] < a[i].bytes) {
            if (o->hbuf[i]) cudaFree(o->hbuf[i]);
            o->hbuf[i] = nullptr; o->hcap[i] = 0;
            PDO_CUDA(cudaMalloc(&o->hbuf[i], a[i].bytes));
            o->hcap[i] = a[i].bytes;
        }
        a[i].dev = o->hbuf[i];
        if (!a[i].out || a[i].inout) PDO_CUDA(cudaMemcpyAsync(a[i].dev, a[i].ptr, a[i].bytes, cudaMemcpyHostToDevice, st));
    }
    if (int rc = body()) return rc;
    if (!any_host) return 0;
    for (int i = 0; i < na; ++i)
        if (a[i].out && a[i].dev != a[i].ptr) PDO_CUDA(cudaMemcpyAsync(const_cast<void*>(a[i].ptr), a[i].dev, a[i].bytes, cudaMemcpyDeviceToHost, st));
    PDO_CUDA(cudaStreamSynchronize(st));
    return 0;
}
#define OPS_D(i) ((double*)a[i].dev)

}  // namespace

extern "C" {

int pdo_operators_init(pdo_operators_t* h, pdo_decomp_t gp, double dx, double dy, double dz, const char* method, int allow_zslab) {
    if (!h || !gp || !method) return fail(PDO_E_BADARG, "null argument");
    *h = nullptr;
    pdo_operators_s* o = new (std::nothrow) pdo_operators_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    o->gp = gp;
    pdo_decomp_get_info(gp, &o->info);
    decomp_grid(gp, &o->p_row, &o->p_col, &o->c1, &o->c2);
    const int n[3] = {o->info.xsz[0], o->info.ysz[1], o->info.zsz[2]};   // global line lengths
    const double d[3] = {dx, dy, dz};
    int rc = 0;
    if (std::strcmp(method, "cd10") == 0) o->method = 0;
    else if (std::strcmp(method, "cd06") == 0) o->method = 1;
    else rc = fail(PDO_E_UNSUPPORTED, "operators: method '%s' is out of scope (cd10, cd06)", method);
    for (int a = 0; a < 3 && !rc; ++a)
        rc = o->method == 0 ? pdo_cd10_init(&o->c10[a], n[a], d[a], 1, 0, 0) : pdo_cd06_init(&o->c06[a], n[a], d[a], 1, 0, 0);
    if (rc) { pdo_operators_destroy(o); return rc; }
    // z-slab mode: every rank of the z-group must hold the same number of planes (collective decision: sizes are global
    // knowledge, peer access is agreed inside comm_sym_alloc)
    if (allow_zslab && o->p_col > 1 && n[2] % o->p_col == 0) {
        const int n_local = n[2] / o->p_col;
        const BandedOp* op = o->op(2);
        const int HW = banded_zslab_halo_chunks(op, n_local);
        if (HW > 0) {
            ZSlab& z = o->zs;
            z.HB = banded_zslab_halo_rows(op); z.HW = HW; z.BW = op->bw; z.n_local = n_local;
            z.n1 = (long long)o->info.ysz[0] * o->info.ysz[1];
            int rank, nproc, p2p;
            comm_info(&rank, &nproc, &p2p);
            // x-extent of the y-pencil must be the same across the z-group (it is: x is split by c1) and even for the tensor maps
            if (p2p && nproc == o->p_row * o->p_col && nproc <= 64 && z.n1 % 2 == 0) {
                std::vector<void*> peers(nproc, nullptr);
                void* local = nullptr;
                z.bytes = z.total();
                if (comm_sym_alloc(z.bytes, &local, peers.data()) == 0) {
                    z.base = (char*)local;
                    const int lo = o->c1 * o->p_col + (o->c2 + o->p_col - 1) % o->p_col;
                    const int up = o->c1 * o->p_col + (o->c2 + 1) % o->p_col;
                    z.lower = (char*)peers[lo];
                    z.upper = (char*)peers[up];
                    z.on = true;
                }
            }
        }
    }
    *h = o;
    return 0;
}

int pdo_operators_destroy(pdo_operators_t o) {
    if (!o) return 0;
    for (int a = 0; a < 3; ++a) { pdo_cd10_destroy(o->c10[a]); pdo_cd06_destroy(o->c06[a]); }
    if (o->zs.base) { cudaDeviceSynchronize(); comm_sym_free(o->zs.base); }
    if (o->side) { cudaStreamDestroy(o->side); cudaEventDestroy(o->ev_fork); cudaEventDestroy(o->ev_join); }
    for (double* p : {o->xtmp, o->xdum, o->ztmp, o->zdum, o->ytmp, o->ytmp2}) if (p) cudaFree(p);
    for (void* p : o->hbuf) if (p) cudaFree(p);
    delete o;
    return 0;
}

/* 0: z is local on this grid, 1: z-slab distributed solve, 2: transposes */
int pdo_operators_zmode(pdo_operators_t o) { return !o ? -1 : (o->p_col == 1 ? 0 : (o->zs.on ? 1 : 2)); }

static int ops_dd_entry(pdo_operators_t o, int axis, const double* f, double* out, void* stream) {
    if (!o) return fail(PDO_E_BADARG, "null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t b = sizeof(double) * (size_t)vol3(o->info.ysz);
    OpsArg a[2] = {{f, b, false, false, nullptr}, {out, b, true, false, nullptr}};
    return with_ops_args(o, a, 2, st, [&]() { return ops_dd(o, axis, OPS_D(0), OPS_D(1), st); });
}
int pdo_operators_ddx(pdo_operators_t o, const double* f, double* out, void* st) { return ops_dd_entry(o, 0, f, out, st); }
int pdo_operators_ddy(pdo_operators_t o, const double* f, double* out, void* st) { return ops_dd_entry(o, 1, f, out, st); }
int pdo_operators_ddz(pdo_operators_t o, const double* f, double* out, void* st) { return ops_dd_entry(o, 2, f, out, st); }

/* operators.F90:17-53 */
int pdo_operators_gradient(pdo_operators_t o, const double* f, double* dfdx, double* dfdy, double* dfdz, void* stream) {
    if (!o) return fail(PDO_E_BADARG, "null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t b = sizeof(double) * (size_t)vol3(o->info.ysz);
    OpsArg a[4] = {{f, b, false, false, nullptr}, {dfdx, b, true, false, nullptr}, {dfdy, b, true, false, nullptr}, {dfdz, b, true, false, nullptr}};
    return with_ops_args(o, a, 4, st, [&]() -> int {
        if (int rc = ops_zbegin(o, OPS_D(0), st)) return rc;          // the z exchange overlaps the y and x derivatives
        if (int rc = ops_dd(o, 1, OPS_D(0), OPS_D(2), st)) return rc;
        if (int rc = ops_dd(o, 0, OPS_D(0), OPS_D(1), st)) return rc;
        return ops_zend(o, OPS_D(0), OPS_D(3), st);
    });
}

/* operators.F90:118-151: div = dv/dy, += du/dx, += dw/dz (same order of additions) */
int pdo_operators_divergence(pdo_operators_t o, const double* u, const double* v, const double* w, double* div, void* stream) {
    if (!o) return fail(PDO_E_BADARG, "null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = vol3(o->info.ysz);
    const size_t b = sizeof(double) * (size_t)n;
    OpsArg a[4] = {{u, b, false, false, nullptr}, {v, b, false, false, nullptr}, {w, b, false, false, nullptr}, {div, b, true, false, nullptr}};
    return with_ops_args(o, a, 4, st, [&]() -> int {
        const double *du = OPS_D(0), *dv = OPS_D(1), *dw = OPS_D(2);
        double* dd = OPS_D(3);
        if (int rc = ensure(&o->ytmp, n)) return rc;
        if (int rc = ops_zbegin(o, dw, st)) return rc;
        if (int rc = ops_dd(o, 1, dv, dd, st)) return rc;
        if (int rc = ops_dd(o, 0, du, o->ytmp, st)) return rc;
        if (int rc = ops_add(dd, o->ytmp, n, 1.0, st)) return rc;
        if (int rc = ops_zend(o, dw, o->ytmp, st)) return rc;
        return ops_add(dd, o->ytmp, n, 1.0, st);
    });
}

/* operators.F90:55-116: curl(:,:,:,c) stored as three consecutive y-pencils */
static int curl_dev(pdo_operators_s* o, const double* u, const double* v, const double* w, double* curl, cudaStream_t st) {
    const long long n = vol3(o->info.ysz);
    if (int rc = ensure(&o->ytmp, n)) return rc;
    double *c1 = curl, *c2 = curl + n, *c3 = curl + 2 * n;
    if (int rc = ops_zbegin(o, v, st)) return rc;
    if (int rc = ops_dd(o, 1, w, c1, st)) return rc;             // dw/dy
    if (int rc = ops_zend(o, v, o->ytmp, st)) return rc;         // dv/dz
    if (int rc = ops_add(c1, o->ytmp, n, -1.0, st)) return rc;
    if (int rc = ops_dd(o, 2, u, c2, st)) return rc;             // du/dz
    if (int rc = ops_dd(o, 0, w, o->ytmp, st)) return rc;        // dw/dx
    if (int rc = ops_add(c2, o->ytmp, n, -1.0, st)) return rc;
    if (int rc = ops_dd(o, 0, v, c3, st)) return rc;             // dv/dx
    if (int rc = ops_dd(o, 1, u, o->ytmp, st)) return rc;        // du/dy
    return ops_add(c3, o->ytmp, n, -1.0, st);
}
int pdo_operators_curl(pdo_operators_t o, const double* u, const double* v, const double* w, double* curl, void* stream) {
    if (!o) return fail(PDO_E_BADARG, "null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t b = sizeof(double) * (size_t)vol3(o->info.ysz);
    OpsArg a[4] = {{u, b, false, false, nullptr}, {v, b, false, false, nullptr}, {w, b, false, false, nullptr}, {curl, 3 * b, true, false, nullptr}};
    return with_ops_args(o, a, 4, st, [&]() { return curl_dev(o, OPS_D(0), OPS_D(1), OPS_D(2), OPS_D(3), st); });
}

/* operators.F90:158-224 filter3D: `numtimes` passes of the y filter, then of the x filter, then of the z filter, on a y-pencil
 * field, in place.  The reference copies the result back into the source before each re-filter and transposes around the x
 * and z stages; here the passes ping-pong between two work pencils (same values, no copies), the last pass of the last stage
 * writes `arr`, and a stage whose axis is already resident in the y-pencil (p_row == 1 / p_col == 1) runs without transposes.
 * bc pairs may be null (periodic / 0, 0). */
static int filter3d_dev(pdo_operators_s* o, pdo_filters_t fil, double* arr, int numtimes, const int* x_bc, const int* y_bc,
                        const int* z_bc, cudaStream_t st) {
    for (int i = 0; i < 3; ++i)
        if (fil->xsz[i] != o->info.xsz[i] || fil->ysz[i] != o->info.ysz[i] || fil->zsz[i] != o->info.zsz[i])
            return fail(234, "filter3D: the filters object was built for another decomposition");   // operators.F90:171-174
    if (numtimes < 1) numtimes = 1;   // "do idx = 1, times2fil-1" runs zero times: one pass is always made
    static const int zero[2] = {0, 0};
    const int* bc[3] = {x_bc ? x_bc : zero, y_bc ? y_bc : zero, z_bc ? z_bc : zero};
    const long long n = vol3(o->info.ysz);
    if (int rc = ensure(&o->ytmp, n)) return rc;
    if (int rc = ensure(&o->ytmp2, n)) return rc;
    typedef int (*fil_fn)(pdo_filters_t, const double*, double*, int, int, void*);
    static const fil_fn fns[3] = {pdo_filters_filterx, pdo_filters_filtery, pdo_filters_filterz};
    const double* src = arr;
    auto pick = [&](bool last) -> double* { return last ? arr : (src == o->ytmp ? o->ytmp2 : o->ytmp); };
    static const int order[3] = {1, 0, 2};
    for (int s = 0; s < 3; ++s) {
        const int axis = order[s];
        const bool resident = axis == 1 || (axis == 0 && o->p_row == 1) || (axis == 2 && o->p_col == 1);
        if (resident) {
            // x on a p_row = 1 grid: the y-pencil IS the x-pencil, the filters object holds the same sizes for both
            for (int t = 0; t < numtimes; ++t) {
                double* dst = pick(s == 2 && t == numtimes - 1);
                if (int rc = fns[axis](fil, src, dst, bc[axis][0], bc[axis][1], st)) return rc;
                src = dst;
            }
            continue;
        }
        const int* ps = axis == 0 ? o->info.xsz : o->info.zsz;
        double** pa = axis == 0 ? &o->xtmp : &o->ztmp;
        double** pb = axis == 0 ? &o->xdum : &o->zdum;
        if (int rc = ensure(pa, vol3(ps))) return rc;
        if (int rc = ensure(pb, vol3(ps))) return rc;
        double *a = *pa, *b = *pb;
        if (int rc = decomp_transpose_device(o->gp, axis == 0 ? 1 : 2, src, a, 1, st)) return rc;
        for (int t = 0; t < numtimes; ++t) {
            if (int rc = fns[axis](fil, a, b, bc[axis][0], bc[axis][1], st)) return rc;
            std::swap(a, b);
        }
        double* dst = pick(s == 2);
        if (int rc = decomp_transpose_device(o->gp, axis == 0 ? 0 : 3, a, dst, 1, st)) return rc;
        src = dst;
    }
    return 0;
}
int pdo_operators_filter3d(pdo_operators_t o, pdo_filters_t fil, double* arr, int numtimes, const int* x_bc, const int* y_bc,
                           const int* z_bc, void* stream) {
    if (!o || !fil || !arr) return fail(PDO_E_BADARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    OpsArg a[1] = {{arr, sizeof(double) * (size_t)vol3(o->info.ysz), true, true, nullptr}};
    return with_ops_args(o, a, 1, st, [&]() { return filter3d_dev(o, fil, OPS_D(0), numtimes, x_bc, y_bc, z_bc, st); });
}

/* Test hook: the z-slab algorithm on ONE GPU.  f(n1, n) holds whole lines; it is cut into `nslabs` slabs that exchange
 * halo planes and edge pieces through ordinary device buffers, exactly as `nslabs` GPUs would through peer memory.
 * `which`: 0 cd10 d1, 1 cd10 d2, 2 cd06 d1 (handle = pdo_cd10_t / pdo_cd06_t created for n). */
}  // extern "C"
namespace pdo { namespace hooks {
int zslab_emulate(void* handle, int which, const double* f, double* out, long long n1, int n, int nslabs, void* stream) {
    if (!handle || nslabs < 2 || n % nslabs) return fail(PDO_E_BADARG, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const BandedOp* op = which == 0 ? &((pdo_cd10_t)handle)->d1 : which == 1 ? &((pdo_cd10_t)handle)->d2 : &((pdo_cd06_t)handle)->d1;
    const int nl = n / nslabs;
    const int HW = banded_zslab_halo_chunks(op, nl);
    if (HW < 0) return fail(PDO_E_UNSUPPORTED, "z-slab mode does not cover this operator / slab size");
    const int HB = banded_zslab_halo_rows(op), BW = op->bw;
    std::vector<double*> planes(nslabs), glo(nslabs), ghi(nslabs);
    for (int s = 0; s < nslabs; ++s) {
        PDO_CUDA(cudaMalloc(&planes[s], sizeof(double) * 2 * HB * n1));
        PDO_CUDA(cudaMalloc(&glo[s], sizeof(double) * 2 * BW * HW * n1));
        PDO_CUDA(cudaMalloc(&ghi[s], sizeof(double) * 2 * BW * HW * n1));
    }
    int rc = 0;
    for (int s = 0; s < nslabs && !rc; ++s) {
        const int lo = (s + nslabs - 1) % nslabs, up = (s + 1) % nslabs;
        zslab_push_planes_kernel<<<grid_for((long long)HB * n1), 256, 0, st>>>(f + (long long)s * nl * n1, n1, nl, HB, planes[lo], planes[up]);
    }
    for (int s = 0; s < nslabs && !rc; ++s) {
        const int lo = (s + nslabs - 1) % nslabs, up = (s + 1) % nslabs;
        if (banded_zslab_edges(op, f + (long long)s * nl * n1, n1, nl, planes[s], ghi[lo], glo[up], st) != cudaSuccess) rc = fail(PDO_E_CUDA, "zslab edges failed");
    }
    for (int s = 0; s < nslabs && !rc; ++s)
        if (banded_zslab_apply(op, f + (long long)s * nl * n1, out + (long long)s * nl * n1, n1, nl, planes[s], glo[s], ghi[s], st) != cudaSuccess)
            rc = fail(PDO_E_CUDA, "zslab apply failed");
    cudaStreamSynchronize(st);
    for (int s = 0; s < nslabs; ++s) { cudaFree(planes[s]); cudaFree(glo[s]); cudaFree(ghi[s]); }
    return rc;
}
}}  // namespace pdo::hooks
extern "C" {

}  // extern "C"
