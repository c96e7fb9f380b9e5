// capi_ops.cu — C ABI for the 1-D line operators and their dispatch types (include/padeops_b200.h).
// Host-side mirror of cd10stuff / cd06stuff / cf90stuff / gaussianstuff / cd06staggstuff /
// DerivativesMod / FiltersMod: same constructor arguments, same error codes, same degenerate cases.
#include <cstdlib>
#include <cstring>
#include <new>

#include "banded.cuh"
#include "common.cuh"
#include "nonperiodic.cuh"
#include "np_chunk_tables.h"

namespace pdo {
thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};
StagePool& stage_pool() {
    static thread_local StagePool p;
    return p;
}
}  // namespace pdo

using namespace pdo;

namespace {

// coefficients: derivatives/cd10.F90:16-27, cd06.F90:14-16, cd06stagg.F90:174-176,307,411,532,
// filters/cf90.F90:16-22, gaussian.F90:16-20
constexpr double alpha10d1 = 1.0 / 2.0, beta10d1 = 1.0 / 20.0;
constexpr double a10d1 = (17.0 / 12.0) / 2.0, b10d1 = (101.0 / 150.0) / 4.0, c10d1 = (1.0 / 100.0) / 6.0;
constexpr double alpha10d2 = 334.0 / 899.0, beta10d2 = 43.0 / 1798.0;
constexpr double a10d2 = (1065.0 / 1798.0) / 1.0, b10d2 = (1038.0 / 899.0) / 4.0, c10d2 = (79.0 / 1798.0) / 9.0;
constexpr double alpha06d1 = 1.0 / 3.0, a06d1 = (14.0 / 9.0) / 2.0, b06d1 = (1.0 / 9.0) / 4.0;
constexpr double alpha90 = 6.6624e-1, beta90 = 1.6688e-1;
constexpr double a90 = 9.9965e-1, b90 = 6.6652e-1, c90 = 1.6674e-1, d90 = 4.0e-5, e90 = -5.0e-6;
constexpr double agf = 3565.0 / 10368.0, bgf = 3091.0 / 12960.0, cgf = 1997.0 / 25920.0, dgf = 149.0 / 12960.0,
                 egf = 107.0 / 103680.0;

int check_bc(int bc1, int bcn) {
    if ((bc1 != 0 && bc1 != 1 && bc1 != -1) || (bcn != 0 && bcn != 1 && bcn != -1))
        return fail(324, "Incorrect boundary specification for bc1/bcn (should be 0, 1 or -1)");  // cd10.F90:2044-2046
    return 0;
}

int ensure_device() {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PDO_E_NODEVICE, "no CUDA device available: padeops_b200 has no CPU fallback");
    }
    return 0;
}

// n == 1: derivative = 0, filter = copy (cd10.F90:2037-2040, cf90.F90:1028-1031)
int degenerate(bool is_filter, const double* f, double* out, size_t count, cudaStream_t st) {
    const size_t bytes = count * sizeof(double);
    return with_device_views(f, bytes, out, bytes, st, [&](const void* din, void* dout) -> int {
        if (is_filter) PDO_CUDA(cudaMemcpyAsync(dout, din, bytes, cudaMemcpyDeviceToDevice, st));
        else PDO_CUDA(cudaMemsetAsync(dout, 0, bytes, st));
        return 0;
    });
}

// ------------------------------------------------------------------------------------------------
// Host-pointer (drop-in) path for large fields: the field is cut into ~64 MB pieces along an index that is
// NOT the solve axis (whole lines stay together), and piece i+1 travels host->device while piece i is solved
// and piece i-1 travels device->host — PCIe is full duplex, so the call costs about one direction's transfer
// time instead of H2D + kernel + D2H back to back.  Three internal streams, rings of three device buffers.
// ------------------------------------------------------------------------------------------------
struct HostPipe {
    static constexpr int R = 3;
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[R] = {}, ev_cmp[R] = {}, ev_out[R] = {};
    void* din[R] = {};
    void* dout[R] = {};
    size_t cap_in = 0, cap_out = 0;
    bool ready = false;
    int init() {
        if (ready) return 0;
        PDO_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        PDO_CUDA(cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking));
        PDO_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int i = 0; i < R; ++i) {
            PDO_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            PDO_CUDA(cudaEventCreateWithFlags(&ev_cmp[i], cudaEventDisableTiming));
            PDO_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
        }
        ready = true;
        return 0;
    }
    int reserve(size_t bin, size_t bout) {
        if (cap_in < bin) {
            for (int i = 0; i < R; ++i) { if (din[i]) cudaFree(din[i]); din[i] = nullptr; }
            cap_in = 0;
            for (int i = 0; i < R; ++i) PDO_CUDA(cudaMalloc(&din[i], bin));
            cap_in = bin;
        }
        if (cap_out < bout) {
            for (int i = 0; i < R; ++i) { if (dout[i]) cudaFree(dout[i]); dout[i] = nullptr; }
            cap_out = 0;
            for (int i = 0; i < R; ++i) PDO_CUDA(cudaMalloc(&dout[i], bout));
            cap_out = bout;
        }
        return 0;
    }
};
HostPipe& host_pipe() {
    static thread_local HostPipe p;
    return p;
}

// PDO_PIPE_CHUNK_MB / PDO_PIPE_MIN_MB override the piece size (64 MB) and the field size from which the pipelined
// path is used (256 MB; smaller fields take the plain staged path).  Tests shrink both to cover it on small fields.
size_t env_mb(const char* name, size_t dflt_mb) {
    const char* e = std::getenv(name);
    const double v = e ? std::atof(e) : 0.0;
    return (size_t)((v > 0.0 ? v : (double)dflt_mb) * 1048576.0);
}
size_t pipe_chunk_bytes() { return env_mb("PDO_PIPE_CHUNK_MB", 64); }
size_t pipe_min_bytes() { return env_mb("PDO_PIPE_MIN_MB", 256); }

// f, out: HOST pointers.  in_rows / out_rows: planes along the solve axis on each side (n or n+1).
int apply_host_pipelined(const BandedOp& op, int axis, const double* f, double* out, long long na, long long nb,
                         int in_rows, int out_rows, cudaStream_t st) {
    HostPipe& hp = host_pipe();
    if (int rc = hp.init()) return rc;
    PDO_CUDA(cudaStreamSynchronize(st));  // host-pointer calls are synchronous with respect to the caller's stream
    // View the field as f(n1, rows, n3), solve along `rows` (axis 0: n1 = 1 and the pieces are groups of lines).
    const long long n1 = axis == 0 ? 1 : (axis == 1 ? na : na * nb);
    const long long n3 = axis == 0 ? na * nb : (axis == 1 ? nb : 1);
    const size_t e = sizeof(double);
    const bool split3 = (axis == 0) || (n3 >= 4);  // cut along the outer index (contiguous pieces), else along n1 (2-D copies)
    const long long mrows = in_rows > out_rows ? in_rows : out_rows;
    const long long unit = split3 ? n1 * mrows : mrows * n3;  // doubles per outer slab / per column
    long long per = (long long)(pipe_chunk_bytes() / e) / (unit > 0 ? unit : 1);  // outer slabs (split3) or columns (else) per piece
    if (per < 1) per = 1;
    if (!split3) { per &= ~1LL; if (per < 2) per = 2; }  // even column counts keep the 16-byte paths
    const long long total = split3 ? n3 : n1;
    const long long npieces = (total + per - 1) / per;
    if (int rc = hp.reserve((size_t)per * unit * e, (size_t)per * unit * e)) return rc;
    for (long long i = 0; i < npieces; ++i) {
        const int slot = (int)(i % HostPipe::R);
        const long long o0 = i * per, cnt = (o0 + per <= total) ? per : total - o0;
        if (i >= HostPipe::R) {
            PDO_CUDA(cudaStreamWaitEvent(hp.s_in, hp.ev_cmp[slot], 0));   // din[slot] was read by the solve of piece i-R
            PDO_CUDA(cudaStreamWaitEvent(hp.s_cmp, hp.ev_out[slot], 0));  // dout[slot] was read by the D2H of piece i-R
        }
        double* di = (double*)hp.din[slot];
        double* d_o = (double*)hp.dout[slot];
        if (split3) {
            PDO_CUDA(cudaMemcpyAsync(di, f + o0 * n1 * in_rows, (size_t)cnt * n1 * in_rows * e, cudaMemcpyHostToDevice, hp.s_in));
        } else {
            PDO_CUDA(cudaMemcpy2DAsync(di, (size_t)cnt * e, f + o0, (size_t)n1 * e, (size_t)cnt * e, (size_t)in_rows * n3,
                                       cudaMemcpyHostToDevice, hp.s_in));
        }
        PDO_CUDA(cudaEventRecord(hp.ev_in[slot], hp.s_in));
        PDO_CUDA(cudaStreamWaitEvent(hp.s_cmp, hp.ev_in[slot], 0));
        // the piece is itself a field of the same kind: f(n, cnt, 1) / f(na, n, cnt) / f(cnt, 1, n)
        const long long pa = axis == 0 ? cnt : (axis == 1 ? (split3 ? na : cnt) : cnt);
        const long long pb = axis == 0 ? 1 : (axis == 1 ? (split3 ? cnt : nb) : 1);
        PDO_CUDA(banded_op_apply(&op, axis, di, d_o, pa, pb, hp.s_cmp, 0));
        g_launches += (op.M ? 1 : (op.bw ? 2 : 1));
        PDO_CUDA(cudaEventRecord(hp.ev_cmp[slot], hp.s_cmp));
        PDO_CUDA(cudaStreamWaitEvent(hp.s_out, hp.ev_cmp[slot], 0));
        if (split3) {
            PDO_CUDA(cudaMemcpyAsync(out + o0 * n1 * out_rows, d_o, (size_t)cnt * n1 * out_rows * e, cudaMemcpyDeviceToHost, hp.s_out));
        } else {
            PDO_CUDA(cudaMemcpy2DAsync(out + o0, (size_t)n1 * e, d_o, (size_t)cnt * e, (size_t)cnt * e, (size_t)out_rows * n3,
                                       cudaMemcpyDeviceToHost, hp.s_out));
        }
        PDO_CUDA(cudaEventRecord(hp.ev_out[slot], hp.s_out));
    }
    PDO_CUDA(cudaStreamSynchronize(hp.s_out));
    PDO_CUDA(cudaStreamSynchronize(hp.s_cmp));
    PDO_CUDA(cudaStreamSynchronize(hp.s_in));
    return 0;
}

int apply(const BandedOp& op, bool is_filter, int axis, const double* f, double* out, long long na, long long nb,
          void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!f || !out) return fail(PDO_E_BADARG, "null field pointer");
    if (na < 0 || nb < 0) return fail(PDO_E_BADARG, "negative extent");
    const int in_rows = op.n + (op.op.edge_in || (op.op.edge_out && op.rk == RK_D2_5) ? 1 : 0);
    const int out_rows = op.n + (op.op.edge_out ? 1 : 0);
    const size_t in_count = (size_t)in_rows * na * nb;
    const size_t out_count = (size_t)out_rows * na * nb;
    if (op.n == 1) return degenerate(is_filter, f, out, out_count, st);
    if (in_count * sizeof(double) >= pipe_min_bytes() && axis >= 0 && axis <= 2 && !is_device_ptr(f) && !is_device_ptr(out))
        return apply_host_pipelined(op, axis, f, out, na, nb, in_rows, out_rows, st);
    return with_device_views(f, in_count * sizeof(double), out, out_count * sizeof(double), st,
                             [&](const void* din, void* dout) -> int {
                                 PDO_CUDA(banded_op_apply(&op, axis, (const double*)din, (double*)dout, na, nb, st, 0));
                                 g_launches += (op.M ? 1 : (op.bw ? 2 : 1));
                                 return 0;
                             });
}

// periodic = .false.: the closures of nonperiodic.cu (correctness path; host pointers take the plain staged route)
int apply_np(const NpOp& op, bool is_filter, int axis, const double* f, double* out, long long na, long long nb, int bc1, int bcn,
             void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!f || !out) return fail(PDO_E_BADARG, "null field pointer");
    if (na < 0 || nb < 0) return fail(PDO_E_BADARG, "negative extent");
    const size_t count = (size_t)op.n * na * nb;
    if (op.n == 1) return degenerate(is_filter, f, out, count, st);
    return with_device_views(f, count * sizeof(double), out, count * sizeof(double), st, [&](const void* din, void* dout) -> int {
        PDO_CUDA(np_op_apply(&op, axis, (const double*)din, (double*)dout, na, nb, bc1, bcn, st));
        g_launches += 2;
        return 0;
    });
}

}  // namespace

// ------------------------------------------------------------------------------------------------
#include "handles.h"

extern "C" {

const char* pdo_last_error(void) { return g_last_error.c_str(); }
int pdo_version(void) { return 100; }
int64_t pdo_launch_count(void) { return (int64_t)g_launches.load(); }

/* Host-only test hook (no device needed): the chunk tables the kernels receive for the cyclic matrix circ[b2 b1 1 b1 b2]
 * of size n cut into chunks of M points.  `out` receives a copy of pdo::ChunkTables (tables.h); returns 0, -1 if (n, M) is
 * not chunkable, PDO_E_BADARG if out_bytes does not match the struct. */
}  // extern "C"
namespace pdo { namespace hooks {
int chunk_tables(int n, int M, int bw, double b1, double b2, void* out, int out_bytes) {
    if (!out || out_bytes != (int)sizeof(ChunkTables)) return fail(PDO_E_BADARG, "out_bytes %d != sizeof(ChunkTables) %d", out_bytes, (int)sizeof(ChunkTables));
    ChunkTables t;
    const int rc = build_chunk_tables(n, M, bw, b1, b2, &t);
    std::memcpy(out, &t, sizeof(t));
    return rc;
}
}}  // namespace pdo::hooks
extern "C" {

/* Host-only test hook: the non-periodic closures computed by the SAME __host__ __device__ routines the kernels execute
 * (nonperiodic.cuh), on the host, so that CPU tests can check the transcription of the boundary rows and the LU sweeps.
 * kind: 0 cd10 d1, 1 cd10 d2, 2 cf90.  Not reachable from any public entry point: the API itself needs a GPU. */
}  // extern "C"
namespace pdo { namespace hooks {
int np_line_host(int kind, int n, double dx, int bc1, int bcn, int axis, const double* f, double* out, long long na,
                           long long nb) {
    if (!f || !out || axis < 0 || axis > 2) return fail(PDO_E_BADARG, "bad argument");
    return np_apply_host(kind, n, dx, bc1, bcn, axis, f, out, na, nb);
}
}}  // namespace pdo::hooks
extern "C" {

/* Host-only test hook: chunk tables of the non-periodic system of `kind` (0 cd10 d1, 1 cd10 d2, 2 cf90) with boundary codes
 * (bc1, bcn): sets = 3 x NpChunkSet (first, mid, last) as doubles, G = [P][2W+1][4], meta = {P, W, doubles per set}. */
}  // extern "C"
namespace pdo { namespace hooks {
int np_chunk_tables(int kind, int n, int M, int bc1, int bcn, double* sets, double* G, int g_capacity, int* meta) {
    if (!sets || !G || !meta) return fail(PDO_E_BADARG, "null argument");
    std::vector<double> rows(5 * (size_t)(n > 0 ? n : 1));
    if (int rc = np_build_rows(kind, n, bc1, bcn, rows.data())) return rc;
    NpChunkTables t;
    if (int rc = build_np_chunk_tables(n, M, rows.data(), &t)) return rc;
    if ((int)t.G.size() > g_capacity) return fail(PDO_E_BADARG, "G capacity %d < %d", g_capacity, (int)t.G.size());
    static_assert(sizeof(NpChunkSet) % sizeof(double) == 0, "NpChunkSet is all doubles");
    std::memcpy(sets, &t.first, sizeof(NpChunkSet));
    std::memcpy(sets + sizeof(NpChunkSet) / sizeof(double), &t.mid, sizeof(NpChunkSet));
    std::memcpy(sets + 2 * sizeof(NpChunkSet) / sizeof(double), &t.last, sizeof(NpChunkSet));
    std::memcpy(G, t.G.data(), sizeof(double) * t.G.size());
    meta[0] = t.P; meta[1] = t.W; meta[2] = (int)(sizeof(NpChunkSet) / sizeof(double));
    return 0;
}
}}  // namespace pdo::hooks
extern "C" {
/* Host-only: chunks per CTA (0 = no launch shape) and dynamic shared memory of the cluster + TMA strided kernel */
}  // extern "C"
namespace pdo { namespace hooks {
int ctma_config(int P, int XT, int HB, int HW, int BW, int pc_max, long long* smem_bytes) {
    return banded_debug_ctma_config(P, XT, HB, HW, BW, pc_max, smem_bytes);
}
}}  // namespace pdo::hooks
extern "C" {
/* Host-only: the rows (bt b d a at, 5n doubles) of the non-periodic system */
}  // extern "C"
namespace pdo { namespace hooks {
int np_rows(int kind, int n, int bc1, int bcn, double* rows5n) { return np_build_rows(kind, n, bc1, bcn, rows5n); }
}}  // namespace pdo::hooks
extern "C" {
/* -1 default (on), 0: non-periodic calls keep the one-thread-per-line sweeps, 1: chunked fast path where the line is chunkable */
}  // extern "C"
namespace pdo { namespace hooks {
int np_fast(int mode) { np_set_fast_path(mode); return 0; }
}}  // namespace pdo::hooks
extern "C" {

int pdo_malloc(void** dptr, size_t bytes) {
    if (int rc = ensure_device()) return rc;
    PDO_CUDA(cudaMalloc(dptr, bytes));
    return 0;
}
int pdo_free(void* dptr) {
    PDO_CUDA(cudaFree(dptr));
    return 0;
}
int pdo_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    PDO_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}
int pdo_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    PDO_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}
int pdo_stream_sync(void* stream) {
    PDO_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

// ---------------- cd10 ----------------
int pdo_cd10_init(pdo_cd10_t* h, int n, double dx, int periodic, int bc1, int bcn) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    if (n < 1) return fail(PDO_E_BADARG, "n < 1");
    if (!periodic) {   // cd10.F90:238-321: the nine penta tables of each derivative
        if ((bc1 != 0 && bc1 != 1 && bc1 != -1) || (bcn != 0 && bcn != 1 && bcn != -1))
            return fail(324, "Incorrect boundary specification for bc1/bcn (should be 0, 1 or -1)");
        if (n != 1 && n < 8) return fail(2, "cd10: non-periodic n must be 1 or >= 8");
        if (int rc = ensure_device()) return rc;
        pdo_cd10_s* o = new (std::nothrow) pdo_cd10_s();
        if (!o) return fail(PDO_E_BADARG, "out of memory");
        o->n = n; o->periodic = false;
        int ie1 = 0, ie2 = 0;
        cudaError_t e = np_op_create(&o->np_d1, NP_CD10_D1, n, dx, &ie1);
        if (e == cudaSuccess) e = np_op_create(&o->np_d2, NP_CD10_D2, n, dx, &ie2);
        if (e != cudaSuccess || ie1 || ie2) {
            np_op_destroy(&o->np_d1); np_op_destroy(&o->np_d2);
            delete o;
            return e != cudaSuccess ? fail(PDO_E_CUDA, "cd10 init: %s", cudaGetErrorString(e)) : fail(ie1 ? ie1 : ie2, "cd10: non-periodic tables");
        }
        *h = o;
        return 0;
    }
    if (n != 1 && n < 8) return fail(2, "cd10: periodic n must be 1 or >= 8");  // cd10.F90:219-226
    if (int rc = ensure_device()) return rc;
    pdo_cd10_s* o = new (std::nothrow) pdo_cd10_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    o->n = n;
    const double onebydx = 1.0 / dx, onebydx2 = onebydx / dx;  // cd10.F90:206-207
    OpParams p1{}, p2{};
    p1.co[0] = a10d1 * onebydx; p1.co[1] = b10d1 * onebydx; p1.co[2] = c10d1 * onebydx;     // :1115-1117
    p2.co[0] = a10d2 * onebydx2; p2.co[1] = b10d2 * onebydx2; p2.co[2] = c10d2 * onebydx2;  // :1601-1603
    cudaError_t e = banded_op_create(&o->d1, n, RK_D1_7, 2, alpha10d1, beta10d1, p1);
    if (e == cudaSuccess) e = banded_op_create(&o->d2, n, RK_D2_7, 2, alpha10d2, beta10d2, p2);
    if (e != cudaSuccess) {
        delete o;
        return fail(PDO_E_CUDA, "cd10 init: %s", cudaGetErrorString(e));
    }
    *h = o;
    return 0;
}
int pdo_cd10_destroy(pdo_cd10_t h) {
    if (!h) return 0;
    if (h->periodic) { banded_op_destroy(&h->d1); banded_op_destroy(&h->d2); }
    else { np_op_destroy(&h->np_d1); np_op_destroy(&h->np_d2); }
    delete h;
    return 0;
}
int pdo_cd10_getsize(pdo_cd10_t h) { return h ? h->n : -1; }

#define PDO_CD10_FN(name, member, axis)                                                                   \
    int name(pdo_cd10_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream) { \
        if (!h) return fail(PDO_E_BADARG, "null handle");                                                 \
        if (int rc = check_bc(bc1, bcn)) return rc;                                                       \
        if (!h->periodic) return apply_np(h->np_##member, false, axis, f, df, na, nb, bc1, bcn, stream);   \
        return apply(h->member, false, axis, f, df, na, nb, stream);                                      \
    }
PDO_CD10_FN(pdo_cd10_dd1, d1, 0)
PDO_CD10_FN(pdo_cd10_dd2, d1, 1)
PDO_CD10_FN(pdo_cd10_dd3, d1, 2)
PDO_CD10_FN(pdo_cd10_d2d1, d2, 0)
PDO_CD10_FN(pdo_cd10_d2d2, d2, 1)
PDO_CD10_FN(pdo_cd10_d2d3, d2, 2)

// ---------------- cd06 ----------------
int pdo_cd06_init(pdo_cd06_t* h, int n, double dx, int periodic, int bc1, int bcn) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    if (n < 1) return fail(PDO_E_BADARG, "n < 1");
    if (!periodic && (bc1 != 0 || bcn != 0))
        return fail(PDO_E_UNSUPPORTED, "cd06: only the one-sided non-periodic closure exists (the reference marks bc = 1 'Incomplete', cd06.F90:288-291)");
    if (n != 1 && n < 6) return fail(3, "cd06: n must be 1 or >= 6");  // cd06.F90:151-160
    if (int rc = ensure_device()) return rc;
    pdo_cd06_s* o = new (std::nothrow) pdo_cd06_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    o->n = n;
    if (!periodic) {   // cd06.F90:172-180: ComputeTri1
        o->periodic = false;
        int ie = 0;
        cudaError_t e = np_op_create(&o->np, NP_CD06_D1, n, dx, &ie);
        if (e != cudaSuccess || ie) {
            np_op_destroy(&o->np);
            delete o;
            return e != cudaSuccess ? fail(PDO_E_CUDA, "cd06 init: %s", cudaGetErrorString(e)) : fail(ie, "cd06: non-periodic tables");
        }
        *h = o;
        return 0;
    }
    const double onebydx = 1.0 / dx;
    OpParams p{};
    p.co[0] = a06d1 * onebydx; p.co[1] = b06d1 * onebydx;  // cd06.F90:530-531
    cudaError_t e = banded_op_create(&o->d1, n, RK_D1_5, 1, alpha06d1, 0.0, p);
    if (e != cudaSuccess) {
        delete o;
        return fail(PDO_E_CUDA, "cd06 init: %s", cudaGetErrorString(e));
    }
    *h = o;
    return 0;
}
int pdo_cd06_destroy(pdo_cd06_t h) {
    if (!h) return 0;
    if (h->periodic) banded_op_destroy(&h->d1);
    else np_op_destroy(&h->np);
    delete h;
    return 0;
}
int pdo_cd06_getsize(pdo_cd06_t h) { return h ? h->n : -1; }
#define PDO_CD06_FN(name, axis)                                                                           \
    int name(pdo_cd06_t h, const double* f, double* df, int na, int nb, int bc1, int bcn, void* stream) { \
        if (!h) return fail(PDO_E_BADARG, "null handle");                                                 \
        if (int rc = check_bc(bc1, bcn)) return rc;                                                       \
        /* the reference's cd06%dd* take no boundary codes (cd06.F90:775-839): always the one-sided closure */ \
        if (!h->periodic) return apply_np(h->np, false, axis, f, df, na, nb, 0, 0, stream);               \
        return apply(h->d1, false, axis, f, df, na, nb, stream);                                          \
    }
PDO_CD06_FN(pdo_cd06_dd1, 0)
PDO_CD06_FN(pdo_cd06_dd2, 1)
PDO_CD06_FN(pdo_cd06_dd3, 2)

// ---------------- cf90 ----------------
int pdo_cf90_init(pdo_cf90_t* h, int n, int periodic) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    if (n < 1) return fail(PDO_E_BADARG, "n < 1");
    if (n != 1 && n < 10) return fail(7, "cf90: n must be 1 or >= 10");  // cf90.F90:121-129
    if (int rc = ensure_device()) return rc;
    pdo_cf90_s* o = new (std::nothrow) pdo_cf90_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    o->n = n;
    if (!periodic) {   // cf90.F90:131-172: the nine penta tables
        o->periodic = false;
        int ie = 0;
        cudaError_t e = np_op_create(&o->np, NP_CF90, n, 1.0, &ie);
        if (e != cudaSuccess || ie) {
            np_op_destroy(&o->np);
            delete o;
            return e != cudaSuccess ? fail(PDO_E_CUDA, "cf90 init: %s", cudaGetErrorString(e)) : fail(ie, "cf90: non-periodic tables");
        }
        *h = o;
        return 0;
    }
    OpParams p{};
    p.co[0] = a90; p.co[1] = b90; p.co[2] = c90; p.co[3] = d90; p.co[4] = e90;
    cudaError_t e = banded_op_create(&o->op, n, RK_SYM_9, 2, alpha90, beta90, p);
    if (e != cudaSuccess) {
        delete o;
        return fail(PDO_E_CUDA, "cf90 init: %s", cudaGetErrorString(e));
    }
    *h = o;
    return 0;
}
int pdo_cf90_destroy(pdo_cf90_t h) {
    if (!h) return 0;
    if (h->periodic) banded_op_destroy(&h->op);
    else np_op_destroy(&h->np);
    delete h;
    return 0;
}
#define PDO_CF90_FN(name, axis)                                                                              \
    int name(pdo_cf90_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream) {   \
        if (!h) return fail(PDO_E_BADARG, "null handle");                                                    \
        if (int rc = check_bc(bc1, bcn)) return rc;                                                          \
        if (!h->periodic) return apply_np(h->np, true, axis, f, fil, na, nb, bc1, bcn, stream);              \
        return apply(h->op, true, axis, f, fil, na, nb, stream);                                             \
    }
PDO_CF90_FN(pdo_cf90_filter1, 0)
PDO_CF90_FN(pdo_cf90_filter2, 1)
PDO_CF90_FN(pdo_cf90_filter3, 2)
// ---------------- gaussian ----------------
int pdo_gaussian_init(pdo_gaussian_t* h, int n, int periodic) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    if (n < 1) return fail(PDO_E_BADARG, "n < 1");
    if (periodic && n != 1 && n < 9) return fail(PDO_E_BADARG, "gaussian: periodic 9-point stencil needs n >= 9");
    if (!periodic && n != 1 && n < 8) return fail(PDO_E_BADARG, "gaussian: the non-periodic closure needs n >= 8 (four boundary rows at each end)");
    if (int rc = ensure_device()) return rc;
    pdo_gaussian_s* o = new (std::nothrow) pdo_gaussian_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    o->n = n;
    if (!periodic) {   // gaussian.F90:215-330: explicit boundary rows (bc 0) or the interior stencil on the reflected line (bc +-1)
        o->periodic = false;
        int ie = 0;
        cudaError_t e = np_op_create(&o->np, NP_GAUSS, n, 1.0, &ie);
        if (e != cudaSuccess || ie) { delete o; return e != cudaSuccess ? fail(PDO_E_CUDA, "gaussian init: %s", cudaGetErrorString(e)) : fail(ie, "gaussian: non-periodic init"); }
        *h = o;
        return 0;
    }
    OpParams p{};
    p.co[0] = agf; p.co[1] = bgf; p.co[2] = cgf; p.co[3] = dgf; p.co[4] = egf;
    cudaError_t e = banded_op_create(&o->op, n, RK_SYM_9, 0, 0.0, 0.0, p);
    if (e != cudaSuccess) {
        delete o;
        return fail(PDO_E_CUDA, "gaussian init: %s", cudaGetErrorString(e));
    }
    *h = o;
    return 0;
}
int pdo_gaussian_destroy(pdo_gaussian_t h) {
    if (!h) return 0;
    if (h->periodic) banded_op_destroy(&h->op);
    else np_op_destroy(&h->np);
    delete h;
    return 0;
}
#define PDO_GAUSS_FN(name, axis)                                                                               \
    int name(pdo_gaussian_t h, const double* f, double* fil, int na, int nb, int bc1, int bcn, void* stream) { \
        if (!h) return fail(PDO_E_BADARG, "null handle");                                                      \
        if (int rc = check_bc(bc1, bcn)) return rc;                                                            \
        if (!h->periodic) return apply_np(h->np, true, axis, f, fil, na, nb, bc1, bcn, stream);                \
        return apply(h->op, true, axis, f, fil, na, nb, stream);                                               \
    }
PDO_GAUSS_FN(pdo_gaussian_filter1, 0)
PDO_GAUSS_FN(pdo_gaussian_filter2, 1)
PDO_GAUSS_FN(pdo_gaussian_filter3, 2)

// ---------------- lstsq (filters/lstsq.F90): the Gaussian filter's structure with the least-squares coefficients; no bc arguments ----------------
int pdo_lstsq_init(pdo_lstsq_t* h, int n, int periodic) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    if (n < 1) return fail(PDO_E_BADARG, "n < 1");
    if (periodic && n != 1 && n < 9) return fail(PDO_E_BADARG, "lstsq: periodic 9-point stencil needs n >= 9");
    if (!periodic && n != 1 && n < 8) return fail(PDO_E_BADARG, "lstsq: the non-periodic closure needs n >= 8 (four boundary rows at each end)");
    if (int rc = ensure_device()) return rc;
    pdo_lstsq_s* o = new (std::nothrow) pdo_lstsq_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    o->n = n;
    if (!periodic) {   // lstsq.F90:169-212: one-sided rows at both ends, always
        o->periodic = false;
        int ie = 0;
        cudaError_t e = np_op_create(&o->np, NP_LSTSQ, n, 1.0, &ie);
        if (e != cudaSuccess || ie) { delete o; return e != cudaSuccess ? fail(PDO_E_CUDA, "lstsq init: %s", cudaGetErrorString(e)) : fail(ie, "lstsq: non-periodic init"); }
        *h = o;
        return 0;
    }
    OpParams p{};
    // lstsq.F90:14-19 writes real(0.6744132, rkind): a default-real (single precision) literal widened to double, i.e.
    // 0.67441320419311523 and -0.17441320419311523, not the decimal values
    p.co[0] = 0.5; p.co[1] = (double)0.6744132f / 2.0; p.co[2] = 0.0 / 2.0; p.co[3] = (double)(-0.1744132f) / 2.0; p.co[4] = 0.0 / 2.0;
    cudaError_t e = banded_op_create(&o->op, n, RK_SYM_9, 0, 0.0, 0.0, p);
    if (e != cudaSuccess) { delete o; return fail(PDO_E_CUDA, "lstsq init: %s", cudaGetErrorString(e)); }
    *h = o;
    return 0;
}
int pdo_lstsq_destroy(pdo_lstsq_t h) {
    if (!h) return 0;
    if (h->periodic) banded_op_destroy(&h->op);
    else np_op_destroy(&h->np);
    delete h;
    return 0;
}
#define PDO_LSTSQ_FN(name, axis)                                                                \
    int name(pdo_lstsq_t h, const double* f, double* fil, int na, int nb, void* stream) {     \
        if (!h) return fail(PDO_E_BADARG, "null handle");                                       \
        if (!h->periodic) return apply_np(h->np, true, axis, f, fil, na, nb, 0, 0, stream);     \
        return apply(h->op, true, axis, f, fil, na, nb, stream);                                \
    }
PDO_LSTSQ_FN(pdo_lstsq_filter1, 0)
PDO_LSTSQ_FN(pdo_lstsq_filter2, 1)
PDO_LSTSQ_FN(pdo_lstsq_filter3, 2)

// ---------------- cd06stagg (periodic) ----------------
int pdo_cd06stagg_init_periodic(pdo_cd06stagg_t* h, int n, double dx) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    if (n <= 4) return fail(21, "CD06_stagg requires at least 4 points");  // cd06stagg.F90:182-184
    if (int rc = ensure_device()) return rc;
    pdo_cd06stagg_s* o = new (std::nothrow) pdo_cd06stagg_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    o->n = n;
    const double onebydx = 1.0 / dx;
    const double aD1 = 9.0 / 62.0, aD2 = 2.0 / 11.0, aI = 3.0 / 10.0;  // :174-176
    OpParams d1{}, in{}, d2{};
    d1.co[0] = (63.0 / 62.0) * onebydx; d1.co[1] = ((17.0 / 62.0) / 3.0) * onebydx;                       // :307-311
    in.co[0] = (3.0 / 2.0) * (1.0 / 2.0); in.co[1] = (1.0 / 10.0) * (1.0 / 2.0);                           // :532-536
    d2.co[0] = (12.0 / 11.0) * (onebydx * onebydx); d2.co[1] = ((3.0 / 11.0) / 4.0) * (onebydx * onebydx);  // :411-415
    cudaError_t e = cudaSuccess;
    auto mk = [&](int idx, int rk, double al, OpParams p, double sg, int ein, int eout) {
        p.co[2] = sg; p.edge_in = ein; p.edge_out = eout;
        if (e == cudaSuccess) e = banded_op_create(&o->ops[idx], n, rk, 1, al, 0.0, p);
    };
    mk(0, RK_STAG_E2C, aD1, d1, -1.0, 1, 0);  // ddz_E2C
    mk(1, RK_STAG_C2E, aD1, d1, -1.0, 0, 1);  // ddz_C2E
    mk(2, RK_STAG_E2C, aI, in, +1.0, 1, 0);   // InterpZ_E2C
    mk(3, RK_STAG_C2E, aI, in, +1.0, 0, 1);   // InterpZ_C2E
    mk(4, RK_D2_5, aD2, d2, 0.0, 0, 0);       // d2dz2_C2C
    mk(5, RK_D2_5, aD2, d2, 0.0, 0, 1);       // d2dz2_E2E
    if (e != cudaSuccess) {
        pdo_cd06stagg_destroy(o);
        return fail(PDO_E_CUDA, "cd06stagg init: %s", cudaGetErrorString(e));
    }
    *h = o;
    return 0;
}
/* cd06stagg%init(nx, dx, isTopEven, isBotEven, isTopSided, isBotSided)  (init_nonperiodic, cd06stagg.F90:197-231) */
int pdo_cd06stagg_init_nonperiodic(pdo_cd06stagg_t* h, int n, double dx, int is_top_even, int is_bot_even, int is_top_sided,
                                   int is_bot_sided) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    if (n <= 4) return fail(21, "CD06_stagg requires at least 4 points");  // :216-218
    if (int rc = ensure_device()) return rc;
    pdo_cd06stagg_s* o = new (std::nothrow) pdo_cd06stagg_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    o->n = n;
    o->periodic = false;
    StaggNpFlags fl{is_bot_even != 0, is_top_even != 0, is_bot_sided != 0, is_top_sided != 0};
    int ierr = 0;
    cudaError_t e = snp_create(&o->np, n, dx, fl, &ierr);
    if (e != cudaSuccess || ierr) {
        delete o;
        return e != cudaSuccess ? fail(PDO_E_CUDA, "cd06stagg init: %s", cudaGetErrorString(e)) : fail(ierr, "cd06stagg init_nonperiodic failed");
    }
    *h = o;
    return 0;
}
int pdo_cd06stagg_destroy(pdo_cd06stagg_t h) {
    if (!h) return 0;
    if (h->periodic) for (int i = 0; i < 6; ++i) banded_op_destroy(&h->ops[i]);
    else snp_destroy(&h->np);
    delete h;
    return 0;
}
// non-periodic operator `op` (stagg_np.cuh) on in(n1, n2, rows_in) -> out(n1, n2, rows_out); host pointers are staged
static int apply_snp(pdo_cd06stagg_t h, int op, const double* in, double* out, int n1, int n2, int is_complex, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!in || !out) return fail(PDO_E_BADARG, "null field pointer");
    if (n1 < 0 || n2 < 0) return fail(PDO_E_BADARG, "negative extent");
    const long long ncols = (long long)n1 * n2 * (is_complex ? 2 : 1);
    const size_t bin = sizeof(double) * (size_t)ncols * snp_rows_in(op, h->n), bout = sizeof(double) * (size_t)ncols * snp_rows_out(op, h->n);
    return with_device_views(in, bin, out, bout, st, [&](const void* di, void* d_o) -> int {
        if (di == d_o) return fail(PDO_E_BADARG, "cd06stagg: input and output must not alias");
        PDO_CUDA(snp_apply(&h->np, op, (const double*)di, (double*)d_o, ncols, st));
        g_launches += 1;
        return 0;
    });
}
#define PDO_STAGG_FN(name, idx, snp)                                                                           \
    int name(pdo_cd06stagg_t h, const double* in, double* out, int n1, int n2, int is_complex, void* stream) { \
        if (!h) return fail(PDO_E_BADARG, "null handle");                                                      \
        if (!h->periodic) return apply_snp(h, snp, in, out, n1, n2, is_complex, stream);                       \
        /* complex data, real LU: re/im are independent lines -> a real field with 2*n1 points in x */         \
        return apply(h->ops[idx], false, 2, in, out, (long long)n1 * (is_complex ? 2 : 1), n2, stream);        \
    }
PDO_STAGG_FN(pdo_cd06stagg_ddz_E2C, 0, SNP_D1_E2C)
PDO_STAGG_FN(pdo_cd06stagg_ddz_C2E, 1, SNP_D1_C2E)
PDO_STAGG_FN(pdo_cd06stagg_interpz_E2C, 2, SNP_INTERP_E2C)
PDO_STAGG_FN(pdo_cd06stagg_interpz_C2E, 3, SNP_INTERP_C2E)
PDO_STAGG_FN(pdo_cd06stagg_d2dz2_C2C, 4, SNP_D2_C2C)
PDO_STAGG_FN(pdo_cd06stagg_d2dz2_E2E, 5, SNP_D2_E2E)
// collocated first derivatives on the staggered grids: tables exist only after init_nonperiodic (cd06stagg.F90:883-925)
int pdo_cd06stagg_ddz_C2C(pdo_cd06stagg_t h, const double* in, double* out, int n1, int n2, int is_complex, void* stream) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    if (h->periodic) return fail(PDO_E_UNSUPPORTED, "cd06stagg ddz_C2C: the reference builds TriD1_C2C only in init_nonperiodic");
    return apply_snp(h, SNP_D1_C2C, in, out, n1, n2, is_complex, stream);
}
int pdo_cd06stagg_ddz_E2E(pdo_cd06stagg_t h, const double* in, double* out, int n1, int n2, int is_complex, void* stream) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    if (h->periodic) return fail(PDO_E_UNSUPPORTED, "cd06stagg ddz_E2E: the reference builds TriD1_E2E only in init_nonperiodic");
    return apply_snp(h, SNP_D1_E2E, in, out, n1, n2, is_complex, stream);
}
// test hooks (host only, not in the public header): the kernel's per-line routine and the tridiagonal rows on the CPU
}  // extern "C"
namespace pdo { namespace hooks {
int stagg_np_host(int op, int n, double dx, int bot_even, int top_even, int bot_sided, int top_sided, const double* in, double* out,
                            long long ncols) {
    StaggNpFlags fl{bot_even != 0, top_even != 0, bot_sided != 0, top_sided != 0};
    return snp_apply_host(op, n, dx, fl, in, out, ncols);
}
}}  // namespace pdo::hooks
extern "C" {
}  // extern "C"
namespace pdo { namespace hooks {
int stagg_np_rows(int op, int n, int bot_even, int top_even, int bot_sided, int top_sided, double* rows3n) {
    StaggNpFlags fl{bot_even != 0, top_even != 0, bot_sided != 0, top_sided != 0};
    return snp_build_rows(op, n, fl, rows3n);
}
}}  // namespace pdo::hooks
extern "C" {

}  // extern "C"

// ---------------- explicit planning (FFTW-planner style; never inside an operator call) ----------------
extern "C" {
static int plan_one(const BandedOp& op, int axis, int na, int nb, int* variant) {
    if (axis < 0 || axis > 2 || na < 1 || nb < 1) return fail(PDO_E_BADARG, "plan: axis must be 0..2 and the extents positive");
    int v = 0;
    PDO_CUDA(banded_op_plan(&op, axis, na, nb, &v));
    if (variant) *variant = v;
    return 0;
}
/* Times the kernel candidates of this operator along `axis` (0 x, 1 y, 2 z) of a pencil whose other two extents are (na, nb) —
   the same (na, nb) the dd* / filter* calls take — and stores the winner in the handle: later calls on that shape use it.
   Without a plan the deterministic table of banded.cu decides.  Synchronises the device; allocates two scratch fields. */
int pdo_cd10_plan(pdo_cd10_t h, int axis, int na, int nb, int* variant_d1, int* variant_d2) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    if (!h->periodic) return 0;
    if (int rc = plan_one(h->d1, axis, na, nb, variant_d1)) return rc;
    return plan_one(h->d2, axis, na, nb, variant_d2);
}
int pdo_cd06_plan(pdo_cd06_t h, int axis, int na, int nb, int* variant) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    return h->periodic ? plan_one(h->d1, axis, na, nb, variant) : 0;
}
int pdo_cf90_plan(pdo_cf90_t h, int axis, int na, int nb, int* variant) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    return h->periodic ? plan_one(h->op, axis, na, nb, variant) : 0;
}
int pdo_gaussian_plan(pdo_gaussian_t h, int axis, int na, int nb, int* variant) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    return h->periodic ? plan_one(h->op, axis, na, nb, variant) : 0;
}
/* 0: deterministic dispatch (default); 1: the first large call of an operator on a shape times the candidates itself (the
   PDO_TUNE=1 behaviour) — synchronises inside that call, so not for captured streams. */
int pdo_plan_on_first_call(int enable) { banded_set_tuning(enable != 0); return 0; }
}  // extern "C"

// ---------------- DerivativesMod::derivatives / FiltersMod::filters ----------------
extern "C" {

int pdo_derivatives_init(pdo_derivatives_t* h, const int xsz[3], const int ysz[3], const int zsz[3], double dx, double dy,
                         double dz, int px, int py, int pz, const char* mx, const char* my, const char* mz) {
    if (!h || !xsz || !ysz || !zsz || !mx || !my || !mz) return fail(PDO_E_BADARG, "null argument");
    *h = nullptr;
    pdo_derivatives_s* o = new (std::nothrow) pdo_derivatives_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    std::memcpy(o->xsz, xsz, sizeof(o->xsz));
    std::memcpy(o->ysz, ysz, sizeof(o->ysz));
    std::memcpy(o->zsz, zsz, sizeof(o->zsz));
    const int n[3] = {xsz[0], ysz[1], zsz[2]};  // derivatives.F90:288, 322, 356
    const double d[3] = {dx, dy, dz};
    const int per[3] = {px, py, pz};
    const char* m[3] = {mx, my, mz};
    for (int a = 0; a < 3; ++a) { o->c10[a] = nullptr; o->c06[a] = nullptr; }
    for (int a = 0; a < 3; ++a) {
        int rc;
        if (std::strncmp(m[a], "cd10", 4) == 0) {
            o->method[a] = 0;
            rc = pdo_cd10_init(&o->c10[a], n[a], d[a], per[a], 0, 0);
        } else if (std::strncmp(m[a], "cd06", 4) == 0) {
            o->method[a] = 1;
            rc = pdo_cd06_init(&o->c06[a], n[a], d[a], per[a], 0, 0);
        } else {
            rc = fail(PDO_E_UNSUPPORTED, "derivatives: method '%s' is out of scope (cd10, cd06 only; SURVEY.md 2.1 #5-6)", m[a]);
        }
        if (rc) {
            pdo_derivatives_destroy(o);
            return rc;
        }
    }
    *h = o;
    return 0;
}
int pdo_derivatives_destroy(pdo_derivatives_t h) {
    if (!h) return 0;
    for (int a = 0; a < 3; ++a) { pdo_cd10_destroy(h->c10[a]); pdo_cd06_destroy(h->c06[a]); }
    delete h;
    return 0;
}
static int der_apply(pdo_derivatives_t h, int axis, int order, const double* f, double* out, int bc1, int bcn, void* st) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    const int* sz = axis == 0 ? h->xsz : axis == 1 ? h->ysz : h->zsz;
    int na, nb;
    if (axis == 0) { na = sz[1]; nb = sz[2]; }
    else if (axis == 1) { na = sz[0]; nb = sz[2]; }
    else { na = sz[0]; nb = sz[1]; }
    if (h->method[axis] == 0) {
        typedef int (*fn_t)(pdo_cd10_t, const double*, double*, int, int, int, int, void*);
        static const fn_t fns[2][3] = {{pdo_cd10_dd1, pdo_cd10_dd2, pdo_cd10_dd3}, {pdo_cd10_d2d1, pdo_cd10_d2d2, pdo_cd10_d2d3}};
        return fns[order - 1][axis](h->c10[axis], f, out, na, nb, bc1, bcn, st);
    }
    if (order == 2) return fail(PDO_E_UNSUPPORTED, "CD06 is incomplete right now");  // derivatives.F90:525
    typedef int (*fn6_t)(pdo_cd06_t, const double*, double*, int, int, int, int, void*);
    static const fn6_t f6[3] = {pdo_cd06_dd1, pdo_cd06_dd2, pdo_cd06_dd3};
    return f6[axis](h->c06[axis], f, out, na, nb, bc1, bcn, st);
}
int pdo_derivatives_ddx(pdo_derivatives_t h, const double* f, double* o, int b1, int bn, void* s) { return der_apply(h, 0, 1, f, o, b1, bn, s); }
int pdo_derivatives_ddy(pdo_derivatives_t h, const double* f, double* o, int b1, int bn, void* s) { return der_apply(h, 1, 1, f, o, b1, bn, s); }
int pdo_derivatives_ddz(pdo_derivatives_t h, const double* f, double* o, int b1, int bn, void* s) { return der_apply(h, 2, 1, f, o, b1, bn, s); }
int pdo_derivatives_d2dx2(pdo_derivatives_t h, const double* f, double* o, int b1, int bn, void* s) { return der_apply(h, 0, 2, f, o, b1, bn, s); }
int pdo_derivatives_d2dy2(pdo_derivatives_t h, const double* f, double* o, int b1, int bn, void* s) { return der_apply(h, 1, 2, f, o, b1, bn, s); }
int pdo_derivatives_d2dz2(pdo_derivatives_t h, const double* f, double* o, int b1, int bn, void* s) { return der_apply(h, 2, 2, f, o, b1, bn, s); }

int pdo_filters_init(pdo_filters_t* h, const int xsz[3], const int ysz[3], const int zsz[3], int px, int py, int pz,
                     const char* mx, const char* my, const char* mz) {
    if (!h || !xsz || !ysz || !zsz || !mx || !my || !mz) return fail(PDO_E_BADARG, "null argument");
    *h = nullptr;
    pdo_filters_s* o = new (std::nothrow) pdo_filters_s();
    if (!o) return fail(PDO_E_BADARG, "out of memory");
    std::memcpy(o->xsz, xsz, sizeof(o->xsz));
    std::memcpy(o->ysz, ysz, sizeof(o->ysz));
    std::memcpy(o->zsz, zsz, sizeof(o->zsz));
    const int n[3] = {xsz[0], ysz[1], zsz[2]};
    const int per[3] = {px, py, pz};
    const char* m[3] = {mx, my, mz};
    for (int a = 0; a < 3; ++a) { o->cf[a] = nullptr; o->ga[a] = nullptr; o->ls[a] = nullptr; }
    for (int a = 0; a < 3; ++a) {
        int rc;
        if (std::strncmp(m[a], "cf90", 4) == 0) {
            o->method[a] = 0;
            rc = pdo_cf90_init(&o->cf[a], n[a], per[a]);
        } else if (std::strncmp(m[a], "gaussian", 8) == 0) {
            o->method[a] = 1;
            rc = pdo_gaussian_init(&o->ga[a], n[a], per[a]);
        } else if (std::strncmp(m[a], "lstsq", 5) == 0) {
            o->method[a] = 2;
            rc = pdo_lstsq_init(&o->ls[a], n[a], per[a]);
        } else {
            rc = fail(52, "Incorrect method select in direction %c", "XYZ"[a]);   // filters.F90:120, 151, 181 ("spectral" is not built)
        }
        if (rc) {
            pdo_filters_destroy(o);
            return rc;
        }
    }
    *h = o;
    return 0;
}
int pdo_filters_destroy(pdo_filters_t h) {
    if (!h) return 0;
    for (int a = 0; a < 3; ++a) { pdo_cf90_destroy(h->cf[a]); pdo_gaussian_destroy(h->ga[a]); pdo_lstsq_destroy(h->ls[a]); }
    delete h;
    return 0;
}
static int fil_apply(pdo_filters_t h, int axis, const double* f, double* out, int bc1, int bcn, void* st) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    const int* sz = axis == 0 ? h->xsz : axis == 1 ? h->ysz : h->zsz;
    int na, nb;
    if (axis == 0) { na = sz[1]; nb = sz[2]; }
    else if (axis == 1) { na = sz[0]; nb = sz[2]; }
    else { na = sz[0]; nb = sz[1]; }
    if (h->method[axis] == 0) {
        typedef int (*fn_t)(pdo_cf90_t, const double*, double*, int, int, int, int, void*);
        static const fn_t fns[3] = {pdo_cf90_filter1, pdo_cf90_filter2, pdo_cf90_filter3};
        return fns[axis](h->cf[axis], f, out, na, nb, bc1, bcn, st);
    }
    if (h->method[axis] == 2) {   // filters.F90:231, 248, 265: the least-squares filter takes no boundary codes
        typedef int (*fnl_t)(pdo_lstsq_t, const double*, double*, int, int, void*);
        static const fnl_t fl[3] = {pdo_lstsq_filter1, pdo_lstsq_filter2, pdo_lstsq_filter3};
        return fl[axis](h->ls[axis], f, out, na, nb, st);
    }
    typedef int (*fng_t)(pdo_gaussian_t, const double*, double*, int, int, int, int, void*);
    static const fng_t fg[3] = {pdo_gaussian_filter1, pdo_gaussian_filter2, pdo_gaussian_filter3};
    return fg[axis](h->ga[axis], f, out, na, nb, bc1, bcn, st);
}
int pdo_filters_filterx(pdo_filters_t h, const double* f, double* o, int b1, int bn, void* s) { return fil_apply(h, 0, f, o, b1, bn, s); }
int pdo_filters_filtery(pdo_filters_t h, const double* f, double* o, int b1, int bn, void* s) { return fil_apply(h, 1, f, o, b1, bn, s); }
int pdo_filters_filterz(pdo_filters_t h, const double* f, double* o, int b1, int bn, void* s) { return fil_apply(h, 2, f, o, b1, bn, s); }

// test hook (not in the public header): run the any-n kernels even when a chunked path exists
}  // extern "C"
namespace pdo { namespace hooks {
int cd10_generic(pdo_cd10_t h, int which, int axis, const double* f, double* df, int na, int nb, void* stream) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    PDO_CUDA(banded_op_apply(which == 1 ? &h->d1 : &h->d2, axis, f, df, na, nb, (cudaStream_t)stream, 1));
    g_launches += 2;
    return 0;
}
}}  // namespace pdo::hooks
extern "C" {

// test hook (not in the public header): force a kernel variant, see banded.cuh
}  // extern "C"
namespace pdo { namespace hooks {
int last_variant(void) { return banded_debug_last_variant(); }
}}  // namespace pdo::hooks
extern "C" {
}  // extern "C"
namespace pdo { namespace hooks {
int set_variant(int strided_mode, int x_threads) {
    banded_debug_set_variant(strided_mode, x_threads);
    return 0;
}
}}  // namespace pdo::hooks
extern "C" {

}  // extern "C"
