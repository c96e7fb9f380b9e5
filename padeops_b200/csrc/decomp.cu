// decomp.cu — 2DECOMP pencil decomposition arithmetic and the four pencil transposes on NCCL.
//
// Replaces (2D» = dependencies/2decomp_fft-1.5.847.tar.gz » 2decomp_fft/src):
//   decomp_2d_init / decomp_info_init / partition / distribute / prepare_buffer   2D» decomp_2d.f90:297-455, 498-580, 622-794
//   transpose_x_to_y / y_to_x / y_to_z / z_to_y (real + complex)                  2D» transpose_*.f90
//   p_maxval / p_sum                                                              utilities/reductions.F90:29-225
// Process model: one process per GPU; rank r sits at (r / p_col, r % p_col) like MPI_CART_CREATE without
// reorder.  The reference's MPI_ALLTOALLV on the COL / ROW sub-communicators becomes one grouped
// ncclSend/ncclRecv exchange addressed by world rank (no communicator split needed), bracketed by a
// single batched pack kernel and a single batched unpack kernel; the block a rank keeps for itself never
// touches NCCL.  With one rank in the sub-communicator the two pencils have the same layout and the
// transpose is a device copy.  Work buffers belong to the decomp handle (the reference's are module
// globals, which is what makes its transposes non-re-entrant).
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

using namespace pdo;

namespace {

// A device buffer every rank registered collectively (same order on all ranks): peer[r] is rank r's copy of it,
// mapped into this process with CUDA IPC.  Registered buffers can be written by peer GPUs over NVLink.
struct SymBuf {
    char* base = nullptr;
    size_t bytes = 0;
    std::vector<char*> peer;
};
struct IpcMapping { cudaIpcMemHandle_t h; char* mapped; };

struct Comm {
    bool inited = false;
    int rank = 0, nproc = 1;
    ncclComm_t comm = nullptr;
    double* d_scalar = nullptr;
    // P2P transposes
    bool p2p = false;
    std::vector<SymBuf> sym;
    std::vector<IpcMapping> maps;
    unsigned long long* flags = nullptr;               // [2][nproc]: entry / exit epochs written by peers
    std::vector<unsigned long long*> peer_flags;       // mapped flags of every rank
    unsigned long long epoch = 0;
    char* d_xchg = nullptr;                            // all-gather staging for registration
};
Comm g_comm;

#define PDO_NCCL(expr)                                                                                          \
    do {                                                                                                        \
        ncclResult_t _r = (expr);                                                                               \
        if (_r != ncclSuccess) return fail(PDO_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_r)); \
    } while (0)

// distribute (2D» decomp_2d.f90:676-708) in closed form: the last `rem` ranks get one extra point.
inline int dist_size(int n, int p, int i) { const int base = n / p, rem = n % p; return base + (i >= p - rem ? 1 : 0); }
inline int dist_start0(int n, int p, int i) {  // 0-based start
    const int base = n / p, rem = n % p;
    const int extra = i - (p - rem);
    return i * base + (extra > 0 ? extra : 0);
}

void fill_info(int nx, int ny, int nz, int p_row, int p_col, int rank, pdo_decomp_info* d) {
    const int c1 = rank / p_col, c2 = rank % p_col;
    auto set = [](int* st, int* en, int* sz, int i, int n, int p, int c) {
        if (p < 0) { st[i] = 1; en[i] = n; sz[i] = n; }
        else { st[i] = dist_start0(n, p, c) + 1; sz[i] = dist_size(n, p, c); en[i] = st[i] + sz[i] - 1; }
    };
    // x-pencil: (nx, ny/p_row, nz/p_col); y-pencil: (nx/p_row, ny, nz/p_col); z-pencil: (nx/p_row, ny/p_col, nz)
    set(d->xst, d->xen, d->xsz, 0, nx, -1, 0); set(d->xst, d->xen, d->xsz, 1, ny, p_row, c1); set(d->xst, d->xen, d->xsz, 2, nz, p_col, c2);
    set(d->yst, d->yen, d->ysz, 0, nx, p_row, c1); set(d->yst, d->yen, d->ysz, 1, ny, -1, 0); set(d->yst, d->yen, d->ysz, 2, nz, p_col, c2);
    set(d->zst, d->zen, d->zsz, 0, nx, p_row, c1); set(d->zst, d->zen, d->zsz, 1, ny, p_col, c2); set(d->zst, d->zen, d->zsz, 2, nz, -1, 0);
}

// ---- batched 3-D box copy: the pack (mem_split_*) and unpack (mem_merge_*) loops of every peer in ONE launch ----
constexpr int kMaxPeers = 16;
struct BoxCopy {
    long long src_off, dst_off;       // element (double) offsets
    long long s_ld2, s_ld3, d_ld2, d_ld3;  // strides of index 2 and 3 on each side (index 1 is contiguous)
    int b1, b2, b3;                   // box extents (b1 in doubles)
};
struct BoxBatch { int count; BoxCopy c[kMaxPeers]; };

template <typename T>
__global__ void __launch_bounds__(256) box_copy_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                       const __grid_constant__ BoxBatch batch) {
    const BoxCopy& c = batch.c[blockIdx.y];
    constexpr int V = sizeof(T) / sizeof(double);
    const int b1v = c.b1 / V;
    const long long rows = (long long)c.b2 * c.b3;
    const T* s = reinterpret_cast<const T*>(src + c.src_off);
    T* d = reinterpret_cast<T*>(dst + c.dst_off);
    for (long long row = (long long)blockIdx.x * blockDim.y + threadIdx.y; row < rows; row += (long long)gridDim.x * blockDim.y) {
        const long long k = row / c.b2;
        const int j = (int)(row - k * c.b2);
        const T* sr = s + (j * c.s_ld2 + k * c.s_ld3) / V;
        T* dr = d + (j * c.d_ld2 + k * c.d_ld3) / V;
        for (int i = threadIdx.x; i < b1v; i += blockDim.x) dr[i] = sr[i];
    }
}

int launch_box_copy(const double* src, double* dst, const BoxBatch& b, cudaStream_t st) {
    if (b.count == 0) return 0;
    bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    long long max_rows = 0;
    int min_b1 = 1 << 30;
    for (int i = 0; i < b.count; ++i) {
        const BoxCopy& c = b.c[i];
        if ((c.b1 | c.src_off | c.dst_off | c.s_ld2 | c.s_ld3 | c.d_ld2 | c.d_ld3) & 1) vec = false;
        const long long rows = (long long)c.b2 * c.b3;
        if (rows > max_rows) max_rows = rows;
        if (c.b1 < min_b1) min_b1 = c.b1;
    }
    if (max_rows == 0 || min_b1 <= 0) return 0;
    const int tx = (min_b1 / (vec ? 2 : 1)) >= 128 ? 128 : ((min_b1 / (vec ? 2 : 1)) >= 64 ? 64 : 32);
    dim3 block(tx, 256 / tx);
    long long gx = (max_rows + block.y - 1) / block.y;
    const long long cap = 148LL * 16;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, (unsigned)b.count);
    if (vec) box_copy_kernel<double2><<<grid, block, 0, st>>>(src, dst, b);
    else box_copy_kernel<double><<<grid, block, 0, st>>>(src, dst, b);
    PDO_CUDA(cudaGetLastError());
    g_launches += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Symmetric buffers (CUDA IPC) and the fused pack + transfer + unpack transposes that use them.
// ------------------------------------------------------------------------------------------------
struct XchgRec { cudaIpcMemHandle_t h; unsigned long long offset; unsigned long long bytes; };

typedef int (*cuMemGetAddressRange_t)(unsigned long long*, size_t*, unsigned long long);
cuMemGetAddressRange_t get_range_fn() {
    static cuMemGetAddressRange_t fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        if (void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL)) fn = (cuMemGetAddressRange_t)dlsym(lib, "cuMemGetAddressRange_v2");
    }
    return fn;
}

// Collective over all ranks.  Returns the registry index, or -1 with p2p left usable for other buffers.
int sym_register(void* ptr, size_t bytes) {
    if (!g_comm.p2p || g_comm.nproc == 1 || !ptr) return -1;
    const int np = g_comm.nproc;
    XchgRec mine;
    std::memset(&mine, 0, sizeof(mine));
    bool ok = true;
    unsigned long long base = 0;
    size_t asz = 0;
    cuMemGetAddressRange_t fn = get_range_fn();
    if (!fn || fn(&base, &asz, (unsigned long long)(uintptr_t)ptr) != 0) ok = false;
    if (ok && cudaIpcGetMemHandle(&mine.h, (void*)(uintptr_t)base) != cudaSuccess) { cudaGetLastError(); ok = false; }
    mine.offset = ok ? (unsigned long long)(uintptr_t)ptr - base : ~0ull;
    mine.bytes = bytes;
    std::vector<XchgRec> all(np);
    if (cudaMemcpy(g_comm.d_xchg, &mine, sizeof(mine), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (ncclAllGather(g_comm.d_xchg, g_comm.d_xchg + sizeof(XchgRec), sizeof(XchgRec), ncclChar, g_comm.comm, 0) != ncclSuccess) return -1;
    if (cudaMemcpy(all.data(), g_comm.d_xchg + sizeof(XchgRec), sizeof(XchgRec) * np, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    for (int r = 0; r < np; ++r) if (all[r].offset == ~0ull) ok = false;  // somebody could not export: nobody uses this buffer
    SymBuf sb;
    sb.base = (char*)ptr; sb.bytes = bytes; sb.peer.assign(np, nullptr);
    for (int r = 0; r < np && ok; ++r) {
        if (r == g_comm.rank) { sb.peer[r] = (char*)ptr; continue; }
        char* mapped = nullptr;
        for (auto& m : g_comm.maps) if (std::memcmp(&m.h, &all[r].h, sizeof(cudaIpcMemHandle_t)) == 0) mapped = m.mapped;
        if (!mapped) {
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
            mapped = (char*)q;
            g_comm.maps.push_back({all[r].h, mapped});
        }
        sb.peer[r] = mapped + all[r].offset;
    }
    // agree on the outcome (an open can fail on one rank only)
    int flag = ok ? 1 : 0, *d = (int*)g_comm.d_xchg;
    cudaMemcpy(d, &flag, sizeof(int), cudaMemcpyHostToDevice);
    ncclAllReduce(d, d + 1, 1, ncclInt, ncclMin, g_comm.comm, 0);
    cudaMemcpy(&flag, d + 1, sizeof(int), cudaMemcpyDeviceToHost);
    if (!flag) return -1;
    g_comm.sym.push_back(sb);
    return (int)g_comm.sym.size() - 1;
}

const SymBuf* sym_find(const void* p, size_t bytes) {
    const char* c = (const char*)p;
    for (const auto& sb : g_comm.sym)
        if (c == sb.base && bytes <= sb.bytes) return &sb;  // exact base: the symmetric-offset rule is then trivially met
    return nullptr;
}

struct PushBatch {
    int count;                  // boxes (one per peer, self included)
    BoxCopy c[kMaxPeers];
    double* dst[kMaxPeers];     // absolute destination base per box (peer memory for the others)
    int npeers;                 // peers to synchronise with (self excluded)
    unsigned long long* peer_flags[kMaxPeers];
    int peer_rank[kMaxPeers];
    unsigned long long* my_flags;
    unsigned int* counter;
    unsigned long long epoch;
    int me, nproc;
};

__device__ __forceinline__ void flag_store(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long flag_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Pack + transfer + unpack in one kernel: every box of `src` is copied to where it belongs in its owner's `dst`
// (peer memory over NVLink, 16-byte stores).  Entry handshake: a rank's dst may be overwritten once that rank's stream
// has reached this call (it announces itself to its peers; everybody waits for everybody in the group).  Exit: the
// last CTA to finish publishes the epoch to the peers after a system-scope fence; box_wait_kernel (next in the stream)
// holds the consumer until every peer has published.
template <typename T>
__global__ void __launch_bounds__(256) box_push_kernel(const double* __restrict__ src, const __grid_constant__ PushBatch b) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid < b.npeers) flag_store(b.peer_flags[tid] + b.me, b.epoch);
    if (tid < b.npeers) {
        const unsigned long long* f = b.my_flags + b.peer_rank[tid];
        while (flag_load(f) < b.epoch) { }
    }
    __syncthreads();
    const BoxCopy& c = b.c[blockIdx.y];
    constexpr int V = sizeof(T) / sizeof(double);
    const int b1v = c.b1 / V;
    const long long rows = (long long)c.b2 * c.b3;
    const T* s = reinterpret_cast<const T*>(src + c.src_off);
    T* d = reinterpret_cast<T*>(b.dst[blockIdx.y] + c.dst_off);
    for (long long row = (long long)blockIdx.x * blockDim.y + threadIdx.y; row < rows; row += (long long)gridDim.x * blockDim.y) {
        const long long k = row / c.b2;
        const int j = (int)(row - k * c.b2);
        const T* sr = s + (j * c.s_ld2 + k * c.s_ld3) / V;
        T* dr = d + (j * c.d_ld2 + k * c.d_ld3) / V;
        int i = threadIdx.x;
        for (; i + 3 * (int)blockDim.x < b1v; i += 4 * blockDim.x) {  // four independent loads in flight per thread
            const T v0 = sr[i], v1 = sr[i + blockDim.x], v2 = sr[i + 2 * blockDim.x], v3 = sr[i + 3 * blockDim.x];
            dr[i] = v0; dr[i + blockDim.x] = v1; dr[i + 2 * blockDim.x] = v2; dr[i + 3 * blockDim.x] = v3;
        }
        for (; i < b1v; i += blockDim.x) dr[i] = sr[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        if (atomicAdd(b.counter, 1u) == total - 1) {
            *b.counter = 0;
            __threadfence_system();
            for (int t = 0; t < b.npeers; ++t) flag_store(b.peer_flags[t] + b.nproc + b.me, b.epoch);
        }
    }
}

// copy-engine variant: the data moves with cudaMemcpy2D/3DAsync; these two kernels are the handshakes around it
__global__ void box_entry_kernel(const __grid_constant__ PushBatch b) {
    const int tid = threadIdx.x;
    if (tid < b.npeers) {
        flag_store(b.peer_flags[tid] + b.me, b.epoch);
        const unsigned long long* f = b.my_flags + b.peer_rank[tid];
        while (flag_load(f) < b.epoch) { }
    }
}
__global__ void box_exit_kernel(const __grid_constant__ PushBatch b) {
    const int tid = threadIdx.x;
    if (tid < b.npeers) {
        __threadfence_system();
        flag_store(b.peer_flags[tid] + b.nproc + b.me, b.epoch);
        const unsigned long long* f = b.my_flags + b.nproc + b.peer_rank[tid];
        while (flag_load(f) < b.epoch) { }
    }
}

__global__ void box_wait_kernel(const __grid_constant__ PushBatch b) {
    const int tid = threadIdx.x;
    if (tid < b.npeers) {
        const unsigned long long* f = b.my_flags + b.nproc + b.peer_rank[tid];
        while (flag_load(f) < b.epoch) { }
    }
}


// ------------------------------------------------------------------------------------------------
// TMA bulk push ("bulk" data plane of the fused transposes).  A persistent CTA streams 32 KB pieces of its source boxes
// through a ring of shared-memory stages: cp.async.bulk global -> shared (one elected thread, transaction mbarrier per
// stage), then cp.async.bulk shared -> the OWNER's dst in peer memory over NVLink.  Every piece is a run that is contiguous
// on both sides (2DECOMP's blocks are contiguous for whole x-rows, for whole (x, y) planes on the y<->z pair), so the
// fabric sees long bursts instead of per-thread 16-byte stores, and no thread touches the data.  Work items interleave the
// destination ranks (item i -> peer order[i % np]) so that all links carry traffic all the time.
// ------------------------------------------------------------------------------------------------
constexpr int kBulkStages = 6;
constexpr int kBulkAhead = 4;                 // loads in flight per CTA; kBulkStages - kBulkAhead stores may still be reading
constexpr unsigned kBulkChunk = 32768;        // bytes per stage

struct BulkBox {
    const char* src; char* dst;               // absolute byte addresses of the box origin on each side
    long long s_ld2, s_ld3, d_ld2, d_ld3;     // byte strides of the piece indices (inner, outer)
    long long piece_bytes;                    // contiguous run on both sides
    int n_inner, n_outer;                     // pieces = n_inner x n_outer
    int cpp;                                  // chunks per piece (long pieces are cut into kBulkChunk-byte chunks) ...
    int ppc;                                  // ... or pieces per chunk (short pieces, e.g. 4 KB x-rows, share a stage)
    int cpo;                                  // chunks per outer index when ppc > 1
    long long nchunks;
};
struct BulkBatch {
    int count;
    BulkBox c[kMaxPeers];
    long long max_chunks;
    int npeers;
    unsigned long long* peer_flags[kMaxPeers];
    int peer_rank[kMaxPeers];
    unsigned long long* my_flags;
    unsigned int* counter;
    unsigned long long epoch;
    int me, nproc;
};

__device__ __forceinline__ unsigned d_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) bulk_push_kernel(const __grid_constant__ BulkBatch b) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) unsigned long long full_bar[kBulkStages];
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && tid < b.npeers) flag_store(b.peer_flags[tid] + b.me, b.epoch);
    if (tid < b.npeers) {
        const unsigned long long* f = b.my_flags + b.peer_rank[tid];
        while (flag_load(f) < b.epoch) { }
    }
    if (tid == 0) {
        for (int s = 0; s < kBulkStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(d_smem_u32(&full_bar[s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const long long total = b.max_chunks * b.count;     // item -> (box = item % count, chunk = item / count)
        // item -> the pieces of one stage: `np` runs of `pb` bytes each, strides (ss, ds) apart, starting at (src, dst)
        auto locate = [&](long long item, const char*& src, char*& dst, unsigned& pb, int& np, long long& ss, long long& ds) -> bool {
            const BulkBox& c = b.c[(int)(item % b.count)];
            const long long ch = item / b.count;
            if (ch >= c.nchunks) return false;
            if (c.ppc > 1) {                                   // several short pieces (consecutive inner indices) per stage
                const long long po = ch / c.cpo;
                const long long pi = (ch - po * c.cpo) * c.ppc;
                const long long left = c.n_inner - pi;
                np = (int)(left < c.ppc ? left : c.ppc);
                pb = (unsigned)c.piece_bytes;
                ss = c.s_ld2; ds = c.d_ld2;
                src = c.src + pi * c.s_ld2 + po * c.s_ld3;
                dst = c.dst + pi * c.d_ld2 + po * c.d_ld3;
                return true;
            }
            const long long piece = ch / c.cpp;
            const long long sub = ch - piece * c.cpp;
            const long long po = piece / c.n_inner, pi = piece - po * c.n_inner;
            const long long off = sub * (long long)kBulkChunk;
            const long long left = c.piece_bytes - off;
            pb = (unsigned)(left < (long long)kBulkChunk ? left : (long long)kBulkChunk);
            np = 1; ss = ds = 0;
            src = c.src + pi * c.s_ld2 + po * c.s_ld3 + off;
            dst = c.dst + pi * c.d_ld2 + po * c.d_ld3 + off;
            return true;
        };
        // my items: blockIdx.x, blockIdx.x + gridDim.x, ...; slot numbers count the items that exist
        long long li = blockIdx.x, si = blockIdx.x;       // next item to load / to store
        long long nl = 0, ns = 0;                          // slots loaded / stored so far
        struct Item { const char* s; char* d; unsigned pb; int np; long long ss, ds; };
        auto next_valid = [&](long long& it, Item& x) -> bool {
            while (it < total) {
                if (locate(it, x.s, x.d, x.pb, x.np, x.ss, x.ds)) return true;
                it += gridDim.x;
            }
            return false;
        };
        Item L, S;
        bool more_l = next_valid(li, L);
        bool more_s = next_valid(si, S);
        while (more_s) {
            // keep kBulkAhead loads in flight
            while (more_l && nl < ns + kBulkAhead) {
                const int st = (int)(nl % kBulkStages);
                // the store that last used this stage (slot nl - kBulkStages) must have read its shared source: every store
                // commits its own group and at least kBulkStages - kBulkAhead groups are newer than that one
                if (nl >= kBulkStages) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kBulkStages - kBulkAhead) : "memory");
                const unsigned bar = d_smem_u32(&full_bar[st]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(L.pb * (unsigned)L.np) : "memory");
                for (int j = 0; j < L.np; ++j)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(d_smem_u32(ring + (size_t)st * kBulkChunk + (size_t)j * L.pb)), "l"(L.s + j * L.ss), "r"(L.pb), "r"(bar) : "memory");
                ++nl;
                li += gridDim.x;
                more_l = next_valid(li, L);
            }
            const int st = (int)(ns % kBulkStages);
            const unsigned bar = d_smem_u32(&full_bar[st]);
            const unsigned parity = (unsigned)((ns / kBulkStages) & 1);
            unsigned ok = 0;
            for (unsigned spin = 0; !ok; ++spin) {
                asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                if (spin > (1u << 26)) __trap();
            }
            for (int j = 0; j < S.np; ++j)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(S.d + j * S.ds),
                             "r"(d_smem_u32(ring + (size_t)st * kBulkChunk + (size_t)j * S.pb)), "r"(S.pb) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            ++ns;
            si += gridDim.x;
            more_s = next_valid(si, S);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every write of this CTA has been performed
        __threadfence_system();
        if (atomicAdd(b.counter, 1u) == gridDim.x - 1) {
            *b.counter = 0;
            __threadfence_system();
            for (int t = 0; t < b.npeers; ++t) flag_store(b.peer_flags[t] + b.nproc + b.me, b.epoch);
        }
    }
}

}  // namespace

struct pdo_decomp_s {
    int nx, ny, nz, p_row, p_col, c1, c2;
    int world_rank = 0;
    unsigned int* push_counter = nullptr;
    cudaStream_t ce_stream[kMaxPeers] = {};   // copy-engine variant of the fused path: one stream per peer
    cudaEvent_t ce_fork = nullptr, ce_join[kMaxPeers] = {};
    bool ce_ready = false;
    pdo_decomp_info info;
    std::vector<int> x1dist, y1dist, y2dist, z2dist;
    double* work_send = nullptr;
    double* work_recv = nullptr;
    size_t work_cap = 0;  // doubles
};

namespace {

int ensure_work(pdo_decomp_s* d, size_t doubles) {
    if (d->work_cap >= doubles) return 0;
    if (d->work_send) cudaFree(d->work_send);
    if (d->work_recv) cudaFree(d->work_recv);
    d->work_send = d->work_recv = nullptr;
    d->work_cap = 0;
    PDO_CUDA(cudaMalloc(&d->work_send, doubles * sizeof(double)));
    PDO_CUDA(cudaMalloc(&d->work_recv, doubles * sizeof(double)));
    d->work_cap = doubles;
    return 0;
}

inline long long vol3(const int* s) { return (long long)s[0] * s[1] * s[2]; }

pdo_decomp_s* decomp_new(int nx, int ny, int nz, int p_row, int p_col, int rank) {
    pdo_decomp_s* d = new (std::nothrow) pdo_decomp_s();
    if (!d) return nullptr;
    d->nx = nx; d->ny = ny; d->nz = nz; d->p_row = p_row; d->p_col = p_col;
    d->world_rank = rank;
    d->c1 = rank / p_col; d->c2 = rank % p_col;
    fill_info(nx, ny, nz, p_row, p_col, rank, &d->info);
    for (int i = 0; i < p_row; ++i) { d->x1dist.push_back(dist_size(nx, p_row, i)); d->y1dist.push_back(dist_size(ny, p_row, i)); }
    for (int i = 0; i < p_col; ++i) { d->y2dist.push_back(dist_size(ny, p_col, i)); d->z2dist.push_back(dist_size(nz, p_col, i)); }
    return d;
}

// Geometry of one transpose as seen from one rank: 2DECOMP's counts / displacements (prepare_buffer, 2D» decomp_2d.f90:739-794)
// and the boxes of the pack (mem_split_*) and unpack (mem_merge_*) loops (2D» transpose_x_to_y.f90:332-513 and siblings).
// dir: 0 x→y, 1 y→x, 2 y→z, 3 z→y.  w = doubles per element.  Pure host arithmetic: the same object drives the real exchange
// and the single-GPU emulation of a p_row x p_col grid.
struct XGeom {
    const pdo_decomp_s* d;
    int dir, w, np, me;
    bool col;
    long long s1, s2, s3, d1, d2, d3;
    std::vector<int> sdist, rdist;
    std::vector<long long> scnt, sdisp, rcnt, rdisp, sst, rst;
    long long stot = 0, rtot = 0;
    XGeom(const pdo_decomp_s* dd, int dir_, int w_) : d(dd), dir(dir_), w(w_) {
        col = (dir == 0 || dir == 1);
        np = col ? d->p_row : d->p_col;
        me = col ? d->c1 : d->c2;
        const int* ssz = (dir == 0) ? d->info.xsz : (dir == 3) ? d->info.zsz : d->info.ysz;
        const int* dsz = (dir == 1) ? d->info.xsz : (dir == 2) ? d->info.zsz : d->info.ysz;
        s1 = (long long)ssz[0] * w; s2 = ssz[1]; s3 = ssz[2];
        d1 = (long long)dsz[0] * w; d2 = dsz[1]; d3 = dsz[2];
        sdist = (dir == 0) ? d->x1dist : (dir == 3) ? d->z2dist : (dir == 1) ? d->y1dist : d->y2dist;
        rdist = (dir == 0) ? d->y1dist : (dir == 3) ? d->y2dist : (dir == 1) ? d->x1dist : d->z2dist;
        // send block to peer m = a range of SRC index (1 for x→y, 2 for y→x / y→z, 3 for z→y);
        // recv block from peer m = a range of DST index (2 for x→y / z→y, 1 for y→x, 3 for y→z).
        scnt.resize(np); sdisp.resize(np); rcnt.resize(np); rdisp.resize(np); sst.resize(np); rst.resize(np);
        long long sa = 0, ra = 0;
        for (int m = 0; m < np; ++m) {
            sst[m] = sa; rst[m] = ra;
            scnt[m] = (dir == 0) ? (long long)sdist[m] * w * s2 * s3 : (dir == 3) ? s1 * s2 * sdist[m] : s1 * sdist[m] * s3;
            rcnt[m] = (dir == 1) ? (long long)rdist[m] * w * d2 * d3 : (dir == 2) ? d1 * d2 * rdist[m] : d1 * rdist[m] * d3;
            sdisp[m] = stot; rdisp[m] = rtot;
            stot += scnt[m]; rtot += rcnt[m];
            sa += sdist[m]; ra += rdist[m];
        }
    }
    int world_rank(int m) const { return col ? (m * d->p_col + d->c2) : (d->c1 * d->p_col + m); }
    void src_box(int m, BoxCopy& c) const {  // where peer m's block lives in src
        c.s_ld2 = s1; c.s_ld3 = s1 * s2;
        if (dir == 0) { c.src_off = sst[m] * w; c.b1 = sdist[m] * w; c.b2 = (int)s2; c.b3 = (int)s3; }
        else if (dir == 3) { c.src_off = sst[m] * s1 * s2; c.b1 = (int)s1; c.b2 = (int)s2; c.b3 = sdist[m]; }
        else { c.src_off = sst[m] * s1; c.b1 = (int)s1; c.b2 = sdist[m]; c.b3 = (int)s3; }
    }
    void dst_box(int m, BoxCopy& c) const {  // where the block from peer m lands in my dst
        c.d_ld2 = d1; c.d_ld3 = d1 * d2;
        if (dir == 1) { c.dst_off = rst[m] * w; c.b1 = rdist[m] * w; c.b2 = (int)d2; c.b3 = (int)d3; }
        else if (dir == 2) { c.dst_off = rst[m] * d1 * d2; c.b1 = (int)d1; c.b2 = (int)d2; c.b3 = rdist[m]; }
        else { c.dst_off = rst[m] * d1; c.b1 = (int)d1; c.b2 = rdist[m]; c.b3 = (int)d3; }
    }
    void push_box(int m, BoxCopy& c) const {  // my block for peer m and where it lands in RANK m's dst (its extents differ from mine)
        src_box(m, c);
        if (dir == 0) { c.d_ld2 = (long long)sdist[m] * w; c.d_ld3 = c.d_ld2 * d->ny; c.dst_off = rst[me] * c.d_ld2; }
        else if (dir == 1) { c.d_ld2 = (long long)d->nx * w; c.d_ld3 = c.d_ld2 * sdist[m]; c.dst_off = rst[me] * w; }
        else if (dir == 2) { c.d_ld2 = s1; c.d_ld3 = s1 * sdist[m]; c.dst_off = rst[me] * c.d_ld3; }
        else { c.d_ld2 = s1; c.d_ld3 = s1 * d->ny; c.dst_off = rst[me] * s1; }
    }
};

// Who the peers are for one call: the real thing (CUDA-IPC mappings + epoch flags) or, in the single-GPU emulation, the
// other simulated ranks' buffers on this device with the handshakes switched off (the emulation orders ranks by stream).
struct PeerView {
    double* const* dst_of_world;          // dst base of every world rank (byte offset already applied)
    bool handshake;
};

enum { kPlaneAuto = 0, kPlaneSm = 1, kPlaneCe = 2, kPlaneBulk = 3 };

int plane_from_env() {
    static int mode = -1;  // PDO_P2P_MODE = sm (store kernel) | ce (copy engines) | bulk (TMA bulk push); default: bulk when aligned
    if (mode < 0) {
        const char* e = std::getenv("PDO_P2P_MODE");
        mode = !e ? kPlaneAuto : !std::strcmp(e, "sm") ? kPlaneSm : !std::strcmp(e, "ce") ? kPlaneCe : !std::strcmp(e, "bulk") ? kPlaneBulk : kPlaneAuto;
    }
    return mode;
}

template <class B>
void fill_sync(B& b, const XGeom& g, const PeerView& pv, pdo_decomp_s* d, unsigned long long epoch) {
    b.npeers = 0; b.me = g_comm.rank; b.nproc = g_comm.nproc;
    b.my_flags = g_comm.flags; b.counter = d->push_counter; b.epoch = epoch;
    if (!pv.handshake) return;
    for (int m = 0; m < g.np; ++m) {
        if (m == g.me) continue;
        const int pw = g.world_rank(m);
        b.peer_flags[b.npeers] = g_comm.peer_flags[pw];
        b.peer_rank[b.npeers] = pw;
        b.npeers++;
    }
}

// Fused path: every rank stores its blocks straight into their final place in the owners' dst.  No pack buffer, no unpack
// pass, no NCCL.  `plane` picks who moves the bytes.
int transpose_push(pdo_decomp_s* d, const XGeom& g, const double* src, const PeerView& pv, int plane, cudaStream_t st) {
    const int np = g.np, me = g.me;
    if (!d->push_counter) {
        PDO_CUDA(cudaMalloc(&d->push_counter, sizeof(unsigned int)));
        PDO_CUDA(cudaMemset(d->push_counter, 0, sizeof(unsigned int)));
    }
    PushBatch b;
    std::memset(&b, 0, sizeof(b));
    b.count = np;
    const unsigned long long epoch = pv.handshake ? ++g_comm.epoch : 0ull;
    fill_sync(b, g, pv, d, epoch);
    bool vec = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    long long max_rows = 0, max_box_bytes = 0;
    int min_b1 = 1 << 30;
    for (int m = 0; m < np; ++m) {
        BoxCopy& c = b.c[m];
        g.push_box(m, c);
        b.dst[m] = pv.dst_of_world[g.world_rank(m)];
        if ((c.b1 | c.src_off | c.dst_off | c.s_ld2 | c.s_ld3 | c.d_ld2 | c.d_ld3) & 1) vec = false;
        if (reinterpret_cast<uintptr_t>(b.dst[m]) & 15) vec = false;
        const long long rows = (long long)c.b2 * c.b3;
        if (rows > max_rows) max_rows = rows;
        if (c.b1 < min_b1) min_b1 = c.b1;
        if (m != me && rows * c.b1 * 8 > max_box_bytes) max_box_bytes = rows * c.b1 * 8;
    }
    if (plane == kPlaneAuto) plane = plane_from_env();
    // measured at 1024^3 (profiles/r02o_transposes_8gpu.jsonl, r02m_transposes_2gpu.jsonl): with 4 or 8 peers the TMA bulk push
    // keeps every link busy (0.70-0.75 of NVLink on 1x8 and 2x4, copy engines 0.45-0.68, store kernel 0.37-0.67); between two
    // GPUs one copy-engine stream per direction already runs at 0.85 (bulk 0.77)
    if (plane == kPlaneAuto) plane = (max_box_bytes < (1LL << 20)) ? kPlaneSm : ((vec && np >= 3) ? kPlaneBulk : kPlaneCe);
    if (plane == kPlaneBulk && !vec) plane = kPlaneSm;    // bulk copies need 16-byte aligned runs
    if (plane == kPlaneBulk) {
        BulkBatch bb;
        std::memset(&bb, 0, sizeof(bb));
        bb.count = np;
        fill_sync(bb, g, pv, d, epoch);
        long long all_chunks = 0;
        for (int q = 0; q < np; ++q) {
            const int m = (me + 1 + q) % np;               // start with my right neighbour: spreads the ingress over the destinations
            const BoxCopy& c = b.c[m];
            BulkBox& x = bb.c[q];
            x.src = (const char*)(src + c.src_off);
            x.dst = (char*)(b.dst[m] + c.dst_off);
            const bool rows_join = (c.b1 == c.s_ld2 && c.b1 == c.d_ld2);             // (i, j) collapse into one run per k
            const bool planes_join = rows_join && c.s_ld3 == (long long)c.b1 * c.b2 && c.d_ld3 == c.s_ld3;  // the whole box is one run
            if (planes_join) { x.piece_bytes = 8LL * c.b1 * c.b2 * c.b3; x.n_inner = 1; x.n_outer = 1; x.s_ld2 = x.d_ld2 = x.s_ld3 = x.d_ld3 = 0; }
            else if (rows_join) { x.piece_bytes = 8LL * c.b1 * c.b2; x.n_inner = 1; x.n_outer = c.b3; x.s_ld2 = x.d_ld2 = 0; x.s_ld3 = 8 * c.s_ld3; x.d_ld3 = 8 * c.d_ld3; }
            else { x.piece_bytes = 8LL * c.b1; x.n_inner = c.b2; x.n_outer = c.b3; x.s_ld2 = 8 * c.s_ld2; x.d_ld2 = 8 * c.d_ld2; x.s_ld3 = 8 * c.s_ld3; x.d_ld3 = 8 * c.d_ld3; }
            x.ppc = 1; x.cpo = 1;
            if (x.piece_bytes <= 0 || c.b2 <= 0 || c.b3 <= 0) { x.piece_bytes = 0; x.nchunks = 0; x.cpp = 1; x.n_inner = x.n_outer = 1; continue; }
            x.cpp = (int)((x.piece_bytes + kBulkChunk - 1) / kBulkChunk);
            x.nchunks = (long long)x.n_inner * x.n_outer * x.cpp;
            if (x.piece_bytes * 2 <= (long long)kBulkChunk && x.n_inner > 1) {   // short rows: several consecutive ones share a stage
                x.ppc = (int)((long long)kBulkChunk / x.piece_bytes);
                if (x.ppc > x.n_inner) x.ppc = x.n_inner;
                x.cpo = (x.n_inner + x.ppc - 1) / x.ppc;
                x.cpp = 1;
                x.nchunks = (long long)x.cpo * x.n_outer;
            }
            if (x.nchunks > bb.max_chunks) bb.max_chunks = x.nchunks;
            all_chunks += x.nchunks;
        }
        static int smem_set = 0;
        const int smem = kBulkStages * (int)kBulkChunk;
        if (!smem_set) {
            PDO_CUDA(cudaFuncSetAttribute(bulk_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            smem_set = 1;
        }
        static int bulk_ctas = -1;
        if (bulk_ctas < 0) { const char* e = std::getenv("PDO_BULK_CTAS"); bulk_ctas = e ? std::atoi(e) : 148; if (bulk_ctas < 1) bulk_ctas = 148; }
        long long grid = all_chunks < bulk_ctas ? (all_chunks > 0 ? all_chunks : 1) : bulk_ctas;
        bulk_push_kernel<<<(unsigned)grid, 128, smem, st>>>(bb);
        PDO_CUDA(cudaGetLastError());
        if (pv.handshake) { box_wait_kernel<<<1, 32, 0, st>>>(b); PDO_CUDA(cudaGetLastError()); }
        g_launches += 2;
        return 0;
    }
    if (plane == kPlaneCe) {
        if (!d->ce_ready) {
            for (int m = 0; m < kMaxPeers; ++m) {
                PDO_CUDA(cudaStreamCreateWithFlags(&d->ce_stream[m], cudaStreamNonBlocking));
                PDO_CUDA(cudaEventCreateWithFlags(&d->ce_join[m], cudaEventDisableTiming));
            }
            PDO_CUDA(cudaEventCreateWithFlags(&d->ce_fork, cudaEventDisableTiming));
            d->ce_ready = true;
        }
        if (pv.handshake) { box_entry_kernel<<<1, 32, 0, st>>>(b); PDO_CUDA(cudaGetLastError()); }
        PDO_CUDA(cudaEventRecord(d->ce_fork, st));
        for (int q = 0; q < np; ++q) {
            const int m = (me + 1 + q) % np;
            const BoxCopy& c = b.c[m];
            if (c.b1 <= 0 || c.b2 <= 0 || c.b3 <= 0) continue;
            cudaStream_t cs = d->ce_stream[m];
            PDO_CUDA(cudaStreamWaitEvent(cs, d->ce_fork, 0));
            const double* sp = src + c.src_off;
            double* dp = b.dst[m] + c.dst_off;
            if (c.b1 == c.s_ld2 && c.b1 == c.d_ld2) {  // rows are contiguous on both sides: (i, j) collapse into wide rows
                PDO_CUDA(cudaMemcpy2DAsync(dp, (size_t)c.d_ld3 * 8, sp, (size_t)c.s_ld3 * 8, (size_t)c.b1 * c.b2 * 8, (size_t)c.b3,
                                           cudaMemcpyDeviceToDevice, cs));
            } else {
                cudaMemcpy3DParms pr;
                std::memset(&pr, 0, sizeof(pr));
                pr.srcPtr = make_cudaPitchedPtr((void*)sp, (size_t)c.s_ld2 * 8, (size_t)c.s_ld2, (size_t)(c.s_ld3 / c.s_ld2));
                pr.dstPtr = make_cudaPitchedPtr((void*)dp, (size_t)c.d_ld2 * 8, (size_t)c.d_ld2, (size_t)(c.d_ld3 / c.d_ld2));
                pr.extent = make_cudaExtent((size_t)c.b1 * 8, (size_t)c.b2, (size_t)c.b3);
                pr.kind = cudaMemcpyDeviceToDevice;
                PDO_CUDA(cudaMemcpy3DAsync(&pr, cs));
            }
            PDO_CUDA(cudaEventRecord(d->ce_join[m], cs));
            PDO_CUDA(cudaStreamWaitEvent(st, d->ce_join[m], 0));
        }
        if (pv.handshake) { box_exit_kernel<<<1, 32, 0, st>>>(b); PDO_CUDA(cudaGetLastError()); }
        g_launches += 2;
        return 0;
    }
    const int tx = (min_b1 / (vec ? 2 : 1)) >= 128 ? 128 : ((min_b1 / (vec ? 2 : 1)) >= 64 ? 64 : 32);
    dim3 block(tx, 256 / tx);
    long long gx = (max_rows + block.y - 1) / block.y;
    const long long cap = (148LL * 8 + np - 1) / np;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)np);
    if (vec) box_push_kernel<double2><<<grid, block, 0, st>>>(src, b);
    else box_push_kernel<double><<<grid, block, 0, st>>>(src, b);
    PDO_CUDA(cudaGetLastError());
    if (pv.handshake) { box_wait_kernel<<<1, 32, 0, st>>>(b); PDO_CUDA(cudaGetLastError()); }
    g_launches += 2;
    return 0;
}

// NCCL path, stage 1: my own block goes src box → dst box directly; the other peers' blocks are packed contiguously, in peer
// order, at the ALLTOALLV displacements (y→z's z-side and z→y's z-side need no pack / unpack: the block is a contiguous k-slab).
int transpose_pack(pdo_decomp_s* d, const XGeom& g, const double* src, double* dst, cudaStream_t st) {
    if (int rc = ensure_work(d, (size_t)((g.stot > g.rtot ? g.stot : g.rtot) + 2))) return rc;
    {
        BoxBatch b{};
        b.count = 1;
        g.src_box(g.me, b.c[0]);
        const int sb1 = b.c[0].b1, sb2 = b.c[0].b2, sb3 = b.c[0].b3;
        g.dst_box(g.me, b.c[0]);
        if (sb1 != b.c[0].b1 || sb2 != b.c[0].b2 || sb3 != b.c[0].b3) return fail(PDO_E_BADARG, "transpose: self block mismatch");
        if (int rc = launch_box_copy(src, dst, b, st)) return rc;
    }
    if (g.dir != 3) {
        BoxBatch b{};
        for (int m = 0; m < g.np; ++m) {
            if (m == g.me) continue;
            BoxCopy& c = b.c[b.count++];
            g.src_box(m, c);
            c.dst_off = g.sdisp[m]; c.d_ld2 = c.b1; c.d_ld3 = (long long)c.b1 * c.b2;
        }
        if (int rc = launch_box_copy(src, d->work_send, b, st)) return rc;
    }
    return 0;
}
inline const double* xchg_send_ptr(const pdo_decomp_s* d, const XGeom& g, const double* src, int m) {
    return g.dir != 3 ? d->work_send + g.sdisp[m] : src + g.sst[m] * g.s1 * g.s2;
}
inline double* xchg_recv_ptr(const pdo_decomp_s* d, const XGeom& g, double* dst, int m) {
    return g.dir != 2 ? d->work_recv + g.rdisp[m] : dst + g.rst[m] * g.d1 * g.d2;
}
int transpose_unpack(pdo_decomp_s* d, const XGeom& g, double* dst, cudaStream_t st) {
    if (g.dir == 2) return 0;
    BoxBatch b{};
    for (int m = 0; m < g.np; ++m) {
        if (m == g.me) continue;
        BoxCopy& c = b.c[b.count++];
        g.dst_box(m, c);
        c.src_off = g.rdisp[m]; c.s_ld2 = c.b1; c.s_ld3 = (long long)c.b1 * c.b2;
    }
    return launch_box_copy(d->work_recv, dst, b, st);
}

// dir: 0 x→y, 1 y→x, 2 y→z, 3 z→y.  Device pointers.  w = doubles per element.
int transpose_device(pdo_decomp_s* d, int dir, const double* src, double* dst, int w, cudaStream_t st) {
    const XGeom g(d, dir, w);
    const int np = g.np, me = g.me;
    if (np == 1) {  // same layout on both sides
        PDO_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * g.s1 * g.s2 * g.s3, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    if (np > kMaxPeers) return fail(PDO_E_UNSUPPORTED, "more than %d ranks in one sub-communicator", kMaxPeers);
    // Fused path: dst is a registered (IPC-shared) buffer on every rank of the group
    if (g_comm.p2p) {
        if (const SymBuf* sb = sym_find(dst, sizeof(double) * (size_t)(g.d1 * g.d2 * g.d3))) {
            const size_t off = (const char*)dst - sb->base;
            std::vector<double*> peers(g_comm.nproc);
            for (int r = 0; r < g_comm.nproc; ++r) peers[r] = (double*)(sb->peer[r] + off);
            PeerView pv{peers.data(), true};
            return transpose_push(d, g, src, pv, kPlaneAuto, st);
        }
    }
    if (int rc = transpose_pack(d, g, src, dst, st)) return rc;
    PDO_NCCL(ncclGroupStart());
    for (int m = 0; m < np; ++m) {
        if (m == me) continue;
        const int peer = g.world_rank(m);
        PDO_NCCL(ncclSend(xchg_send_ptr(d, g, src, m), (size_t)g.scnt[m], ncclDouble, peer, g_comm.comm, st));
        PDO_NCCL(ncclRecv(xchg_recv_ptr(d, g, dst, m), (size_t)g.rcnt[m], ncclDouble, peer, g_comm.comm, st));
    }
    PDO_NCCL(ncclGroupEnd());
    g_launches += 1;
    return transpose_unpack(d, g, dst, st);
}

int transpose_any(pdo_decomp_t h, int dir, const double* src, double* dst, int w, void* stream) {
    if (!h || !src || !dst) return fail(PDO_E_BADARG, "null argument");
    if (w != 1 && w != 2) return fail(PDO_E_BADARG, "elem_doubles must be 1 (real) or 2 (complex)");
    cudaStream_t st = (cudaStream_t)stream;
    const int* ssz = (dir == 0) ? h->info.xsz : (dir == 3) ? h->info.zsz : h->info.ysz;
    const int* dsz = (dir == 1) ? h->info.xsz : (dir == 2) ? h->info.zsz : h->info.ysz;
    return with_device_views(src, sizeof(double) * vol3(ssz) * w, dst, sizeof(double) * vol3(dsz) * w, st,
                             [&](const void* ds, void* dd) { return transpose_device(h, dir, (const double*)ds, (double*)dd, w, st); });
}

}  // namespace

namespace pdo {
// Collective: allocates `bytes` of zeroed device memory on every rank and maps every rank's copy into this process.
// peers[r] (r = world rank) is rank r's copy; peers[me] == *local.  Returns 0, or -1 when peer access is unavailable
// (the allocation is freed again, *local = nullptr) — identically on all ranks.
// Symmetric allocations are pooled until pdo_comm_finalize: a peer keeps its CUDA-IPC mapping of a buffer open, and
// freeing memory that is still mapped elsewhere (then getting the same address back from cudaMalloc) is a hazard.
// alloc / free are collective and happen in the same order on every rank, so every rank picks the same pool entry.
struct SymPoolEntry { void* base; size_t bytes; std::vector<void*> peers; bool used; };
std::vector<SymPoolEntry> g_sym_pool;

static int comm_barrier() {
    int* d = (int*)g_comm.d_xchg;
    if (ncclAllReduce(d, d + 1, 1, ncclInt, ncclMin, g_comm.comm, 0) != ncclSuccess) return -1;
    return cudaStreamSynchronize(0) == cudaSuccess ? 0 : -1;
}

int comm_sym_alloc(size_t bytes, void** local, void** peers) {
    *local = nullptr;
    if (!g_comm.inited || g_comm.nproc == 1 || !g_comm.p2p) return -1;
    {   // pencils are uneven: agree on the largest request so that every rank takes the same pool decision
        unsigned long long mine = bytes, all = 0;
        unsigned long long* d = (unsigned long long*)g_comm.d_xchg;
        if (cudaMemcpy(d, &mine, sizeof(mine), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
        if (ncclAllReduce(d, d + 1, 1, ncclUint64, ncclMax, g_comm.comm, 0) != ncclSuccess) return -1;
        if (cudaMemcpy(&all, d + 1, sizeof(all), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        bytes = (size_t)((all + 255) & ~255ull);
    }
    SymPoolEntry* best = nullptr;
    for (auto& e : g_sym_pool)
        if (!e.used && e.bytes >= bytes && (!best || e.bytes < best->bytes)) best = &e;
    if (best) {
        SymPoolEntry& e = *best;
        if (cudaMemset(e.base, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); return -1; }
        if (comm_barrier() != 0) return -1;   // nobody touches a peer's copy before every copy has been cleared
        e.used = true;
        if (peers) for (int r = 0; r < g_comm.nproc; ++r) peers[r] = e.peers[r];
        *local = e.base;
        return 0;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); p = nullptr; }
    if (p) { cudaMemset(p, 0, bytes); cudaDeviceSynchronize(); }
    if (!p) return -1;  // sizes are symmetric: an allocation failure is taken to be symmetric too
    const int id = sym_register(p, bytes);
    if (id < 0) { cudaFree(p); return -1; }
    SymPoolEntry e;
    e.base = p; e.bytes = bytes; e.used = true;
    e.peers.assign(g_comm.sym[id].peer.begin(), g_comm.sym[id].peer.end());
    if (peers) for (int r = 0; r < g_comm.nproc; ++r) peers[r] = e.peers[r];
    g_sym_pool.push_back(e);
    *local = p;
    return 0;
}
void comm_sym_free(void* p) {
    for (auto& e : g_sym_pool) if (e.base == p) e.used = false;
}
// Library-owned buffers that peers write into (transpose destinations of fft_3d, spectral, padepoisson, igrid).  With more than
// one rank and peer access they come from the symmetric pool, so a create / destroy cycle never cudaFree's memory that is still
// mapped in a peer process (importers keep their CUDA-IPC mappings until pdo_comm_finalize) and never hands a stale mapping to
// a later allocation at the same address.  Collective in that case: every rank allocates and frees in the same order (the
// constructors and destructors that call this are SPMD).  Otherwise plain cudaMalloc / cudaFree.
int comm_shared_malloc(void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) bytes = 8;
    if (g_comm.inited && g_comm.nproc > 1 && g_comm.p2p && comm_sym_alloc(bytes, p, nullptr) == 0) return 0;
    PDO_CUDA(cudaMalloc(p, bytes));
    return 0;
}
void comm_shared_free(void* p) {
    if (!p) return;
    for (auto& e : g_sym_pool) if (e.base == p) { cudaDeviceSynchronize(); e.used = false; return; }
    cudaFree(p);
}
void decomp_grid(pdo_decomp_t h, int* p_row, int* p_col, int* c1, int* c2) {
    *p_row = h->p_row; *p_col = h->p_col; *c1 = h->c1; *c2 = h->c2;
}
void comm_info(int* rank, int* nproc, int* p2p) {
    *rank = g_comm.rank; *nproc = g_comm.nproc; *p2p = (g_comm.inited && g_comm.p2p) ? 1 : 0;
}
void comm_deregister_buffer(void* p) { if (p) pdo_comm_deregister_buffer(p); }
// in-place sum of a small device array over all ranks, stream-ordered (no-op on one rank)
int comm_allreduce_sum(double* dev, int count, cudaStream_t st) {
    if (g_comm.nproc == 1 || count <= 0) return 0;
    PDO_NCCL(ncclAllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, g_comm.comm, st));
    return 0;
}
void comm_register_buffer_quiet(void* p, size_t bytes) { if (g_comm.inited && g_comm.nproc > 1 && g_comm.p2p) sym_register(p, bytes); }
// Single-GPU emulation of a p_row x p_col grid (test hook; tests/test_transpose_emulated_gpu.py): every simulated rank owns a
// device buffer pair, runs the SAME geometry, pack / unpack kernels and push kernels the multi-GPU path runs, and "peer memory"
// is simply the other ranks' buffers on this device.  path: 1 SM store kernel, 2 copy-engine copies, 3 TMA bulk push,
// 4 pack -> exchange -> unpack (cudaMemcpyAsync stands in for the grouped ncclSend / ncclRecv).
int decomp_transpose_emulate(int nx, int ny, int nz, int p_row, int p_col, int dir, int w, int path, const double* const* src,
                             double* const* dst, cudaStream_t st) {
    const int R = p_row * p_col;
    if (R < 1 || R > 64 || dir < 0 || dir > 3 || (w != 1 && w != 2) || path < 1 || path > 4) return fail(PDO_E_BADARG, "bad argument");
    if (nx < p_row || ny < p_row || ny < p_col || nz < p_col) return fail(6, "Invalid 2D processor grid");
    std::vector<pdo_decomp_s*> ds(R, nullptr);
    int rc = 0;
    for (int r = 0; r < R && !rc; ++r) if (!(ds[r] = decomp_new(nx, ny, nz, p_row, p_col, r))) rc = fail(PDO_E_BADARG, "out of memory");
    std::vector<double*> dstv(dst, dst + R);
    PeerView pv{dstv.data(), false};
    if (!rc && path != 4) {
        for (int r = 0; r < R && !rc; ++r) {
            const XGeom g(ds[r], dir, w);
            if (g.np > kMaxPeers) { rc = fail(PDO_E_UNSUPPORTED, "too many peers"); break; }
            rc = transpose_push(ds[r], g, src[r], pv, path, st);
        }
    } else if (!rc) {
        std::vector<XGeom> gs;
        for (int r = 0; r < R; ++r) gs.emplace_back(ds[r], dir, w);
        for (int r = 0; r < R && !rc; ++r) rc = transpose_pack(ds[r], gs[r], src[r], dst[r], st);
        for (int r = 0; r < R && !rc; ++r) {
            const XGeom& g = gs[r];
            for (int m = 0; m < g.np && !rc; ++m) {
                if (m == g.me) continue;
                const int peer = g.world_rank(m);
                const XGeom& gp = gs[peer];
                if (g.scnt[m] != gp.rcnt[g.me]) { rc = fail(PDO_E_BADARG, "count mismatch"); break; }
                if (cudaMemcpyAsync(xchg_recv_ptr(ds[peer], gp, dst[peer], g.me), xchg_send_ptr(ds[r], g, src[r], m),
                                    sizeof(double) * (size_t)g.scnt[m], cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                    rc = fail(PDO_E_CUDA, "emulated exchange failed");
            }
        }
        for (int r = 0; r < R && !rc; ++r) rc = transpose_unpack(ds[r], gs[r], dst[r], st);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && !rc) rc = fail(PDO_E_CUDA, "emulated transpose failed: %s", cudaGetErrorString(cudaGetLastError()));
    for (auto* d : ds) if (d) pdo_decomp_destroy(d);
    return rc;
}
// used by spectral.cu: transposes on device pointers without the host-pointer probe
int decomp_transpose_device(pdo_decomp_t h, int dir, const double* src, double* dst, int w, cudaStream_t st) {
    return transpose_device(h, dir, src, dst, w, st);
}
}  // namespace pdo

extern "C" {

int pdo_comm_unique_id(char id[128]) {
    ncclUniqueId uid;
    static_assert(sizeof(uid) == 128, "ncclUniqueId is 128 bytes");
    PDO_NCCL(ncclGetUniqueId(&uid));
    std::memcpy(id, &uid, 128);
    return 0;
}

int pdo_comm_init(int rank, int nproc, const char unique_id[128]) {
    if (g_comm.inited) return fail(PDO_E_BADARG, "communicator already initialised");
    if (nproc < 1 || rank < 0 || rank >= nproc) return fail(PDO_E_BADARG, "bad rank/nproc");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PDO_E_NODEVICE, "no CUDA device available: padeops_b200 has no CPU fallback");
    }
    g_comm.rank = rank; g_comm.nproc = nproc;
    if (nproc > 1) {
        if (!unique_id) return fail(PDO_E_BADARG, "unique_id required for nproc > 1");
        ncclUniqueId uid;
        std::memcpy(&uid, unique_id, 128);
        PDO_NCCL(ncclCommInitRank(&g_comm.comm, nproc, uid, rank));
        PDO_CUDA(cudaMalloc(&g_comm.d_scalar, 2 * sizeof(double)));
        // P2P transposes (PDO_P2P=0 keeps everything on NCCL): epoch flags every peer can write, shared through CUDA IPC
        const char* e = std::getenv("PDO_P2P");
        g_comm.p2p = !(e && std::atoi(e) == 0);
        if (g_comm.p2p) {
            PDO_CUDA(cudaMalloc(&g_comm.d_xchg, sizeof(XchgRec) * (size_t)(nproc + 1)));
            PDO_CUDA(cudaMalloc(&g_comm.flags, sizeof(unsigned long long) * 2 * (size_t)nproc));
            PDO_CUDA(cudaMemset(g_comm.flags, 0, sizeof(unsigned long long) * 2 * (size_t)nproc));
            PDO_CUDA(cudaDeviceSynchronize());
            const int id = sym_register(g_comm.flags, sizeof(unsigned long long) * 2 * (size_t)nproc);
            if (id < 0) {
                g_comm.p2p = false;  // no peer access between these GPUs: NCCL path everywhere
            } else {
                g_comm.peer_flags.resize(nproc);
                for (int r = 0; r < nproc; ++r) g_comm.peer_flags[r] = (unsigned long long*)g_comm.sym[id].peer[r];
            }
        }
    }
    g_comm.inited = true;
    return 0;
}

int pdo_comm_register_buffer(void* dev_ptr, size_t bytes) {
    if (!g_comm.inited) return fail(PDO_E_BADARG, "communicator not initialised");
    if (g_comm.nproc == 1 || !g_comm.p2p) return 0;
    sym_register(dev_ptr, bytes);  // a buffer that cannot be shared simply keeps the NCCL path
    return 0;
}
// Local: forget a registered buffer (call before freeing it; peers' mappings stay open until pdo_comm_finalize).  A
// freed-but-still-registered buffer is a hazard: a later allocation at the same address would be taken for it.
int pdo_comm_deregister_buffer(void* dev_ptr) {
    for (size_t i = 0; i < g_comm.sym.size();) {
        if (g_comm.sym[i].base == (char*)dev_ptr && (void*)g_comm.flags != dev_ptr) g_comm.sym.erase(g_comm.sym.begin() + i);
        else ++i;
    }
    return 0;
}
int pdo_comm_p2p_enabled(void) { return g_comm.p2p ? 1 : 0; }

int pdo_comm_finalize(void) {
    if (g_comm.comm) { ncclCommDestroy(g_comm.comm); g_comm.comm = nullptr; }
    if (g_comm.d_scalar) { cudaFree(g_comm.d_scalar); g_comm.d_scalar = nullptr; }
    cudaDeviceSynchronize();
    for (auto& m : g_comm.maps) cudaIpcCloseMemHandle(m.mapped);
    for (auto& e : pdo::g_sym_pool) cudaFree(e.base);
    pdo::g_sym_pool.clear();
    if (g_comm.flags) cudaFree(g_comm.flags);
    if (g_comm.d_xchg) cudaFree(g_comm.d_xchg);
    g_comm = Comm();
    return 0;
}
int pdo_comm_rank(void) { return g_comm.rank; }
int pdo_comm_size(void) { return g_comm.nproc; }

int pdo_decomp_info_for(int nx, int ny, int nz, int p_row, int p_col, int rank, pdo_decomp_info* info) {
    if (!info || p_row < 1 || p_col < 1 || rank < 0 || rank >= p_row * p_col) return fail(PDO_E_BADARG, "bad argument");
    // decomp_info_init's check (2D» decomp_2d.f90:507-514), error code 6
    if (nx < p_row || ny < p_row || ny < p_col || nz < p_col)
        return fail(6, "Invalid 2D processor grid. Make sure that min(nx,ny) >= p_row and min(ny,nz) >= p_col");
    fill_info(nx, ny, nz, p_row, p_col, rank, info);
    return 0;
}

int pdo_decomp_init(pdo_decomp_t* h, int nx, int ny, int nz, int p_row, int p_col) {
    if (!h) return fail(PDO_E_BADARG, "null handle");
    *h = nullptr;
    const int nproc = g_comm.nproc;
    if (p_row == 0 && p_col == 0) { p_row = 1; p_col = nproc; }
    if (p_row * p_col != nproc) return fail(1, "Invalid 2D processor grid - nproc /= p_row*p_col");  // 2D» decomp_2d.f90:331-334
    pdo_decomp_info info;
    if (int rc = pdo_decomp_info_for(nx, ny, nz, p_row, p_col, g_comm.rank, &info)) return rc;
    pdo_decomp_s* d = decomp_new(nx, ny, nz, p_row, p_col, g_comm.rank);
    if (!d) return fail(PDO_E_BADARG, "out of memory");
    *h = d;
    return 0;
}
int pdo_decomp_destroy(pdo_decomp_t h) {
    if (!h) return 0;
    if (h->work_send) cudaFree(h->work_send);
    if (h->work_recv) cudaFree(h->work_recv);
    if (h->push_counter) cudaFree(h->push_counter);
    if (h->ce_ready) {
        for (int m = 0; m < kMaxPeers; ++m) { cudaStreamDestroy(h->ce_stream[m]); cudaEventDestroy(h->ce_join[m]); }
        cudaEventDestroy(h->ce_fork);
    }
    delete h;
    return 0;
}
int pdo_decomp_get_info(pdo_decomp_t h, pdo_decomp_info* info) {
    if (!h || !info) return fail(PDO_E_BADARG, "null argument");
    *info = h->info;
    return 0;
}

int pdo_transpose_x_to_y(pdo_decomp_t h, const double* s, double* d, int w, void* st) { return transpose_any(h, 0, s, d, w, st); }
int pdo_transpose_y_to_x(pdo_decomp_t h, const double* s, double* d, int w, void* st) { return transpose_any(h, 1, s, d, w, st); }
int pdo_transpose_y_to_z(pdo_decomp_t h, const double* s, double* d, int w, void* st) { return transpose_any(h, 2, s, d, w, st); }
int pdo_transpose_z_to_y(pdo_decomp_t h, const double* s, double* d, int w, void* st) { return transpose_any(h, 3, s, d, w, st); }

static int allreduce1(double local, double* global, ncclRedOp_t op) {
    if (!global) return fail(PDO_E_BADARG, "null argument");
    if (g_comm.nproc == 1) { *global = local; return 0; }
    PDO_CUDA(cudaMemcpy(g_comm.d_scalar, &local, sizeof(double), cudaMemcpyHostToDevice));
    PDO_NCCL(ncclAllReduce(g_comm.d_scalar, g_comm.d_scalar + 1, 1, ncclDouble, op, g_comm.comm, 0));
    PDO_CUDA(cudaMemcpy(global, g_comm.d_scalar + 1, sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
}  // extern "C"
namespace pdo { namespace hooks {
int transpose_emulate(int nx, int ny, int nz, int p_row, int p_col, int dir, int w, int path, const double* const* src,
                                double* const* dst, void* stream) {
    return pdo::decomp_transpose_emulate(nx, ny, nz, p_row, p_col, dir, w, path, src, dst, (cudaStream_t)stream);
}
}}  // namespace pdo::hooks
extern "C" {
int pdo_p_maxval(double local, double* global) { return allreduce1(local, global, ncclMax); }
int pdo_p_sum(double local, double* global) { return allreduce1(local, global, ncclSum); }

}  // extern "C"
