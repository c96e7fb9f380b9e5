// banded.cuh — host-side handle + launch interface of the compact (banded) line operators.
#pragma once
#include <cuda_runtime.h>

#include "tables.h"

namespace pdo {

// Right-hand-side stencil families (one per reference RHS loop nest; file:line in banded.cu).
enum RhsKind {
    RK_D1_7 = 0,      // a(f+1 - f-1) + b(f+2 - f-2) + c(f+3 - f-3)                  CD10 first derivative
    RK_D2_7 = 1,      // a(f+1 - 2f + f-1) + b(f+2 - 2f + f-2) + c(f+3 - 2f + f-3)    CD10 second derivative
    RK_D1_5 = 2,      // a(f+1 - f-1) + b(f+2 - f-2)                                  CD06 first derivative
    RK_SYM_9 = 3,     // a f + b(f+1 + f-1) + ... + e(f+4 + f-4)                      CF90 / Gaussian
    RK_D2_5 = 4,      // a(f+1 - 2f + f-1) + b(f+2 - 2f + f-2)                        staggered d2 (C2C / E2E)
    RK_STAG_E2C = 5,  // a(f[k+1] ± f[k]) + b(f[k+2] ± f[k-1])                        edge → cell (ddz / interp)
    RK_STAG_C2E = 6,  // a(f[k] ± f[k-1]) + b(f[k+1] ± f[k-2])                        cell → edge (ddz / interp)
};

struct OpParams {
    double co[5];  // stencil coefficients with the grid spacing folded in exactly as the reference does
    int edge_in;   // input has n+1 planes and plane n (0-based) is read un-wrapped (E2C quirk)
    int edge_out;  // output has n+1 planes; plane n := plane 0 after the solve (C2E / E2E)
};

struct BandedOp {
    int n = 0;
    int rk = 0;
    int bw = 0;  // 0 explicit stencil, 1 cyclic tridiagonal, 2 cyclic pentadiagonal
    OpParams op{};
    int M = 0;   // chunk length of the fast path; 0 → generic any-n kernels
    ChunkTables tab{};
    ChunkTables tab16{};  // the same matrix cut into 16-point chunks, for the TMA x kernel (valid when has_tab16)
    int has_tab16 = 0;
    double* d_line = nullptr;  // generic path tables (device)
    // planner memory: winning kernel variant per (axis, n1, n3), filled by banded_op_plan (or, with PDO_TUNE=1, by the first
    // large call on a shape); without a plan the deterministic table of banded.cu: default_variant decides
    static constexpr int kMaxPlans = 8;
    struct Plan { int axis; long long n1, n3; int choice; };
    mutable Plan plans[kMaxPlans];
    mutable int nplans = 0;
};

// Builds tables (host) and uploads what the generic path needs.  LHS = circ[b2 b1 1 b1 b2].
cudaError_t banded_op_create(BandedOp* h, int n, int rk, int bw, double b1, double b2, const OpParams& op);
void banded_op_destroy(BandedOp* h);

// Applies the operator along `axis` of a column-major field: axis 0: f(n,na,nb); 1: f(na,n,nb); 2: f(na,nb,n).
// Device pointers; f and out must not alias.  For edge_in / edge_out ops the solve axis has n+1 planes on
// that side.  force_generic != 0 routes through the any-n kernels (used by tests to cross-check).
cudaError_t banded_op_apply(const BandedOp* h, int axis, const double* f, double* out, long long na, long long nb,
                            cudaStream_t stream, int force_generic = 0);

// Times the kernel candidates for `axis` of a pencil with extents (na, nb) on scratch arrays and stores the winner in the handle
// (FFTW-planner style; explicit, never inside a user call).  Synchronises the device.  *chosen: the variant code.
cudaError_t banded_op_plan(const BandedOp* h, int axis, long long na, long long nb, int* chosen);
void banded_set_tuning(bool plan_on_first_call);   // default off (PDO_TUNE=1 turns it on)

// z-slab (distributed line) mode of the strided operators: the periodic line of h->n points (the operator is created for
// the GLOBAL length) is cut across GPUs into slabs of n_local rows each, f(n1, n_local).  Instead of transposing the field
// so that one GPU holds whole lines, neighbours exchange only (i) `halo_rows` stencil rows and (ii) the reduced-system
// pieces of `halo_chunks` chunks on each side; the separator systems then close locally.  Buffers:
//   planes  [2*halo_rows][n1]         rows below the slab (from the lower GPU), then rows above it (from the upper GPU)
//   glo/ghi [2][bw][halo_chunks][n1]  (gA, gB) of the lower GPU's last / the upper GPU's first halo_chunks chunks
// banded_zslab_edges computes this slab's own edge pieces and stores them into to_lower (= the lower GPU's ghi) and
// to_upper (= the upper GPU's glo); both may be peer memory.  banded_zslab_apply is the fused solve.
// halo_chunks < 0: this operator / slab size is not supported in z-slab mode (callers fall back to transposes).
int banded_zslab_halo_rows(const BandedOp* h);
int banded_zslab_halo_chunks(const BandedOp* h, int n_local);
cudaError_t banded_zslab_edges(const BandedOp* h, const double* f, long long n1, int n_local, const double* planes, double* to_lower,
                               double* to_upper, cudaStream_t stream);
cudaError_t banded_zslab_apply(const BandedOp* h, const double* f, double* out, long long n1, int n_local, const double* planes,
                               const double* glo, const double* ghi, cudaStream_t stream);

// Test hook: force a kernel variant (strided_mode: -1 env/auto, 0 auto, 1 t512, 2 t256, 3 cluster, 4 cluster4,
// 5 cpipe, 6 pipe1; x_threads: -1 env/default, 128, 256).
void banded_debug_set_variant(int strided_mode, int x_threads);
int banded_debug_last_variant();
int banded_debug_ctma_config(int P, int XT, int HB, int HW, int BW, int pc_max, long long* smem_bytes);

}  // namespace pdo
