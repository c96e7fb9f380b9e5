// testhooks.cu — extern "C" pdo_debug_* wrappers around pdo::hooks::* (see hooks.h).  Built into libpadeops_b200_testhooks.so, which
// links against libpadeops_b200.so; used by tests/ only.
#include "hooks.h"

extern "C" {
int pdo_debug_chunk_tables(int n, int M, int bw, double b1, double b2, void* out, int out_bytes) { return pdo::hooks::chunk_tables(n, M, bw, b1, b2, out, out_bytes); }
int pdo_debug_np_line_host(int kind, int n, double dx, int bc1, int bcn, int axis, const double* f, double* out, long long na, long long nb) { return pdo::hooks::np_line_host(kind, n, dx, bc1, bcn, axis, f, out, na, nb); }
int pdo_debug_np_chunk_tables(int kind, int n, int M, int bc1, int bcn, double* sets, double* G, int g_capacity, int* meta) { return pdo::hooks::np_chunk_tables(kind, n, M, bc1, bcn, sets, G, g_capacity, meta); }
int pdo_debug_ctma_config(int P, int XT, int HB, int HW, int BW, int pc_max, long long* smem_bytes) { return pdo::hooks::ctma_config(P, XT, HB, HW, BW, pc_max, smem_bytes); }
int pdo_debug_np_rows(int kind, int n, int bc1, int bcn, double* rows5n) { return pdo::hooks::np_rows(kind, n, bc1, bcn, rows5n); }
int pdo_debug_np_fast(int mode) { return pdo::hooks::np_fast(mode); }
int pdo_debug_stagg_np_host(int op, int n, double dx, int bot_even, int top_even, int bot_sided, int top_sided, const double* in, double* out, long long ncols) { return pdo::hooks::stagg_np_host(op, n, dx, bot_even, top_even, bot_sided, top_sided, in, out, ncols); }
int pdo_debug_stagg_np_rows(int op, int n, int bot_even, int top_even, int bot_sided, int top_sided, double* rows3n) { return pdo::hooks::stagg_np_rows(op, n, bot_even, top_even, bot_sided, top_sided, rows3n); }
int pdo_debug_cd10_generic(pdo_cd10_t h, int which, int axis, const double* f, double* df, int na, int nb, void* stream) { return pdo::hooks::cd10_generic(h, which, axis, f, df, na, nb, stream); }
int pdo_debug_last_variant(void) { return pdo::hooks::last_variant(); }
int pdo_debug_set_variant(int strided_mode, int x_threads) { return pdo::hooks::set_variant(strided_mode, x_threads); }
int pdo_debug_transpose_emulate(int nx, int ny, int nz, int p_row, int p_col, int dir, int w, int path, const double* const* src, double* const* dst, void* stream) { return pdo::hooks::transpose_emulate(nx, ny, nz, p_row, p_col, dir, w, path, src, dst, stream); }
int pdo_debug_zslab_emulate(void* handle, int which, const double* f, double* out, long long n1, int n, int nslabs, void* stream) { return pdo::hooks::zslab_emulate(handle, which, f, out, n1, n, nslabs, stream); }
int pdo_debug_hit_draw(double kmin, double kmax, int nwaves, int tid_start, int rand_seed_to_add, int updates, long long seeds[4], int* wx, int* wy, int* wz) { return pdo::hooks::hit_draw(kmin, kmax, nwaves, tid_start, rand_seed_to_add, updates, seeds, wx, wy, wz); }
int pdo_debug_ztables(int nz, double dz, double* out) { return pdo::hooks::ztables(nz, dz, out); }
int pdo_debug_igrid_bcs(int bot_wall, int top_wall, int* out24) { return pdo::hooks::igrid_bcs(bot_wall, top_wall, out24); }
int pdo_debug_sgs_point(int mid, double cmodel, double cx, double cy, double cz, const double* d9, double* nu, double* S6) { return pdo::hooks::sgs_point(mid, cmodel, cx, cy, cz, d9, nu, S6); }
int pdo_debug_fft_plan(int log2n, int loge, int* out20) { return pdo::hooks::fft_plan(log2n, loge, out20); }
int pdo_debug_fft_tables(int n, int loge, double* out, int capacity) { return pdo::hooks::fft_tables(n, loge, out, capacity); }
}  // extern "C"
