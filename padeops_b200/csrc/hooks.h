// hooks.h — host-only TEST hooks of libpadeops_b200.so.  They are C++ functions in namespace pdo::hooks, NOT part of the C ABI:
// the product library exports no pdo_debug_* symbol.  libpadeops_b200_testhooks.so (testhooks.cu) wraps each one in an extern "C"
// entry point for the Python tests; nothing under padeops_b200/*.py or include/ reaches them.
#pragma once
#include "../../include/padeops_b200.h"

namespace pdo { namespace hooks {
int chunk_tables(int n, int M, int bw, double b1, double b2, void* out, int out_bytes);
int np_line_host(int kind, int n, double dx, int bc1, int bcn, int axis, const double* f, double* out, long long na, long long nb);
int np_chunk_tables(int kind, int n, int M, int bc1, int bcn, double* sets, double* G, int g_capacity, int* meta);
int ctma_config(int P, int XT, int HB, int HW, int BW, int pc_max, long long* smem_bytes);
int np_rows(int kind, int n, int bc1, int bcn, double* rows5n);
int np_fast(int mode);
int stagg_np_host(int op, int n, double dx, int bot_even, int top_even, int bot_sided, int top_sided, const double* in, double* out, long long ncols);
int stagg_np_rows(int op, int n, int bot_even, int top_even, int bot_sided, int top_sided, double* rows3n);
int cd10_generic(pdo_cd10_t h, int which, int axis, const double* f, double* df, int na, int nb, void* stream);
int last_variant(void);
int set_variant(int strided_mode, int x_threads);
int transpose_emulate(int nx, int ny, int nz, int p_row, int p_col, int dir, int w, int path, const double* const* src, double* const* dst, void* stream);
int zslab_emulate(void* handle, int which, const double* f, double* out, long long n1, int n, int nslabs, void* stream);
int hit_draw(double kmin, double kmax, int nwaves, int tid_start, int rand_seed_to_add, int updates, long long seeds[4], int* wx, int* wy, int* wz);
int ztables(int nz, double dz, double* out);
int igrid_bcs(int bot_wall, int top_wall, int* out24);
int fft_plan(int log2n, int loge, int* out20);
int fft_tables(int n, int loge, double* out, int capacity);
int sgs_point(int mid, double cmodel, double cx, double cy, double cz, const double* d9, double* nu, double* S6);
}}  // namespace pdo::hooks
