/*
 * decomp_oracle.c — CPU restatement of the 2DECOMP&FFT decomposition arithmetic and the four
 * pencil transposes, with all MPI ranks simulated in one process.
 *
 * TEST INFRASTRUCTURE ONLY (see padeops_oracle.c header).
 *
 * Follows ("2D»" = dependencies/2decomp_fft-1.5.847.tar.gz » 2decomp_fft/src):
 *   2D» decomp_2d.f90:676-708   distribute
 *   2D» decomp_2d.f90:622-670   partition
 *   2D» decomp_2d.f90:739-771   prepare_buffer (ALLTOALLV counts / displacements)
 *   2D» transpose_x_to_y.f90:14-91, 332-371, 424-463  (pack, MPI_ALLTOALLV on COL, unpack)
 *   2D» transpose_y_to_x.f90    (mirror)
 *   2D» transpose_y_to_z.f90:14-100, 342-381          (pack, ALLTOALLV on ROW straight into dst)
 *   2D» transpose_z_to_y.f90    (ALLTOALLV straight from src, unpack)
 * The build the reference uses defines only -DDOUBLE_PREC (travis/2decomp_fft_Makefile.inc:15), i.e.
 * the ALLTOALLV branch: no EVEN padding, no SHM.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* 2D» decomp_2d.f90:676-708.  st/en are 1-based like the Fortran. */
void pdo_oracle_distribute(int data1, int proc, int *st, int *en, int *sz)
{
    int size1 = data1 / proc;
    int nu = data1 - size1 * proc;
    int nl = proc - nu;
    int i;
    st[0] = 1;
    sz[0] = size1;
    en[0] = size1;
    for (i = 1; i <= nl - 1; ++i) {
        st[i] = st[i - 1] + size1;
        sz[i] = size1;
        en[i] = en[i - 1] + size1;
    }
    size1 = size1 + 1;
    for (i = nl; i <= proc - 1; ++i) {
        st[i] = en[i - 1] + 1;
        sz[i] = size1;
        en[i] = en[i - 1] + size1;
    }
    en[proc - 1] = data1;
    sz[proc - 1] = data1 - st[proc - 1] + 1;
}

typedef struct {
    int xst[3], xen[3], xsz[3];
    int yst[3], yen[3], ysz[3];
    int zst[3], zen[3], zsz[3];
} pdo_oracle_decomp;

/* partition (2D» decomp_2d.f90:622-670) for rank = coord1*p_col + coord2 (MPI_CART_CREATE row-major,
   no reorder, :336-341) with the pdim triples of decomp_info_init (:529-534). */
void pdo_oracle_decomp_info(int nx, int ny, int nz, int p_row, int p_col, int rank, pdo_oracle_decomp *d)
{
    const int g[3] = { nx, ny, nz };
    const int dims[2] = { p_row, p_col };
    const int coord[2] = { rank / p_col, rank % p_col };
    const int pdims[3][3] = { { 1, 2, 3 }, { 2, 1, 3 }, { 2, 3, 1 } };
    int *outs[3][3] = { { d->xst, d->xen, d->xsz }, { d->yst, d->yen, d->ysz }, { d->zst, d->zen, d->zsz } };
    for (int pen = 0; pen < 3; ++pen)
        for (int i = 0; i < 3; ++i) {
            int pd = pdims[pen][i];
            if (pd == 1) {
                outs[pen][0][i] = 1; outs[pen][1][i] = g[i]; outs[pen][2][i] = g[i];
            } else {
                int p = dims[pd - 2];
                int *st = (int *)malloc(sizeof(int) * 3 * (size_t)p), *en = st + p, *sz = en + p;
                pdo_oracle_distribute(g[i], p, st, en, sz);
                outs[pen][0][i] = st[coord[pd - 2]]; outs[pen][1][i] = en[coord[pd - 2]]; outs[pen][2][i] = sz[coord[pd - 2]];
                free(st);
            }
        }
}

static int64_t vol(const int *s) { return (int64_t)s[0] * s[1] * s[2]; }

/*
 * Transposes with every rank simulated.  `src_all` / `dst_all` hold the per-rank pencils back to
 * back in rank order; `w` = doubles per element (1 real, 2 complex).  dir: 0 x→y, 1 y→x, 2 y→z, 3 z→y.
 * Each rank packs its send buffer exactly as mem_split_* does, the ALLTOALLV is emulated block by
 * block with the reference's counts/displacements, then mem_merge_* unpacks.
 */
int pdo_oracle_transpose(int dir, int nx, int ny, int nz, int p_row, int p_col, int w,
                         const double *src_all, double *dst_all)
{
    const int P = p_row * p_col;
    pdo_oracle_decomp *D = (pdo_oracle_decomp *)malloc(sizeof(pdo_oracle_decomp) * (size_t)P);
    int64_t *soff = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(P + 1)), *doff = soff + P + 1;
    double **sendbuf = (double **)calloc((size_t)P, sizeof(double *));
    double **recvbuf = (double **)calloc((size_t)P, sizeof(double *));
    const int use_col = (dir == 0 || dir == 1);
    const int np = use_col ? p_row : p_col; /* sub-communicator size */
    int *dist_s = (int *)malloc(sizeof(int) * 4 * (size_t)np), *dist_r = dist_s + np, *tst = dist_r + np,
        *ten = tst + np;
    soff[0] = doff[0] = 0;
    for (int r = 0; r < P; ++r) {
        pdo_oracle_decomp_info(nx, ny, nz, p_row, p_col, r, &D[r]);
        const int *ss = (dir == 0) ? D[r].xsz : (dir == 3) ? D[r].zsz : D[r].ysz;
        const int *ds = (dir == 1) ? D[r].xsz : (dir == 2) ? D[r].zsz : D[r].ysz;
        soff[r + 1] = soff[r] + vol(ss) * w;
        doff[r + 1] = doff[r] + vol(ds) * w;
    }
    /* get_dist (2D» decomp_2d.f90:715-733): x1dist = dist(nx,p_row), y1dist = dist(ny,p_row),
       y2dist = dist(ny,p_col), z2dist = dist(nz,p_col) */
    if (dir == 0) { pdo_oracle_distribute(nx, np, tst, ten, dist_s); pdo_oracle_distribute(ny, np, tst, ten, dist_r); }
    if (dir == 1) { pdo_oracle_distribute(ny, np, tst, ten, dist_s); pdo_oracle_distribute(nx, np, tst, ten, dist_r); }
    if (dir == 2) { pdo_oracle_distribute(ny, np, tst, ten, dist_s); pdo_oracle_distribute(nz, np, tst, ten, dist_r); }
    if (dir == 3) { pdo_oracle_distribute(nz, np, tst, ten, dist_s); pdo_oracle_distribute(ny, np, tst, ten, dist_r); }

    /* ---- pack (mem_split_*) ---- */
    for (int r = 0; r < P; ++r) {
        const int *ss = (dir == 0) ? D[r].xsz : (dir == 3) ? D[r].zsz : D[r].ysz;
        const int64_t n1 = ss[0], n2 = ss[1], n3 = ss[2];
        const double *in = src_all + soff[r];
        double *out = (double *)malloc(sizeof(double) * (size_t)(vol(ss) * w + 1));
        sendbuf[r] = out;
        int64_t pos = 0;
        int i1 = 0, i2 = 0;
        for (int m = 0; m < np; ++m) {
            if (m == 0) { i1 = 1; i2 = dist_s[0]; } else { i1 = i2 + 1; i2 = i1 + dist_s[m] - 1; }
            if (dir == 0) { /* mem_split_xy: in(i1:i2, :, :) */
                for (int64_t k = 1; k <= n3; ++k) for (int64_t j = 1; j <= n2; ++j) for (int64_t i = i1; i <= i2; ++i) {
                    memcpy(out + pos, in + (((k - 1) * n2 + (j - 1)) * n1 + (i - 1)) * w, sizeof(double) * (size_t)w); pos += w; }
            } else if (dir == 1 || dir == 2) { /* mem_split_yx / mem_split_yz: in(:, i1:i2, :) */
                for (int64_t k = 1; k <= n3; ++k) for (int64_t j = i1; j <= i2; ++j) for (int64_t i = 1; i <= n1; ++i) {
                    memcpy(out + pos, in + (((k - 1) * n2 + (j - 1)) * n1 + (i - 1)) * w, sizeof(double) * (size_t)w); pos += w; }
            } else { /* z→y sends straight from src: block m is the k-slab (:,:,i1:i2), already contiguous */
                int64_t cnt = n1 * n2 * (int64_t)(i2 - i1 + 1) * w;
                memcpy(out + pos, in + n1 * n2 * (int64_t)(i1 - 1) * w, sizeof(double) * (size_t)cnt); pos += cnt;
            }
        }
    }
    /* ---- ALLTOALLV within each sub-communicator ---- */
    for (int r = 0; r < P; ++r) {
        const int *ds = (dir == 1) ? D[r].xsz : (dir == 2) ? D[r].zsz : D[r].ysz;
        recvbuf[r] = (double *)malloc(sizeof(double) * (size_t)(vol(ds) * w + 1));
    }
    for (int r = 0; r < P; ++r) {
        const int c1 = r / p_col, c2 = r % p_col;
        const int *ss = (dir == 0) ? D[r].xsz : (dir == 3) ? D[r].zsz : D[r].ysz;
        int64_t sdisp = 0;
        for (int m = 0; m < np; ++m) {
            /* send count to peer m (prepare_buffer): x1cnts = x1dist(m)*xsz2*xsz3, y1cnts = ysz1*y1dist(m)*ysz3,
               y2cnts = ysz1*y2dist(m)*ysz3, z2cnts = zsz1*zsz2*z2dist(m) */
            int64_t scnt = (dir == 0) ? (int64_t)dist_s[m] * ss[1] * ss[2]
                          : (dir == 3) ? (int64_t)ss[0] * ss[1] * dist_s[m]
                                       : (int64_t)ss[0] * dist_s[m] * ss[2];
            const int peer = use_col ? (m * p_col + c2) : (c1 * p_col + m);
            const int me_in_peer = use_col ? c1 : c2;
            /* receive displacement on the peer for data coming from me */
            const int *pd = (dir == 1) ? D[peer].xsz : (dir == 2) ? D[peer].zsz : D[peer].ysz;
            int64_t rdisp = 0;
            for (int q = 0; q < me_in_peer; ++q)
                rdisp += (dir == 1) ? (int64_t)dist_r[q] * pd[1] * pd[2]
                        : (dir == 2) ? (int64_t)pd[0] * pd[1] * dist_r[q]
                                     : (int64_t)pd[0] * dist_r[q] * pd[2];
            memcpy(recvbuf[peer] + rdisp * w, sendbuf[r] + sdisp * w, sizeof(double) * (size_t)(scnt * w));
            sdisp += scnt;
        }
    }
    /* ---- unpack (mem_merge_*) ---- */
    for (int r = 0; r < P; ++r) {
        const int *ds = (dir == 1) ? D[r].xsz : (dir == 2) ? D[r].zsz : D[r].ysz;
        const int64_t n1 = ds[0], n2 = ds[1], n3 = ds[2];
        double *out = dst_all + doff[r];
        const double *in = recvbuf[r];
        int64_t pos = 0;
        int i1 = 0, i2 = 0;
        for (int m = 0; m < np; ++m) {
            if (m == 0) { i1 = 1; i2 = dist_r[0]; } else { i1 = i2 + 1; i2 = i1 + dist_r[m] - 1; }
            if (dir == 0 || dir == 3) { /* mem_merge_xy / mem_merge_zy: out(:, i1:i2, :) */
                for (int64_t k = 1; k <= n3; ++k) for (int64_t j = i1; j <= i2; ++j) for (int64_t i = 1; i <= n1; ++i) {
                    memcpy(out + (((k - 1) * n2 + (j - 1)) * n1 + (i - 1)) * w, in + pos, sizeof(double) * (size_t)w); pos += w; }
            } else if (dir == 1) { /* mem_merge_yx: out(i1:i2, :, :) */
                for (int64_t k = 1; k <= n3; ++k) for (int64_t j = 1; j <= n2; ++j) for (int64_t i = i1; i <= i2; ++i) {
                    memcpy(out + (((k - 1) * n2 + (j - 1)) * n1 + (i - 1)) * w, in + pos, sizeof(double) * (size_t)w); pos += w; }
            } else { /* y→z receives straight into dst: block from m is the k-slab (:,:,i1:i2) */
                int64_t cnt = n1 * n2 * (int64_t)(i2 - i1 + 1) * w;
                memcpy(out + n1 * n2 * (int64_t)(i1 - 1) * w, in + pos, sizeof(double) * (size_t)cnt); pos += cnt;
            }
        }
    }
    for (int r = 0; r < P; ++r) { free(sendbuf[r]); free(recvbuf[r]); }
    free(sendbuf); free(recvbuf); free(dist_s); free(soff); free(D);
    return 0;
}
