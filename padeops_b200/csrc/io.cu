// io.cu — the reference's on-disk field format: decomp_2d_write_one / decomp_2d_read_one
// (2DECOMP&FFT io_write_one.f90:23-80, io_read_one.f90) write ONE distributed 3-D array as the flat GLOBAL array in Fortran
// order, native-endian doubles, no header: rank r owns the sub-box (subsizes, starts) of its pencil and MPI-IO's subarray
// file view scatters it.  Here every rank writes its own sub-box with positioned writes (one per contiguous run; runs that
// are adjacent in the file are merged), after rank 0 has truncated the file ("MPI_FILE_SET_SIZE(fh, 0): guarantee
// overwriting").  Device arrays are staged through host memory: a restart dump is PCIe- and disk-bound by nature, it is
// not on the hot path, but it lets the GPU path consume and produce the reference's restart / field files unchanged.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "spectral_internal.cuh"

namespace pdo {
namespace {

struct Fd {
    int fd = -1;
    ~Fd() { if (fd >= 0) ::close(fd); }
};

// (subsizes, starts 0-based) inside (sizes), Fortran order, elements of `esz` bytes: calls fn(file_offset, mem_offset, bytes)
// once per maximal contiguous run of the sub-box in the file
template <class Fn>
int for_each_run(const int sizes[3], const int sub[3], const int st[3], size_t esz, Fn fn) {
    for (int a = 0; a < 3; ++a)
        if (sizes[a] < 1 || sub[a] < 0 || st[a] < 0 || st[a] + sub[a] > sizes[a]) return fail(PDO_E_BADARG, "io: sub-box outside the global array");
    if (sub[0] == 0 || sub[1] == 0 || sub[2] == 0) return 0;
    const bool full_x = sub[0] == sizes[0], full_xy = full_x && sub[1] == sizes[1];
    const size_t row = (size_t)sub[0] * esz;
    if (full_xy) {   // whole planes: one run
        return fn(((size_t)st[2] * sizes[1] * sizes[0]) * esz, (size_t)0, row * sub[1] * sub[2]);
    }
    for (int k = 0; k < sub[2]; ++k) {
        const size_t fplane = ((size_t)(st[2] + k) * sizes[1] + st[1]) * sizes[0] + st[0];
        const size_t mplane = (size_t)k * sub[1] * sub[0];
        if (full_x) {   // whole rows: the j-range of one plane is one run
            if (int rc = fn(fplane * esz, mplane * esz, row * sub[1])) return rc;
            continue;
        }
        for (int j = 0; j < sub[1]; ++j)
            if (int rc = fn((fplane + (size_t)j * sizes[0]) * esz, (mplane + (size_t)j * sub[0]) * esz, row)) return rc;
    }
    return 0;
}

int pwrite_all(int fd, const char* p, size_t n, off_t off) {
    while (n) {
        const ssize_t w = ::pwrite(fd, p, n, off);
        if (w < 0) { if (errno == EINTR) continue; return fail(PDO_E_BADARG, "io: write failed: %s", std::strerror(errno)); }
        p += w; n -= (size_t)w; off += w;
    }
    return 0;
}
int pread_all(int fd, char* p, size_t n, off_t off) {
    while (n) {
        const ssize_t r = ::pread(fd, p, n, off);
        if (r < 0) { if (errno == EINTR) continue; return fail(PDO_E_BADARG, "io: read failed: %s", std::strerror(errno)); }
        if (r == 0) return fail(PDO_E_BADARG, "io: file is shorter than the global array");
        p += r; n -= (size_t)r; off += r;
    }
    return 0;
}

int barrier() {
    double g = 0.0;
    return pdo_p_sum(1.0, &g);
}

int pencil_box(const pdo_decomp_info& d, int ipencil, int sizes[3], int sub[3], int st[3]) {
    sizes[0] = d.xsz[0]; sizes[1] = d.ysz[1]; sizes[2] = d.zsz[2];   // io_write_one.f90:24-26
    const int* s = ipencil == 1 ? d.xsz : ipencil == 2 ? d.ysz : d.zsz;
    const int* b = ipencil == 1 ? d.xst : ipencil == 2 ? d.yst : d.zst;
    if (ipencil < 1 || ipencil > 3) return fail(PDO_E_BADARG, "io: ipencil must be 1, 2 or 3");
    for (int a = 0; a < 3; ++a) { sub[a] = s[a]; st[a] = b[a] - 1; }   // "0-based index" :32-34
    return 0;
}

}  // namespace
}  // namespace pdo

using namespace pdo;

extern "C" {

/* One rank's share of decomp_2d_write_one, host data: writes the sub-box (sub, st 0-based) of the global (sizes) array.
 * create != 0: create the file if needed and set its size to zero first (what rank 0 does for everyone). */
int pdo_io_write_block(const char* filename, const int sizes[3], const int sub[3], const int st[3], int elem_doubles, const double* data,
                       int create) {
    if (!filename || !sizes || !sub || !st || (elem_doubles != 1 && elem_doubles != 2)) return fail(PDO_E_BADARG, "io: bad argument");
    Fd f;
    f.fd = ::open(filename, create ? (O_WRONLY | O_CREAT | O_TRUNC) : O_WRONLY, 0644);
    if (f.fd < 0) return fail(PDO_E_BADARG, "io: cannot open '%s' for writing: %s", filename, std::strerror(errno));
    if ((long long)sub[0] * sub[1] * sub[2] > 0 && !data) return fail(PDO_E_BADARG, "io: null data");
    const size_t esz = sizeof(double) * (size_t)elem_doubles;
    const char* base = (const char*)data;
    return for_each_run(sizes, sub, st, esz, [&](size_t foff, size_t moff, size_t bytes) { return pwrite_all(f.fd, base + moff, bytes, (off_t)foff); });
}
int pdo_io_read_block(const char* filename, const int sizes[3], const int sub[3], const int st[3], int elem_doubles, double* data) {
    if (!filename || !sizes || !sub || !st || (elem_doubles != 1 && elem_doubles != 2)) return fail(PDO_E_BADARG, "io: bad argument");
    Fd f;
    f.fd = ::open(filename, O_RDONLY);
    if (f.fd < 0) return fail(321, "File not found: %s", filename);   // igrid_operators_periodic.F90:176-180
    if ((long long)sub[0] * sub[1] * sub[2] > 0 && !data) return fail(PDO_E_BADARG, "io: null data");
    const size_t esz = sizeof(double) * (size_t)elem_doubles;
    char* base = (char*)data;
    return for_each_run(sizes, sub, st, esz, [&](size_t foff, size_t moff, size_t bytes) { return pread_all(f.fd, base + moff, bytes, (off_t)foff); });
}

/* decomp_2d_write_one(ipencil, var, filename, opt_decomp): collective.  var: this rank's pencil, host or device. */
int pdo_decomp_write_one(pdo_decomp_t h, int ipencil, const double* var, int elem_doubles, const char* filename) {
    if (!h || !filename) return fail(PDO_E_BADARG, "null argument");
    pdo_decomp_info d;
    if (int rc = pdo_decomp_get_info(h, &d)) return rc;
    int sizes[3], sub[3], st[3];
    if (int rc = pencil_box(d, ipencil, sizes, sub, st)) return rc;
    const size_t bytes = sizeof(double) * (size_t)elem_doubles * (size_t)sub[0] * sub[1] * sub[2];
    std::vector<char> stage;
    const double* host = var;
    if (bytes && is_device_ptr(var)) {
        stage.resize(bytes);
        PDO_CUDA(cudaMemcpy(stage.data(), var, bytes, cudaMemcpyDeviceToHost));
        host = (const double*)stage.data();
    }
    int rc = 0;
    if (pdo_comm_rank() == 0) rc = pdo_io_write_block(filename, sizes, sub, st, elem_doubles, host, 1);
    if (int b = barrier()) return b;              // the file exists and is empty before anyone else touches it
    if (pdo_comm_rank() != 0) rc = pdo_io_write_block(filename, sizes, sub, st, elem_doubles, host, 0);
    if (int b = barrier()) return b;              // MPI_FILE_WRITE_ALL / MPI_FILE_CLOSE are collective: complete on return
    return rc;
}
int pdo_decomp_read_one(pdo_decomp_t h, int ipencil, double* var, int elem_doubles, const char* filename) {
    if (!h || !filename) return fail(PDO_E_BADARG, "null argument");
    pdo_decomp_info d;
    if (int rc = pdo_decomp_get_info(h, &d)) return rc;
    int sizes[3], sub[3], st[3];
    if (int rc = pencil_box(d, ipencil, sizes, sub, st)) return rc;
    const size_t bytes = sizeof(double) * (size_t)elem_doubles * (size_t)sub[0] * sub[1] * sub[2];
    if (!bytes) return 0;
    if (!is_device_ptr(var)) return pdo_io_read_block(filename, sizes, sub, st, elem_doubles, var);
    std::vector<char> stage(bytes);
    if (int rc = pdo_io_read_block(filename, sizes, sub, st, elem_doubles, (double*)stage.data())) return rc;
    PDO_CUDA(cudaMemcpy(var, stage.data(), bytes, cudaMemcpyHostToDevice));
    return 0;
}

/* Fortran's G15.5 edit descriptor, the format of the RESTART info file (igrid.F90:2797 write(10,"(100g15.5)") tsim): 15
 * characters, 5 significant digits — F layout plus four blanks when 0.1 <= |x| < 1e5 (after rounding), E layout otherwise.
 * out: at least 16 bytes. */
int pdo_io_format_g15_5(double x, char out[16]) {
    if (!out) return fail(PDO_E_BADARG, "null argument");
    char buf[64];
    const int w = 15, d = 5;
    const double ax = std::fabs(x);
    if (!std::isfinite(x)) {
        std::snprintf(buf, sizeof(buf), "%*s", w, std::isnan(x) ? "NaN" : (x < 0 ? "-Infinity" : "Infinity"));
    } else if (ax == 0.0) {
        std::snprintf(buf, sizeof(buf), "%*.*f    ", w - 4, d - 1, x);                  // F(w-4).(d-1), 4X
    } else {
        // decimal exponent e with 10^(e-1) <= |x| rounded to d significant digits < 10^e
        char sci[64];
        std::snprintf(sci, sizeof(sci), "%.*e", d - 1, ax);                             // d.dddde+XX after rounding
        const int e10 = std::atoi(std::strchr(sci, 'e') + 1) + 1;
        if (e10 >= 0 && e10 <= d) {
            std::snprintf(buf, sizeof(buf), "%#*.*f    ", w - 4, d - e10, x);           // F(w-4).(d-e), 4X ("12346." keeps its point)
        } else {
            // 0.dddddE+ee (two exponent digits; three drop the letter E: 0.ddddd+eee)
            char digits[16];
            int n = 0;
            for (const char* p = sci; *p && *p != 'e'; ++p) if (*p != '.') digits[n++] = *p;
            digits[n] = 0;
            char ex[16];
            if (std::abs(e10) < 100) std::snprintf(ex, sizeof(ex), "E%c%02d", e10 < 0 ? '-' : '+', std::abs(e10));
            else std::snprintf(ex, sizeof(ex), "%c%03d", e10 < 0 ? '-' : '+', std::abs(e10));
            char body[48];
            std::snprintf(body, sizeof(body), "%s0.%s%s", x < 0 ? "-" : "", digits, ex);
            std::snprintf(buf, sizeof(buf), "%*s", w, body);
        }
    }
    if ((int)std::strlen(buf) != w) std::snprintf(buf, sizeof(buf), "%s", "***************");  // field overflow, as Fortran prints it
    std::memcpy(out, buf, 16);
    return 0;
}

}  // extern "C"
