"""CPU: the reference's on-disk field format.  decomp_2d_write_one writes one distributed array as the
flat GLOBAL (nx, ny, nz) array in Fortran order; here the per-rank piece of the product (pure host code) is driven for every
rank of a process grid in turn, like the ranks of examples/io_test/io_test.f90 (17 x 13 x 11 on 4 x 3, values 1, 2, 3 ... in
Fortran order): the file must be the global array byte for byte for every pencil orientation, and reading gives every rank its
pencil back.  Plus Fortran's G15.5 formatting of the restart info file."""
import ctypes as C

import numpy as np
import pytest


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def _boxes(pdo, nx, ny, nz, pr, pc, pen):
    for r in range(pr * pc):
        d = pdo.decomp_info.for_rank(nx, ny, nz, pr, pc, r)
        sz, st = getattr(d, pen + "sz") if not isinstance(d, dict) else d[pen + "sz"], getattr(d, pen + "st") if not isinstance(d, dict) else d[pen + "st"]
        yield r, tuple(sz), tuple(s - 1 for s in st)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("grid", [(4, 3), (1, 4), (3, 1), (1, 1)])
def test_write_one_read_one_like_io_test(pdo, tmp_path, grid, cplx):
    nx, ny, nz = 17, 13, 11
    pr, pc = grid
    L = pdo.lib()
    m = np.arange(1, nx * ny * nz + 1, dtype=np.float64)
    data = (m + 1j * (nx * ny * nz - m)) if cplx else m            # io_test.f90:33-46
    G = data.reshape(nz, ny, nx)                                    # data1(i,j,k), i fastest
    w = 2 if cplx else 1
    for pen in "xyz":
        fn = str(tmp_path / f"u_{pen}.dat").encode()
        order = list(_boxes(pdo, nx, ny, nz, pr, pc, pen))
        for r, sz, st in order[:1] + order[:0:-1]:                  # rank 0 creates, the others arrive in any order
            blk = np.ascontiguousarray(G[st[2]:st[2] + sz[2], st[1]:st[1] + sz[1], st[0]:st[0] + sz[0]])
            rc = L.pdo_io_write_block(fn, _i3((nx, ny, nz)), _i3(sz), _i3(st), w, C.c_void_p(blk.ctypes.data), int(r == 0))
            assert rc == 0, L.pdo_last_error()
        assert open(fn, "rb").read() == G.tobytes()
        for r, sz, st in order:
            back = np.full(tuple(reversed(sz)), -1, dtype=G.dtype)
            assert L.pdo_io_read_block(fn, _i3((nx, ny, nz)), _i3(sz), _i3(st), w, C.c_void_p(back.ctypes.data)) == 0
            assert np.array_equal(back, G[st[2]:st[2] + sz[2], st[1]:st[1] + sz[1], st[0]:st[0] + sz[0]])


def test_overwrite_truncates_and_errors(pdo, tmp_path):
    L = pdo.lib()
    fn = str(tmp_path / "f.dat").encode()
    big = np.arange(4 * 4 * 4, dtype=np.float64)
    assert L.pdo_io_write_block(fn, _i3((4, 4, 4)), _i3((4, 4, 4)), _i3((0, 0, 0)), 1, C.c_void_p(big.ctypes.data), 1) == 0
    small = np.arange(8, dtype=np.float64)
    assert L.pdo_io_write_block(fn, _i3((2, 2, 2)), _i3((2, 2, 2)), _i3((0, 0, 0)), 1, C.c_void_p(small.ctypes.data), 1) == 0
    assert open(fn, "rb").read() == small.tobytes()                # MPI_FILE_SET_SIZE(fh, 0): "guarantee overwriting"
    out = np.zeros(64)
    assert L.pdo_io_read_block(fn, _i3((4, 4, 4)), _i3((4, 4, 4)), _i3((0, 0, 0)), 1, C.c_void_p(out.ctypes.data)) != 0   # short file
    assert L.pdo_io_read_block(str(tmp_path / "nope").encode(), _i3((2, 2, 2)), _i3((2, 2, 2)), _i3((0, 0, 0)), 1, C.c_void_p(out.ctypes.data)) == 321
    assert L.pdo_io_write_block(fn, _i3((4, 4, 4)), _i3((3, 4, 4)), _i3((2, 0, 0)), 1, C.c_void_p(big.ctypes.data), 1) != 0  # outside the box


def test_single_rank_decomp_handles_host_arrays(pdo, tmp_path):
    """decomp_2d_write_one / read_one through a real decomp handle (one rank: no communicator, no GPU needed for host arrays)."""
    nx, ny, nz = 10, 6, 7
    gp = pdo.decomp_info(nx, ny, nz, 1, 1)
    f = np.random.default_rng(0).standard_normal((nz, ny, nx))
    fn = tmp_path / "Run01_uVel_t000012.out"
    pdo.decomp_2d_write_one(1, f, fn, gp)
    assert fn.read_bytes() == f.tobytes()
    back = np.empty_like(f)
    pdo.decomp_2d_read_one(3, back, fn, gp)
    assert np.array_equal(back, f)
    c = f + 1j * f[::-1]
    pdo.decomp_2d_write_one(2, np.ascontiguousarray(c), fn, gp)
    assert fn.read_bytes() == c.tobytes()
    gp.destroy()


# gfortran's output for write(*,"(100g15.5)") x  (G15.5: F layout + 4 blanks for 0.1 <= |x| < 1e5, else E layout)
G15_5 = [(0.0, "     0.0000    "), (0.25, "    0.25000    "), (1.0, "     1.0000    "), (12345.6, "     12346.    "),
         (99999.6, "    0.10000E+06"), (123456.0, "    0.12346E+06"), (0.05, "    0.50000E-01"), (-3.14159265, "    -3.1416    "),
         (1e-7, "    0.10000E-06"), (-2.5e10, "   -0.25000E+11"), (0.099999999, "    0.10000    "), (6.2831853, "     6.2832    "),
         (100.0, "     100.00    "), (0.1, "    0.10000    ")]


@pytest.mark.parametrize("x,text", G15_5)
def test_fortran_g15_5(pdo, x, text):
    buf = C.create_string_buffer(16)
    assert pdo.lib().pdo_io_format_g15_5(float(x), buf) == 0
    assert buf.value.decode() == text and len(text) == 15
    assert abs(float(buf.value) - x) <= 5.1e-5 * abs(x)          # list-directed / g15.5 input reads it back to 5 digits
