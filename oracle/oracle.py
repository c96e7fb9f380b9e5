"""ctypes front-end to the CPU oracle (oracle/libpadeops_oracle.so).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never from padeops_b200/.

Arrays follow the reference's Fortran layout: a field f(n1,n2,n3) is a numpy array of shape
(n3, n2, n1), C-contiguous, so that the first Fortran index is the fastest one in memory.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpadeops_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)


def build(force=False):
    srcs = [os.path.join(_HERE, s) for s in ("padeops_oracle.c", "decomp_oracle.c", "spectral_oracle.c", "nonperiodic_oracle.c")]
    if (not force) and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


# ---------------------------------------------------------------- LU tables
def penta_lu(n, e, a, d, c, f):
    LU = np.zeros((9, n))
    lib().pdo_oracle_penta_lu(C.c_int(n), C.c_double(e), C.c_double(a), C.c_double(d), C.c_double(c), C.c_double(f), _p(LU))
    return LU


def tri_lu(n, b, d, a):
    LU = np.zeros((5, n))
    lib().pdo_oracle_tri_lu(C.c_int(n), C.c_double(b), C.c_double(d), C.c_double(a), _p(LU))
    return LU


def cd10_lu(n, which):
    LU = np.zeros((9, n))
    ierr = lib().pdo_oracle_cd10_lu(C.c_int(n), C.c_int(which), _p(LU))
    return ierr, LU


def cd06_lu(n):
    LU = np.zeros((5, n))
    ierr = lib().pdo_oracle_cd06_lu(C.c_int(n), _p(LU))
    return ierr, LU


def cf90_lu(n):
    LU = np.zeros((9, n))
    ierr = lib().pdo_oracle_cf90_lu(C.c_int(n), _p(LU))
    return ierr, LU


def stagg_lu(n, which):
    LU = np.zeros((5, n))
    ierr = lib().pdo_oracle_stagg_lu(C.c_int(n), C.c_int(which), _p(LU))
    return ierr, LU


# ---------------------------------------------------------------- operators
def _na_nb(f, axis):
    n3, n2, n1 = f.shape
    if axis == 0:
        return n1, n2, n3
    if axis == 1:
        return n2, n1, n3
    return n3, n1, n2


def cd10(f, dx, axis, which=1, out=None):
    """cd10%dd1/dd2/dd3 (which=1) or d2d1/d2d2/d2d3 (which=2) along `axis` (0=x fastest).  `out`: caller-owned result array
    (the reference's df is the caller's too; the timed CPU baseline passes it so that no page faults sit in the timed region)."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    n, na, nb = _na_nb(f, axis)
    ierr, LU = cd10_lu(n, which)
    assert ierr == 0
    if out is None:
        out = np.empty_like(f)
    assert out.shape == f.shape and out.dtype == np.float64 and out.flags.c_contiguous
    lib().pdo_oracle_cd10(_p(LU), C.c_int(n), C.c_double(dx), C.c_int(which), C.c_int(axis), _p(f), _p(out),
                          C.c_int64(na), C.c_int64(nb))
    return out


def cd06(f, dx, axis):
    f = np.ascontiguousarray(f, dtype=np.float64)
    n, na, nb = _na_nb(f, axis)
    ierr, LU = cd06_lu(n)
    assert ierr == 0
    out = np.empty_like(f)
    lib().pdo_oracle_cd06(_p(LU), C.c_int(n), C.c_double(dx), C.c_int(axis), _p(f), _p(out), C.c_int64(na), C.c_int64(nb))
    return out


def cf90(f, axis):
    f = np.ascontiguousarray(f, dtype=np.float64)
    n, na, nb = _na_nb(f, axis)
    ierr, LU = cf90_lu(n)
    assert ierr == 0
    out = np.empty_like(f)
    lib().pdo_oracle_cf90(_p(LU), C.c_int(n), C.c_int(axis), _p(f), _p(out), C.c_int64(na), C.c_int64(nb))
    return out


def gaussian(f, axis):
    f = np.ascontiguousarray(f, dtype=np.float64)
    n, na, nb = _na_nb(f, axis)
    out = np.empty_like(f)
    lib().pdo_oracle_gaussian(C.c_int(n), C.c_int(axis), _p(f), _p(out), C.c_int64(na), C.c_int64(nb))
    return out


STAGG_OPS = ("ddz_E2C", "ddz_C2E", "interp_E2C", "interp_C2E", "d2dz2_C2C", "d2dz2_E2E")
_STAGG_LU = (0, 0, 2, 2, 1, 1)
_STAGG_IN_E = (True, False, True, False, False, True)
_STAGG_OUT_E = (False, True, False, True, False, True)


def stagg(op, f, n, dz):
    """Periodic cd06stagg op along z.  f: (nz_in, n2, n1) real or complex; n = number of cells."""
    if isinstance(op, str):
        op = STAGG_OPS.index(op)
    cplx = np.iscomplexobj(f)
    f = np.ascontiguousarray(f, dtype=np.complex128 if cplx else np.float64)
    nin = n + 1 if _STAGG_IN_E[op] else n
    nout = n + 1 if _STAGG_OUT_E[op] else n
    assert f.shape[0] == nin, (f.shape, nin)
    out = np.zeros((nout,) + f.shape[1:], dtype=f.dtype)
    m = f.shape[1] * f.shape[2] * (2 if cplx else 1)
    ierr, LU = stagg_lu(n, _STAGG_LU[op])
    assert ierr == 0
    lib().pdo_oracle_stagg(_p(LU), C.c_int(n), C.c_double(dz), C.c_int(op), _p(f.view(np.float64)), _p(out.view(np.float64)),
                           C.c_int64(m))
    return out


# ---------------------------------------------------------------- decomposition / transposes
class _Decomp(C.Structure):
    _fields_ = [(nm, C.c_int * 3) for nm in ("xst", "xen", "xsz", "yst", "yen", "ysz", "zst", "zen", "zsz")]


def distribute(n, p):
    st = (C.c_int * p)()
    en = (C.c_int * p)()
    sz = (C.c_int * p)()
    lib().pdo_oracle_distribute(C.c_int(n), C.c_int(p), st, en, sz)
    return list(st), list(en), list(sz)


def decomp_info(nx, ny, nz, p_row, p_col, rank):
    d = _Decomp()
    lib().pdo_oracle_decomp_info(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(p_row), C.c_int(p_col), C.c_int(rank), C.byref(d))
    return {nm: tuple(getattr(d, nm)) for nm, _ in _Decomp._fields_}


_PEN_OF_DIR = {0: ("x", "y"), 1: ("y", "x"), 2: ("y", "z"), 3: ("z", "y")}


def transpose(direction, nx, ny, nz, p_row, p_col, src_list):
    """Run the reference's pack → ALLTOALLV → unpack with all ranks simulated.

    direction: 0 x→y, 1 y→x, 2 y→z, 3 z→y (or the strings 'x2y','y2x','y2z','z2y').
    src_list[r]: rank r's pencil, shape (s3, s2, s1), float64 or complex128.  Returns dst_list.
    """
    if isinstance(direction, str):
        direction = ("x2y", "y2x", "y2z", "z2y").index(direction)
    P = p_row * p_col
    cplx = np.iscomplexobj(src_list[0])
    dt = np.complex128 if cplx else np.float64
    w = 2 if cplx else 1
    infos = [decomp_info(nx, ny, nz, p_row, p_col, r) for r in range(P)]
    sp, dp_ = _PEN_OF_DIR[direction]
    for r in range(P):
        assert tuple(src_list[r].shape) == tuple(reversed(infos[r][sp + "sz"])), (src_list[r].shape, infos[r][sp + "sz"])
    src_all = np.concatenate([np.ascontiguousarray(s, dtype=dt).ravel() for s in src_list])
    dsz = [infos[r][dp_ + "sz"] for r in range(P)]
    tot = sum(int(np.prod(s)) for s in dsz)
    dst_all = np.zeros(tot, dtype=dt)
    lib().pdo_oracle_transpose(C.c_int(direction), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(p_row), C.c_int(p_col),
                               C.c_int(w), _p(src_all.view(np.float64)), _p(dst_all.view(np.float64)))
    out, off = [], 0
    for r in range(P):
        cnt = int(np.prod(dsz[r]))
        out.append(dst_all[off:off + cnt].reshape(tuple(reversed(dsz[r]))).copy())
        off += cnt
    return out


def scatter_global(g, nx, ny, nz, p_row, p_col, pencil):
    """Cut a global array g (nz, ny, nx) into the per-rank pencils 2DECOMP would hold."""
    out = []
    for r in range(p_row * p_col):
        d = decomp_info(nx, ny, nz, p_row, p_col, r)
        st, en = d[pencil + "st"], d[pencil + "en"]
        out.append(np.ascontiguousarray(g[st[2] - 1:en[2], st[1] - 1:en[1], st[0] - 1:en[0]]))
    return out


# ---------------------------------------------------------------- spectral
def wavenums(n, dx):
    k = np.zeros(n)
    lib().pdo_oracle_wavenums(C.c_int(n), C.c_double(dx), _p(k))
    return k


def poisson_multiply(rhs_hat, kx, ky, kz, have_zero=True):
    a = np.ascontiguousarray(rhs_hat, dtype=np.complex128).copy()
    nzh, nyh, nxh = a.shape
    lib().pdo_oracle_poisson_multiply(_p(a.view(np.float64)), C.c_int64(nxh), C.c_int64(nyh), C.c_int64(nzh),
                                      _p(np.ascontiguousarray(kx)), _p(np.ascontiguousarray(ky)), _p(np.ascontiguousarray(kz)),
                                      C.c_int(1 if have_zero else 0))
    return a


def poisson_solve(rhs, dx, dy, dz):
    """PoissonPeriodic%poisson_solve on one rank (dir_id=1): fft3_x2z → multiply → ifft3_z2x
    (utilities/PoissonPeriodic.F90:62-74; fft_3d.F90:588-613, 670-696).  rhs: (nz, ny, nx) real.
    FFT arithmetic: numpy pocketfft standing in for FFTW (unnormalised forward, 1/(nx ny nz) on inverse)."""
    nz, ny, nx = rhs.shape
    h = np.fft.rfft(rhs, axis=2)
    h = np.fft.fft(h, axis=1)
    h = np.fft.fft(h, axis=0)
    kx = wavenums(nx, dx)[: nx // 2 + 1]
    ky = wavenums(ny, dy)
    kz = wavenums(nz, dz)
    h = poisson_multiply(h, kx, ky, kz, True)
    h = np.fft.ifft(h, axis=0) * nz  # FFTW backward is unnormalised
    h = np.fft.ifft(h, axis=1) * ny
    f = np.fft.irfft(h, n=nx, axis=2) * nx
    return f * (1.0 / (float(nx) * float(ny) * float(nz)))


# ---- utilities/operators.F90: vector calculus on the GLOBAL field (the distributed results are its pencil slices) ----
def _dd(f, d, axis, method):
    return cd10(f, d, axis, 1) if method == "cd10" else cd06(f, d, axis)


def gradient(f, dx, dy, dz, method="cd10"):
    """operators.F90:17-53 (periodic): ddy, then ddx, then ddz of the same field."""
    return _dd(f, dx, 0, method), _dd(f, dy, 1, method), _dd(f, dz, 2, method)


def divergence(u, v, w, dx, dy, dz, method="cd10"):
    """operators.F90:118-151: div = dv/dy; div = div + du/dx; div = div + dw/dz (this order of additions)."""
    div = _dd(v, dy, 1, method)
    div = div + _dd(u, dx, 0, method)
    div = div + _dd(w, dz, 2, method)
    return div


def curl(u, v, w, dx, dy, dz, method="cd10"):
    """operators.F90:55-116: (dw/dy - dv/dz, du/dz - dw/dx, dv/dx - du/dy), returned as an array (3, nz, ny, nx)."""
    c1 = _dd(w, dy, 1, method) - _dd(v, dz, 2, method)
    c2 = _dd(u, dz, 2, method) - _dd(w, dx, 0, method)
    c3 = _dd(v, dx, 0, method) - _dd(u, dy, 1, method)
    return np.stack([c1, c2, c3])


def filter3D(f, numtimes=1, methods=("cf90", "cf90", "cf90"), periodic=(True, True, True), x_bc=(0, 0), y_bc=(0, 0), z_bc=(0, 0)):
    """operators.F90:158-224: numtimes passes of fil%filtery, then of fil%filterx, then of fil%filterz (this order; the
    transposes between the stages are pure permutations).  f: global array (nz, ny, nx); returns the filtered copy."""
    def one(a, axis, bc):
        if methods[axis].startswith("gaussian"):
            return gaussian(a, axis) if periodic[axis] else gaussian_np(a, axis, bc[0], bc[1])
        if methods[axis].startswith("lstsq"):
            return lstsq(a, axis) if periodic[axis] else lstsq_np(a, axis)
        return cf90(a, axis) if periodic[axis] else cf90_np(a, axis, bc[0], bc[1])
    out = np.ascontiguousarray(f, dtype=np.float64)
    for axis, bc in ((1, y_bc), (0, x_bc), (2, z_bc)):
        for _ in range(max(int(numtimes), 1)):
            out = one(out, axis, bc)
    return out


# ---- non-periodic CD10 closures (cd10.F90:29-96, 429-707, 823-851, 1143-1262, 1636-1731); SURVEY.md 8f rank 2 groundwork ----
def cd10_np(f, dx, axis, which=1, bc1=0, bcn=0):
    """cd10%dd* (which=1) / d2d* (which=2) with periodic=.false. and boundary codes bc1, bcn in {0, 1, -1}."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    n, na, nb = _na_nb(f, axis)
    out = np.empty_like(f)
    rc = lib().pdo_oracle_cd10_np(C.c_int(n), C.c_double(dx), C.c_int(which), C.c_int(bc1), C.c_int(bcn), C.c_int(axis), _p(f), _p(out),
                                  C.c_int64(na), C.c_int64(nb))
    assert rc == 0, rc
    return out


def cd10_np_penta(n, which, bc1, bcn):
    """(ierr, penta(n,11)) of ComputePenta1/2: columns bt, b, d, a, at, e, obc, f, g, eobc (Fortran order: column-major)."""
    P = np.zeros((11, n))
    rc = lib().pdo_oracle_cd10_np_penta(C.c_int(n), C.c_int(which), C.c_int(bc1), C.c_int(bcn), _p(P))
    return rc, P


def cd10_np_solve_line(P, y):
    y = np.ascontiguousarray(y, dtype=np.float64).copy()
    lib().pdo_oracle_cd10_np_solve_line(C.c_int(y.size), _p(np.ascontiguousarray(P)), _p(y))
    return y


def cf90_np(f, axis, bc1=0, bcn=0):
    """cf90%filter1/2/3 with periodic=.false. (filters/cf90.F90:276-418, 532-558, 672-801)."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    n, na, nb = _na_nb(f, axis)
    out = np.empty_like(f)
    rc = lib().pdo_oracle_cf90_np(C.c_int(n), C.c_int(bc1), C.c_int(bcn), C.c_int(axis), _p(f), _p(out), C.c_int64(na), C.c_int64(nb))
    assert rc == 0, rc
    return out


GAUSS_COEFS = (3565.0 / 10368.0, 3091.0 / 12960.0, 1997.0 / 25920.0, 149.0 / 12960.0, 107.0 / 103680.0)    # gaussian.F90:15-19
# lstsq.F90:14-19: real(0.6744132, rkind) is a default-real (single precision) literal widened to double
LSTSQ_COEFS = (0.5, float(np.float32(0.6744132)) / 2.0, 0.0 / 2.0, float(np.float32(-0.1744132)) / 2.0, 0.0 / 2.0)


def lstsq(f, axis):
    """lstsq%filter1/2/3, periodic (filters/lstsq.F90:117-167): the explicit symmetric 9-point stencil with the least-squares
    coefficients, periodic wrap."""
    a, b, c, d, e = LSTSQ_COEFS
    ax = {0: 2, 1: 1, 2: 0}[axis]
    g = np.asarray(f, dtype=np.float64)
    R = lambda k: np.roll(g, -k, axis=ax)
    return a * (g) + b * (R(1) + R(-1)) + c * (R(2) + R(-2)) + d * (R(3) + R(-3)) + e * (R(4) + R(-4))


def lstsq_np(f, axis):
    """lstsq%filter* with periodic=.false. (lstsq.F90:169-212): always the one-sided rows b1..b4 (the same as the Gaussian filter's)"""
    return gaussian_np(f, axis, 0, 0, LSTSQ_COEFS)


def gaussian_np(f, axis, bc1=0, bcn=0, coefs=GAUSS_COEFS):
    """gaussian%filter1/2/3 with periodic=.false. (filters/gaussian.F90:22-46, 215-330): explicit filter; bc = 0: the four boundary
    rows b1..b4 at that end, bc = +1 / -1: the interior 9-point stencil on the even / odd reflection about the end point.
    Statement-by-statement numpy restatement (the line axis is moved to the front)."""
    agf, bgf, cgf, dgf, egf = coefs
    b1 = (5.0 / 6.0, 1.0 / 6.0)
    b2 = (2.0 / 3.0, 1.0 / 6.0)
    b3 = (31.0 / 64.0, 7.0 / 32.0, 5.0 / 128.0)
    b4 = (17.0 / 48.0, 15.0 / 64.0, 7.0 / 96.0, 1.0 / 64.0)
    if bc1 not in (0, 1, -1) or bcn not in (0, 1, -1):
        raise ValueError("Incorrect boundary specification (324)")
    ax = {0: 2, 1: 1, 2: 0}[axis]                      # numpy axis of the Fortran index `axis + 1`
    g = np.moveaxis(np.asarray(f, dtype=np.float64), ax, 0)
    n = g.shape[0]
    assert n >= 8
    out = np.empty_like(g)
    F = lambda i: g[i - 1]                              # 1-based like the Fortran

    def interior(i, L, R):
        """row i with left neighbours L(k) = f(i-k) and right neighbours R(k) = f(i+k), possibly reflected"""
        return agf * (F(i)) + bgf * (R(1) + L(1)) + cgf * (R(2) + L(2)) + dgf * (R(3) + L(3)) + egf * (R(4) + L(4))
    for i in range(5, n - 3):
        out[i - 1] = interior(i, lambda k: F(i - k), lambda k: F(i + k))
    if bc1 == 0:
        out[0] = b1[0] * (F(1)) + b1[1] * (F(2))
        out[1] = b2[0] * (F(2)) + b2[1] * (F(3) + F(1))
        out[2] = b3[0] * (F(3)) + b3[1] * (F(4) + F(2)) + b3[2] * (F(5) + F(1))
        out[3] = b4[0] * (F(4)) + b4[1] * (F(5) + F(3)) + b4[2] * (F(6) + F(2)) + b4[3] * (F(7) + F(1))
    else:
        for i in range(1, 5):                           # f(1-m) := +- f(1+m): "f(2) + f(2)", "- f(2) + ..." written out in the reference
            out[i - 1] = interior(i, lambda k: F(i - k) if i - k >= 1 else bc1 * F(2 - (i - k)), lambda k: F(i + k))
    if bcn == 0:
        out[n - 4] = b4[0] * (F(n - 3)) + b4[1] * (F(n - 2) + F(n - 4)) + b4[2] * (F(n - 1) + F(n - 5)) + b4[3] * (F(n) + F(n - 6))
        out[n - 3] = b3[0] * (F(n - 2)) + b3[1] * (F(n - 1) + F(n - 3)) + b3[2] * (F(n) + F(n - 4))
        out[n - 2] = b2[0] * (F(n - 1)) + b2[1] * (F(n) + F(n - 2))
        out[n - 1] = b1[0] * (F(n)) + b1[1] * (F(n - 1))
    else:
        for i in range(n - 3, n + 1):
            out[i - 1] = interior(i, lambda k: F(i - k), lambda k: F(i + k) if i + k <= n else bcn * F(2 * n - (i + k)))
    return np.ascontiguousarray(np.moveaxis(out, 0, ax))


def cf90_np_penta(n, bc1, bcn):
    P = np.zeros((11, n))
    rc = lib().pdo_oracle_cf90_np_penta(C.c_int(n), C.c_int(bc1), C.c_int(bcn), _p(P))
    return rc, P


def cd06_np(f, dx, axis):
    """cd06%dd1/dd2/dd3 with periodic=.false. (one-sided closure at both ends; cd06.F90:27-58, 264-327, 432-449, 551-590)."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    n, na, nb = _na_nb(f, axis)
    out = np.empty_like(f)
    rc = lib().pdo_oracle_cd06_np(C.c_int(n), C.c_double(dx), C.c_int(axis), _p(f), _p(out), C.c_int64(na), C.c_int64(nb))
    assert rc == 0, rc
    return out


def cd06_np_tri(n):
    """(ierr, T(6, n)): rows 0-2 = Tri1 columns (a*den, den, cp), rows 3-5 = the raw sub / diagonal / super entries."""
    T = np.zeros((6, n))
    rc = lib().pdo_oracle_cd06_np_tri(C.c_int(n), _p(T))
    return rc, T
