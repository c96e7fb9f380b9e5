"""fft_3d / PoissonPeriodic executed on the REFERENCE's own FFT arithmetic: FFTW 3.3.5 compiled from the tarball the reference
vendors (dependencies/fftw-3.3.5.tar.gz -> oracle/_ref/lib/libfftw3.so, `make -C oracle ref`), driven through the very plan tuples
utilities/fft_3d.F90 builds for base "x" (:256-306).

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  This is the one piece of genuine reference arithmetic that can be built in
this image: it pins oracle.py's numpy/pocketfft stand-in (tests/test_oracle_fftw_ref.py), generates the committed fixture
tests/golden/fftw_ref_golden.npz (tests/golden/make_fftw_ref_golden.py) and is the FFT of the CPU Poisson baseline.

Arrays follow oracle.py: f(n1,n2,n3) Fortran == numpy shape (n3,n2,n1) C-contiguous.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "lib", "libfftw3.so")
_lib = None

FFTW_FORWARD, FFTW_BACKWARD = -1, 1
FFTW_MEASURE, FFTW_ESTIMATE, FFTW_EXHAUSTIVE = 0, 1 << 6, 1 << 3   # fftw3.f: the reference passes FFTW_MEASURE / FFTW_EXHAUSTIVE


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        vp, ip, i = C.c_void_p, C.POINTER(C.c_int), C.c_int
        # fftw_plan_many_dft*(rank, n, howmany, in, inembed, istride, idist, out, onembed, ostride, odist, [sign,] flags)
        L.fftw_plan_many_dft.restype = vp
        L.fftw_plan_many_dft.argtypes = [i, ip, i, vp, ip, i, i, vp, ip, i, i, i, C.c_uint]
        L.fftw_plan_many_dft_r2c.restype = vp
        L.fftw_plan_many_dft_r2c.argtypes = [i, ip, i, vp, ip, i, i, vp, ip, i, i, C.c_uint]
        L.fftw_plan_many_dft_c2r.restype = vp
        L.fftw_plan_many_dft_c2r.argtypes = [i, ip, i, vp, ip, i, i, vp, ip, i, i, C.c_uint]
        for nm in ("fftw_execute_dft", "fftw_execute_dft_r2c", "fftw_execute_dft_c2r"):
            getattr(L, nm).restype = None
            getattr(L, nm).argtypes = [vp, vp, vp]
        L.fftw_destroy_plan.argtypes = [vp]
        _lib = L
    return _lib


def version():
    return (C.c_char * 32).in_dll(lib(), "fftw_version").value.decode()   # `const char fftw_version[]`


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class FFT3D:
    """fft_3d%init(nx, ny, nz, "x", ...) on one rank: same scratch pencils, same five plans (fft_3d.F90:243-306; the Fortran
    legacy interface dfftw_plan_many_* forwards its arguments to these C entry points unchanged)."""

    def __init__(self, nx, ny, nz, flags=FFTW_MEASURE):
        L = lib()
        self.nx, self.ny, self.nz = nx, ny, nz
        self.nxh = nxh = nx // 2 + 1
        self.normfactor = 1.0 / (float(nx) * float(ny) * float(nz))     # fft_3d.F90:359-361
        self.normfactor2d = 1.0 / (float(nx) * float(ny))
        self.f_xhat_in_xD = np.zeros((nz, ny, nxh), np.complex128)
        self.f_xyhat_in_yD = np.zeros((nz, ny, nxh), np.complex128)
        self.f_xyzhat_in_zD = np.zeros((nz, ny, nxh), np.complex128)
        one = lambda v: (C.c_int * 1)(v)
        temp = np.zeros((nz, ny, nx))
        # r2c-x (:256-259) / c2r-x (:262-265): n = nx, howmany = xsz2*xsz3, stride 1, dist nx -> nx/2+1
        self.plan_r2c_x = L.fftw_plan_many_dft_r2c(1, one(nx), ny * nz, _ptr(temp), one(nx), 1, nx, _ptr(self.f_xhat_in_xD), one(nxh), 1, nxh, flags)
        self.plan_c2r_x = L.fftw_plan_many_dft_c2r(1, one(nx), ny * nz, _ptr(self.f_xhat_in_xD), one(nxh), 1, nxh, _ptr(temp), one(nx), 1, nx, flags)
        # c2c-y per z-plane (:274-289): n = ysz2, howmany = ysz1, stride = ysz1, dist = 1
        a2 = np.zeros((ny, nxh), np.complex128)
        b2 = np.zeros((ny, nxh), np.complex128)
        self.plan_c2c_fwd_y = L.fftw_plan_many_dft(1, one(ny), nxh, _ptr(a2), one(ny), nxh, 1, _ptr(a2), one(ny), nxh, 1, FFTW_FORWARD, flags)
        self.plan_c2c_bwd_y = L.fftw_plan_many_dft(1, one(ny), nxh, _ptr(a2), one(ny), nxh, 1, _ptr(a2), one(ny), nxh, 1, FFTW_BACKWARD, flags)
        self.plan_c2c_bwd_y_oop = L.fftw_plan_many_dft(1, one(ny), nxh, _ptr(a2), one(ny), nxh, 1, _ptr(b2), one(ny), nxh, 1, FFTW_BACKWARD, flags)
        # c2c-z (:295-306): n = zsz3, howmany = zsz1*zsz2, stride = zsz1*zsz2, dist = 1; backward out of place, forward in place
        dz_ = np.zeros((nz, ny, nxh), np.complex128)
        self.plan_c2c_bwd_z = L.fftw_plan_many_dft(1, one(nz), nxh * ny, _ptr(dz_), one(nz), nxh * ny, 1, _ptr(self.f_xyzhat_in_zD), one(nz), nxh * ny, 1, FFTW_BACKWARD, flags)
        self.plan_c2c_fwd_z = L.fftw_plan_many_dft(1, one(nz), nxh * ny, _ptr(self.f_xyzhat_in_zD), one(nz), nxh * ny, 1, _ptr(self.f_xyzhat_in_zD), one(nz), nxh * ny, 1, FFTW_FORWARD, flags)
        self._keep = (temp, a2, b2, dz_)
        # planning with FFTW_MEASURE overwrites its arrays
        for a in (self.f_xhat_in_xD, self.f_xyhat_in_yD, self.f_xyzhat_in_zD):
            a[...] = 0

    def _yplanes(self, plan, src, dst):
        L = lib()
        for k in range(self.nz):
            L.fftw_execute_dft(plan, _ptr(src[k]), _ptr(dst[k]))

    def fft2_x2y(self, inp):
        """fft_3d.F90:645-663 (one rank: transpose_x_to_y is a copy)"""
        L = lib()
        inp = np.ascontiguousarray(inp, np.float64)
        L.fftw_execute_dft_r2c(self.plan_r2c_x, _ptr(inp), _ptr(self.f_xhat_in_xD))
        out = self.f_xhat_in_xD.copy()
        self._yplanes(self.plan_c2c_fwd_y, out, out)
        return out

    def ifft2_y2x(self, inp, setOddBall=False):
        """fft_3d.F90:616-643"""
        L = lib()
        inp = np.ascontiguousarray(inp, np.complex128)
        self._yplanes(self.plan_c2c_bwd_y_oop, inp, self.f_xyhat_in_yD)
        self.f_xhat_in_xD[...] = self.f_xyhat_in_yD
        if setOddBall:
            self.f_xhat_in_xD[:, :, self.nx // 2] = 0.0
        out = np.empty((self.nz, self.ny, self.nx))
        L.fftw_execute_dft_c2r(self.plan_c2r_x, _ptr(self.f_xhat_in_xD), _ptr(out))
        return out * self.normfactor2d

    def fft3_x2z(self, inp):
        """fft_3d.F90:588-613"""
        L = lib()
        inp = np.ascontiguousarray(inp, np.float64)
        L.fftw_execute_dft_r2c(self.plan_r2c_x, _ptr(inp), _ptr(self.f_xhat_in_xD))
        self.f_xyhat_in_yD[...] = self.f_xhat_in_xD
        self._yplanes(self.plan_c2c_fwd_y, self.f_xyhat_in_yD, self.f_xyhat_in_yD)
        out = self.f_xyhat_in_yD.copy()
        L.fftw_execute_dft(self.plan_c2c_fwd_z, _ptr(out), _ptr(out))
        return out

    def ifft3_z2x(self, inp):
        """fft_3d.F90:670-696"""
        L = lib()
        inp = np.ascontiguousarray(inp, np.complex128).copy()   # FFTW's out-of-place c2c preserves its input; the copy is for safety
        L.fftw_execute_dft(self.plan_c2c_bwd_z, _ptr(inp), _ptr(self.f_xyzhat_in_zD))
        self.f_xyhat_in_yD[...] = self.f_xyzhat_in_zD
        self._yplanes(self.plan_c2c_bwd_y, self.f_xyhat_in_yD, self.f_xyhat_in_yD)
        self.f_xhat_in_xD[...] = self.f_xyhat_in_yD
        out = np.empty((self.nz, self.ny, self.nx))
        L.fftw_execute_dft_c2r(self.plan_c2r_x, _ptr(self.f_xhat_in_xD), _ptr(out))
        return out * self.normfactor

    def destroy(self):
        L = lib()
        for nm in ("plan_r2c_x", "plan_c2r_x", "plan_c2c_fwd_y", "plan_c2c_bwd_y", "plan_c2c_bwd_y_oop", "plan_c2c_bwd_z", "plan_c2c_fwd_z"):
            L.fftw_destroy_plan(getattr(self, nm))


def poisson_solve(rhs, dx, dy, dz, flags=FFTW_MEASURE, fft=None):
    """PoissonPeriodic%poisson_solve, dir_id = 1 (PoissonPeriodic.F90:62-74): fft3_x2z -> poisson3D_multiply -> ifft3_z2x with FFTW."""
    from . import oracle as O
    nz, ny, nx = rhs.shape
    F = fft or FFT3D(nx, ny, nz, flags)
    h = F.fft3_x2z(rhs)
    h = O.poisson_multiply(h, O.wavenums(nx, dx)[: nx // 2 + 1], O.wavenums(ny, dy), O.wavenums(nz, dz), True)
    out = F.ifft3_z2x(h)
    if fft is None:
        F.destroy()
    return out


class npfft:
    """numpy.fft's four 1-D entry points (fft / ifft / rfft / irfft along one axis, numpy's normalisation) executed by FFTW 3.3.5:
    lets a test run a whole numpy restatement (igrid_oracle.py's substep) on the reference's FFT arithmetic instead of pocketfft."""

    @staticmethod
    def _c2c(a, axis, sign):
        L = lib()
        a = np.moveaxis(np.asarray(a, dtype=np.complex128), axis, -1)
        src = np.ascontiguousarray(a).copy()
        n = src.shape[-1]
        how = src.size // n
        dst = np.empty_like(src)
        nn = (C.c_int * 1)(n)
        p = L.fftw_plan_many_dft(1, nn, how, _ptr(src), nn, 1, n, _ptr(dst), nn, 1, n, sign, FFTW_ESTIMATE)
        L.fftw_execute_dft(p, _ptr(src), _ptr(dst))
        L.fftw_destroy_plan(p)
        return np.moveaxis(dst, -1, axis)

    @staticmethod
    def fft(a, axis=-1):
        return npfft._c2c(a, axis, FFTW_FORWARD)

    @staticmethod
    def ifft(a, axis=-1):
        a = np.asarray(a)
        return npfft._c2c(a, axis, FFTW_BACKWARD) / a.shape[axis]

    @staticmethod
    def rfft(a, axis=-1):
        L = lib()
        src = np.ascontiguousarray(np.moveaxis(np.asarray(a, dtype=np.float64), axis, -1)).copy()
        n = src.shape[-1]
        nh = n // 2 + 1
        how = src.size // n
        dst = np.empty(src.shape[:-1] + (nh,), np.complex128)
        p = L.fftw_plan_many_dft_r2c(1, (C.c_int * 1)(n), how, _ptr(src), (C.c_int * 1)(n), 1, n, _ptr(dst), (C.c_int * 1)(nh), 1, nh, FFTW_ESTIMATE)
        L.fftw_execute_dft_r2c(p, _ptr(src), _ptr(dst))
        L.fftw_destroy_plan(p)
        return np.moveaxis(dst, -1, axis)

    @staticmethod
    def irfft(a, n=None, axis=-1):
        L = lib()
        src = np.ascontiguousarray(np.moveaxis(np.asarray(a, dtype=np.complex128), axis, -1)).copy()
        nh = src.shape[-1]
        n = n if n is not None else 2 * (nh - 1)
        assert nh == n // 2 + 1
        how = src.size // nh
        dst = np.empty(src.shape[:-1] + (n,), np.float64)
        p = L.fftw_plan_many_dft_c2r(1, (C.c_int * 1)(n), how, _ptr(src), (C.c_int * 1)(nh), 1, nh, _ptr(dst), (C.c_int * 1)(n), 1, n, FFTW_ESTIMATE)
        L.fftw_execute_dft_c2r(p, _ptr(src), _ptr(dst))
        L.fftw_destroy_plan(p)
        return np.moveaxis(dst, -1, axis) / n
