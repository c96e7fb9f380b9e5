"""Mirror of fft_3d_stuff::fft_3d ("x" base pencil; utilities/fft_3d.F90:69-91) and
PoissonPeriodicMod::PoissonPeriodic (utilities/PoissonPeriodic.F90:27-34) on the C ABI."""
import ctypes as C

from ._lib import DecompInfo, check, lib, ptr, stream_ptr
from .decomp import decomp_2d


def _empty(like, shape, complex_):
    import torch
    if hasattr(like, "new_empty"):
        dt = torch.complex128 if complex_ else torch.float64
        return like.new_empty(tuple(shape), dtype=dt)
    import numpy as np
    return np.empty(tuple(shape), dtype=np.complex128 if complex_ else np.float64)


class fft_3d:
    def __init__(self):
        self._h = C.c_void_p(None)

    def init(self, nx, ny, nz, base="x", dx=1.0, dy=1.0, dz=1.0, exhaustive=False, p_row=0, p_col=0):
        if base != "x":
            raise NotImplementedError("only the 'x' base pencil is in scope (SURVEY.md 2.1 #11)")
        decomp_2d.comm_init()
        rc = lib().pdo_fft3d_init(C.byref(self._h), int(nx), int(ny), int(nz), float(dx), float(dy), float(dz), int(p_row), int(p_col))
        if rc == 0:
            s, p = DecompInfo(), DecompInfo()
            check(lib().pdo_fft3d_get_spectral_info(self._h, C.byref(s)))
            check(lib().pdo_fft3d_get_physical_info(self._h, C.byref(p)))
            self.spectral = {nm: tuple(getattr(s, nm)) for nm, _ in DecompInfo._fields_}
            self.physical = {nm: tuple(getattr(p, nm)) for nm, _ in DecompInfo._fields_}
        return rc

    def destroy(self):
        if self._h:
            lib().pdo_fft3d_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def get_complex_output_size(self):
        sz = (C.c_int * 3)()
        check(lib().pdo_fft3d_get_complex_output_size(self._h, sz))
        return tuple(sz)

    def fft3_x2z(self, input, output=None, stream=None):
        if output is None:
            output = _empty(input, reversed(self.spectral["zsz"]), True)
        check(lib().pdo_fft3d_fft3_x2z(self._h, ptr(input), ptr(output), stream_ptr(stream)))
        return output

    def ifft3_z2x(self, input, output=None, stream=None):
        if output is None:
            output = _empty(input, reversed(self.physical["xsz"]), False)
        check(lib().pdo_fft3d_ifft3_z2x(self._h, ptr(input), ptr(output), stream_ptr(stream)))
        return output

    def fft2_x2y(self, input, output=None, stream=None):
        if output is None:
            output = _empty(input, reversed(self.spectral["ysz"]), True)
        check(lib().pdo_fft3d_fft2_x2y(self._h, ptr(input), ptr(output), stream_ptr(stream)))
        return output

    def ifft2_y2x(self, input, output=None, setOddBall=False, stream=None):
        if output is None:
            output = _empty(input, reversed(self.physical["xsz"]), False)
        check(lib().pdo_fft3d_ifft2_y2x(self._h, ptr(input), ptr(output), int(bool(setOddBall)), stream_ptr(stream)))
        return output


class PoissonPeriodic:
    def __init__(self):
        self._h = C.c_void_p(None)

    def init(self, dx, dy, dz, gp, dir_id=1, useExhaustiveFFT=False, modkx=None, modky=None, modkz=None, p_row=0, p_col=0):
        """gp: (nx, ny, nz) global sizes or a decomp_info.  modk*: optional already-modified wavenumber arrays
        (what the reference's Get_ModK* callbacks produce)."""
        import numpy as np
        decomp_2d.comm_init()
        nx, ny, nz = (gp.nx, gp.ny, gp.nz) if hasattr(gp, "nx") else gp
        ks = [None if k is None else np.ascontiguousarray(k, dtype=np.float64) for k in (modkx, modky, modkz)]
        self._keep = ks
        kp = [C.c_void_p(0) if k is None else C.c_void_p(k.ctypes.data) for k in ks]
        check(lib().pdo_poisson_init(C.byref(self._h), int(nx), int(ny), int(nz), float(dx), float(dy), float(dz), int(p_row),
                                     int(p_col), int(dir_id), kp[0], kp[1], kp[2]))

    def destroy(self):
        if self._h:
            lib().pdo_poisson_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def poisson_solve(self, rhs, f=None, stream=None):
        """Out-of-place poisson_solve(rhs, f); f=None → in place like the one-argument specific."""
        if f is None:
            f = rhs
        check(lib().pdo_poisson_solve(self._h, ptr(rhs), ptr(f), stream_ptr(stream)))
        return f
