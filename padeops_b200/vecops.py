"""Mirror of utilities/operators.F90 (gradient :17-53, curl :55-116, divergence :118-151, filter3D :158-224) on the C ABI.

Every field is a y-pencil array of `gp` (shape (ysz[2], ysz[1], ysz[0]) in torch's row-major view of the Fortran
layout), device-resident.  `vector_ops.init` is collective.  Along z the library distributes the compact solve across
the GPUs of a z-group (zmode == 1) instead of transposing, when it can; allow_zslab=False forces the reference's
transpose choreography (zmode == 2)."""
import ctypes as C

from ._lib import check, lib, ptr, stream_ptr


class vector_ops:
    def __init__(self):
        self._h = C.c_void_p(None)
        self.gp = None

    def init(self, gp, dx, dy, dz, method="cd10", allow_zslab=True):
        self.destroy()
        self.gp = gp
        check(lib().pdo_operators_init(C.byref(self._h), gp._h, float(dx), float(dy), float(dz), method.encode(), int(bool(allow_zslab))))
        self.zmode = int(lib().pdo_operators_zmode(self._h))
        return 0

    def destroy(self):
        if self._h:
            lib().pdo_operators_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def _shape(self):
        return tuple(reversed(self.gp.ysz))

    def _chk(self, *arrs):
        for a in arrs:
            assert tuple(a.shape) == self._shape() and a.is_contiguous(), (tuple(a.shape), self._shape())

    def _dd(self, name, f, out, stream):
        out = f.new_empty(f.shape) if out is None else out
        self._chk(f, out)
        check(getattr(lib(), "pdo_operators_" + name)(self._h, ptr(f), ptr(out), stream_ptr(stream)))
        return out

    def ddx(self, f, out=None, stream=None):
        return self._dd("ddx", f, out, stream)

    def ddy(self, f, out=None, stream=None):
        return self._dd("ddy", f, out, stream)

    def ddz(self, f, out=None, stream=None):
        return self._dd("ddz", f, out, stream)

    def gradient(self, f, dfdx=None, dfdy=None, dfdz=None, stream=None):
        dfdx, dfdy, dfdz = (f.new_empty(f.shape) if a is None else a for a in (dfdx, dfdy, dfdz))
        self._chk(f, dfdx, dfdy, dfdz)
        check(lib().pdo_operators_gradient(self._h, ptr(f), ptr(dfdx), ptr(dfdy), ptr(dfdz), stream_ptr(stream)))
        return dfdx, dfdy, dfdz

    def divergence(self, u, v, w, div=None, stream=None):
        div = u.new_empty(u.shape) if div is None else div
        self._chk(u, v, w, div)
        check(lib().pdo_operators_divergence(self._h, ptr(u), ptr(v), ptr(w), ptr(div), stream_ptr(stream)))
        return div

    def curl(self, u, v, w, curlu=None, stream=None):
        """curlu(:,:,:,c) of the reference = curlu[c] here (three consecutive y-pencils)."""
        curlu = u.new_empty((3,) + tuple(u.shape)) if curlu is None else curlu
        self._chk(u, v, w)
        assert tuple(curlu.shape) == (3,) + self._shape() and curlu.is_contiguous()
        check(lib().pdo_operators_curl(self._h, ptr(u), ptr(v), ptr(w), ptr(curlu), stream_ptr(stream)))
        return curlu

    def filter3D(self, fil, arr, numtimes=1, x_bc=None, y_bc=None, z_bc=None, stream=None):
        """filter3D(decomp, fil, arr, numtimes, x_bc_, y_bc_, z_bc_) (operators.F90:158-224): numtimes passes of the y, then x,
        then z filter of `fil` (a `filters` object built on the same gp), in place on the y-pencil field `arr`."""
        self._chk(arr)

        def bc(p):
            return None if p is None else (C.c_int * 2)(int(p[0]), int(p[1]))
        check(lib().pdo_operators_filter3d(self._h, fil._h, ptr(arr), int(numtimes), bc(x_bc), bc(y_bc), bc(z_bc), stream_ptr(stream)))
        return arr
