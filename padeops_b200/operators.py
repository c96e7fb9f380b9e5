"""Host-side mirror of the reference's 1-D operator types on top of the C ABI.

Same names, constructor arguments and error behaviour as the Fortran types they stand for:
  cd10      derivatives/cd10.F90:108-183   (init :195, dd1/dd2/dd3 :2029-2237, d2d1/d2d2/d2d3 :2239-2447)
  cd06      derivatives/cd06.F90           (init :129, dd1/dd2/dd3 :775-839)
  cf90      filters/cf90.F90               (init :107, filter1/2/3 :1020-1228)
  gaussian  filters/gaussian.F90           (init :74,  filter1/2/3 :104, 336, 564)
  cd06stagg derivatives/cd06stagg.F90      (init_periodic :170, six z-ops :820-1059)
  derivatives  derivatives/derivatives.F90 (init :189-236, ddx..d2dz2 :447-569)
  filters      filters/filters.F90         (init :274-297, filterx/y/z :220-269)

Fields are torch CUDA tensors (device-resident fast path) or numpy arrays / CPU tensors (the library
stages them): a Fortran array f(n1,n2,n3) is a C-contiguous array of shape (n3,n2,n1).  `init`
returns the reference's ierr instead of raising, exactly like the Fortran `function init`.
"""
import ctypes as C

from . import _lib
from ._lib import PadeOpsError, check, lib, ptr, stream_ptr


def _alloc_like(f, shape=None):
    shape = tuple(f.shape) if shape is None else tuple(shape)
    if hasattr(f, "new_empty"):
        return f.new_empty(shape)
    import numpy as np
    return np.empty(shape, dtype=f.dtype)


def _contig(f):
    if hasattr(f, "is_contiguous"):
        assert f.is_contiguous(), "fields must be contiguous"
    else:
        assert f.flags.c_contiguous, "fields must be contiguous"
    return f


class _LineOp:
    _prefix = None

    def __init__(self):
        self._h = C.c_void_p(None)
        self.n = 0

    def destroy(self):
        if self._h:
            getattr(lib(), f"pdo_{self._prefix}_destroy")(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def GetSize(self):
        return self.n

    def _call(self, fn, f, out, na, nb, bc1, bcn, stream):
        _contig(f)
        if out is None:
            out = _alloc_like(f)
        _contig(out)
        check(getattr(lib(), f"pdo_{self._prefix}_{fn}")(self._h, ptr(f), ptr(out), int(na), int(nb), int(bc1), int(bcn),
                                                       stream_ptr(stream)))
        return out

    @staticmethod
    def _extents(f, axis):
        n3, n2, n1 = f.shape
        return ((n2, n3), (n1, n3), (n1, n2))[axis]

    def plan(self, axis, na, nb):
        """Time the kernel candidates for `axis` (0 x, 1 y, 2 z) of a pencil with the other extents (na, nb) and keep the winner
        in the handle (pdo_*_plan: explicit, set-up time; without it a fixed table picks the kernel).  Returns the variant code(s)."""
        if self._prefix == "cd10":
            v1, v2 = C.c_int(0), C.c_int(0)
            check(lib().pdo_cd10_plan(self._h, int(axis), int(na), int(nb), C.byref(v1), C.byref(v2)))
            return v1.value, v2.value
        v = C.c_int(0)
        check(getattr(lib(), f"pdo_{self._prefix}_plan")(self._h, int(axis), int(na), int(nb), C.byref(v)))
        return v.value


class cd10(_LineOp):
    _prefix = "cd10"

    def init(self, n_, dx_, periodic_=True, bc1_=0, bcn_=0):
        self.destroy()
        self.n = n_
        return lib().pdo_cd10_init(C.byref(self._h), int(n_), float(dx_), int(bool(periodic_)), int(bc1_), int(bcn_))

    def dd1(self, f, df=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 0)
        return self._call("dd1", f, df, na, nb, bc1_, bcn_, stream)

    def dd2(self, f, df=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 1)
        return self._call("dd2", f, df, na, nb, bc1_, bcn_, stream)

    def dd3(self, f, df=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 2)
        return self._call("dd3", f, df, na, nb, bc1_, bcn_, stream)

    def d2d1(self, f, df=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 0)
        return self._call("d2d1", f, df, na, nb, bc1_, bcn_, stream)

    def d2d2(self, f, df=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 1)
        return self._call("d2d2", f, df, na, nb, bc1_, bcn_, stream)

    def d2d3(self, f, df=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 2)
        return self._call("d2d3", f, df, na, nb, bc1_, bcn_, stream)


class cd06(_LineOp):
    _prefix = "cd06"

    def init(self, n_, dx_, periodic_=True, bc1_=0, bcn_=0):
        self.destroy()
        self.n = n_
        return lib().pdo_cd06_init(C.byref(self._h), int(n_), float(dx_), int(bool(periodic_)), int(bc1_), int(bcn_))

    def dd1(self, f, df=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 0)
        return self._call("dd1", f, df, na, nb, bc1_, bcn_, stream)

    def dd2(self, f, df=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 1)
        return self._call("dd2", f, df, na, nb, bc1_, bcn_, stream)

    def dd3(self, f, df=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 2)
        return self._call("dd3", f, df, na, nb, bc1_, bcn_, stream)


class _Filter(_LineOp):
    def init(self, n_, periodic_=True):
        self.destroy()
        self.n = n_
        return getattr(lib(), f"pdo_{self._prefix}_init")(C.byref(self._h), int(n_), int(bool(periodic_)))

    def filter1(self, f, fil=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 0)
        return self._call("filter1", f, fil, na, nb, bc1_, bcn_, stream)

    def filter2(self, f, fil=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 1)
        return self._call("filter2", f, fil, na, nb, bc1_, bcn_, stream)

    def filter3(self, f, fil=None, na=None, nb=None, bc1_=0, bcn_=0, stream=None):
        na, nb = (na, nb) if na is not None else self._extents(f, 2)
        return self._call("filter3", f, fil, na, nb, bc1_, bcn_, stream)


class cf90(_Filter):
    _prefix = "cf90"


class gaussian(_Filter):
    _prefix = "gaussian"


class lstsq(_Filter):
    """lstsqstuff::lstsq (filters/lstsq.F90): filter1/2/3(f, fil, na, nb) take no boundary codes"""
    _prefix = "lstsq"

    def _call(self, fn, f, out, na, nb, bc1, bcn, stream):
        _contig(f)
        if out is None:
            out = _alloc_like(f)
        check(getattr(lib(), f"pdo_lstsq_{fn}")(self._h, ptr(f), ptr(out), int(na), int(nb), stream_ptr(stream)))
        return out


class cd06stagg:
    """Staggered CD06 in z, periodic or with walls.  Cells: n planes; edges: n+1 planes (periodic: plane n+1 == plane 1)."""

    def __init__(self):
        self._h = C.c_void_p(None)
        self.n = 0

    def init(self, nx, dx, isTopEven=None, isBotEven=None, isTopSided=False, isBotSided=False):
        """The generic init: init(nx, dx) is the periodic overload (cd06stagg.F90:92, 170-195); with isTopEven / isBotEven it is
        init_nonperiodic (:197-231): the field is even / odd about that wall, or the wall takes the one-sided closure.
        Raises code 21 for nx <= 4 where the reference calls GracefulExit."""
        self.destroy()
        self.n = nx
        if isTopEven is None and isBotEven is None:
            check(lib().pdo_cd06stagg_init_periodic(C.byref(self._h), int(nx), float(dx)))
        else:
            check(lib().pdo_cd06stagg_init_nonperiodic(C.byref(self._h), int(nx), float(dx), int(bool(isTopEven)), int(bool(isBotEven)),
                                                       int(bool(isTopSided)), int(bool(isBotSided))))

    def destroy(self):
        if self._h:
            lib().pdo_cd06stagg_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def _call(self, fn, fin, out, n_in, n_out, stream):
        _contig(fin)
        assert fin.shape[0] == n_in, f"{fn}: expected {n_in} planes, got {fin.shape[0]}"
        n2, n1 = fin.shape[1], fin.shape[2]
        is_c = 1 if ("complex" in str(fin.dtype)) else 0
        if out is None:
            out = _alloc_like(fin, (n_out, n2, n1))
        _contig(out)
        check(getattr(lib(), f"pdo_cd06stagg_{fn}")(self._h, ptr(fin), ptr(out), int(n1), int(n2), is_c, stream_ptr(stream)))
        return out

    def ddz_E2C(self, fE, dfC=None, stream=None):
        return self._call("ddz_E2C", fE, dfC, self.n + 1, self.n, stream)

    def ddz_C2E(self, fC, dfE=None, stream=None):
        return self._call("ddz_C2E", fC, dfE, self.n, self.n + 1, stream)

    def InterpZ_E2C(self, fE, fC=None, stream=None):
        return self._call("interpz_E2C", fE, fC, self.n + 1, self.n, stream)

    def InterpZ_C2E(self, fC, fE=None, stream=None):
        return self._call("interpz_C2E", fC, fE, self.n, self.n + 1, stream)

    def d2dz2_C2C(self, fC, d2fC=None, stream=None):
        return self._call("d2dz2_C2C", fC, d2fC, self.n, self.n, stream)

    def d2dz2_E2E(self, fE, d2fE=None, stream=None):
        return self._call("d2dz2_E2E", fE, d2fE, self.n + 1, self.n + 1, stream)

    def ddz_C2C(self, fC, dfC=None, stream=None):
        """non-periodic handles only (the reference builds TriD1_C2C / TriD1_E2E in init_nonperiodic)"""
        return self._call("ddz_C2C", fC, dfC, self.n, self.n, stream)

    def ddz_E2E(self, fE, dfE=None, stream=None):
        return self._call("ddz_E2E", fE, dfE, self.n + 1, self.n + 1, stream)


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


class derivatives:
    """DerivativesMod::derivatives.  init(gp, dx,dy,dz, periodicx,y,z, methodx,y,z) with gp a decomp_info
    (anything with xsz/ysz/zsz) or the serial overload init(nx,ny,nz, ...)."""

    def __init__(self):
        self._h = C.c_void_p(None)

    def init(self, gp, dx, dy, dz, periodicx, periodicy, periodicz, methodx, methody, methodz):
        self.destroy()
        if isinstance(gp, (tuple, list)):
            xsz = ysz = zsz = tuple(gp)
        else:
            xsz, ysz, zsz = gp.xsz, gp.ysz, gp.zsz
        self.xsz, self.ysz, self.zsz = tuple(xsz), tuple(ysz), tuple(zsz)
        # the reference turns a non-zero ierr into GracefulExit (derivatives.F90:291-293): raise
        check(lib().pdo_derivatives_init(C.byref(self._h), _i3(xsz), _i3(ysz), _i3(zsz), float(dx), float(dy), float(dz),
                                         int(bool(periodicx)), int(bool(periodicy)), int(bool(periodicz)),
                                         methodx.encode(), methody.encode(), methodz.encode()))

    def destroy(self):
        if self._h:
            lib().pdo_derivatives_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def _call(self, fn, f, out, bc1, bcn, stream):
        _contig(f)
        if out is None:
            out = _alloc_like(f)
        check(getattr(lib(), f"pdo_derivatives_{fn}")(self._h, ptr(f), ptr(out), int(bc1), int(bcn), stream_ptr(stream)))
        return out

    def ddx(self, f, dfdx=None, bc1=0, bcn=0, stream=None):
        return self._call("ddx", f, dfdx, bc1, bcn, stream)

    def ddy(self, f, dfdy=None, bc1=0, bcn=0, stream=None):
        return self._call("ddy", f, dfdy, bc1, bcn, stream)

    def ddz(self, f, dfdz=None, bc1=0, bcn=0, stream=None):
        return self._call("ddz", f, dfdz, bc1, bcn, stream)

    def d2dx2(self, f, d2f=None, bc1=0, bcn=0, stream=None):
        return self._call("d2dx2", f, d2f, bc1, bcn, stream)

    def d2dy2(self, f, d2f=None, bc1=0, bcn=0, stream=None):
        return self._call("d2dy2", f, d2f, bc1, bcn, stream)

    def d2dz2(self, f, d2f=None, bc1=0, bcn=0, stream=None):
        return self._call("d2dz2", f, d2f, bc1, bcn, stream)


class filters:
    """FiltersMod::filters.  init(gp, periodicx,y,z, methodx,y,z)."""

    def __init__(self):
        self._h = C.c_void_p(None)

    def init(self, gp, periodicx, periodicy, periodicz, methodx, methody, methodz):
        self.destroy()
        if isinstance(gp, (tuple, list)):
            xsz = ysz = zsz = tuple(gp)
        else:
            xsz, ysz, zsz = gp.xsz, gp.ysz, gp.zsz
        check(lib().pdo_filters_init(C.byref(self._h), _i3(xsz), _i3(ysz), _i3(zsz), int(bool(periodicx)), int(bool(periodicy)),
                                     int(bool(periodicz)), methodx.encode(), methody.encode(), methodz.encode()))

    def destroy(self):
        if self._h:
            lib().pdo_filters_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def _call(self, fn, f, out, bc1, bcn, stream):
        _contig(f)
        if out is None:
            out = _alloc_like(f)
        check(getattr(lib(), f"pdo_filters_{fn}")(self._h, ptr(f), ptr(out), int(bc1), int(bcn), stream_ptr(stream)))
        return out

    def filterx(self, f, fil=None, bc1=0, bcn=0, stream=None):
        return self._call("filterx", f, fil, bc1, bcn, stream)

    def filtery(self, f, fil=None, bc1=0, bcn=0, stream=None):
        return self._call("filtery", f, fil, bc1, bcn, stream)

    def filterz(self, f, fil=None, bc1=0, bcn=0, stream=None):
        return self._call("filterz", f, fil, bc1, bcn, stream)
