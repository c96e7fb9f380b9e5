"""Mirror of the igrid-side module interfaces on the C ABI (include/padeops_b200.h):
spectralMod::spectral ("x" pencil, dimTransform = 2; incompressible/spectral.F90), PadeDerOps::Pade6stagg (periodic;
PadeDerOps.F90), PadePoissonMod::padepoisson (PeriodicInZ; PadePoisson.F90) and IncompressibleGrid::igrid's periodic
substep (igrid.F90).  Arrays are torch tensors (device) or numpy arrays (host) in the reference's Fortran layout:
f(n1,n2,n3) has shape (n3, n2, n1); complex arrays are complex128."""
import ctypes as C

from ._lib import DecompInfo, IgridParams, check, lib, ptr, stream_ptr
from .decomp import decomp_2d


def _info(fn, h, *a):
    d = DecompInfo()
    check(fn(h, *a, C.byref(d)))
    return {nm: tuple(getattr(d, nm)) for nm, _ in DecompInfo._fields_}


def _empty(like, shape, complex_):
    import torch
    if hasattr(like, "new_empty"):
        return like.new_empty(tuple(shape), dtype=torch.complex128 if complex_ else torch.float64)
    import numpy as np
    return np.empty(tuple(shape), dtype=np.complex128 if complex_ else np.float64)


class spectral:
    def __init__(self):
        self._h = C.c_void_p(None)

    def init(self, pencil, nx_g, ny_g, nz_g, dx, dy, dz, scheme="four", filt="2/3rd", dimTransform=2, fixOddball=True,
             init_periodicInZ=False, dealiasF=2.0 / 3.0, p_row=0, p_col=0):
        if pencil != "x" or dimTransform != 2:
            raise NotImplementedError("only pencil='x', dimTransform=2 (the igrid configuration) is in scope")
        decomp_2d.comm_init()
        check(lib().pdo_spectral_init(C.byref(self._h), int(nx_g), int(ny_g), int(nz_g), float(dx), float(dy), float(dz), int(p_row),
                                      int(p_col), int(bool(fixOddball)), int(bool(init_periodicInZ)), float(dealiasF)))
        self.nx_g, self.ny_g, self.nz_g = nx_g, ny_g, nz_g
        self.physdecomp = _info(lib().pdo_spectral_get_physical_info, self._h)
        self.spectdecomp = _info(lib().pdo_spectral_get_spectral_info, self._h)
        return 0

    def destroy(self):
        if self._h:
            lib().pdo_spectral_destroy(self._h)
            self._h = C.c_void_p(None)

    def tables(self):
        import numpy as np
        nxh = self.nx_g // 2 + 1
        t = [np.zeros(nxh), np.zeros(self.ny_g), np.zeros(nxh), np.zeros(self.ny_g), np.zeros(self.nz_g)]
        check(lib().pdo_spectral_get_tables(self._h, *[C.c_void_p(a.ctypes.data) for a in t]))
        return dict(zip(("k1", "k2", "gx", "gy", "gz"), t))

    def fft(self, arr_in, arr_out=None, stream=None):
        if arr_out is None:
            arr_out = _empty(arr_in, reversed(self.spectdecomp["ysz"]), True)
        check(lib().pdo_spectral_fft(self._h, ptr(arr_in), ptr(arr_out), stream_ptr(stream)))
        return arr_out

    def ifft(self, arr_in, arr_out=None, setOddball=False, stream=None):
        if arr_out is None:
            arr_out = _empty(arr_in, reversed(self.physdecomp["xsz"]), False)
        check(lib().pdo_spectral_ifft(self._h, ptr(arr_in), ptr(arr_out), int(bool(setOddball)), stream_ptr(stream)))
        return arr_out

    def mTimes_ik1_oop(self, fin, fout=None, stream=None):
        fout = _empty(fin, fin.shape, True) if fout is None else fout
        check(lib().pdo_spectral_mtimes_ik1_oop(self._h, ptr(fin), ptr(fout), stream_ptr(stream)))
        return fout

    def mTimes_ik2_oop(self, fin, fout=None, stream=None):
        fout = _empty(fin, fin.shape, True) if fout is None else fout
        check(lib().pdo_spectral_mtimes_ik2_oop(self._h, ptr(fin), ptr(fout), stream_ptr(stream)))
        return fout

    def mTimes_ik1_ip(self, f, stream=None):
        check(lib().pdo_spectral_mtimes_ik1_ip(self._h, ptr(f), stream_ptr(stream)))
        return f

    def mTimes_ik2_ip(self, f, stream=None):
        check(lib().pdo_spectral_mtimes_ik2_ip(self._h, ptr(f), stream_ptr(stream)))
        return f

    def dealias(self, fhat, stream=None):
        check(lib().pdo_spectral_dealias(self._h, ptr(fhat), stream_ptr(stream)))
        return fhat

    def dealias_edgeField(self, fhat, stream=None):
        check(lib().pdo_spectral_dealias_edgefield(self._h, ptr(fhat), stream_ptr(stream)))
        return fhat

    def take_fft1d_z2z_ip(self, a, stream=None):
        check(lib().pdo_spectral_take_fft1d_z2z_ip(self._h, ptr(a), stream_ptr(stream)))
        return a

    def take_ifft1d_z2z_ip(self, a, stream=None):
        check(lib().pdo_spectral_take_ifft1d_z2z_ip(self._h, ptr(a), stream_ptr(stream)))
        return a

    def ddz_C2C_real_inplace(self, a, stream=None):
        """real z-pencil of physdecomp; the oddball mode passes through (spectral.F90:507-526)"""
        check(lib().pdo_spectral_ddz_c2c_real_ip(self._h, ptr(a), stream_ptr(stream)))
        return a

    def ddz_C2C_complex_inplace(self, a, stream=None):
        check(lib().pdo_spectral_ddz_c2c_complex_ip(self._h, ptr(a), stream_ptr(stream)))
        return a

    def shiftz_E2C(self, a, stream=None):
        check(lib().pdo_spectral_shiftz_e2c(self._h, ptr(a), stream_ptr(stream)))
        return a

    def shiftz_C2E(self, a, stream=None):
        check(lib().pdo_spectral_shiftz_c2e(self._h, ptr(a), stream_ptr(stream)))
        return a


class HIT_shell_forcing:
    """forcingmod::HIT_shell_forcing (forcingIsotropic.F90:45-314); the &HIT_Forcing namelist enters as keyword arguments.
    Right-hand sides and fields: device tensors, complex y-pencils of the cell / edge spectral decompositions."""

    def __init__(self):
        self._h = C.c_void_p(None)
        self.Nwaves = 0

    def init(self, spectC, spectE, kmin=2.0, kmax=10.0, Nwaves=20, EpsAmplitude=0.1, tidStart=0, RandSeedToAdd=0):
        self.destroy()
        self._keep = (spectC, spectE)
        self.Nwaves = int(Nwaves)
        check(lib().pdo_hit_forcing_init(C.byref(self._h), spectC._h, spectE._h, float(kmin), float(kmax), int(Nwaves), float(EpsAmplitude),
                                         int(tidStart), int(RandSeedToAdd)))
        return 0

    def destroy(self):
        if self._h:
            lib().pdo_hit_forcing_destroy(self._h)
            self._h = C.c_void_p(None)

    def set_wavenumbers(self, wave_x, wave_y, wave_z):
        a = [(C.c_int * self.Nwaves)(*[int(v) for v in w]) for w in (wave_x, wave_y, wave_z)]
        check(lib().pdo_hit_forcing_set_wavenumbers(self._h, *a))

    def get_wavenumbers(self):
        a = [(C.c_int * self.Nwaves)() for _ in range(3)]
        check(lib().pdo_hit_forcing_get_wavenumbers(self._h, *a))
        return tuple(list(x) for x in a)

    def getRHS_HITforcing(self, urhs_xy, vrhs_xy, wrhs_xy, uhat_xy, vhat_xy, what_xy, newTimestep, stream=None):
        check(lib().pdo_hit_forcing_get_rhs(self._h, ptr(urhs_xy), ptr(vrhs_xy), ptr(wrhs_xy), ptr(uhat_xy), ptr(vhat_xy), ptr(what_xy),
                                            int(bool(newTimestep)), stream_ptr(stream)))
        return urhs_xy, vrhs_xy, wrhs_xy


class Ops_Periodic:
    """igrid_Operators_Periodic::Ops_Periodic (igrid_operators_periodic.F90:13-161).  Real arrays: x-pencils of the physical
    decomposition; `gp` enters as its process grid (p_row, p_col; 0, 0 = 1 x nproc)."""

    def __init__(self):
        self._h = C.c_void_p(None)
        self.spect = None

    def init(self, nx, ny, nz, dx, dy, dz, p_row=0, p_col=0, InputDir="", OutputDir=""):
        self.destroy()
        decomp_2d.comm_init()
        check(lib().pdo_ops_periodic_init(C.byref(self._h), int(nx), int(ny), int(nz), float(dx), float(dy), float(dz), int(p_row), int(p_col)))
        sp = spectral()                      # link_spect: a borrowed view of the type the handle owns
        sp._h = C.c_void_p(lib().pdo_ops_periodic_spect(self._h))
        sp.nx_g, sp.ny_g, sp.nz_g = nx, ny, nz
        sp.physdecomp = _info(lib().pdo_spectral_get_physical_info, sp._h)
        sp.spectdecomp = _info(lib().pdo_spectral_get_spectral_info, sp._h)
        sp.destroy = lambda: None
        self.spect = sp
        self.inputdir, self.outputdir = InputDir, OutputDir
        return 0

    def destroy(self):
        if self._h:
            lib().pdo_ops_periodic_destroy(self._h)
            self._h = C.c_void_p(None)
            self.spect = None

    def link_spect(self):
        return self.spect

    def _rr(self, name, f, out, stream):
        out = _empty(f, f.shape, False) if out is None else out
        check(getattr(lib(), "pdo_ops_periodic_" + name)(self._h, ptr(f), ptr(out), stream_ptr(stream)))
        return out

    def ddx(self, f, dfdx=None, stream=None):
        return self._rr("ddx", f, dfdx, stream)

    def ddy(self, f, dfdy=None, stream=None):
        return self._rr("ddy", f, dfdy, stream)

    def ddz(self, f, dfdz=None, stream=None):
        return self._rr("ddz", f, dfdz, stream)

    def ddz_cmplx2cmplx(self, fhat, stream=None):
        check(lib().pdo_ops_periodic_ddz_cmplx2cmplx(self._h, ptr(fhat), stream_ptr(stream)))
        return fhat

    def SolvePoisson_oop(self, rhs, p=None, stream=None):
        return self._rr("solve_poisson", rhs, p, stream)

    def SolvePoisson_ip(self, rhs, stream=None):
        return self._rr("solve_poisson", rhs, rhs, stream)

    def dealiasField(self, f, stream=None):
        check(lib().pdo_ops_periodic_dealias_field(self._h, ptr(f), stream_ptr(stream)))
        return f

    def WriteField3D(self, field, label, tidx, runID, newOutputDir=None):
        assert len(label) == 4, "label is character(len=4) in the reference"
        check(lib().pdo_ops_periodic_write_field3d(self._h, ptr(field), label.encode(), int(tidx), int(runID),
                                                   str(newOutputDir if newOutputDir is not None else self.outputdir).encode()))

    def ReadField3D(self, field, label, tidx, runID, newinputdir=None):
        assert len(label) == 4
        check(lib().pdo_ops_periodic_read_field3d(self._h, ptr(field), label.encode(), int(tidx), int(runID),
                                                  str(newinputdir if newinputdir is not None else self.inputdir).encode()))
        return field

    def allocate3Dfield(self, device="cuda"):
        import torch
        return torch.empty(tuple(reversed(self.spect.physdecomp["xsz"])), dtype=torch.float64, device=device)


fd02, cd06, fourierColl = 0, 1, 2


class Pade6stagg:
    def __init__(self):
        self._h = C.c_void_p(None)

    def init(self, gpC, sp_gpC, gpE=None, sp_gpE=None, dz=1.0, scheme=cd06, isPeriodic=True, spectC=None):
        """gpC / sp_gpC: dicts (or objects) with the z-pencil sizes `zsz` of the physical / spectral cell decompositions."""
        g = gpC["zsz"] if isinstance(gpC, dict) else gpC.zsz
        s = sp_gpC["zsz"] if isinstance(sp_gpC, dict) else sp_gpC.zsz
        self.gp_zsz, self.sp_zsz = tuple(g), tuple(s)
        # spectC is the spectral type whose z transforms scheme = fourierColl uses (PadeDerOps.F90:73-78; code 43 without it)
        self._spectC = spectC
        check(lib().pdo_pade6stagg_init2(C.byref(self._h), (C.c_int * 3)(*g), (C.c_int * 3)(*s), float(dz), int(scheme), int(bool(isPeriodic)),
                                         spectC._h if spectC is not None else C.c_void_p(None)))
        return 0

    def destroy(self):
        if self._h:
            lib().pdo_pade6stagg_destroy(self._h)
            self._h = C.c_void_p(None)

    def _call(self, name, inp, out, edge_out, bot, top, stream):
        cplx = "complex" in str(inp.dtype)
        z = self.sp_zsz if cplx else self.gp_zsz
        if out is None:
            out = _empty(inp, (z[2] + (1 if edge_out else 0), z[1], z[0]), cplx)
        check(getattr(lib(), "pdo_pade6stagg_" + name)(self._h, ptr(inp), ptr(out), int(cplx), int(bot), int(top), stream_ptr(stream)))
        return out

    def ddz_C2E(self, inp, out=None, bot=0, top=0, stream=None): return self._call("ddz_C2E", inp, out, True, bot, top, stream)
    def ddz_E2C(self, inp, out=None, bot=0, top=0, stream=None): return self._call("ddz_E2C", inp, out, False, bot, top, stream)
    def interpz_C2E(self, inp, out=None, bot=0, top=0, stream=None): return self._call("interpz_C2E", inp, out, True, bot, top, stream)
    def interpz_E2C(self, inp, out=None, bot=0, top=0, stream=None): return self._call("interpz_E2C", inp, out, False, bot, top, stream)
    def d2dz2_C2C(self, inp, out=None, bot=0, top=0, stream=None): return self._call("d2dz2_C2C", inp, out, False, bot, top, stream)
    def d2dz2_E2E(self, inp, out=None, bot=0, top=0, stream=None): return self._call("d2dz2_E2E", inp, out, True, bot, top, stream)

    def getModifiedWavenumbers(self, k):
        import numpy as np
        k = np.ascontiguousarray(k, dtype=np.float64)
        kp = np.empty_like(k)
        check(lib().pdo_pade6stagg_get_modified_wavenumbers(self._h, C.c_void_p(k.ctypes.data), C.c_void_p(kp.ctypes.data), int(k.size)))
        return kp


class padepoisson:
    def __init__(self):
        self._h = C.c_void_p(None)

    def init(self, dx, dy, dz, sp, spE, computeStokesPressure=False, Lz=None, storePressure=True, gpC=None, derivZ=None, PeriodicInZ=True):
        self._keep = (sp, spE, derivZ)
        self._sp = sp
        check(lib().pdo_padepoisson_init3(C.byref(self._h), float(dx), float(dy), float(dz), sp._h, spE._h, derivZ._h, int(bool(PeriodicInZ)),
                                          int(bool(computeStokesPressure) and not PeriodicInZ), float(Lz) if Lz is not None else 0.0))
        return 0

    def destroy(self):
        if self._h:
            lib().pdo_padepoisson_destroy(self._h)
            self._h = C.c_void_p(None)

    def PressureProjection(self, uhat, vhat, what, stream=None):
        check(lib().pdo_padepoisson_pressure_projection(self._h, ptr(uhat), ptr(vhat), ptr(what), stream_ptr(stream)))

    def getPressure(self, uhat, vhat, what, pressure=None, stream=None):
        if pressure is None:
            pressure = _empty(uhat, reversed(self._sp.physdecomp["xsz"]), False)
        check(lib().pdo_padepoisson_get_pressure(self._h, ptr(uhat), ptr(vhat), ptr(what), ptr(pressure), stream_ptr(stream)))
        return pressure

    def getPressureAndUpdateRHS(self, uhat, vhat, what, pressure=None, stream=None):
        if pressure is None:
            pressure = _empty(uhat, reversed(self._sp.physdecomp["xsz"]), False)
        check(lib().pdo_padepoisson_get_pressure_and_update_rhs(self._h, ptr(uhat), ptr(vhat), ptr(what), ptr(pressure), stream_ptr(stream)))
        return pressure

    def DivergenceCheck(self, uhat, vhat, what, divergence=None, fixDiv=False, stream=None):
        """Returns (divergence, maxDiv) with maxDiv = p_maxval(maxval(divergence))."""
        if divergence is None:
            divergence = _empty(uhat, reversed(self._sp.physdecomp["xsz"]), False)
        md = C.c_double(0.0)
        check(lib().pdo_padepoisson_divergence_check(self._h, ptr(uhat), ptr(vhat), ptr(what), ptr(divergence), int(bool(fixDiv)),
                                                     C.byref(md), stream_ptr(stream)))
        return divergence, md.value


class igrid:
    """igrid%init / timeAdvance for the periodic substep; the namelist input file becomes keyword arguments."""
    FIELDS = {"u": 0, "v": 1, "w": 2, "wC": 3, "uE": 4, "vE": 5, "divergence": 6, "uhat": 10, "vhat": 11, "what": 12}

    def __init__(self):
        self._h = C.c_void_p(None)

    def init(self, nx, ny, nz, Lx, Ly, Lz, Re, u, v, w, isInviscid=False, dealiasFact=2.0 / 3.0, t_DivergenceCheck=10,
             TimeSteppingScheme=1, prow=0, pcol=0, use_d2dz2_C2C=True, computeAllGradients=False, AdvectionTerm=1, NumericalSchemeVert=1,
             PeriodicInZ=True, topWall=2, botWall=2, ComputeStokesPressure=True):
        """AdvectionTerm: 1 skew-symmetric (igrid.F90:1572-1679), 0 rotational u x omega (:1527-1555); NumericalSchemeVert: 1 cd06
        staggered compact operators, 2 Fourier collocation in z (PadeDerOps.F90:16-18) — as in the namelist."""
        if AdvectionTerm not in (0, 1):
            raise ValueError("AdvectionTerm must be 0 (rotational) or 1 (skew-symmetric)")
        if NumericalSchemeVert not in (1, 2):
            raise ValueError("NumericalSchemeVert must be 1 (cd06) or 2 (fourierColl); fd02 is not built")
        decomp_2d.comm_init()
        p = IgridParams(int(nx), int(ny), int(nz), float(Lx), float(Ly), float(Lz), float(Re), int(bool(isInviscid)), float(dealiasFact),
                        int(t_DivergenceCheck), int(TimeSteppingScheme), int(prow), int(pcol), int(bool(use_d2dz2_C2C)),
                        int(bool(computeAllGradients)), int(AdvectionTerm == 0), int(NumericalSchemeVert == 2),
                        int(not PeriodicInZ), int(topWall), int(botWall), int(not ComputeStokesPressure))
        check(lib().pdo_igrid_init(C.byref(self._h), C.byref(p), ptr(u), ptr(v), ptr(w)))
        self.gpC = _info(lib().pdo_igrid_get_decomp_info, self._h, 0)
        self.gpE = _info(lib().pdo_igrid_get_decomp_info, self._h, 1)
        self.sp_gpC = _info(lib().pdo_igrid_get_decomp_info, self._h, 2)
        self.sp_gpE = _info(lib().pdo_igrid_get_decomp_info, self._h, 3)
        return 0

    def destroy(self):
        if self._h:
            lib().pdo_igrid_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def timeAdvance(self, dtforced, stream=None):
        check(lib().pdo_igrid_time_advance(self._h, float(dtforced), stream_ptr(stream)))

    def get(self, name, out=None, stream=None):
        import numpy as np
        which = self.FIELDS[name]
        if out is None:
            if which >= 10:
                sz = self.sp_gpE["ysz"] if which == 12 else self.sp_gpC["ysz"]
                out = np.empty(tuple(reversed(sz)), dtype=np.complex128)
            else:
                sz = self.gpE["xsz"] if which in (2, 4, 5) else self.gpC["xsz"]
                out = np.empty(tuple(reversed(sz)), dtype=np.float64)
        check(lib().pdo_igrid_get_field(self._h, which, ptr(out), stream_ptr(stream)))
        return out

    @property
    def step(self):
        s = C.c_int(0)
        check(lib().pdo_igrid_get_state(self._h, C.byref(s), None))
        return s.value

    @property
    def tsim(self):
        t = C.c_double(0.0)
        check(lib().pdo_igrid_get_state(self._h, None, C.byref(t)))
        return t.value

    def compute_deltaT(self, CFL, stream=None):
        dt = C.c_double(0.0)
        check(lib().pdo_igrid_compute_delta_t(self._h, float(CFL), C.byref(dt), stream_ptr(stream)))
        return dt.value

    def maxDivergence(self, stream=None):
        md = C.c_double(0.0)
        check(lib().pdo_igrid_max_divergence(self._h, C.byref(md), stream_ptr(stream)))
        return md.value

    # ---- the reference's restart / field files (igrid.F90:2719-2823) ----
    def dumpRestartFile(self, OutputDir, runID=1):
        check(lib().pdo_igrid_dump_restart(self._h, str(OutputDir).encode(), int(runID)))

    def readRestartFile(self, tid, rid, InputDir):
        check(lib().pdo_igrid_read_restart(self._h, str(InputDir).encode(), int(rid), int(tid)))

    def dumpFullField(self, name, label, OutputDir, runID=1):
        assert len(label) == 4, "label is character(len=4) in the reference"
        check(lib().pdo_igrid_dump_full_field(self._h, self.FIELDS[name], label.encode(), str(OutputDir).encode(), int(runID)))

    def enableHITForcing(self, kmin=2.0, kmax=10.0, Nwaves=20, EpsAmplitude=0.1, RandSeedToAdd=0):
        """useHITForcing = .true. with the &HIT_Forcing namelist (igrid.F90:940-944); call once after init"""
        check(lib().pdo_igrid_enable_hit_forcing(self._h, float(kmin), float(kmax), int(Nwaves), float(EpsAmplitude), int(RandSeedToAdd)))
        self._hit_nwaves = int(Nwaves)

    def setHITWavenumbers(self, wave_x, wave_y, wave_z):
        """inject the draw of the next time step (e.g. the reference RNG's) into the handle's forcing object"""
        n = self._hit_nwaves
        a = [(C.c_int * n)(*[int(v) for v in w]) for w in (wave_x, wave_y, wave_z)]
        check(lib().pdo_hit_forcing_set_wavenumbers(C.c_void_p(lib().pdo_igrid_hit_forcing(self._h)), *a))

    def enableSGS(self, SGSModelID=2, Csgs=0.17, explicitCalcEdgeEddyViscosity=False):
        """useSGS = .true. with the &SGS_MODEL entries in scope (igrid.F90:1866-1871); init with computeAllGradients=True"""
        check(lib().pdo_igrid_enable_sgs(self._h, int(SGSModelID), float(Csgs), int(bool(explicitCalcEdgeEddyViscosity))))
