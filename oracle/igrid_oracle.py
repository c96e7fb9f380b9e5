"""CPU restatement of the igrid periodic substep and the spectral / projection pieces it is made of.

TEST INFRASTRUCTURE ONLY (see oracle.py header): imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.

Follows (paths relative to /root/reference/src/incompressible):
  spectral.F90:235-341     mTimes_ik1/ik2 (ip, oop)
  spectral.F90:314-363     dealias (init_periodicInZ branch), dealias_edgeField
  spectral.F90:755-865     init_periodic_inZ_procedures (3-D dealias mask with `>=`, k3 tables, normfactZ)
  spectral.F90:933-1200    initializeEverything, TwoPeriodic branch (Nyquist sign flip :1030-1032, kabs_sq = k1^2+k2^2,
                           2-D mask with strict `<` and a hard-wired 2/3 :1147-1159, fixOddball :1189-)
  spectral.F90:1413-1509   fft / ifft / take_fftz / take_ifftz / take_(i)fft1d_z2z_ip
  PadeDerOps.F90:57-88, 997-1053   Pade6stagg periodic dispatch (scheme cd06) and getmodCD06stagg
  PadePoisson.F90:76-128, 386-432, 716-750, 900-949, 1165-1244   periodic Poisson init, PeriodicProjection,
                           Periodic_getPressure(AndUpdateRHS), DivergenceCheck (with the fixDiv re-projection)
  igrid.F90:1020-1037 dealiasFields, :1105-1173 TVD_RK3, :1176-1299 SSP_RK45, :1423-1447 interp_PrimitiveVars,
  :1572-1679 addNonLinearTerm_skewSymm, :1914-1941 addViscousTerm, :1961-1990 project_and_prep,
  :2553-2683 compute_duidxj, :625-655 the initial fft / dealias / projection sequence of igrid%init.

Everything is written for ONE rank holding the global arrays: 2DECOMP transposes are pure relabelings of the
same global array, so results do not depend on the processor grid.  Arrays follow the
reference's Fortran layout: f(nx,ny,nz) is a numpy array of shape (nz, ny, nx); spectral arrays are
(nz, ny, nx/2+1) complex128.  FFT arithmetic: numpy's pocketfft standing in for FFTW 3.3.5 (unnormalised forward,
normalisation applied on the inverse as the reference does).
"""
import numpy as np

from . import oracle as O

imi = 1j


def _wavenums(n, d, flip_nyquist):
    k = O.wavenums(n, d)
    if flip_nyquist:
        k = k.copy()
        k[n // 2] = -k[n // 2]  # spectral.F90:1030-1032 (1-based n/2+1)
    return k


class Spectral:
    """spectralMod::spectral with pencil "x", dimTransform = 2 (igrid.F90:487-495)."""

    def __init__(self, nx, ny, nz, dx, dy, dz, init_periodicInZ=False, dealiasF=2.0 / 3.0, fixOddball=False):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.nxh = nx // 2 + 1
        self.dx, self.dy, self.dz = dx, dy, dz
        k1 = _wavenums(nx, dx, True)[: self.nxh].copy()
        k2 = _wavenums(ny, dy, True).copy()
        if fixOddball:  # spectral.F90:1189-1199 (TwoPeriodic branch)
            k2[ny // 2] = 0.0
            k1[nx // 2] = 0.0
        self.k1_1d, self.k2_1d = k1, k2
        self.k1 = np.broadcast_to(k1[None, None, :], (nz, ny, self.nxh))
        self.k2 = np.broadcast_to(k2[None, :, None], (nz, ny, self.nxh))
        self.kabs_sq = np.ascontiguousarray(self.k1 ** 2 + self.k2 ** 2)
        kdx, kdy = ((2.0 / 3.0) * np.pi / dx), ((2.0 / 3.0) * np.pi / dy)
        self.Gdealias = ((np.abs(self.k1) < kdx) & (np.abs(self.k2) < kdy)).astype(np.float64)
        self.normfact2d = 1.0 / (float(nx) * float(ny))
        self.init_periodicInZ = init_periodicInZ
        if init_periodicInZ:
            assert nz % 2 == 0
            k3 = _wavenums(nz, dz, True)
            f = dealiasF
            cut = (np.abs(self.k1) >= f * np.pi / dx) | (np.abs(self.k2) >= f * np.pi / dy) | \
                  (np.abs(k3)[:, None, None] >= f * np.pi / dz)
            self.Gdealias = np.where(cut, 0.0, 1.0)
            self.normfactz = 1.0 / float(nz)
            k3_1d = O.wavenums(nz, dz)  # GetWaveNums, no sign flip (spectral.F90:845)
            self.k3inZ = k3
            self.mk3sq = -(k3_1d ** 2)
            self.k3_C2Eshift = imi * k3_1d * np.exp(-imi * k3_1d * dz / 2.0)
            self.k3_E2Cshift = imi * k3_1d * np.exp(imi * k3_1d * dz / 2.0)
            self.k3_C2Cder = imi * k3_1d
            self.C2Eshift = np.exp(-imi * k3_1d * dz / 2.0)
            self.E2Cshift = np.exp(imi * k3_1d * dz / 2.0)

    # fft_3d%fft2_x2y / ifft2_y2x (utilities/fft_3d.F90:645-663, 616-643)
    def fft(self, a):
        return np.fft.fft(np.fft.rfft(a, axis=2), axis=1)

    def ifft(self, ahat, setOddball=False):
        a = np.array(ahat, dtype=np.complex128, copy=True)
        if setOddball:
            a[:, :, self.nx // 2] = 0.0
        a = np.fft.ifft(a, axis=1) * self.ny
        r = np.fft.irfft(a, n=self.nx, axis=2) * self.nx
        return r * self.normfact2d

    def mTimes_ik1(self, f):
        return (-self.k1 * f.imag) + 1j * (self.k1 * f.real)

    def mTimes_ik2(self, f):
        return (-self.k2 * f.imag) + 1j * (self.k2 * f.real)

    def dealias(self, fhat):
        if self.init_periodicInZ:
            c = np.fft.fft(fhat, axis=0)
            c = c * self.Gdealias
            c = np.fft.ifft(c, axis=0) * self.nz  # FFTW backward is unnormalised ...
            return self.normfactz * c             # ... take_ifftz scales by normfactz
        return fhat * self.Gdealias

    def dealias_edgeField(self, fhatE):
        """fhatE: z-pencil edge field with nz+1 planes; spectC's (periodic) tables (spectral.F90:343-363)."""
        nz = self.nz
        out = np.array(fhatE, dtype=np.complex128, copy=True)
        c = np.fft.fft(out[:nz], axis=0) * self.Gdealias
        c = np.fft.ifft(c, axis=0) * nz
        out[:nz] = self.normfactz * c
        out[nz] = out[0]
        return out

    # ---- z-Fourier operators on whole arrays (spectral.F90:409-437, 507-547) ----
    def _real_z(self, a, table):
        """The REAL procedures: r2c along z, modes 0 .. nz/2-1 times the table, the oddball (Nyquist) mode left as it came
        out of the transform ("Note that the oddball is ignored"), c2r, x normfactz."""
        nz = self.nz
        c = np.fft.rfft(np.asarray(a, dtype=np.float64)[:nz], axis=0)
        c[: nz // 2] = c[: nz // 2] * table[: nz // 2, None, None]
        return np.fft.irfft(c, n=nz, axis=0)        # numpy's irfft = FFTW c2r x 1/nz; both drop Im of the Nyquist coefficient

    def ddz_C2C_real_inplace(self, a):
        return self._real_z(a, self.k3_C2Cder)

    def ddz_C2C_complex_inplace(self, a):
        return self.normfactz * (np.fft.ifft(np.fft.fft(a, axis=0) * self.k3_C2Cder[:, None, None], axis=0) * self.nz)

    def shiftz_E2C(self, ahat_z):
        """in place on an array that is ALREADY Fourier-transformed in z (spectral.F90:409-422)"""
        return ahat_z * self.E2Cshift[:, None, None]

    def shiftz_C2E(self, ahat_z):
        return ahat_z * self.C2Eshift[:, None, None]

    def take_fft1d_z2z(self, a):
        return np.fft.fft(a, axis=0)

    def take_ifft1d_z2z(self, a):
        return self.normfactz * (np.fft.ifft(a, axis=0) * self.nz)


def getmodCD06stagg(k, dx):
    """PadeDerOps.F90:1034-1053."""
    alpha, beta, a, b, c = 9.0 / 62.0, 0.0, 63.0 / 62.0, 17.0 / 62.0, 0.0
    omega = k * dx
    kp = (2.0 * a * np.sin(omega / 2.0) + (2.0 / 3.0) * b * np.sin(3.0 * omega / 2.0) + (2.0 / 5.0) * c * np.sin(5.0 * omega / 2.0)) / \
         (1.0 + 2.0 * alpha * np.cos(omega) + 2.0 * beta * np.cos(2.0 * omega))
    return kp / dx


class Pade6stagg:
    """PadeDerOps::Pade6stagg, isPeriodic = .true. (BC integers are ignored on this branch).  scheme = 1: cd06 (the compact
    staggered operators); scheme = 2: fourierColl — every operator is `c2c-z forward, x table(k3), c2c-z backward, x 1/nz`
    with the tables of spectral.F90:843-856 (k3 = GetWaveNums(nz, dz), shifts e^{+-i k3 dz/2} between cells and edges), the
    complex procedures of spectral.F90:387-407, 462-482, 528-568, 596-680 and their real twins (r2c / c2r, oddball mode untouched);
    edge outputs copy plane 1 into plane nz+1."""

    def __init__(self, nz, dz, scheme=1, isPeriodic=True):
        assert scheme in (1, 2)
        self.nz, self.dz, self.scheme = nz, dz, scheme
        self.isPeriodic = isPeriodic
        if not isPeriodic:
            # PadeDerOps.F90:92-110: derOO .. derSS, keyed here by (bot, top) in {-1 odd, 0 one-sided, +1 even}^2; the Even flag of
            # a one-sided wall never reaches a row
            assert scheme == 1, "Invalid choice for numerical scheme in vertical direction (323)"
            from . import stagg_np_oracle as SN
            self.wall = {(b, t): SN.CD06StaggNP(nz, dz, isTopEven=(t == 1), isBotEven=(b == 1), isTopSided=(t == 0), isBotSided=(b == 0))
                         for b in (-1, 0, 1) for t in (-1, 0, 1)}
        if scheme == 2:
            k3 = O.wavenums(nz, dz)
            self.mk3sq = -(k3 ** 2)
            self.k3_C2Eshift = 1j * k3 * np.exp(-1j * k3 * dz / 2.0)
            self.k3_E2Cshift = 1j * k3 * np.exp(1j * k3 * dz / 2.0)
            self.C2Eshift = np.exp(-1j * k3 * dz / 2.0)
            self.E2Cshift = np.exp(1j * k3 * dz / 2.0)

    def _spect(self, f, table, edge_out):
        nz = self.nz
        if not np.iscomplexobj(f):
            # the REAL procedures (spectral.F90:365-385, 439-459, 484-505, 572-593, 639-658, 682-702): r2c, modes 0 .. nz/2-1
            # times the table, the oddball mode passes through untouched, c2r, x 1/nz
            c = np.fft.rfft(np.asarray(f, dtype=np.float64)[:nz], axis=0)
            c[: nz // 2] = c[: nz // 2] * table[: nz // 2, None, None]
            out = np.fft.irfft(c, n=nz, axis=0)
        else:
            out = np.fft.ifft(np.fft.fft(np.asarray(f)[:nz], axis=0) * table[:, None, None], axis=0)   # backward x normfactz = numpy's ifft
        if edge_out:
            out = np.concatenate([out, out[:1]], axis=0)
        return out

    def _wall(self, name, f, bot, top, edge_out, second):
        """isPeriodic = .false. (PadeDerOps.F90:185-205, 449-482, ...): unsupported (bot, top) combinations give output = 0"""
        f = np.asarray(f)
        ok = bot in (-1, 0, 1) and top in (-1, 0, 1) and not (second and (bot == 0 or top == 0))
        if not ok:
            return np.zeros((self.nz + (1 if edge_out else 0),) + f.shape[1:], dtype=f.dtype)
        return getattr(self.wall[(bot, top)], name)(f)

    def ddz_E2C(self, fE, bot=0, top=0):
        if not self.isPeriodic:
            return self._wall("ddz_E2C", fE, bot, top, False, False)
        return O.stagg("ddz_E2C", fE, self.nz, self.dz) if self.scheme == 1 else self._spect(fE, self.k3_E2Cshift, False)

    def ddz_C2E(self, fC, bot=0, top=0):
        if not self.isPeriodic:
            return self._wall("ddz_C2E", fC, bot, top, True, False)
        return O.stagg("ddz_C2E", fC, self.nz, self.dz) if self.scheme == 1 else self._spect(fC, self.k3_C2Eshift, True)

    def interpz_E2C(self, fE, bot=0, top=0):
        if not self.isPeriodic:
            return self._wall("InterpZ_E2C", fE, bot, top, False, False)
        return O.stagg("interp_E2C", fE, self.nz, self.dz) if self.scheme == 1 else self._spect(fE, self.E2Cshift, False)

    def interpz_C2E(self, fC, bot=0, top=0):
        if not self.isPeriodic:
            return self._wall("InterpZ_C2E", fC, bot, top, True, False)
        return O.stagg("interp_C2E", fC, self.nz, self.dz) if self.scheme == 1 else self._spect(fC, self.C2Eshift, True)

    def d2dz2_C2C(self, fC, bot=0, top=0):
        if not self.isPeriodic:
            return self._wall("d2dz2_C2C", fC, bot, top, False, True)
        return O.stagg("d2dz2_C2C", fC, self.nz, self.dz) if self.scheme == 1 else self._spect(fC, self.mk3sq, False)

    def d2dz2_E2E(self, fE, bot=0, top=0):
        if not self.isPeriodic:
            return self._wall("d2dz2_E2E", fE, bot, top, True, True)
        return O.stagg("d2dz2_E2E", fE, self.nz, self.dz) if self.scheme == 1 else self._spect(fE, self.mk3sq, True)

    def getModifiedWavenumbers(self, k):
        return getmodCD06stagg(k, self.dz) if self.scheme == 1 else np.array(k, dtype=float)     # PadeDerOps.F90:1003-1006


class PadePoisson:
    """PadePoissonMod::padepoisson.  PeriodicInZ = .true.: the z-periodic solver (PadePoisson.F90:76-128, 386-432, ...).
    PeriodicInZ = .false., computeStokesPressure = .false. (:130-230, 434-624): walls at z = 0 and z = Lz; the divergence is
    extended evenly and w oddly to 2 nz points, transformed in z, solved with the z-scheme's modified wavenumber carrying the
    half-cell shifts (k3modcm / k3modcp), and cut back; w is zero on both walls afterwards."""

    def __init__(self, dx, dy, dz, spC, spE, derivZ, PeriodicInZ=True, computeStokesPressure=False, Lz=None):
        self.sp, self.spE, self.derivZ = spC, spE, derivZ
        self.PeriodicInZ = PeriodicInZ
        self.computeStokesPressure = computeStokesPressure and not PeriodicInZ
        nx, ny, nz = spC.nx, spC.ny, spC.nz
        k1 = O.wavenums(nx, dx)[: spC.nxh]
        k2 = O.wavenums(ny, dy)
        if not PeriodicInZ:
            nzExt = 2 * nz
            k3 = O.wavenums(nzExt, dz)
            k3mod = derivZ.getModifiedWavenumbers(k3)
            tfm, tfp = np.exp(imi * (-dz / 2.0) * k3), np.exp(imi * (dz / 2.0) * k3)
            self.k3modcm, self.k3modcp = k3mod * tfm, k3mod * tfp
            kradsq = k1[None, None, :] ** 2 + k2[None, :, None] ** 2 + k3mod[:, None, None] ** 2
            with np.errstate(divide="ignore"):
                self.kradsq_inv = np.where(kradsq <= 1e-14, 0.0, 1.0 / kradsq)
            self.mfact = 1.0 / float(nzExt)
            if self.computeStokesPressure:       # PadePoisson.F90:232-296
                Lz = float(nz) * dz if Lz is None else Lz
                zEdge = np.linspace(0.0, Lz, nz + 1)
                zCell = 0.5 * (zEdge[:nz] + zEdge[1:])
                self.k1inZ = np.broadcast_to(k1[None, :], (ny, spC.nxh)).copy()           # GetWaveNums, no oddball flip (:248-249)
                self.k2inZ = np.broadcast_to(k2[:, None], (ny, spC.nxh)).copy()
                lam = np.sqrt(self.k1inZ ** 2 + self.k2inZ ** 2)
                temp = lam * Lz
                with np.errstate(over="ignore"):
                    den = np.where(temp < 500.0, 1.0 / (lam * np.sinh(np.minimum(lam * Lz, 700.0)) + 1.0e-13), 0.0)
                den = np.where(den < 1e-16, 0.0, den)
                den[0, 0] = 0.0                                                          # nrank == 0 owns the mean mode
                self.denFact = den
                t = lam[None] * (Lz - zCell)[:, None, None]
                self.cosh_bot = np.where(t < 32.0, np.cosh(np.minimum(t, 32.0)), 4.0e13)
                t = lam[None] * zCell[:, None, None]
                self.cosh_top = np.where(t < 32.0, np.cosh(np.minimum(t, 32.0)), 1.0e13)
                t = lam[None] * (Lz - zEdge)[:, None, None]
                self.sinh_bot = np.where(t < 32.0, -lam[None] * np.sinh(np.minimum(t, 32.0)), -4.0e13)
                t = lam[None] * zEdge[:, None, None]
                self.sinh_top = np.where(t < 32.0, lam[None] * np.sinh(np.minimum(t, 32.0)), 4.0e13)
            return
        k3mod = derivZ.getModifiedWavenumbers(O.wavenums(nz, dz))
        kradsq = k1[None, None, :] ** 2 + k2[None, :, None] ** 2 + k3mod[:, None, None] ** 2
        with np.errstate(divide="ignore"):
            self.kradsq_inv = np.where(kradsq <= 1e-14, 0.0, 1.0 / kradsq)
        self.mfact = 1.0 / float(nz)

    def _solve(self, uhat, vhat, what):
        sp = self.sp
        f2dy = sp.k1 * uhat
        f2dy = f2dy + sp.k2 * vhat
        f2dy = -f2dy.imag + 1j * f2dy.real
        w2 = what                                   # transpose_y_to_z: same global array
        f2d = self.derivZ.ddz_E2C(w2)
        f2d = f2d + f2dy
        f2d = np.fft.fft(f2d, axis=0)
        f2d = -self.kradsq_inv * f2d
        f2d = np.fft.ifft(f2d, axis=0) * sp.nz
        f2d = f2d * self.mfact
        return f2d, w2

    def ProjectStokesPressure(self, uhat, vhat, what, store=False):
        """PadePoisson.F90:320-384: the harmonic (Stokes) pressure that cancels w on the two walls, bottom first, then top with the
        already corrected top plane; returns (uhatInZ, vhatInZ, w2).  store = True is GetStokesPressure (:641-714): the same
        arithmetic, and the two pressure pieces are kept in phat_z1 / phat_z2."""
        nz = self.sp.nz
        u, v, w2 = np.array(uhat, dtype=complex), np.array(vhat, dtype=complex), np.array(what, dtype=complex)
        chat = -w2[0] * self.denFact
        phat = imi * chat[None] * self.cosh_bot
        if store:
            self.phat_z1 = phat.copy()
        u = u - self.k1inZ[None] * phat
        v = v - self.k2inZ[None] * phat
        w2[0] = 0.0
        w2[1:] = w2[1:] - chat[None] * self.sinh_bot[1:]
        chat = w2[nz] * self.denFact
        phat = imi * chat[None] * self.cosh_top
        if store:
            self.phat_z2 = phat.copy()
        u = u - self.k1inZ[None] * phat
        v = v - self.k2inZ[None] * phat
        w2[:nz] = w2[:nz] - chat[None] * self.sinh_top[:nz]
        w2[nz] = 0.0
        return u, v, w2

    def _wall_projection(self, uhat, vhat, what, want_pressure=False, store_stokes=False):
        """PadePoisson.F90:444-623 (PressureProjection), :762-896 (getPressure), :963-1160 (getPressureAndUpdateRHS): the three
        share Steps 0-7.  want_pressure: also return phat (z-pencil, before the Stokes terms are added)."""
        sp = self.sp
        nz = sp.nz
        if self.computeStokesPressure:
            uZ, vZ, w2 = self.ProjectStokesPressure(uhat, vhat, what, store=store_stokes)
            f2d = self.k1inZ[None] * uZ
            f2d = f2d + self.k2inZ[None] * vZ
            f2d = -f2d.imag + 1j * f2d.real
        else:
            f2dy = sp.k1 * uhat
            f2dy = f2dy + sp.k2 * vhat
            f2d = -f2dy.imag + 1j * f2dy.real
            w2 = what
        f2dext = np.concatenate([f2d[::-1], f2d], axis=0)                      # Step 3: even extension, 2 nz planes
        wext = np.empty_like(f2dext)
        wext[0:nz - 1] = -w2[nz - 1:0:-1]                                      # wext(kk-1) = -w2(nzG-kk+2), kk = 2..nzG
        wext[nz - 1:2 * nz] = w2                                               # wext(nzG+kk-1) = w2(kk), kk = 1..nzG+1
        f2dext = np.fft.fft(f2dext, axis=0)
        wext = np.fft.fft(wext, axis=0)
        f2dext = f2dext + imi * self.k3modcm[:, None, None] * wext             # Step 5
        f2dext = -f2dext * self.kradsq_inv
        wext = wext - imi * self.k3modcp[:, None, None] * f2dext               # Step 6
        f2dext = self.mfact * (np.fft.ifft(f2dext, axis=0) * (2 * nz))         # Step 7
        wext = (np.fft.ifft(wext, axis=0) * (2 * nz)) * self.mfact
        f2d = f2dext[nz:]
        w2 = wext[nz - 1:].copy()
        w2[0] = 0.0
        w2[nz] = 0.0
        if self.computeStokesPressure:                                         # :597-609
            g = -f2d.imag + 1j * f2d.real
            out = (uZ - g * self.k1inZ[None], vZ - g * self.k2inZ[None], w2)
        else:
            out = (uhat - imi * sp.k1 * f2d, vhat - imi * sp.k2 * f2d, w2)
        return out + (f2d,) if want_pressure else out

    def PressureProjection(self, uhat, vhat, what):
        if not self.PeriodicInZ:
            return self._wall_projection(uhat, vhat, what)
        sp = self.sp
        f2d, w2 = self._solve(uhat, vhat, what)
        dwdz = self.derivZ.ddz_C2E(f2d)
        what_new = w2 - dwdz
        uhat_new = uhat - imi * sp.k1 * f2d
        vhat_new = vhat - imi * sp.k2 * f2d
        return uhat_new, vhat_new, what_new

    def getPressure(self, uhat, vhat, what):
        if not self.PeriodicInZ:                                               # :762-896
            f2d = self._wall_projection(uhat, vhat, what, want_pressure=True, store_stokes=True)[3]
            if self.computeStokesPressure:
                f2d = f2d + self.phat_z1 + self.phat_z2
            return self.sp.ifft(f2d)
        f2d, _ = self._solve(uhat, vhat, what)
        return self.sp.ifft(f2d)

    def getPressureAndUpdateRHS(self, uhat, vhat, what):
        sp = self.sp
        if not self.PeriodicInZ:
            # :963-1160.  With computeStokesPressure this routine calls ProjectStokesPressure, which does NOT refresh phat_z1 /
            # phat_z2, and then adds them (:1146-1156): the pressure it returns carries the Stokes pieces of the LAST getPressure
            # call (zero here before any; unallocated-array garbage in the reference).  The updated right-hand sides do not depend
            # on that.
            u, v, w, f2d = self._wall_projection(uhat, vhat, what, want_pressure=True, store_stokes=False)
            if self.computeStokesPressure:
                z1 = getattr(self, "phat_z1", None)
                if z1 is not None:
                    f2d = f2d + self.phat_z1 + self.phat_z2
            return u, v, w, sp.ifft(f2d)
        f2d, w2 = self._solve(uhat, vhat, what)
        dwdz = self.derivZ.ddz_C2E(f2d)
        return uhat - imi * sp.k1 * f2d, vhat - imi * sp.k2 * f2d, w2 - dwdz, sp.ifft(f2d)

    def divergence(self, uhat, vhat, what):
        sp = self.sp
        f2dy = self.derivZ.ddz_E2C(what) if self.PeriodicInZ else self.derivZ.ddz_E2C(what, -1, -1)    # :1188
        f2dy = f2dy + imi * sp.k1 * uhat + imi * sp.k2 * vhat
        return sp.ifft(f2dy)

    def DivergenceCheck(self, uhat, vhat, what, fixDiv=False):
        """Returns (uhat, vhat, what, divergence); re-projects like PadePoisson.F90:1203-1241 when fixDiv."""
        div = self.divergence(uhat, vhat, what)
        if fixDiv and div.max() > 1e-13:
            uhat, vhat, what = self.PressureProjection(uhat, vhat, what)
            div = self.divergence(uhat, vhat, what)
            if div.max() > 1e-10:
                uhat, vhat, what = self.PressureProjection(uhat, vhat, what)
        return uhat, vhat, what, div


def splitmix64_uniform(seed, count):
    """`count` doubles in [0, 1) from SplitMix64 seeded with `seed`.  Fortran's random_seed(put = seed) / random_number, which
    utilities/random.F90:154-174 uses, is compiler-specific (gfortran: xoshiro256**, ifort: L'Ecuyer) — no drop-in can
    reproduce its stream, so the GPU library and this oracle share this documented generator instead, and both accept the
    reference's own draw through set_wavenumbers."""
    mask = (1 << 64) - 1
    state = int(seed) & mask
    out = np.empty(count)
    for i in range(count):
        state = (state + 0x9E3779B97F4A7C15) & mask
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        z ^= z >> 31
        out[i] = (z >> 11) * (1.0 / 9007199254740992.0)
    return out


class HITForcing:
    """forcingmod::HIT_shell_forcing (incompressible/forcingIsotropic.F90:45-314): every time step Nwaves integer wavenumber
    triplets are drawn on the shell kmin <= |k| <= kmax; each forced mode receives EpsAmplitude / Nwaves of energy injection
    rate, f_hat(mode) += normfact * Eps / (|u_hat|^2 + |v_hat|^2 + |w_hat|^2 + 1e-14) / Nwaves * conjg(u_hat(mode)), in the
    fully (x, y, z)-transformed space; w lives on edges and is shifted to cells and back with the spectral type's tables."""

    def __init__(self, spectC, kmin=2.0, kmax=10.0, Nwaves=20, EpsAmplitude=0.1, tidStart=0, RandSeedToAdd=0):
        self.spectC = spectC
        self.kmin, self.kmax, self.Nwaves, self.Eps = kmin, kmax, int(Nwaves), EpsAmplitude
        self.seed0 = tidStart + RandSeedToAdd                        # :86
        self.update_seeds()
        self.normfact = (float(spectC.nx) * float(spectC.ny) * float(spectC.nz)) ** 2    # :89
        self.wave_x = self.wave_y = self.wave_z = None

    def update_seeds(self):                                          # :122-128
        self.seed0 = abs(self.seed0 + 2223345)
        self.seed1 = abs(self.seed0 + 1423246)
        self.seed2 = abs(self.seed0 + 8723446)
        self.seed3 = abs(self.seed0 + 3423444)

    @staticmethod
    def _uniform(n, left, right, seed):                              # unrand1R, random.F90:154-174
        a = splitmix64_uniform(seed, n)
        a = (right - left) * a
        return a + left

    def wavenumbers_from_samples(self, kabs, zeta, theta):           # :137-147
        t = kabs * np.sqrt(1 - zeta ** 2) * np.cos(theta)
        self.wave_x = np.ceil(np.abs(t)).astype(int)
        t = kabs * np.sqrt(1 - zeta ** 2) * np.sin(theta)
        self.wave_y = np.ceil(np.abs(t)).astype(int)
        t = kabs * zeta
        self.wave_z = np.ceil(np.abs(t)).astype(int)

    def pick_random_wavenumbers(self):                               # :131-149
        n = self.Nwaves
        self.wavenumbers_from_samples(self._uniform(n, self.kmin, self.kmax, self.seed1), self._uniform(n, -1.0, 1.0, self.seed2),
                                      self._uniform(n, 0.0, 2.0 * np.pi, self.seed3))

    def set_wavenumbers(self, wx, wy, wz):
        """inject a draw (the reference RNG's, say): the next new time step keeps it instead of drawing; its seeds still advance"""
        self.wave_x, self.wave_y, self.wave_z = (np.asarray(a, dtype=int) for a in (wx, wy, wz))
        self._injected = True

    def getRHS_HITforcing(self, urhs, vrhs, wrhs, uhat_xy, vhat_xy, what_xy, newTimestep):     # :254-311
        sp = self.spectC
        nz = sp.nz
        if newTimestep:
            if not getattr(self, "_injected", False):
                self.pick_random_wavenumbers()
            self._injected = False
            self.update_seeds()
        uh = sp.take_fft1d_z2z(uhat_xy)
        vh = sp.take_fft1d_z2z(vhat_xy)
        wh = sp.shiftz_E2C(sp.take_fft1d_z2z(what_xy[:nz]))
        fx, fy, fz = (np.zeros_like(uh) for _ in range(3))
        ny = sp.ny
        for kx, ky, kz in zip(self.wave_x, self.wave_y, self.wave_z):       # compute_forcing / embed_forcing_mode :203-251
            if not (0 <= kx < sp.nxh and 0 <= ky < ny):
                continue                                                      # "not on this processor" for every rank
            den = abs(uh[kz, ky, kx]) ** 2 + abs(vh[kz, ky, kx]) ** 2 + abs(wh[kz, ky, kx]) ** 2 + 1.0e-14
            fac = self.normfact * self.Eps / den / float(self.Nwaves)
            fx[kz, ky, kx] += fac * np.conj(uh[kz, ky, kx])
            fy[kz, ky, kx] += fac * np.conj(vh[kz, ky, kx])
            fz[kz, ky, kx] += fac * np.conj(wh[kz, ky, kx])
        urhs = urhs + sp.take_ifft1d_z2z(fx)
        vrhs = vrhs + sp.take_ifft1d_z2z(fy)
        fzE = sp.take_ifft1d_z2z(sp.shiftz_C2E(fz))
        wrhs = wrhs + np.concatenate([fzE, fzE[:1]], axis=0)
        return urhs, vrhs, wrhs


class SGS:
    """sgs_igrid (incompressible/sgsmod_igrid.F90:156-251 + sgs_models/{smagorinsky, sigma, AMD, eddyViscosity}.F90) for the
    periodic box: eddy-viscosity models 0 Smagorinsky, 1 sigma, 2 AMD with a global constant — no wall damping, no dynamic
    procedure, no wall model (the HIT deck's settings: SGSModelID = 2, DynamicProcedureType = 0, WallModelType = 0).
    duidxj arrays: lists of the nine gradients dudx, dudy, dudz, dvdx, ... in the reference's order, on cells (C) and edges (E)."""

    def __init__(self, spectC, spectE, ops, SGSModelID=2, Csgs=0.17, explicitCalcEdgeEddyViscosity=False):
        self.spectC, self.spectE, self.ops = spectC, spectE, ops
        self.mid, self.explicitE = SGSModelID, explicitCalcEdgeEddyViscosity
        dx, dy, dz = spectC.dx, spectC.dy, spectC.dz
        deltaLES = (1.5 * dx * 1.5 * dy * 1.5 * dz) ** (1.0 / 3.0)            # isPeriodic branch (smagorinsky.F90:17, sigma.F90:15)
        if SGSModelID in (0, 1):
            self.cmodel_global = (Csgs * deltaLES) ** 2
        elif SGSModelID == 2:                                               # AMD.F90:1-14
            poincare = 1.0 / np.sqrt(10.0) if ops.scheme == 1 else 1.0 / np.sqrt(12.0)     # PadeDerOps.F90:1018-1031
            self.camd_x = Csgs * dx * np.sqrt(1.0 / 12.0)
            self.camd_y = Csgs * dy * np.sqrt(1.0 / 12.0)
            self.camd_z = Csgs * dz * poincare
            self.cmodel_global = 1.0
        else:
            raise ValueError("Incorrect choice for SGS model ID. (213)")

    @staticmethod
    def get_Sij(d):                                                          # eddyViscosity.F90:1-24: S11 S12 S13 S22 S23 S33
        return [d[0], 0.5 * (d[1] + d[3]), 0.5 * (d[2] + d[6]), d[4], 0.5 * (d[5] + d[7]), d[8]]

    def kernel(self, d, S):
        if self.mid == 0:                                                    # smagorinsky.F90:44-66
            t = S[0] * S[0]
            t = t + 2.0 * (S[1] * S[1])
            t = t + 2.0 * (S[2] * S[2])
            t = t + (S[3] * S[3])
            t = t + 2.0 * (S[4] * S[4])
            t = t + (S[5] * S[5])
            return np.sqrt(2.0 * t)
        if self.mid == 1:                                                    # sigma.F90:31-98
            G11 = d[0] * d[0] + d[3] * d[3] + d[6] * d[6]
            G12 = d[0] * d[1] + d[3] * d[4] + d[6] * d[7]
            G13 = d[0] * d[2] + d[3] * d[5] + d[6] * d[8]
            G22 = d[1] * d[1] + d[4] * d[4] + d[7] * d[7]
            G23 = d[1] * d[2] + d[4] * d[5] + d[7] * d[8]
            G33 = d[2] * d[2] + d[5] * d[5] + d[8] * d[8]
            I1 = G11 + G22 + G33
            I1sq = I1 * I1
            I1cu = I1sq * I1
            I2 = -G11 * G11 - G22 * G22 - G33 * G33
            I2 = I2 - 2.0 * G12 * G12 - 2.0 * G13 * G13
            I2 = I2 - 2.0 * G23 * G23
            I2 = I2 + I1sq
            I2 = 0.5 * I2
            I3 = G11 * (G22 * G33 - G23 * G23)
            I3 = I3 + G12 * (G13 * G23 - G12 * G33)
            I3 = I3 + G13 * (G12 * G23 - G22 * G13)
            alpha1 = np.maximum(I1sq / 9.0 - I2 / 3.0, 0.0)
            alpha2 = I1cu / 27.0 - I1 * I2 / 6.0 + I3 / 2.0
            a1s = np.sqrt(alpha1)
            t = alpha2 / (alpha1 * a1s + 1.0e-13)
            t = np.arccos(np.maximum(np.minimum(t, 1.0), -1.0))
            alpha3 = (1.0 / 3.0) * t
            s1sq = np.maximum(I1 / 3.0 + 2.0 * a1s * np.cos(alpha3), 0.0)
            s1 = np.sqrt(s1sq)
            s2 = np.sqrt(np.maximum((-2.0) * a1s * np.cos(np.pi / 3.0 + alpha3) + I1 / 3.0, 0.0))
            s3 = np.sqrt(np.maximum((-2.0) * a1s * np.cos(np.pi / 3.0 - alpha3) + I1 / 3.0, 0.0))
            return s3 * (s1 - s2) * (s2 - s3) / (s1sq + 1.0e-15)
        cx, cy, cz = self.camd_x, self.camd_y, self.camd_z                  # AMD.F90:24-72
        row = lambda a, b: (d[a] * cx) * (d[b] * cx) + (d[a + 1] * cy) * (d[b + 1] * cy) + (d[a + 2] * cz) * (d[b + 2] * cz)
        num = row(0, 0) * S[0]
        num = num + row(3, 3) * S[3]
        num = num + row(6, 6) * S[5]
        num = num + 2.0 * row(0, 3) * S[1]
        num = num + 2.0 * row(0, 6) * S[2]
        num = num + 2.0 * row(3, 6) * S[4]
        den = sum(x * x for x in d)
        return np.maximum(-num / (den + 1.0e-32), 0.0)

    def getTauSGS(self, duidxjC, duidxjE):                                   # sgsmod_igrid.F90:156-203
        C, E = self.spectC, self.spectE
        SC, SE = self.get_Sij(duidxjC), self.get_Sij(duidxjE)
        nuC = self.cmodel_global * self.kernel(duidxjC, SC)
        if self.explicitE:
            nuE = self.cmodel_global * self.kernel(duidxjE, SE)
        else:                                                                # interpolate_eddy_viscosity(.true.): eddyViscosity.F90:97-113
            nuE = self.ops.interpz_C2E(nuC)
            nuE = np.where(nuE < 0.0, 0.0, nuE)
        self.nu_sgs_C, self.nu_sgs_E = nuC, nuE
        return (-2.0 * nuC * SC[0], -2.0 * nuC * SC[1], -2.0 * nuE * SE[2], -2.0 * nuC * SC[3], -2.0 * nuE * SE[4], -2.0 * nuC * SC[5])

    def getRHS_SGS(self, urhs, vrhs, wrhs, duidxjC, duidxjE):                # :206-268
        C, E, ops = self.spectC, self.spectE, self.ops
        t11, t12, t13, t22, t23, t33 = self.getTauSGS(duidxjC, duidxjE)
        urhs = urhs - C.mTimes_ik1(C.fft(t11))
        vrhs = vrhs - C.mTimes_ik2(C.fft(t22))
        wrhs = wrhs - ops.ddz_C2E(C.fft(t33))
        c = C.fft(t12)
        vrhs = vrhs - C.mTimes_ik1(c)
        urhs = urhs - C.mTimes_ik2(c)
        c = E.fft(t13)
        urhs = urhs - ops.ddz_E2C(c)
        wrhs = wrhs - E.mTimes_ik1(c)
        c = E.fft(t23)
        vrhs = vrhs - ops.ddz_E2C(c)
        wrhs = wrhs - E.mTimes_ik2(c)
        return urhs, vrhs, wrhs


class IGrid:
    """igrid, periodic in x, y, z; NumericalSchemeVert = 1 (CD06) or 2 (Fourier collocation in z), AdvectionTerm = 1 (skew-symmetric) or 0 (rotational), no SGS / forcing /
    Coriolis / stratification; viscous unless isInviscid.  u, v: (nz, ny, nx); w: (nz+1, ny, nx) with plane nz == plane 0."""

    def __init__(self, nx, ny, nz, Lx, Ly, Lz, Re, u, v, w, isInviscid=False, dealiasFact=2.0 / 3.0, t_DivergenceCheck=10,
                 TimeSteppingScheme=1, use_d2dz2_C2C=True, AdvectionTerm=1, NumericalSchemeVert=1, HITForcing_=None, SGS_=None,
                 PeriodicInZ=True, topWall=2, botWall=2, ComputeStokesPressure=True):
        """HITForcing_: None, or the &HIT_Forcing namelist as a dict (kmin, kmax, Nwaves, EpsAmplitude, RandSeedToAdd) for
        useHITForcing = .true. (igrid.F90:940-944, 1908-1910).  SGS_: None, or the &SGS_MODEL namelist entries in scope as a dict
        (SGSModelID, Csgs, explicitCalcEdgeEddyViscosity) for useSGS = .true. (:1866-1871).
        PeriodicInZ = .false. (&BCs namelist): walls at z = 0 and z = Lz, topWall / botWall = 1 no-slip, 2 slip (3, wall model, is
        out of scope); the stencil codes of get_boundary_conditions_stencil (:5148-5204) then reach every z-operator, the
        staggered operators use their wall closures, the Poisson solver its even / odd extension, and w is dealiased in 2-D."""
        assert AdvectionTerm in (0, 1)      # 0 rotational (igrid.F90:1527-1555), 1 skew-symmetric (:1572-1679)
        assert NumericalSchemeVert in (1, 2)  # 1 cd06, 2 fourierColl (PadeDerOps.F90:16-18)
        self.AdvectionTerm = AdvectionTerm
        self.nx, self.ny, self.nz = nx, ny, nz
        self.dx, self.dy, self.dz = Lx / nx, Ly / ny, Lz / nz
        self.Re, self.isInviscid = Re, isInviscid
        self.t_DivergenceCheck, self.scheme = t_DivergenceCheck, TimeSteppingScheme
        self.use_d2dz2_C2C = use_d2dz2_C2C
        self.PeriodicInZ = PeriodicInZ
        if not PeriodicInZ:
            assert NumericalSchemeVert == 1, "If you use Fourier Collocation in Z, the problem must be periodic in Z. (123)"
            assert HITForcing_ is None and SGS_ is None
            self.use_d2dz2_C2C = True      # uBC, vBC are +-1 for slip / no-slip walls (:2653, 2671)
        self.bc = self.get_boundary_conditions_stencil(topWall, botWall)
        self.spectC = Spectral(nx, ny, nz, self.dx, self.dy, self.dz, PeriodicInZ, dealiasFact, False)
        self.spectE = Spectral(nx, ny, nz + 1, self.dx, self.dy, self.dz, False, dealiasFact, False)
        self.ops = Pade6stagg(nz, self.dz, NumericalSchemeVert, isPeriodic=PeriodicInZ)
        self.poiss = PadePoisson(self.dx, self.dy, self.dz, self.spectC, self.spectE, self.ops, PeriodicInZ=PeriodicInZ,
                                 computeStokesPressure=ComputeStokesPressure, Lz=Lz)     # igrid.F90:585-586 (&NUMERICS, default .true.)
        self.step, self.tsim = 0, 0.0
        self.newTimeStep = True
        self.hitforce = HITForcing(self.spectC, tidStart=self.step, **HITForcing_) if HITForcing_ is not None else None
        self.sgsmodel = SGS(self.spectC, self.spectE, self.ops, **SGS_) if SGS_ is not None else None
        # igrid.F90:625-655
        self.uhat = self.spectC.fft(u)
        self.vhat = self.spectC.fft(v)
        self.what = self.spectE.fft(w)
        self.dealiasFields()
        _, _, _, self.divergence = self.poiss.DivergenceCheck(self.uhat, self.vhat, self.what)
        self.uhat, self.vhat, self.what = self.poiss.PressureProjection(self.uhat, self.vhat, self.what)
        self._to_physical()
        self.interp_PrimitiveVars()
        self.compute_duidxj()

    @staticmethod
    def get_boundary_conditions_stencil(topWall, botWall):
        """igrid.F90:5148-5204: (bottom, top) stencil codes per quantity; -1 odd, +1 even, 0 one-sided"""
        bc = {"w": [-1, -1], "WdWdz": [-1, -1], "WW": [1, 1], "dWdz": [0, 0]}
        for side, wall in ((0, botWall), (1, topWall)):
            if wall == 1:      # no-slip: w = 0 and dwdz = 0, so w is extended evenly
                vals = {"u": -1, "v": -1, "dUdz": 0, "dVdz": 0, "WdUdz": 0, "WdVdz": 0, "UW": 1, "VW": 1, "w": 1, "WdWdz": -1, "WW": 1, "dWdz": -1}
            elif wall == 2:    # slip
                vals = {"u": 1, "v": 1, "dUdz": -1, "dVdz": -1, "WdUdz": 1, "WdVdz": 1, "UW": -1, "VW": -1}
            else:
                raise ValueError("Invalid choice for wall BCs (423 / 13); the wall model (3) is out of scope")
            for k, v in vals.items():
                bc.setdefault(k, [0, 0])[side] = v
        return {k: tuple(v) for k, v in bc.items()}

    # ---- igrid.F90:1020-1037
    def dealiasFields(self):
        self.uhat = self.spectC.dealias(self.uhat)
        self.vhat = self.spectC.dealias(self.vhat)
        self.what = self.spectC.dealias_edgeField(self.what) if self.PeriodicInZ else self.spectE.dealias(self.what)

    def _to_physical(self):
        self.u = self.spectC.ifft(self.uhat)
        self.v = self.spectC.ifft(self.vhat)
        self.w = self.spectE.ifft(self.what)

    # ---- igrid.F90:1423-1447
    def interp_PrimitiveVars(self):
        self.whatC = self.ops.interpz_E2C(self.what, *self.bc["w"])
        self.wC = self.spectC.ifft(self.whatC)
        self.uEhat = self.ops.interpz_C2E(self.uhat, *self.bc["u"])
        self.uE = self.spectE.ifft(self.uEhat)
        self.vEhat = self.ops.interpz_C2E(self.vhat, *self.bc["v"])
        self.vE = self.spectE.ifft(self.vEhat)

    # ---- igrid.F90:2553-2683
    def compute_duidxj(self):
        C, E, ops = self.spectC, self.spectE, self.ops
        d = {}
        d["dudx"] = C.ifft(C.mTimes_ik1(self.uhat)); d["dudxE"] = E.ifft(E.mTimes_ik1(self.uEhat))
        d["dudy"] = C.ifft(C.mTimes_ik2(self.uhat)); d["dudyE"] = E.ifft(E.mTimes_ik2(self.uEhat))
        d["dvdx"] = C.ifft(C.mTimes_ik1(self.vhat)); d["dvdxE"] = E.ifft(E.mTimes_ik1(self.vEhat))
        d["dvdy"] = C.ifft(C.mTimes_ik2(self.vhat)); d["dvdyE"] = E.ifft(E.mTimes_ik2(self.vEhat))
        d["dwdxC"] = C.ifft(C.mTimes_ik1(self.whatC)); d["dwdx"] = E.ifft(E.mTimes_ik1(self.what))
        d["dwdyC"] = C.ifft(C.mTimes_ik2(self.whatC)); d["dwdy"] = E.ifft(E.mTimes_ik2(self.what))
        dwdzH = ops.ddz_E2C(self.what, *self.bc["w"])
        d["dwdz"] = C.ifft(dwdzH)
        d["dwdzE"] = E.ifft(ops.interpz_C2E(dwdzH, *self.bc["dWdz"]))
        if not self.isInviscid:
            self.d2wdz2hatE = ops.d2dz2_E2E(self.what, *self.bc["w"])
        for nm, fhat in (("u", self.uhat), ("v", self.vhat)):
            bcf, bcd = self.bc[nm], self.bc["d%sdz" % nm.upper()]
            dEH = ops.ddz_C2E(fhat, *bcf)
            d["d%sdz" % nm] = E.ifft(dEH)
            if not self.isInviscid:
                d2 = ops.d2dz2_C2C(fhat, *bcf) if self.use_d2dz2_C2C else ops.ddz_E2C(ops.ddz_C2E(fhat, *bcf), *bcd)
                setattr(self, "d2%sdz2hatC" % nm, d2)
            d["d%sdzC" % nm] = C.ifft(ops.interpz_E2C(dEH, *bcd))
        self.duidxj = d

    # ---- igrid.F90:1572-1679
    def addNonLinearTerm_skewSymm(self):
        C, E, ops, d = self.spectC, self.spectE, self.ops, self.duidxj
        u, v, w, wC, uE, vE = self.u, self.v, self.w, self.wC, self.uE, self.vE
        T1C = d["dudx"] * u; T2C = d["dudy"] * v; T1C = T1C + T2C
        T1E = d["dudz"] * w
        fT1C = C.fft(T1C); fT1E = E.fft(T1E)
        u_rhs = ops.interpz_E2C(fT1E, *self.bc["WdUdz"]) + fT1C
        T1C = d["dvdx"] * u; T2C = d["dvdy"] * v; T1C = T1C + T2C
        T1E = d["dvdz"] * w
        fT1C = C.fft(T1C); fT1E = E.fft(T1E)
        v_rhs = ops.interpz_E2C(fT1E, *self.bc["WdVdz"]) + fT1C
        T1E = d["dwdx"] * uE; T2E = d["dwdy"] * vE; T2E = T1E + T2E
        fT2E = E.fft(T2E)
        T1C = d["dwdz"] * wC
        fT1C = C.fft(T1C)
        w_rhs = ops.interpz_C2E(fT1C, *self.bc["WdWdz"]) + fT2E
        fT1C = C.mTimes_ik1(C.fft(u * u)); u_rhs = u_rhs + fT1C
        fT1C = C.mTimes_ik2(C.fft(v * v)); v_rhs = v_rhs + fT1C
        fT1C = C.fft(wC * wC)
        w_rhs = w_rhs + ops.ddz_C2E(fT1C, *self.bc["WW"])
        fT1C = C.fft(u * v)
        u_rhs = u_rhs + C.mTimes_ik2(fT1C)
        v_rhs = v_rhs + C.mTimes_ik1(fT1C)
        fT1E = E.fft(uE * w)
        u_rhs = u_rhs + ops.ddz_E2C(fT1E, *self.bc["UW"])
        w_rhs = w_rhs + E.mTimes_ik1(fT1E)
        fT1E = E.fft(vE * w)
        v_rhs = v_rhs + ops.ddz_E2C(fT1E, *self.bc["VW"])
        w_rhs = w_rhs + E.mTimes_ik2(fT1E)
        return -0.5 * u_rhs, -0.5 * v_rhs, -0.5 * w_rhs

    # ---- igrid.F90:1527-1555: u x omega, the z-components multiplied on the edge grid and interpolated back
    def addNonLinearTerm_Rot(self):
        C, E, ops, d = self.spectC, self.spectE, self.ops, self.duidxj
        T1C = d["dvdx"] - d["dudy"]; T1C = T1C * self.v
        fT1C = C.fft(T1C)
        T2E = d["dwdx"] - d["dudz"]; T2E = T2E * self.w
        fT2E = E.fft(T2E)
        u_rhs = ops.interpz_E2C(fT2E, 0, 0) + fT1C
        T1C = d["dudy"] - d["dvdx"]; T1C = T1C * self.u
        fT1C = C.fft(T1C)
        T2E = d["dwdy"] - d["dvdz"]; T2E = T2E * self.w
        fT2E = E.fft(T2E)
        v_rhs = ops.interpz_E2C(fT2E, 0, 0) + fT1C
        T1E = d["dudz"] - d["dwdx"]; T1E = T1E * self.uE
        T2E = d["dvdz"] - d["dwdy"]; T2E = T2E * self.vE
        T1E = T1E + T2E
        w_rhs = E.fft(T1E)
        return u_rhs, v_rhs, w_rhs

    # ---- igrid.F90:1793-1912 (branches in scope) + 1914-1941
    def populate_rhs(self):
        u_rhs, v_rhs, w_rhs = self.addNonLinearTerm_skewSymm() if self.AdvectionTerm == 1 else self.addNonLinearTerm_Rot()
        if not self.isInviscid:
            oneByRe = 1.0 / self.Re
            u_rhs = u_rhs + oneByRe * (-self.spectC.kabs_sq * self.uhat + self.d2udz2hatC)
            v_rhs = v_rhs + oneByRe * (-self.spectC.kabs_sq * self.vhat + self.d2vdz2hatC)
            w_rhs = w_rhs + oneByRe * (-self.spectE.kabs_sq * self.what + self.d2wdz2hatE)
        if self.sgsmodel is not None:       # Step 6 (:1866-1871)
            d = self.duidxj
            dC = [d["dudx"], d["dudy"], d["dudzC"], d["dvdx"], d["dvdy"], d["dvdzC"], d["dwdxC"], d["dwdyC"], d["dwdz"]]
            dE = [d["dudxE"], d["dudyE"], d["dudz"], d["dvdxE"], d["dvdyE"], d["dvdz"], d["dwdx"], d["dwdy"], d["dwdzE"]]
            u_rhs, v_rhs, w_rhs = self.sgsmodel.getRHS_SGS(u_rhs, v_rhs, w_rhs, dC, dE)
        if self.hitforce is not None:       # Step 8 (:1907-1910)
            u_rhs, v_rhs, w_rhs = self.hitforce.getRHS_HITforcing(u_rhs, v_rhs, w_rhs, self.uhat, self.vhat, self.what, self.newTimeStep)
        self.newTimeStep = False            # :1128, 1203: cleared after the first stage's right-hand side
        return u_rhs, v_rhs, w_rhs

    # ---- igrid.F90:1961-1990
    def project_and_prep(self, AlreadyProjected=False):
        self.dealiasFields()
        if not AlreadyProjected:
            self.uhat, self.vhat, self.what = self.poiss.PressureProjection(self.uhat, self.vhat, self.what)
            if self.step % self.t_DivergenceCheck == 0:
                self.uhat, self.vhat, self.what, self.divergence = self.poiss.DivergenceCheck(self.uhat, self.vhat, self.what, True)
        self._to_physical()
        self.interp_PrimitiveVars()
        self.compute_duidxj()

    def timeAdvance(self, dt):
        self.dt = dt
        (self.TVD_RK3 if self.scheme == 1 else self.SSP_RK45)(dt)
        self.step += 1       # wrapup_timestep
        self.tsim += dt
        self.newTimeStep = True   # :2075

    # ---- igrid.F90:1105-1173
    def TVD_RK3(self, dt):
        u0, v0, w0 = self.uhat, self.vhat, self.what
        ur, vr, wr = self.populate_rhs()
        u1, v1, w1 = u0 + dt * ur, v0 + dt * vr, w0 + dt * wr
        self.uhat, self.vhat, self.what = u1, v1, w1
        self.project_and_prep()
        u1, v1, w1 = self.uhat, self.vhat, self.what   # uhat1 aliases the projected stage array
        ur, vr, wr = self.populate_rhs()
        u1 = (3.0 / 4.0) * u0 + (1.0 / 4.0) * u1 + (1.0 / 4.0) * dt * ur
        v1 = (3.0 / 4.0) * v0 + (1.0 / 4.0) * v1 + (1.0 / 4.0) * dt * vr
        w1 = (3.0 / 4.0) * w0 + (1.0 / 4.0) * w1 + (1.0 / 4.0) * dt * wr
        self.uhat, self.vhat, self.what = u1, v1, w1
        self.project_and_prep()
        u1, v1, w1 = self.uhat, self.vhat, self.what
        ur, vr, wr = self.populate_rhs()
        self.uhat = (1.0 / 3.0) * u0 + (2.0 / 3.0) * u1 + (2.0 / 3.0) * dt * ur
        self.vhat = (1.0 / 3.0) * v0 + (2.0 / 3.0) * v1 + (2.0 / 3.0) * dt * vr
        self.what = (1.0 / 3.0) * w0 + (2.0 / 3.0) * w1 + (2.0 / 3.0) * dt * wr
        self.project_and_prep()

    # ---- igrid.F90:1176-1299
    def SSP_RK45(self, dt):
        b01, b12, b23, b34 = 0.39175222657189, 0.368410593050371, 0.25189177427169, 0.54497475022852
        b35, b45 = 0.06369246866629, 0.22600748323690
        a20, a21 = 0.444370493651235, 0.555629506348765
        a30, a32 = 0.620101851488403, 0.379898148511597
        a40, a43 = 0.17807995439313, 0.821920045606868
        a52, a53, a54 = 0.517231671970585, 0.096059710526147, 0.386708617503269
        S0 = (self.uhat, self.vhat, self.what)
        R = self.populate_rhs()
        self.uhat, self.vhat, self.what = (s + b01 * dt * r for s, r in zip(S0, R))
        self.project_and_prep()
        S1 = (self.uhat, self.vhat, self.what)
        R = self.populate_rhs()
        self.uhat, self.vhat, self.what = (a20 * s0 + a21 * s1 + b12 * dt * r for s0, s1, r in zip(S0, S1, R))
        self.project_and_prep()
        S2 = (self.uhat, self.vhat, self.what)
        R = self.populate_rhs()
        self.uhat, self.vhat, self.what = (a30 * s0 + a32 * s2 + b23 * dt * r for s0, s2, r in zip(S0, S2, R))
        self.project_and_prep()
        S3 = (self.uhat, self.vhat, self.what)
        R3 = self.populate_rhs()
        # stage 4 overwrites the base arrays (uhat4 => uhat)
        self.uhat, self.vhat, self.what = (a40 * s0 + a43 * s3 + b34 * dt * r for s0, s3, r in zip(S0, S3, R3))
        self.project_and_prep()
        S4 = (self.uhat, self.vhat, self.what)
        R4 = self.populate_rhs()
        self.uhat, self.vhat, self.what = (a52 * s2 + a53 * s3 + b35 * dt * r3 + a54 * s4 + b45 * dt * r4
                                           for s2, s3, r3, s4, r4 in zip(S2, S3, R3, S4, R4))
        self.project_and_prep()
