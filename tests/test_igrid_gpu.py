"""GPU parity: spectral ops, periodic pressure projection and the igrid RK substep (through the C ABI) against the
CPU oracle on the same seeded inputs.  Bar: 1e-12 relative to max|ref| (north_star)."""
import numpy as np
import pytest

from conftest import broadband

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(got, ref):
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)


def _cplx(shape, seed):
    return broadband(shape, seed) + 1j * broadband(shape, seed + 100)


@pytest.fixture(scope="module")
def IG(oracle):
    from oracle import igrid_oracle
    return igrid_oracle


SHAPES = [(32, 24, 16), (16, 16, 32)]


@pytest.mark.parametrize("shape", SHAPES)
def test_spectral_ops_match_oracle(pdo, IG, shape):
    nx, ny, nz = shape
    d = [2 * np.pi / n for n in shape]
    sp = pdo.spectral()
    sp.init("x", nx, ny, nz, *d, "four", "2/3rd", 2, fixOddball=False, init_periodicInZ=True)
    ref = IG.Spectral(nx, ny, nz, *d, True, 2.0 / 3.0, False)
    t = sp.tables()
    assert np.array_equal(t["k1"], ref.k1_1d) and np.array_equal(t["k2"], ref.k2_1d)
    G = t["gz"][:, None, None] * t["gy"][None, :, None] * t["gx"][None, None, :]
    assert np.array_equal(G, ref.Gdealias)
    f = broadband((nz, ny, nx), seed=1)
    fh = sp.fft(_dev(f))
    assert _rel(fh.cpu().numpy(), ref.fft(f)) < TOL
    assert _rel(sp.ifft(fh).cpu().numpy(), f) < TOL
    fh_np = ref.fft(f)
    assert _rel(sp.mTimes_ik1_oop(fh).cpu().numpy(), ref.mTimes_ik1(fh_np)) < TOL
    assert _rel(sp.mTimes_ik2_oop(fh).cpu().numpy(), ref.mTimes_ik2(fh_np)) < TOL
    g = fh.clone()
    sp.mTimes_ik1_ip(g)
    assert _rel(g.cpu().numpy(), ref.mTimes_ik1(fh_np)) < TOL
    c = _cplx((nz, ny, nx // 2 + 1), 3)
    got = sp.dealias(_dev(c)).cpu().numpy()
    assert _rel(got, ref.dealias(c)) < TOL
    cE = _cplx((nz + 1, ny, nx // 2 + 1), 5)
    got = sp.dealias_edgeField(_dev(cE)).cpu().numpy()
    assert _rel(got, ref.dealias_edgeField(cE)) < TOL
    got = sp.take_ifft1d_z2z_ip(sp.take_fft1d_z2z_ip(_dev(c))).cpu().numpy()
    assert _rel(got, c) < TOL
    # setOddball zeroes the x-Nyquist column before the inverse
    assert _rel(sp.ifft(fh, setOddball=True).cpu().numpy(), ref.ifft(fh_np, True)) < TOL
    # the edge-grid type (nz+1 planes, not periodic in z): 2-D mask with the strict inequality
    spE = pdo.spectral()
    spE.init("x", nx, ny, nz + 1, *d, "four", "2/3rd", 2, fixOddball=False, init_periodicInZ=False)
    refE = IG.Spectral(nx, ny, nz + 1, *d, False)
    assert _rel(spE.dealias(_dev(cE)).cpu().numpy(), refE.dealias(cE)) < TOL
    # host pointers in, host pointers out
    out = np.empty((nz, ny, nx // 2 + 1), dtype=np.complex128)
    sp.fft(f, out)
    assert _rel(out, fh_np) < TOL


@pytest.mark.parametrize("shape", SHAPES)
def test_pade6stagg_and_projection_match_oracle(pdo, IG, shape):
    nx, ny, nz = shape
    d = [2 * np.pi / n for n in shape]
    spC, spE = pdo.spectral(), pdo.spectral()
    spC.init("x", nx, ny, nz, *d, fixOddball=False, init_periodicInZ=True)
    spE.init("x", nx, ny, nz + 1, *d, fixOddball=False, init_periodicInZ=False)
    der = pdo.Pade6stagg()
    der.init(spC.physdecomp, spC.spectdecomp, dz=d[2], scheme=1, isPeriodic=True)
    po = pdo.padepoisson()
    po.init(*d, spC, spE, derivZ=der)
    rC, rE = IG.Spectral(nx, ny, nz, *d, True), IG.Spectral(nx, ny, nz + 1, *d, False)
    rops = IG.Pade6stagg(nz, d[2])
    rpo = IG.PadePoisson(*d, rC, rE, rops)
    k = np.linspace(-3.0, 3.0, 7)
    assert np.allclose(der.getModifiedWavenumbers(k), rops.getModifiedWavenumbers(k), rtol=1e-15, atol=0)
    nxh = nx // 2 + 1
    uh, vh, wh = _cplx((nz, ny, nxh), 1), _cplx((nz, ny, nxh), 2), _cplx((nz + 1, ny, nxh), 3)
    wh[nz] = wh[0]
    # real and complex staggered ops through the dispatcher
    fr = broadband((nz, ny, nx), 9)
    assert _rel(der.ddz_C2E(_dev(fr)).cpu().numpy(), rops.ddz_C2E(fr)) < TOL
    assert _rel(der.interpz_E2C(_dev(wh)).cpu().numpy(), rops.interpz_E2C(wh)) < TOL
    assert _rel(der.d2dz2_E2E(_dev(wh)).cpu().numpy(), rops.d2dz2_E2E(wh)) < TOL
    # divergence, projection, pressure
    div, md = po.DivergenceCheck(_dev(uh), _dev(vh), _dev(wh))
    rdiv = rpo.divergence(uh, vh, wh)
    assert _rel(div.cpu().numpy(), rdiv) < TOL and abs(md - rdiv.max()) < TOL * np.abs(rdiv).max()
    ud, vd, wd = _dev(uh), _dev(vh), _dev(wh)
    po.PressureProjection(ud, vd, wd)
    ru, rv, rw = rpo.PressureProjection(uh, vh, wh)
    scale = max(np.abs(ru).max(), np.abs(rw).max())
    for got, ref in ((ud, ru), (vd, rv), (wd, rw)):
        assert np.abs(got.cpu().numpy() - ref).max() < TOL * scale
    div2, md2 = po.DivergenceCheck(ud, vd, wd)
    assert np.abs(div2.cpu().numpy()).max() < 1e-12 * np.abs(rdiv).max()
    p = po.getPressure(_dev(uh), _dev(vh), _dev(wh)).cpu().numpy()
    assert _rel(p, rpo.getPressure(uh, vh, wh)) < TOL
    ud, vd, wd = _dev(uh), _dev(vh), _dev(wh)
    p2 = po.getPressureAndUpdateRHS(ud, vd, wd).cpu().numpy()
    assert _rel(p2, p) < TOL and np.abs(ud.cpu().numpy() - ru).max() < TOL * scale
    # host arrays: updated in place
    uh2, vh2, wh2 = uh.copy(), vh.copy(), wh.copy()
    po.PressureProjection(uh2, vh2, wh2)
    assert np.abs(uh2 - ru).max() < TOL * scale and np.abs(wh2 - rw).max() < TOL * scale


def _tg_fields(n, direction):
    d = 2 * np.pi / n
    x = np.arange(n) * d
    zC, zE = (np.arange(n) + 0.5) * d, np.arange(n + 1) * d
    X, Y, ZC, ZE = x[None, None, :], x[None, :, None], zC[:, None, None], zE[:, None, None]
    if direction == 1:
        return (np.sin(X) * np.cos(Y) * np.ones((n, 1, 1)), -np.cos(X) * np.sin(Y) * np.ones((n, 1, 1)), np.zeros((n + 1, n, n)))
    return (np.sin(X) * np.cos(ZC) * np.ones((1, n, 1)), np.zeros((n, n, n)), -np.cos(X) * np.sin(ZE) * np.ones((1, n, 1)))


# (24, 16, 32): cuFFT passes; power-of-two extents: the hand-written passes of csrc/fft2d.cu with the products, i k and the dealiasing
# mask fused into their first loads
@pytest.mark.parametrize("shape", [(24, 16, 32), (32, 16, 16), (16, 64, 32)])
@pytest.mark.parametrize("scheme", [1, 2])
@pytest.mark.parametrize("inviscid", [False, True])
def test_igrid_substep_matches_oracle_broadband(pdo, IG, scheme, inviscid, shape):
    """Two full time steps (6 / 10 RK substeps) from a broadband, non-solenoidal start: every branch of the substep
    (dealiasing, projection with the divergence re-check, interpolation, gradients, skew-symmetric advection, viscous term)."""
    nx, ny, nz = shape
    L = (2 * np.pi, 2 * np.pi, 2 * np.pi)
    u, v = broadband((nz, ny, nx), 1), broadband((nz, ny, nx), 2)
    w = broadband((nz + 1, ny, nx), 3)
    w[nz] = w[0]
    Re = 50.0
    ref = IG.IGrid(nx, ny, nz, *L, Re, u, v, w, isInviscid=inviscid, TimeSteppingScheme=scheme)
    g = pdo.igrid()
    g.init(nx, ny, nz, *L, Re, u, v, w, isInviscid=inviscid, TimeSteppingScheme=scheme)
    scale = max(np.abs(ref.u).max(), np.abs(ref.w).max())
    for nm in ("u", "v", "w", "wC", "uE", "vE"):
        assert np.abs(g.get(nm) - getattr(ref, nm)).max() < TOL * scale, ("init", nm)
    dt = 0.01
    for it in range(2):
        ref.timeAdvance(dt)
        g.timeAdvance(dt)
        for nm in ("u", "v", "w", "wC", "uhat", "vhat", "what"):
            r = getattr(ref, nm)
            # measured: 1e-15 .. 2e-15 after 15 substeps (profiles/r01_igrid_parity.jsonl); the bar is north_star's 1e-12
            assert np.abs(g.get(nm) - r).max() < TOL * np.abs(r).max(), (it, nm, np.abs(g.get(nm) - r).max() / np.abs(r).max())
    assert g.step == 2 and abs(g.tsim - 2 * dt) < 1e-15
    assert g.maxDivergence() < 1e-11 * scale


@pytest.mark.parametrize("shape", [(24, 16, 32), (16, 32, 32)])
@pytest.mark.parametrize("scheme,adv,vert", [(1, 1, 1), (2, 0, 1), (1, 1, 2)])
def test_igrid_decomposed_code_path_on_one_gpu(pdo, IG, monkeypatch, shape, scheme, adv, vert):
    """PDO_IG_FORCE_TRANSPOSES=1 runs what a decomposed column communicator runs — explicit y <-> z transposes (device copies on one
    rank) and the z-resident dealias + projection of project_and_prep — on a single GPU: against the oracle, and bit for bit against
    the default single-GPU path (the same operations per element in another order of passes)."""
    nx, ny, nz = shape
    L = (2 * np.pi, 2 * np.pi, 2 * np.pi)
    u, v = broadband((nz, ny, nx), 11), broadband((nz, ny, nx), 12)
    w = broadband((nz + 1, ny, nx), 13)
    w[nz] = w[0]
    kw = dict(TimeSteppingScheme=scheme, AdvectionTerm=adv, NumericalSchemeVert=vert)
    ref = IG.IGrid(nx, ny, nz, *L, 80.0, u, v, w, **kw)
    ga = pdo.igrid()
    ga.init(nx, ny, nz, *L, 80.0, u, v, w, **kw)
    monkeypatch.setenv("PDO_IG_FORCE_TRANSPOSES", "1")
    gt = pdo.igrid()
    gt.init(nx, ny, nz, *L, 80.0, u, v, w, **kw)
    monkeypatch.delenv("PDO_IG_FORCE_TRANSPOSES")
    for it in range(2):
        ref.timeAdvance(0.01)
        ga.timeAdvance(0.01)
        gt.timeAdvance(0.01)
        for nm in ("u", "v", "w", "wC", "uhat", "vhat", "what"):
            r = getattr(ref, nm)
            assert np.abs(gt.get(nm) - r).max() < TOL * np.abs(r).max(), (it, nm)
            assert np.array_equal(gt.get(nm), ga.get(nm)), (it, nm, "decomposed-path result differs from the single-GPU path")
    assert gt.maxDivergence() < 1e-11 * max(np.abs(ref.u).max(), np.abs(ref.w).max())


@pytest.mark.parametrize("direction", [1, 2])
def test_igrid_taylor_green_decay_on_gpu(pdo, IG, direction):
    """problems/incompressible/TaylorGreenPeriodic: 32^3, Re = 100 — analytic decay exp(-2t/Re)."""
    n, Re = 32, 100.0
    u, v, w = _tg_fields(n, direction)
    g = pdo.igrid()
    g.init(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, Re, u, v, w, TimeSteppingScheme=2)
    dt = 0.25 * (2 * np.pi / n)
    for _ in range(4):
        g.timeAdvance(dt)
    decay = np.exp(-2.0 * g.tsim / Re)
    tol = 1e-9 if direction == 1 else 5e-7
    assert np.abs(g.get("u") - u * decay).max() < tol
    assert np.abs(g.get("w") - w * decay).max() < tol
    cfl_dt = g.compute_deltaT(0.4)
    umax = (np.abs(g.get("u")) / (2 * np.pi / n) + np.abs(g.get("v")) / (2 * np.pi / n) + np.abs(g.get("wC")) / (2 * np.pi / n)).max()
    assert abs(cfl_dt - min(0.4 / umax, 0.4 * Re * (2 * np.pi / n) ** 2)) < 1e-12 * cfl_dt


def test_igrid_all_gradients_flag_does_not_change_the_solution(pdo):
    n = 16
    u, v = broadband((n, n, n), 1), broadband((n, n, n), 2)
    w = broadband((n + 1, n, n), 3)
    w[n] = w[0]
    outs = []
    for flag in (False, True):
        g = pdo.igrid()
        g.init(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, 100.0, u, v, w, computeAllGradients=flag)
        g.timeAdvance(0.01)
        outs.append((g.get("u"), g.get("w")))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("shape", [(24, 16, 32), (32, 32, 16)])
@pytest.mark.parametrize("scheme", [1, 2])
def test_igrid_substep_rotational_form_matches_oracle(pdo, IG, scheme, shape):
    """AdvectionTerm = 0 (u x omega, igrid.F90:1527-1555 — what the authors' HIT deck runs) against the oracle."""
    nx, ny, nz = shape
    L = (2 * np.pi, 2 * np.pi, 2 * np.pi)
    u, v = broadband((nz, ny, nx), 1), broadband((nz, ny, nx), 2)
    w = broadband((nz + 1, ny, nx), 3)
    w[nz] = w[0]
    ref = IG.IGrid(nx, ny, nz, *L, 50.0, u, v, w, TimeSteppingScheme=scheme, AdvectionTerm=0)
    g = pdo.igrid()
    g.init(nx, ny, nz, *L, 50.0, u, v, w, TimeSteppingScheme=scheme, AdvectionTerm=0)
    for it in range(2):
        ref.timeAdvance(0.01)
        g.timeAdvance(0.01)
        for nm in ("u", "v", "w", "uhat", "vhat", "what"):
            r = getattr(ref, nm)
            assert np.abs(g.get(nm) - r).max() < TOL * np.abs(r).max(), (it, nm)
    assert g.maxDivergence() < 1e-11 * max(np.abs(ref.u).max(), np.abs(ref.w).max())


def test_pade6stagg_fourier_collocation_matches_oracle(pdo, IG):
    """scheme = fourierColl (PadeDerOps.F90:57-78, spectral.F90:387-680, 843-856): the six z-operators on complex arrays."""
    nx, ny, nz = 16, 12, 16
    d = [2 * np.pi / n for n in (nx, ny, nz)]
    spC = pdo.spectral()
    spC.init("x", nx, ny, nz, *d, fixOddball=False, init_periodicInZ=True)
    der = pdo.Pade6stagg()
    der.init(spC.physdecomp, spC.spectdecomp, dz=d[2], scheme=2, isPeriodic=True, spectC=spC)
    rops = IG.Pade6stagg(nz, d[2], scheme=2)
    nxh = nx // 2 + 1
    fC, fE = _cplx((nz, ny, nxh), 4), _cplx((nz + 1, ny, nxh), 5)
    fE[nz] = fE[0]
    for name in ("ddz_E2C", "interpz_E2C", "d2dz2_E2E"):
        assert _rel(getattr(der, name)(_dev(fE)).cpu().numpy(), getattr(rops, name)(fE)) < TOL, name
    for name in ("ddz_C2E", "interpz_C2E", "d2dz2_C2C"):
        assert _rel(getattr(der, name)(_dev(fC)).cpu().numpy(), getattr(rops, name)(fC)) < TOL, name
    k = np.linspace(-3.0, 3.0, 7)
    assert np.array_equal(der.getModifiedWavenumbers(k), k)
    with pytest.raises(pdo.PadeOpsError) as e:
        pdo.Pade6stagg().init(spC.physdecomp, spC.spectdecomp, dz=d[2], scheme=2, isPeriodic=True)    # no spectC
    assert e.value.code == 43


@pytest.mark.parametrize("adv", [0, 1])
def test_igrid_substep_fourier_z_matches_oracle(pdo, IG, adv):
    """NumericalSchemeVert = 2 with either advection form (rotational + Fourier-z is the authors' HIT deck's choice)."""
    nx, ny, nz = 24, 16, 32
    L = (2 * np.pi, 2 * np.pi, 2 * np.pi)
    u, v = broadband((nz, ny, nx), 1), broadband((nz, ny, nx), 2)
    w = broadband((nz + 1, ny, nx), 3)
    w[nz] = w[0]
    ref = IG.IGrid(nx, ny, nz, *L, 50.0, u, v, w, TimeSteppingScheme=2, AdvectionTerm=adv, NumericalSchemeVert=2)
    g = pdo.igrid()
    g.init(nx, ny, nz, *L, 50.0, u, v, w, TimeSteppingScheme=2, AdvectionTerm=adv, NumericalSchemeVert=2)
    for it in range(2):
        ref.timeAdvance(0.01)
        g.timeAdvance(0.01)
        for nm in ("u", "v", "w", "uhat", "vhat", "what"):
            r = getattr(ref, nm)
            assert np.abs(g.get(nm) - r).max() < TOL * np.abs(r).max(), (it, nm)
    assert g.maxDivergence() < 1e-11 * max(np.abs(ref.u).max(), np.abs(ref.w).max())


def test_restart_files_round_trip(pdo, IG, tmp_path):
    """dumpRestartFile / readRestartFile / dumpFullField in the reference's format (igrid.F90:2719-2823): flat global
    Fortran-order doubles per field + the g15.5 info file; a run restarted from them continues like the original."""
    nx, ny, nz = 16, 16, 16
    L = (2 * np.pi,) * 3
    u, v = broadband((nz, ny, nx), 1), broadband((nz, ny, nx), 2)
    w = broadband((nz + 1, ny, nx), 3)
    w[nz] = w[0]
    g = pdo.igrid()
    g.init(nx, ny, nz, *L, 50.0, u, v, w, TimeSteppingScheme=1)
    for _ in range(2):
        g.timeAdvance(0.01)
    g.dumpRestartFile(tmp_path, runID=7)
    for nm in ("u", "v", "w"):
        assert (tmp_path / f"RESTART_Run07_{nm}.000002").read_bytes() == g.get(nm).tobytes()
    assert (tmp_path / "RESTART_Run07_info.000002").read_text() == "    0.20000E-01\n"
    g.dumpFullField("wC", "wCel", tmp_path, runID=7)
    assert (tmp_path / "Run07_wCel_t000002.out").read_bytes() == g.get("wC").tobytes()
    h = pdo.igrid()
    h.init(nx, ny, nz, *L, 50.0, np.zeros_like(u), np.zeros_like(v), np.zeros_like(w), TimeSteppingScheme=1)
    h.readRestartFile(2, 7, tmp_path)
    assert h.step == 2 and abs(h.tsim - 0.02) < 1e-15
    for nm in ("u", "v", "w", "wC"):
        assert np.abs(h.get(nm) - g.get(nm)).max() < TOL * np.abs(g.get(nm)).max(), nm     # re-projection of a projected field
    g.timeAdvance(0.01)
    h.timeAdvance(0.01)
    for nm in ("u", "v", "w"):
        assert np.abs(h.get(nm) - g.get(nm)).max() < TOL * np.abs(g.get(nm)).max(), nm
    with pytest.raises(pdo.PadeOpsError) as e:
        h.readRestartFile(3, 7, tmp_path)
    assert e.value.code == 321


def test_hit_forcing_matches_oracle(pdo, IG):
    """HIT_shell_forcing%getRHS_HITforcing (forcingIsotropic.F90:254-311) on device-resident right-hand sides: the library's
    direct-DFT evaluation against the oracle's whole-field FFT formulation, with injected and with drawn wavenumbers."""
    import torch
    nx, ny, nz = 16, 12, 16
    d = [2 * np.pi / n for n in (nx, ny, nz)]
    spC, spE = pdo.spectral(), pdo.spectral()
    spC.init("x", nx, ny, nz, *d, fixOddball=False, init_periodicInZ=True)
    spE.init("x", nx, ny, nz + 1, *d, fixOddball=False, init_periodicInZ=False)
    rC = IG.Spectral(nx, ny, nz, *d, True, 2.0 / 3.0, False)
    nxh = nx // 2 + 1
    uh, vh, wh = _cplx((nz, ny, nxh), 1), _cplx((nz, ny, nxh), 2), _cplx((nz + 1, ny, nxh), 3)
    r = [_cplx((nz, ny, nxh), 4), _cplx((nz, ny, nxh), 5), _cplx((nz + 1, ny, nxh), 6)]
    kw = dict(kmin=2.0, kmax=4.0, Nwaves=8, EpsAmplitude=0.3, tidStart=3, RandSeedToAdd=1)
    f = pdo.HIT_shell_forcing()
    f.init(spC, spE, **kw)
    ref = IG.HITForcing(rC, **kw)
    waves = ([2, 3, 2, 0, 6, 40, 1, 1], [1, 9, 1, 0, 3, 1, 2, 11], [4, 0, 4, 7, 15, 1, 2, 3])
    f.set_wavenumbers(*waves)
    ref.set_wavenumbers(*waves)
    got = f.getRHS_HITforcing(*[_dev(a) for a in r], _dev(uh), _dev(vh), _dev(wh), False)
    want = ref.getRHS_HITforcing(r[0], r[1], r[2], uh, vh, wh, False)
    for a, b in zip(got, want):
        assert _rel(a.cpu().numpy(), b) < TOL
    for _ in range(2):          # newTimestep: both sides draw from the shared generator and advance their seeds
        got = f.getRHS_HITforcing(*[_dev(a) for a in r], _dev(uh), _dev(vh), _dev(wh), True)
        want = ref.getRHS_HITforcing(r[0], r[1], r[2], uh, vh, wh, True)
        assert f.get_wavenumbers() == (ref.wave_x.tolist(), ref.wave_y.tolist(), ref.wave_z.tolist())
        for a, b in zip(got, want):
            assert _rel(a.cpu().numpy(), b) < TOL
    with pytest.raises(pdo.PadeOpsError):
        f.getRHS_HITforcing(r[0], r[1], r[2], uh, vh, wh, False)     # host arrays: the forcing lives on the device
    assert torch.cuda.is_available()


@pytest.mark.parametrize("vert", [1, 2])
def test_igrid_with_hit_forcing_matches_oracle(pdo, IG, vert):
    """useHITForcing = .true. (igrid.F90:940-944, 1907-1910): a new draw per step, the same waves through the RK stages."""
    n = 16
    L = (2 * np.pi,) * 3
    rng = np.random.default_rng(7)
    u, v = 0.3 * rng.standard_normal((n, n, n)), 0.3 * rng.standard_normal((n, n, n))
    w = 0.3 * rng.standard_normal((n + 1, n, n))
    w[n] = w[0]
    hit = dict(kmin=1.0, kmax=2.5, Nwaves=12, EpsAmplitude=0.5, RandSeedToAdd=3)
    ref = IG.IGrid(n, n, n, *L, 1.0e3, u, v, w, TimeSteppingScheme=2, NumericalSchemeVert=vert, HITForcing_=hit)
    g = pdo.igrid()
    g.init(n, n, n, *L, 1.0e3, u, v, w, TimeSteppingScheme=2, NumericalSchemeVert=vert)
    g.enableHITForcing(**hit)
    for it in range(2):
        ref.timeAdvance(0.005)
        g.timeAdvance(0.005)
        for nm in ("u", "v", "w"):
            r = getattr(ref, nm)
            assert np.abs(g.get(nm) - r).max() < 1e-10 * np.abs(r).max(), (it, nm)    # den amplifies rounding of weak modes
    assert g.maxDivergence() < 1e-10


@pytest.mark.parametrize("mid,Csgs,explicitE", [(0, 0.17, False), (1, 1.5, False), (2, 1.67, False), (2, 1.67, True)])
def test_igrid_with_sgs_matches_oracle(pdo, IG, mid, Csgs, explicitE):
    """useSGS = .true. (igrid.F90:1866-1871): Smagorinsky / sigma / AMD with a global constant, interpolated or explicit edge
    viscosity."""
    n = 16
    L = (2 * np.pi,) * 3
    rng = np.random.default_rng(3)
    u, v = 0.3 * rng.standard_normal((n, n, n)), 0.3 * rng.standard_normal((n, n, n))
    w = 0.3 * rng.standard_normal((n + 1, n, n))
    w[n] = w[0]
    sgs = dict(SGSModelID=mid, Csgs=Csgs, explicitCalcEdgeEddyViscosity=explicitE)
    ref = IG.IGrid(n, n, n, *L, 1.0e3, u, v, w, TimeSteppingScheme=1, SGS_=sgs)
    g = pdo.igrid()
    g.init(n, n, n, *L, 1.0e3, u, v, w, TimeSteppingScheme=1, computeAllGradients=True)
    g.enableSGS(**sgs)
    for it in range(2):
        ref.timeAdvance(0.01)
        g.timeAdvance(0.01)
        for nm in ("u", "v", "w"):
            r = getattr(ref, nm)
            assert np.abs(g.get(nm) - r).max() < 1e-11 * np.abs(r).max(), (it, nm)     # sqrt / acos of the kernels at rounding level
    h = pdo.igrid()
    h.init(n, n, n, *L, 1.0e3, u, v, w, TimeSteppingScheme=1)
    with pytest.raises(pdo.PadeOpsError):
        h.enableSGS(**sgs)            # the models need all eighteen gradients
    with pytest.raises(pdo.PadeOpsError) as e:
        g.enableSGS(SGSModelID=7)
    assert e.value.code == 213


def test_hit_periodic_deck_configuration(pdo, IG):
    """The combination the authors' HIT_Periodic deck runs (problems/incompressible/HIT_Periodic_moving_files/input_fourier.dat):
    rotational advection, Fourier collocation in z, SSP-RK45, AMD model (Csgs = 1.67), HIT shell forcing (kmin 4 ... here 1-2.5 on
    a 16^3 box, Nwaves = 20, EpsAmplitude = 0.05)."""
    n = 16
    L = (2 * np.pi,) * 3
    rng = np.random.default_rng(9)
    u, v = 0.3 * rng.standard_normal((n, n, n)), 0.3 * rng.standard_normal((n, n, n))
    w = 0.3 * rng.standard_normal((n + 1, n, n))
    w[n] = w[0]
    hit = dict(kmin=1.0, kmax=2.5, Nwaves=20, EpsAmplitude=0.05, RandSeedToAdd=0)
    sgs = dict(SGSModelID=2, Csgs=1.67, explicitCalcEdgeEddyViscosity=False)
    ref = IG.IGrid(n, n, n, *L, 1.0e10, u, v, w, TimeSteppingScheme=2, AdvectionTerm=0, NumericalSchemeVert=2, HITForcing_=hit, SGS_=sgs)
    g = pdo.igrid()
    g.init(n, n, n, *L, 1.0e10, u, v, w, TimeSteppingScheme=2, AdvectionTerm=0, NumericalSchemeVert=2, computeAllGradients=True)
    g.enableSGS(**sgs)
    g.enableHITForcing(**hit)
    for it in range(2):
        ref.timeAdvance(0.005)
        g.timeAdvance(0.005)
        for nm in ("u", "v", "w"):
            r = getattr(ref, nm)
            assert np.abs(g.get(nm) - r).max() < 1e-10 * np.abs(r).max(), (it, nm)
    assert g.maxDivergence() < 1e-10


def _channel_fields(nx, ny, nz, Lz, slip):
    x = np.arange(nx) * 2 * np.pi / nx
    y = np.arange(ny) * 2 * np.pi / ny
    zc = (np.arange(nz) + 0.5) * Lz / nz
    ze = np.arange(nz + 1) * Lz / nz
    X, Y = x[None, None, :], y[None, :, None]
    a = np.pi / Lz
    cz, sz = (np.cos, np.sin) if slip else (np.sin, np.cos)
    u = np.sin(X) * np.cos(Y) * cz(a * zc)[:, None, None] + 0.3 * np.cos(2 * Y) * cz(2 * a * zc)[:, None, None] + 0 * X
    v = -np.cos(X) * np.sin(Y) * cz(a * zc)[:, None, None] + 0.2 * np.sin(2 * X) * cz(3 * a * zc)[:, None, None] + 0 * Y
    wz = np.sin(2 * a * ze) if slip else np.sin(a * ze) ** 2
    w = 0.25 * np.sin(X) * np.sin(2 * Y) * wz[:, None, None]
    return u, v, w


@pytest.mark.parametrize("walls", [(2, 2), (1, 1), (1, 2)])
@pytest.mark.parametrize("adv,scheme,stokes", [(1, 1, True), (0, 2, True), (1, 2, False)])
def test_wall_bounded_igrid_matches_oracle(pdo, IG, walls, adv, scheme, stokes):
    """PeriodicInZ = .false. with slip / no-slip walls: the stencil codes reach every z-operator, the staggered operators take their
    wall closures, the projection its even / odd extension (igrid.F90:5148-5204, PadeDerOps.F90:92-110, PadePoisson.F90:459-623)."""
    nx, ny, nz, Lz = 16, 12, 24, 2.0
    L = 2 * np.pi
    u, v, w = _channel_fields(nx, ny, nz, Lz, slip=(walls == (2, 2)))
    kw = dict(TimeSteppingScheme=scheme, AdvectionTerm=adv, PeriodicInZ=False, botWall=walls[0], topWall=walls[1], ComputeStokesPressure=stokes)
    ref = IG.IGrid(nx, ny, nz, L, L, Lz, 100.0, u, v, w, **kw)
    g = pdo.igrid()
    g.init(nx, ny, nz, L, L, Lz, 100.0, u, v, w, **kw)
    for nm in ("u", "v", "w", "wC"):
        r = getattr(ref, nm)
        assert np.abs(g.get(nm) - r).max() < TOL * max(np.abs(r).max(), 1e-3), ("init", nm)
    for it in range(2):
        ref.timeAdvance(0.005)
        g.timeAdvance(0.005)
        for nm in ("u", "v", "w"):
            r = getattr(ref, nm)
            assert np.abs(g.get(nm) - r).max() < 1e-11 * np.abs(ref.u).max(), (it, nm)
    assert not np.any(g.get("w")[0]) and not np.any(g.get("w")[nz])
    with pytest.raises(pdo.PadeOpsError) as e:
        pdo.igrid().init(nx, ny, nz, L, L, Lz, 100.0, u, v, w, PeriodicInZ=False, NumericalSchemeVert=2)
    assert e.value.code == 123
    with pytest.raises(pdo.PadeOpsError):
        pdo.igrid().init(nx, ny, nz, L, L, Lz, 100.0, u, v, w, PeriodicInZ=False, botWall=3)
