"""GPU parity of the NON-PERIODIC staggered compact operators (cd06stagg%init_nonperiodic) and of
Pade6stagg's wall dispatch (PadeDerOps.F90:92-110, 185-205, 449-482) against the oracle.  Bar: 1e-12 relative.

The kernel's per-line routine
is also verified on the host for every operator and wall combination (tests/test_stagg_nonperiodic_cpu.py)."""
import itertools

import numpy as np
import pytest

from conftest import broadband

pytestmark = [pytest.mark.gpu]
TOL = 1e-12
OPS = [("ddz_E2C", 1, 0), ("ddz_C2E", 0, 1), ("ddz_C2C", 0, 0), ("ddz_E2E", 1, 1), ("InterpZ_E2C", 1, 0), ("InterpZ_C2E", 0, 1),
       ("d2dz2_C2C", 0, 0), ("d2dz2_E2E", 1, 1)]


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(got, ref):
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("walls", list(itertools.product([0, 1], [0, 1], [0, 1], [0, 1])))
@pytest.mark.parametrize("cplx", [False, True])
def test_cd06stagg_nonperiodic_operators(pdo, oracle, walls, cplx):
    from oracle import stagg_np_oracle as SN
    be, te, bs, ts = walls
    n, n2, n1 = 48, 9, 37          # odd n1: partial warp of columns
    dz = 1.0 / n
    st = pdo.cd06stagg()
    st.init(n, dz, isTopEven=bool(te), isBotEven=bool(be), isTopSided=bool(ts), isBotSided=bool(bs))
    ref = SN.CD06StaggNP(n, dz, bool(te), bool(be), bool(ts), bool(bs))
    for name, ein, eout in OPS:
        f = broadband((n + ein, n2, n1), seed=len(name) + ein)
        if cplx:
            f = f + 1j * broadband((n + ein, n2, n1), seed=50 + eout)
        got = getattr(st, name)(_dev(f)).cpu().numpy()
        want = getattr(ref, name)(f)
        assert got.shape == want.shape == (n + eout, n2, n1)
        assert _rel(got, want) < TOL, (name, walls, cplx)
    # host arrays take the same entry points
    f = broadband((n + 1, n2, n1), seed=3)
    assert _rel(st.ddz_E2C(f.copy()), ref.ddz_E2C(f)) < TOL
    st.destroy()


def test_periodic_handle_has_no_collocated_first_derivative(pdo):
    st = pdo.cd06stagg()
    st.init(32, 0.1)
    with pytest.raises(pdo.PadeOpsError):
        st.ddz_C2C(_dev(np.zeros((32, 2, 4))))
    st.destroy()


def test_pade6stagg_wall_dispatch(pdo, oracle):
    from oracle import igrid_oracle as IG
    nx, ny, nz = 16, 12, 24
    d = [2 * np.pi / nx, 2 * np.pi / ny, 1.0 / nz]
    spC = pdo.spectral()
    spC.init("x", nx, ny, nz, *d, fixOddball=False, init_periodicInZ=False)
    der = pdo.Pade6stagg()
    der.init(spC.physdecomp, spC.spectdecomp, dz=d[2], scheme=1, isPeriodic=False)
    rops = IG.Pade6stagg(nz, d[2], scheme=1, isPeriodic=False)
    nxh = nx // 2 + 1
    fC = broadband((nz, ny, nxh), 4) + 1j * broadband((nz, ny, nxh), 5)
    fE = broadband((nz + 1, ny, nxh), 6) + 1j * broadband((nz + 1, ny, nxh), 7)
    rC, rE = broadband((nz, ny, nx), 8), broadband((nz + 1, ny, nx), 9)
    for bot, top in itertools.product((-1, 0, 1), (-1, 0, 1)):
        for name, inC, inE in (("ddz_E2C", None, fE), ("interpz_E2C", None, fE), ("d2dz2_E2E", None, fE), ("ddz_C2E", fC, None),
                               ("interpz_C2E", fC, None), ("d2dz2_C2C", fC, None)):
            f = inC if inC is not None else inE
            got = getattr(der, name)(_dev(f), bot=bot, top=top).cpu().numpy()
            want = getattr(rops, name)(f, bot, top)
            assert got.shape == want.shape
            if not np.any(want):
                assert not np.any(got), (name, bot, top)        # unsupported combination: output = 0
            else:
                assert _rel(got, want) < TOL, (name, bot, top)
        assert _rel(der.ddz_C2E(_dev(rC), bot=bot, top=top).cpu().numpy(), rops.ddz_C2E(rC, bot, top)) < TOL
        assert _rel(der.interpz_E2C(_dev(rE), bot=bot, top=top).cpu().numpy(), rops.interpz_E2C(rE, bot, top)) < TOL
    with pytest.raises(pdo.PadeOpsError) as e:
        pdo.Pade6stagg().init(spC.physdecomp, spC.spectdecomp, dz=d[2], scheme=2, isPeriodic=False, spectC=spC)
    assert e.value.code == 323
    der.destroy()


@pytest.mark.parametrize("shape", [(16, 12, 24), (12, 16, 10)])
def test_wall_bounded_projection(pdo, oracle, shape):
    """padepoisson PressureProjection / DivergenceCheck with PeriodicInZ = .false. (PadePoisson.F90:459-623, 1165-1244)"""
    from oracle import igrid_oracle as IG
    nx, ny, nz = shape
    d = [2 * np.pi / nx, 2 * np.pi / ny, 1.0 / nz]
    spC, spE = pdo.spectral(), pdo.spectral()
    spC.init("x", nx, ny, nz, *d, fixOddball=False, init_periodicInZ=False)
    spE.init("x", nx, ny, nz + 1, *d, fixOddball=False, init_periodicInZ=False)
    der = pdo.Pade6stagg()
    der.init(spC.physdecomp, spC.spectdecomp, dz=d[2], scheme=1, isPeriodic=False)
    po = pdo.padepoisson()
    po.init(*d, spC, spE, derivZ=der, PeriodicInZ=False)
    rC, rE = IG.Spectral(nx, ny, nz, *d), IG.Spectral(nx, ny, nz + 1, *d)
    rP = IG.PadePoisson(*d, rC, rE, IG.Pade6stagg(nz, d[2], 1, isPeriodic=False), PeriodicInZ=False)
    rng = np.random.default_rng(nz)
    u, v = rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz, ny, nx))
    w = rng.standard_normal((nz + 1, ny, nx))
    w[0] = 0.0
    w[nz] = 0.0        # no penetration: the odd extension of w is continuous only then
    uh, vh, wh = rC.fft(u), rC.fft(v), rE.fft(w)
    want = rP.PressureProjection(uh, vh, wh)
    du, dv, dw = _dev(uh), _dev(vh), _dev(wh)
    po.PressureProjection(du, dv, dw)
    for got, ref in zip((du, dv, dw), want):
        assert _rel(got.cpu().numpy(), ref) < TOL
    div, _ = po.DivergenceCheck(du, dv, dw)
    assert np.abs(div.cpu().numpy()).max() < 1e-11 * np.abs(rP.divergence(uh, vh, wh)).max()
    # getPressure / getPressureAndUpdateRHS with walls (PadePoisson.F90:762-896, 963-1160): the inputs of getPressure are intent(in)
    a, b, c = _dev(uh), _dev(vh), _dev(wh)
    pr = po.getPressure(a, b, c).cpu().numpy()
    assert _rel(pr, rP.getPressure(uh, vh, wh)) < TOL
    assert np.array_equal(a.cpu().numpy(), uh) and np.array_equal(c.cpu().numpy(), wh)
    pr2 = po.getPressureAndUpdateRHS(a, b, c).cpu().numpy()
    wu, wv, ww, wp = rP.getPressureAndUpdateRHS(uh, vh, wh)
    assert _rel(pr2, wp) < TOL
    for got, ref in zip((a, b, c), (wu, wv, ww)):
        assert _rel(got.cpu().numpy(), ref) < TOL
    pr_host = np.empty_like(pr)
    po.getPressure(uh.copy(), vh.copy(), wh.copy(), pr_host)                     # host-pointer (drop-in) path
    assert _rel(pr_host, rP.getPressure(uh, vh, wh)) < TOL
    with pytest.raises(pdo.PadeOpsError):
        pdo.padepoisson().init(*d, spC, spE, derivZ=der, PeriodicInZ=True)      # periodicity of derivZ and the solver must agree
    # computeStokesPressure = .true. (:320-384): w* nonzero on the walls on input
    ps = pdo.padepoisson()
    ps.init(*d, spC, spE, computeStokesPressure=True, Lz=1.0, derivZ=der, PeriodicInZ=False)
    rS = IG.PadePoisson(*d, rC, rE, IG.Pade6stagg(nz, d[2], 1, isPeriodic=False), PeriodicInZ=False, computeStokesPressure=True, Lz=1.0)
    uhd, vhd = rC.dealias(uh), rC.dealias(vh)
    whs = rE.dealias(rE.fft(rng.standard_normal((nz + 1, ny, nx))))
    want = rS.PressureProjection(uhd, vhd, whs)
    du, dv, dw = _dev(uhd), _dev(vhd), _dev(whs)
    ps.PressureProjection(du, dv, dw)
    for got, ref in zip((du, dv, dw), want):
        assert _rel(got.cpu().numpy(), ref) < TOL
    div, _ = ps.DivergenceCheck(du, dv, dw)
    assert np.abs(div.cpu().numpy()).max() < 1e-11 * np.abs(rS.divergence(uhd, vhd, whs)).max()
    # pressure getters with the Stokes pressure.  Reference quirk kept: getPressureAndUpdateRHS adds the Stokes pieces of the LAST
    # getPressure call (it runs ProjectStokesPressure, which does not refresh phat_z1 / phat_z2; :1146-1156) — zero before any
    a, b, c = _dev(uhd), _dev(vhd), _dev(whs)
    p0 = ps.getPressureAndUpdateRHS(a, b, c).cpu().numpy()
    wu, wv, ww, wp0 = rS.getPressureAndUpdateRHS(uhd, vhd, whs)
    assert _rel(p0, wp0) < TOL
    for got, ref in zip((a, b, c), (wu, wv, ww)):
        assert _rel(got.cpu().numpy(), ref) < TOL
    a, b, c = _dev(uhd), _dev(vhd), _dev(whs)
    pr = ps.getPressure(a, b, c).cpu().numpy()
    assert _rel(pr, rS.getPressure(uhd, vhd, whs)) < TOL
    assert np.array_equal(a.cpu().numpy(), uhd) and np.array_equal(c.cpu().numpy(), whs)
    p1 = ps.getPressureAndUpdateRHS(a, b, c).cpu().numpy()                       # now carries the pieces getPressure stored
    assert _rel(p1, rS.getPressureAndUpdateRHS(uhd, vhd, whs)[3]) < TOL
    assert np.abs(p1 - p0).max() > 1e-6 * np.abs(p1).max()
