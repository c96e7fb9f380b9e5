"""Known-answer tests that pin the igrid / projection oracle (oracle/igrid_oracle.py): the reference ships no golden
vectors, so the restatement is held to the analytic identities its own tests and problems rely on
(divergence after projection, PadePoisson.F90:1209; 2-D Taylor-Green decay,
problems/incompressible/TaylorGreenPeriodic_files/initialize.F90:145-155; modified wavenumbers,
tests/test_PoissonPeriodic.F90:12-25)."""
import numpy as np
import pytest

from oracle import igrid_oracle as IG
from oracle import oracle as O


def _grid(nx, ny, nz):
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    x = np.arange(nx) * dx
    y = np.arange(ny) * dy
    zC = (np.arange(nz) + 0.5) * dz
    zE = np.arange(nz + 1) * dz
    return x, y, zC, zE


def _rand_fields(nx, ny, nz, seed=0):
    rng = np.random.default_rng(seed)
    u = rng.standard_normal((nz, ny, nx))
    v = rng.standard_normal((nz, ny, nx))
    w = rng.standard_normal((nz + 1, ny, nx))
    w[nz] = w[0]
    return u, v, w


def test_spectral_roundtrip_and_derivative_symbols():
    nx, ny, nz = 16, 12, 8
    sp = IG.Spectral(nx, ny, nz, 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz, True)
    x, y, zC, _ = _grid(nx, ny, nz)
    f = np.sin(3 * x)[None, None, :] * np.cos(2 * y)[None, :, None] * np.cos(zC)[:, None, None]
    fh = sp.fft(f)
    assert np.abs(sp.ifft(fh) - f).max() < 1e-14
    dfdx = sp.ifft(sp.mTimes_ik1(fh))
    dfdy = sp.ifft(sp.mTimes_ik2(fh))
    assert np.abs(dfdx - 3 * np.cos(3 * x)[None, None, :] * np.cos(2 * y)[None, :, None] * np.cos(zC)[:, None, None]).max() < 1e-13
    assert np.abs(dfdy + 2 * np.sin(3 * x)[None, None, :] * np.sin(2 * y)[None, :, None] * np.cos(zC)[:, None, None]).max() < 1e-13
    # 2/3 rule: a mode at k = 5 <= (2/3)*8 survives, k = 6 >= 16/3 is removed (x has 16 points -> cutoff 16/3)
    g = np.sin(5 * x)[None, None, :] * np.ones((nz, ny, 1)) + np.sin(6 * x)[None, None, :] * np.ones((nz, ny, 1))
    gd = sp.ifft(sp.dealias(sp.fft(g)))
    assert np.abs(gd - np.sin(5 * x)[None, None, :]).max() < 1e-13


def test_modified_wavenumber_matches_the_compact_operator():
    """getmodCD06stagg is the symbol of ddz_C2E / ddz_E2C: d/dz of exp(ikz) on the staggered grid = i k' exp(ikz)."""
    nz = 32
    dz = 2 * np.pi / nz
    _, _, zC, zE = _grid(4, 4, nz)
    ops = IG.Pade6stagg(nz, dz)
    for k in (1, 5, 11):
        kp = IG.getmodCD06stagg(np.array([float(k)]), dz)[0]
        fC = np.exp(1j * k * zC)[:, None, None] * np.ones((1, 3, 2))
        dE = ops.ddz_C2E(fC)
        assert np.abs(dE - 1j * kp * np.exp(1j * k * zE)[:, None, None]).max() < 1e-12 * max(1.0, kp)


@pytest.mark.parametrize("shape", [(16, 12, 8), (12, 16, 16)])
def test_projection_makes_the_field_divergence_free(shape):
    nx, ny, nz = shape
    d = [2 * np.pi / n for n in shape]
    spC = IG.Spectral(nx, ny, nz, *d, True)
    spE = IG.Spectral(nx, ny, nz + 1, *d, False)
    ops = IG.Pade6stagg(nz, d[2])
    po = IG.PadePoisson(*d, spC, spE, ops)
    u, v, w = _rand_fields(nx, ny, nz)
    uh, vh, wh = spC.fft(u), spC.fft(v), spE.fft(w)
    div0 = po.divergence(uh, vh, wh)
    uh, vh, wh = po.PressureProjection(uh, vh, wh)
    div1 = po.divergence(uh, vh, wh)
    assert np.abs(div0).max() > 1.0
    assert np.abs(div1).max() < 1e-12 * np.abs(div0).max()       # PadePoisson.F90:1209 threshold scale
    assert np.abs(wh[nz] - wh[0]).max() < 1e-12 * np.abs(wh).max()  # edge plane nz+1 stays the periodic image
    # projecting twice changes nothing (idempotence)
    u2, v2, w2 = po.PressureProjection(uh, vh, wh)
    assert max(np.abs(u2 - uh).max(), np.abs(v2 - vh).max(), np.abs(w2 - wh).max()) < 1e-11 * np.abs(uh).max()


@pytest.mark.parametrize("scheme", [1, 2])
@pytest.mark.parametrize("direction", [1, 2])
def test_taylor_green_decay(scheme, direction):
    """2-D Taylor-Green vortex, Re = 100: u = sin x cos y exp(-2t/Re) (directionID 1) and its x-z variant (directionID 2,
    which drives every staggered z operator)."""
    n = 16
    Re = 100.0
    x, y, zC, zE = _grid(n, n, n)
    X, Y, ZC, ZE = x[None, None, :], y[None, :, None], zC[:, None, None], zE[:, None, None]
    if direction == 1:
        u = np.sin(X) * np.cos(Y) * np.ones((n, 1, 1))
        v = -np.cos(X) * np.sin(Y) * np.ones((n, 1, 1))
        w = np.zeros((n + 1, n, n))
    else:
        u = np.sin(X) * np.cos(ZC) * np.ones((1, n, 1))
        v = np.zeros((n, n, n))
        w = -np.cos(X) * np.sin(ZE) * np.ones((1, n, 1))
    g = IG.IGrid(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, Re, u, v, w, TimeSteppingScheme=scheme)
    dt = 0.25 * (2 * np.pi / n)
    for _ in range(4):
        g.timeAdvance(dt)
    decay = np.exp(-2.0 * g.tsim / Re)
    tol = 1e-9 if direction == 1 else 2e-5   # directionID 2 carries the 6th-order truncation error of the z schemes
    # (measured: 2.7e-6 at n=16 -> 4.2e-8 at n=32 for equal t, ratio 63 = 2^6)
    assert np.abs(g.u - u * decay).max() < tol
    assert np.abs(g.w - w * decay).max() < tol
    assert np.abs(g.v - v * decay).max() < tol
    assert np.abs(g.poiss.divergence(g.uhat, g.vhat, g.what)).max() < 1e-12


@pytest.mark.parametrize("direction", [1, 2])
def test_taylor_green_decay_rotational_form(direction):
    """AdvectionTerm = 0 (u x omega, igrid.F90:1527-1555 — the form the authors' HIT deck runs): Taylor-Green is an exact
    Navier-Stokes solution whatever the form of the advection term, the gradient part goes into the pressure."""
    n = 16
    Re = 100.0
    x, y, zC, zE = _grid(n, n, n)
    X, Y, ZC, ZE = x[None, None, :], y[None, :, None], zC[:, None, None], zE[:, None, None]
    if direction == 1:
        u = np.sin(X) * np.cos(Y) * np.ones((n, 1, 1)); v = -np.cos(X) * np.sin(Y) * np.ones((n, 1, 1)); w = np.zeros((n + 1, n, n))
    else:
        u = np.sin(X) * np.cos(ZC) * np.ones((1, n, 1)); v = np.zeros((n, n, n)); w = -np.cos(X) * np.sin(ZE) * np.ones((1, n, 1))
    g = IG.IGrid(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, Re, u, v, w, TimeSteppingScheme=1, AdvectionTerm=0)
    dt = 0.25 * (2 * np.pi / n)
    for _ in range(4):
        g.timeAdvance(dt)
    decay = np.exp(-2.0 * g.tsim / Re)
    tol = 1e-9 if direction == 1 else 2e-5
    assert np.abs(g.u - u * decay).max() < tol and np.abs(g.w - w * decay).max() < tol and np.abs(g.v - v * decay).max() < tol
    assert np.abs(g.poiss.divergence(g.uhat, g.vhat, g.what)).max() < 1e-12


def test_rotational_and_skew_symmetric_forms_agree_to_truncation_error():
    """u x omega = -(u . grad) u + grad(|u|^2 / 2): after the projection the two forms differ only by the discretisation
    (aliasing and the z schemes' truncation), which shrinks with resolution."""
    diffs = []
    for n in (16, 32):
        x, y, zC, zE = _grid(n, n, n)
        X, Y, ZC, ZE = x[None, None, :], y[None, :, None], zC[:, None, None], zE[:, None, None]
        # a smooth multi-mode field (the initial projection makes it solenoidal)
        u = np.sin(2 * X) * np.cos(Y) * np.cos(ZC) + np.cos(3 * Y) * np.sin(ZC)
        v = np.cos(X) * np.sin(2 * Y) * np.sin(2 * ZC) + np.sin(X)
        w = np.sin(X) * np.cos(2 * Y) * np.sin(ZE) + np.cos(2 * X) * np.sin(3 * Y) * np.cos(2 * ZE)
        out = []
        for adv in (0, 1):
            g = IG.IGrid(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, 50.0, u, v, w, TimeSteppingScheme=1, AdvectionTerm=adv)
            g.timeAdvance(0.02)
            out.append((g.u.copy(), g.v.copy(), g.w.copy()))
        diffs.append(max(np.abs(a - b).max() for a, b in zip(*out)))
    # measured: 5.6e-5 (n = 16), 1.0e-6 (n = 32), 1.5e-8 (n = 64): sixth order, the z schemes' truncation error
    assert diffs[0] < 2e-4 and diffs[1] < diffs[0] / 30, diffs


def test_fourier_collocation_z_operators_are_exact_on_resolved_modes():
    """scheme = fourierColl (spectral.F90:843-856 tables): derivative, interpolation and second derivative between the cell
    and edge grids are exact for every resolved z-mode, real or complex."""
    nz = 16
    dz = 2 * np.pi / nz
    zC, zE = (np.arange(nz) + 0.5) * dz, np.arange(nz + 1) * dz
    ops = IG.Pade6stagg(nz, dz, scheme=2)
    for m in (1, 3, 5):
        fC = (np.exp(1j * m * zC))[:, None, None] * np.ones((1, 2, 3))
        fE = (np.exp(1j * m * zE))[:, None, None] * np.ones((1, 2, 3))
        assert np.abs(ops.ddz_E2C(fE) - 1j * m * fC).max() < 1e-13 * m
        assert np.abs(ops.ddz_C2E(fC) - 1j * m * fE).max() < 1e-13 * m
        assert np.abs(ops.interpz_E2C(fE) - fC).max() < 1e-13
        assert np.abs(ops.interpz_C2E(fC) - fE).max() < 1e-13
        assert np.abs(ops.d2dz2_C2C(fC) + m * m * fC).max() < 1e-12 * m * m
        assert np.abs(ops.d2dz2_E2E(fE) + m * m * fE).max() < 1e-12 * m * m
    k = np.array([0.0, 1.0, -3.0])
    assert np.array_equal(ops.getModifiedWavenumbers(k), k)      # PadeDerOps.F90:1003-1004


@pytest.mark.parametrize("adv", [0, 1])
def test_taylor_green_decay_fourier_z(adv):
    """NumericalSchemeVert = 2: the x-z Taylor-Green vortex decays at the exact rate to rounding-level accuracy (no z
    truncation error any more), with either form of the advection term."""
    n = 16
    Re = 100.0
    x, y, zC, zE = _grid(n, n, n)
    X, ZC, ZE = x[None, None, :], zC[:, None, None], zE[:, None, None]
    u = np.sin(X) * np.cos(ZC) * np.ones((1, n, 1)); v = np.zeros((n, n, n)); w = -np.cos(X) * np.sin(ZE) * np.ones((1, n, 1))
    g = IG.IGrid(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, Re, u, v, w, TimeSteppingScheme=1, AdvectionTerm=adv, NumericalSchemeVert=2)
    dt = 0.25 * (2 * np.pi / n)
    for _ in range(4):
        g.timeAdvance(dt)
    decay = np.exp(-2.0 * g.tsim / Re)
    assert np.abs(g.u - u * decay).max() < 1e-9 and np.abs(g.w - w * decay).max() < 1e-9 and np.abs(g.v).max() < 1e-12
    assert np.abs(g.poiss.divergence(g.uhat, g.vhat, g.what)).max() < 1e-12


def test_stokes_pressure_projection():
    """computeStokesPressure = .true. (PadePoisson.F90:232-296, 320-384, 444-458, 597-609): w* need not vanish on the walls — the
    harmonic pressure removes the wall-normal velocity first, bottom then top; the projected field is solenoidal to rounding,
    w = 0 on both walls, the projection is idempotent, and it reduces to the plain wall projection when w* is already zero there."""
    nx, ny, nz, Lz = 16, 12, 24, 2.0
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, Lz / nz
    spC, spE = IG.Spectral(nx, ny, nz, dx, dy, dz), IG.Spectral(nx, ny, nz + 1, dx, dy, dz)
    ops = IG.Pade6stagg(nz, dz, 1, isPeriodic=False)
    P = IG.PadePoisson(dx, dy, dz, spC, spE, ops, PeriodicInZ=False, computeStokesPressure=True, Lz=Lz)
    P0 = IG.PadePoisson(dx, dy, dz, spC, spE, ops, PeriodicInZ=False)
    rng = np.random.default_rng(0)
    u, v, w = rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz + 1, ny, nx))
    uh, vh, wh = spC.dealias(spC.fft(u)), spC.dealias(spC.fft(v)), spE.dealias(spE.fft(w))   # the oddball mode is never alive in a run
    scale = np.abs(P.divergence(uh, vh, wh)).max()
    a, b, c = P.PressureProjection(uh, vh, wh)
    assert np.abs(P.divergence(a, b, c)).max() < 1e-12 * scale
    assert not np.any(c[0]) and not np.any(c[nz])
    a0, b0, c0 = P0.PressureProjection(uh, vh, wh)
    assert np.abs(P0.divergence(a0, b0, c0)).max() > 1e-3 * scale          # without it the wall values are simply cut off
    a2, b2, c2 = P.PressureProjection(a, b, c)
    assert np.abs(a2 - a).max() < 1e-12 * np.abs(a).max() and np.abs(c2 - c).max() < 1e-12 * np.abs(c).max()
    wh0 = wh.copy()
    wh0[0] = 0
    wh0[nz] = 0
    for x, y in zip(P.PressureProjection(uh, vh, wh0), P0.PressureProjection(uh, vh, wh0)):
        assert np.abs(x - y).max() < 1e-12 * np.abs(y).max()


@pytest.mark.parametrize("shape", [(16, 12, 24), (12, 16, 10)])
def test_wall_bounded_projection(shape):
    """padepoisson with PeriodicInZ = .false., computeStokesPressure = .false. (PadePoisson.F90:180-230, 459-623): after the
    projection the field is discretely divergence-free for the odd / odd wall operator DivergenceCheck uses (:1188), w vanishes on
    both walls, the projection is idempotent, and a field that is already solenoidal with w = 0 on the walls passes unchanged."""
    nx, ny, nz = shape
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 1.0 / nz
    spC, spE = IG.Spectral(nx, ny, nz, dx, dy, dz), IG.Spectral(nx, ny, nz + 1, dx, dy, dz)
    ops = IG.Pade6stagg(nz, dz, 1, isPeriodic=False)
    P = IG.PadePoisson(dx, dy, dz, spC, spE, ops, PeriodicInZ=False)
    rng = np.random.default_rng(nz)
    u, v = rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz, ny, nx))
    w = rng.standard_normal((nz + 1, ny, nx))
    w[0] = 0.0
    w[nz] = 0.0        # no penetration: the odd extension of w is continuous only then
    uh, vh, wh = spC.fft(u), spC.fft(v), spE.fft(w)
    scale = np.abs(P.divergence(uh, vh, wh)).max()
    u2, v2, w2 = P.PressureProjection(uh, vh, wh)
    assert np.abs(P.divergence(u2, v2, w2)).max() < 1e-12 * scale
    assert not np.any(w2[0]) and not np.any(w2[nz])
    u3, v3, w3 = P.PressureProjection(u2, v2, w2)
    for a, b in ((u3, u2), (v3, v2), (w3, w2)):
        assert np.abs(a - b).max() < 1e-12 * np.abs(b).max()
    _, _, _, div = P.DivergenceCheck(uh, vh, wh, fixDiv=True)
    assert np.abs(div).max() < 1e-12 * scale
    # the mean horizontal flow (kx = ky = 0) is untouched: kradsq_inv = 0 there only for k3 = 0, and its divergence is zero anyway
    assert np.abs(u2[:, 0, 0] - uh[:, 0, 0]).max() < 1e-12 * np.abs(uh[:, 0, 0]).max()


def _channel_fields(nx, ny, nz, Lz, slip=True):
    x = np.arange(nx) * 2 * np.pi / nx
    y = np.arange(ny) * 2 * np.pi / ny
    zc = (np.arange(nz) + 0.5) * Lz / nz
    ze = np.arange(nz + 1) * Lz / nz
    X, Y = x[None, None, :], y[None, :, None]
    a = np.pi / Lz
    if slip:      # u, v even about both walls, w odd
        u = np.sin(X) * np.cos(Y) * np.cos(a * zc)[:, None, None] + 0.3 * np.cos(2 * Y) * np.cos(2 * a * zc)[:, None, None] + 0 * X
        v = -np.cos(X) * np.sin(Y) * np.cos(a * zc)[:, None, None] + 0.2 * np.sin(2 * X) * np.cos(3 * a * zc)[:, None, None] + 0 * Y
        w = 0.25 * np.sin(X) * np.sin(2 * Y) * np.sin(2 * a * ze)[:, None, None]
    else:         # u, v vanish on the walls, w and dw/dz too
        u = np.sin(X) * np.cos(Y) * np.sin(a * zc)[:, None, None] + 0.3 * np.cos(2 * Y) * np.sin(2 * a * zc)[:, None, None] + 0 * X
        v = -np.cos(X) * np.sin(Y) * np.sin(a * zc)[:, None, None] + 0 * X
        w = 0.25 * np.sin(X) * np.sin(2 * Y) * (np.sin(a * ze) ** 2)[:, None, None]
    return u, v, w


@pytest.mark.parametrize("adv", [1, 0])
def test_slip_wall_channel_equals_the_periodic_run_on_the_doubled_box(adv):
    """PeriodicInZ = .false. with slip walls (topWall = botWall = 2): u, v even and w odd about both walls, so the wall-bounded run
    on [0, Lz] is the lower half of a periodic run on [0, 2 Lz) started from the extended fields — the wall closures of the
    staggered operators, the stencil codes of get_boundary_conditions_stencil and the even / odd Poisson solver all have to be
    right for that.  (The periodic run's z-dealiasing is switched off: the wall-bounded code dealiases in x and y only.)"""
    nx, ny, nz, Lz = 16, 12, 16, np.pi
    L = 2 * np.pi
    u, v, w = _channel_fields(nx, ny, nz, Lz)
    g = IG.IGrid(nx, ny, nz, L, L, Lz, 200.0, u, v, w, TimeSteppingScheme=1, AdvectionTerm=adv, PeriodicInZ=False, topWall=2, botWall=2)
    ue = np.concatenate([u, u[::-1]], axis=0)
    ve = np.concatenate([v, v[::-1]], axis=0)
    we = np.concatenate([w[:nz], -w[nz:0:-1], w[:1]], axis=0)
    p = IG.IGrid.__new__(IG.IGrid)
    # build the periodic twin, then widen its dealiasing mask to the 2-D one before any z-dealiasing can act on products
    p.__init__(nx, ny, 2 * nz, L, L, 2 * Lz, 200.0, ue, ve, we, TimeSteppingScheme=1, AdvectionTerm=adv)
    p.spectC.Gdealias = np.broadcast_to(p.spectE.Gdealias[:1], p.spectC.Gdealias.shape).copy()
    for it in range(2):
        g.timeAdvance(0.01)
        p.timeAdvance(0.01)
        # the rotational form interpolates its edge products with the ONE-SIDED closures (interpz_E2C(.., 0, 0), igrid.F90:1534, 1546):
        # there the two runs agree to the truncation error of those rows only
        tol = 2e-10 if adv == 1 else 5e-3
        assert np.abs(g.u - p.u[:nz]).max() < tol * np.abs(p.u).max(), it
        assert np.abs(g.v - p.v[:nz]).max() < tol * np.abs(p.v).max(), it
        assert np.abs(g.w - p.w[:nz + 1]).max() < tol * np.abs(p.u).max(), it
    assert not np.any(g.what[0]) and not np.any(g.what[nz])


@pytest.mark.parametrize("walls", [(1, 1), (1, 2)])
def test_no_slip_channel_stays_solenoidal_and_decays(walls):
    nx, ny, nz, Lz = 16, 12, 24, 2.0
    L = 2 * np.pi
    u, v, w = _channel_fields(nx, ny, nz, Lz, slip=False)
    g = IG.IGrid(nx, ny, nz, L, L, Lz, 50.0, u, v, w, TimeSteppingScheme=2, PeriodicInZ=False, botWall=walls[0], topWall=walls[1])
    assert g.bc["u"] == (-1, -1 if walls[1] == 1 else 1) and g.bc["w"] == (1, 1 if walls[1] == 1 else -1) and g.bc["dWdz"][0] == -1
    e0 = (g.u ** 2 + g.v ** 2 + g.wC ** 2).mean()
    for _ in range(3):
        g.timeAdvance(0.005)
    # the right-hand side leaves w* nonzero on a no-slip wall; the Stokes-pressure step of the projection (ComputeStokesPressure, the
    # reference's default) removes it harmonically, so the projected field is solenoidal to rounding; without it only approximately
    _, _, _, div = g.poiss.DivergenceCheck(g.uhat, g.vhat, g.what)
    assert np.abs(div).max() < 1e-11 * np.abs(g.duidxj["dudx"]).max()
    assert not np.any(g.what[0]) and not np.any(g.what[nz])
    h = IG.IGrid(nx, ny, nz, L, L, Lz, 50.0, u, v, w, TimeSteppingScheme=2, PeriodicInZ=False, botWall=walls[0], topWall=walls[1],
                 ComputeStokesPressure=False)
    for _ in range(3):
        h.timeAdvance(0.005)
    _, _, _, div0 = h.poiss.DivergenceCheck(h.uhat, h.vhat, h.what)
    assert 1e-9 < np.abs(div0).max() < 1e-3 * np.abs(h.duidxj["dudx"]).max()
    e1 = (g.u ** 2 + g.v ** 2 + g.wC ** 2).mean()
    assert 0.5 * e0 < e1 < e0
    with pytest.raises(ValueError):
        IG.IGrid(nx, ny, nz, L, L, Lz, 50.0, u, v, w, PeriodicInZ=False, botWall=3)


@pytest.mark.parametrize("shape", [(16, 12, 16), (12, 16, 10)])
def test_wall_pressure_getters_are_consistent_with_the_projection(shape):
    """getPressure / getPressureAndUpdateRHS with walls (PadePoisson.F90:762-896, 963-1160).  Pins: (i) without the Stokes pressure
    the projected right-hand sides ARE u - i k1 p, v - i k2 p for the pressure getPressure returns (every mode the c2r keeps);
    (ii) getPressureAndUpdateRHS updates exactly as PressureProjection does; (iii) with computeStokesPressure the returned pressure
    is phat + phat_z1 + phat_z2 with the pieces in the form the reference stores them (i chat cosh: its velocity correction is
    u - k1 (phat_z1 + phat_z2), without another i — the pressure output keeps that factor, as the reference's does), and
    getPressureAndUpdateRHS adds the pieces of the LAST getPressure call (:1146-1156)."""
    nx, ny, nz = shape
    L = 2 * np.pi
    d = [L / nx, L / ny, L / nz]
    spC, spE = IG.Spectral(nx, ny, nz, *d), IG.Spectral(nx, ny, nz + 1, *d)
    ops = IG.Pade6stagg(nz, d[2], scheme=1, isPeriodic=False)
    rng = np.random.default_rng(nx)
    u, v, w = rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz + 1, ny, nx))
    uh, vh, wh = spC.fft(u), spC.fft(v), spE.fft(w)
    keep = np.ones((ny, nx // 2 + 1), bool)
    keep[ny // 2, :] = False
    keep[:, nx // 2] = False            # the oddball modes: ifft drops their imaginary parts
    P = IG.PadePoisson(*d, spC, spE, ops, PeriodicInZ=False)
    pr = P.getPressure(uh, vh, wh)
    un, vn, wn = P.PressureProjection(uh, vh, wh)
    ph = spC.fft(pr)
    assert np.abs((un - (uh - 1j * spC.k1 * ph))[:, keep]).max() < 1e-13 * np.abs(un).max()
    assert np.abs((vn - (vh - 1j * spC.k2 * ph))[:, keep]).max() < 1e-13 * np.abs(vn).max()
    u2, v2, w2, p2 = P.getPressureAndUpdateRHS(uh, vh, wh)
    assert np.array_equal(u2, un) and np.array_equal(v2, vn) and np.array_equal(w2, wn) and np.array_equal(p2, pr)
    S = IG.PadePoisson(*d, spC, spE, ops, PeriodicInZ=False, computeStokesPressure=True, Lz=L)
    u3, v3, w3, p3 = S.getPressureAndUpdateRHS(uh, vh, wh)          # no getPressure yet: no Stokes pieces in the pressure
    us, vs, ws = S.PressureProjection(uh, vh, wh)
    assert np.array_equal(u3, us) and np.array_equal(w3, ws)
    prs = S.getPressure(uh, vh, wh)
    pieces = S.phat_z1 + S.phat_z2
    assert np.abs(prs - p3 - spC.ifft(pieces)).max() < 1e-12 * np.abs(prs).max()
    f2d = spC.fft(p3)
    assert np.abs((us - (uh - spC.k1 * pieces - 1j * spC.k1 * f2d))[:, keep]).max() < 1e-12 * np.abs(us).max()
    assert np.array_equal(S.getPressureAndUpdateRHS(uh, vh, wh)[3], prs)


@pytest.mark.parametrize("scheme", [1, 2])
@pytest.mark.parametrize("shape", [(16, 12, 16), (8, 8, 32)])
def test_dealias_and_projection_in_kz_space_equal_the_reference_sequence(scheme, shape):
    """What csrc/ig_padepoisson.inc.cuh: poiss_dealias_project_kz does, in numpy: the periodic z-operators are circulant, so with their
    symbols (the transform of the response to a unit pulse — whatever the scheme) dealias (spectral.F90:343-363) followed by
    PeriodicProjection (PadePoisson.F90:386-432) is ONE pointwise pass between a forward and a backward z transform per field.  It
    must reproduce the statement-by-statement sequence to rounding."""
    nx, ny, nz = shape
    L = 2 * np.pi
    d = [L / nx, L / ny, L / nz]
    spC, spE = IG.Spectral(nx, ny, nz, *d, init_periodicInZ=True), IG.Spectral(nx, ny, nz + 1, *d)
    ops = IG.Pade6stagg(nz, d[2], scheme=scheme)
    P = IG.PadePoisson(*d, spC, spE, ops)
    rng = np.random.default_rng(nz)
    u, v, w = rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz + 1, ny, nx))
    w[nz] = w[0]
    uh, vh, wh = spC.fft(u), spC.fft(v), spE.fft(w)
    want = P.PressureProjection(spC.dealias(uh), spC.dealias(vh), spC.dealias_edgeField(wh))
    pulseE = np.zeros((nz + 1, 1, 1), complex)
    pulseE[0] = pulseE[nz] = 1.0
    pulseC = np.zeros((nz, 1, 1), complex)
    pulseC[0] = 1.0
    dE2C = np.fft.fft(ops.ddz_E2C(pulseE)[:, 0, 0])[:, None, None]
    dC2E = np.fft.fft(ops.ddz_C2E(pulseC)[:nz, 0, 0])[:, None, None]
    U, V, W = (spC.Gdealias * np.fft.fft(a, axis=0) for a in (uh, vh, wh[:nz]))
    f = dE2C * W + 1j * spC.k1 * U + 1j * spC.k2 * V
    p = -P.kradsq_inv * f
    got_u = np.fft.ifft(U - 1j * spC.k1 * p, axis=0)
    got_v = np.fft.ifft(V - 1j * spC.k2 * p, axis=0)
    got_w = np.fft.ifft(W - dC2E * p, axis=0)
    got_w = np.concatenate([got_w, got_w[:1]], axis=0)
    for got, ref in zip((got_u, got_v, got_w), want):
        assert np.abs(got - ref).max() < 1e-13 * np.abs(ref).max()


def test_wall_projection_kernels_index_arithmetic():
    """csrc/igrid.cu poiss_wall_projection re-enacted in numpy with the kernels' own flat-index expressions (extension, fused
    solve + project with the complex products written out in components, extraction) against the oracle's array formulation."""
    nx, ny, nz = 8, 6, 10
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 1.0 / nz
    spC, spE = IG.Spectral(nx, ny, nz, dx, dy, dz), IG.Spectral(nx, ny, nz + 1, dx, dy, dz)
    ops = IG.Pade6stagg(nz, dz, 1, isPeriodic=False)
    P = IG.PadePoisson(dx, dy, dz, spC, spE, ops, PeriodicInZ=False)
    rng = np.random.default_rng(1)
    nxh = nx // 2 + 1
    uh = rng.standard_normal((nz, ny, nxh)) + 1j * rng.standard_normal((nz, ny, nxh))
    vh = rng.standard_normal((nz, ny, nxh)) + 1j * rng.standard_normal((nz, ny, nxh))
    wh = rng.standard_normal((nz + 1, ny, nxh)) + 1j * rng.standard_normal((nz + 1, ny, nxh))
    wh[0] = 0
    wh[nz] = 0
    want = P.PressureProjection(uh, vh, wh)
    # --- kernels ---
    n1, n2 = nxh, ny
    cols = n1 * n2
    f2dy = spC.k1 * uh + spC.k2 * vh
    uz = (-f2dy.imag + 1j * f2dy.real).reshape(-1)             # poiss_div_xy; flat index = c + cols * k
    wz = wh.reshape(-1)
    fe = np.zeros(cols * 2 * nz, dtype=complex)
    we = np.zeros(cols * 2 * nz, dtype=complex)
    for i in range(cols * 2 * nz):
        c, kk = i % cols, i // cols
        fe[i] = uz[c + cols * (nz - 1 - kk)] if kk < nz else uz[c + cols * (kk - nz)]
        we[i] = -wz[c + cols * (nz - 1 - kk)] if kk < nz - 1 else wz[c + cols * (kk - (nz - 1))]
    fe = np.fft.fft(fe.reshape(2 * nz, cols), axis=0).reshape(-1)
    we = np.fft.fft(we.reshape(2 * nz, cols), axis=0).reshape(-1)
    k1sq = (IG.O.wavenums(nx, dx)[:nxh]) ** 2
    k2sq = IG.O.wavenums(ny, dy) ** 2
    k3e = IG.O.wavenums(2 * nz, dz)
    k3m = ops.getModifiedWavenumbers(k3e)
    cm = np.stack([k3m * np.cos((dz / 2) * k3e), -k3m * np.sin((dz / 2) * k3e)], axis=1)
    cp = np.stack([k3m * np.cos((dz / 2) * k3e), k3m * np.sin((dz / 2) * k3e)], axis=1)
    mfact = 1.0 / (2 * nz)
    for i in range(cols * 2 * nz):
        ii = i % n1
        t = i // n1
        jj, kk = t % n2, t // n2
        kradsq = k1sq[ii] + k2sq[jj] + k3m[kk] ** 2
        kinv = 0.0 if kradsq <= 1e-14 else 1.0 / kradsq
        fx, fy, wx, wy = fe[i].real, fe[i].imag, we[i].real, we[i].imag
        ax, ay, bx, by = cm[kk, 0], cm[kk, 1], cp[kk, 0], cp[kk, 1]
        fx += -(ax * wy + ay * wx)
        fy += ax * wx - ay * wy
        fx, fy = -fx * kinv, -fy * kinv
        wx -= -(bx * fy + by * fx)
        wy -= bx * fx - by * fy
        fe[i] = (fx + 1j * fy) * mfact
        we[i] = (wx + 1j * wy) * mfact
    fe = (np.fft.ifft(fe.reshape(2 * nz, cols), axis=0) * (2 * nz)).reshape(-1)
    we = (np.fft.ifft(we.reshape(2 * nz, cols), axis=0) * (2 * nz)).reshape(-1)
    f2d = np.zeros(cols * nz, dtype=complex)
    w2 = np.zeros(cols * (nz + 1), dtype=complex)
    for i in range(cols * (nz + 1)):
        kk = i // cols
        if kk < nz:
            f2d[i] = fe[i + cols * nz]
        w2[i] = 0.0 if kk in (0, nz) else we[i + cols * (nz - 1)]
    f2d = f2d.reshape(nz, ny, nxh)
    got = (uh - 1j * spC.k1 * f2d, vh - 1j * spC.k2 * f2d, w2.reshape(nz + 1, ny, nxh))
    for a, b in zip(got, want):
        assert np.abs(a - b).max() < 1e-12 * np.abs(b).max()


def test_stokes_kernel_arithmetic():
    """csrc/ig_padepoisson.inc.cuh stokes_kernel re-enacted in numpy column by column (cosh / sinh on the fly with the reference's
    clipping, bottom wall first, top wall from the corrected top plane) against the oracle's table formulation; large lambda Lz
    included so that the clipped branches are exercised."""
    nx, ny, nz, Lz = 16, 12, 10, 6.0
    dx, dy, dz = 2 * np.pi / nx / 8, 2 * np.pi / ny / 8, Lz / nz          # small box in x, y: lambda up to ~90, lambda Lz >> 32
    spC, spE = IG.Spectral(nx, ny, nz, dx, dy, dz), IG.Spectral(nx, ny, nz + 1, dx, dy, dz)
    P = IG.PadePoisson(dx, dy, dz, spC, spE, IG.Pade6stagg(nz, dz, 1, isPeriodic=False), PeriodicInZ=False, computeStokesPressure=True, Lz=Lz)
    rng = np.random.default_rng(3)
    nxh = nx // 2 + 1
    u = rng.standard_normal((nz, ny, nxh)) + 1j * rng.standard_normal((nz, ny, nxh))
    v = rng.standard_normal((nz, ny, nxh)) + 1j * rng.standard_normal((nz, ny, nxh))
    w = rng.standard_normal((nz + 1, ny, nxh)) + 1j * rng.standard_normal((nz + 1, ny, nxh))
    wu, wv, ww = P.ProjectStokesPressure(u, v, w)
    k1 = IG.O.wavenums(nx, dx)[:nxh]
    k2 = IG.O.wavenums(ny, dy)
    gu, gv, gw = u.copy(), v.copy(), w.copy()
    dzl = Lz / nz
    for jj in range(ny):
        for ii in range(nxh):
            lam = np.sqrt(k1[ii] ** 2 + k2[jj] ** 2)
            den = 1.0 / (lam * np.sinh(lam * Lz) + 1e-13) if lam * Lz < 500.0 else 0.0
            den = 0.0 if den < 1e-16 else den
            if ii == 0 and jj == 0:
                den = 0.0
            ch = -gw[0, jj, ii] * den
            for k in range(nz):
                zc = 0.5 * (k * dzl + (k + 1) * dzl)
                t = lam * (Lz - zc)
                cb = np.cosh(t) if t < 32.0 else 4.0e13
                ph = 1j * ch * cb
                gu[k, jj, ii] -= k1[ii] * ph
                gv[k, jj, ii] -= k2[jj] * ph
            gw[0, jj, ii] = 0.0
            for k in range(1, nz + 1):
                t = lam * (Lz - k * dzl)
                sb = -lam * np.sinh(t) if t < 32.0 else -4.0e13
                gw[k, jj, ii] -= ch * sb
            ch = gw[nz, jj, ii] * den
            for k in range(nz):
                zc = 0.5 * (k * dzl + (k + 1) * dzl)
                t = lam * zc
                ct = np.cosh(t) if t < 32.0 else 1.0e13
                ph = 1j * ch * ct
                gu[k, jj, ii] -= k1[ii] * ph
                gv[k, jj, ii] -= k2[jj] * ph
                t = lam * (k * dzl)
                stp = lam * np.sinh(t) if t < 32.0 else 4.0e13
                gw[k, jj, ii] -= ch * stp
            gw[nz, jj, ii] = 0.0
    for a, b in ((gu, wu), (gv, wv), (gw, ww)):
        assert np.abs(a - b).max() < 1e-12 * max(np.abs(b).max(), 1.0)


@pytest.mark.parametrize("bot,top", [(1, 1), (1, 2), (2, 1), (2, 2)])
def test_product_stencil_codes_match_the_oracle(bot, top):
    """The PRODUCT's get_boundary_conditions_stencil table (host-only hook) against the oracle's restatement of igrid.F90:5148-5204."""
    import ctypes as C
    import padeops_b200 as pdo
    out = (C.c_int * 24)()
    L = pdo.lib()
    L.pdo_debug_igrid_bcs.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    assert L.pdo_debug_igrid_bcs(bot, top, out) == 0
    ref = IG.IGrid.get_boundary_conditions_stencil(top, bot)
    order = ("w", "u", "v", "WdUdz", "WdVdz", "WdWdz", "WW", "UW", "VW", "dUdz", "dVdz", "dWdz")
    for q, name in enumerate(order):
        assert (out[2 * q], out[2 * q + 1]) == ref[name], name
