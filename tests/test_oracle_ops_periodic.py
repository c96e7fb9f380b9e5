"""Pins the oracle's restatement of Ops_Periodic (igrid_operators_periodic.F90:13-161), of the spectral type's z-Fourier
procedures it uses (spectral.F90:409-437, 507-547) and of the REAL fourierColl procedures of Pade6stagg
(spectral.F90:365-385, 439-459, 484-505, 572-593, 639-702) with analytic fields: Fourier differentiation is exact on every
resolved mode, the real procedures leave the oddball (Nyquist) mode untouched, the Poisson solve inverts the Laplacian."""
import numpy as np
import pytest

from oracle import igrid_oracle as IG
from oracle import ops_periodic_oracle as OP


def _grid(nx, ny, nz):
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    x, y, z = np.arange(nx) * dx, np.arange(ny) * dy, np.arange(nz) * dz
    return (dx, dy, dz), x[None, None, :], y[None, :, None], z[:, None, None]


def test_ops_periodic_derivatives_are_exact_on_resolved_modes():
    nx, ny, nz = 24, 16, 20
    d, X, Y, Z = _grid(nx, ny, nz)
    op = OP.OpsPeriodic(nx, ny, nz, *d)
    f = np.sin(3 * X + 0.2) * np.cos(5 * Y - 0.4) * np.sin(7 * Z + 0.9) + 0.5 * np.cos(11 * X) + 0.25
    assert np.abs(op.ddx(f) - (3 * np.cos(3 * X + 0.2) * np.cos(5 * Y - 0.4) * np.sin(7 * Z + 0.9) - 5.5 * np.sin(11 * X))).max() < 1e-12
    assert np.abs(op.ddy(f) + 5 * np.sin(3 * X + 0.2) * np.sin(5 * Y - 0.4) * np.sin(7 * Z + 0.9)).max() < 1e-12
    assert np.abs(op.ddz(f) - 7 * np.sin(3 * X + 0.2) * np.cos(5 * Y - 0.4) * np.cos(7 * Z + 0.9)).max() < 1e-12


def test_real_z_derivative_passes_the_oddball_mode_through():
    """spectral.F90:514 "Note that the oddball is ignored": the loop stops at nz/2, the Nyquist coefficient goes back through
    the c2r transform as it came — cos(nz/2 z) = (-1)^k comes out UNCHANGED, not differentiated and not zeroed."""
    nx, ny, nz = 8, 8, 12
    d, X, Y, Z = _grid(nx, ny, nz)
    op = OP.OpsPeriodic(nx, ny, nz, *d)
    odd = np.cos((nz // 2) * Z) + 0 * X + 0 * Y
    assert np.abs(op.ddz(odd) - odd).max() < 1e-13
    f = np.sin(2 * Z) * np.cos(X) + 0 * Y
    assert np.abs(op.ddz(f + 0.3 * odd) - (2 * np.cos(2 * Z) * np.cos(X) + 0.3 * odd)).max() < 1e-13


def test_complex_z_derivative_and_shifts():
    nx, ny, nz = 12, 10, 16
    d, X, Y, Z = _grid(nx, ny, nz)
    op = OP.OpsPeriodic(nx, ny, nz, *d)
    sp = op.spect
    f = np.cos(2 * X) * np.sin(3 * Y) * np.sin(5 * Z + 0.3)
    fhat = sp.fft(f)
    ref = sp.fft(np.cos(2 * X) * np.sin(3 * Y) * 5 * np.cos(5 * Z + 0.3))
    assert np.abs(op.ddz_cmplx2cmplx(fhat) - ref).max() < 1e-12 * np.abs(ref).max()
    # the complex procedure multiplies the oddball by i k3 like any other mode (spectral.F90:535)
    odd = sp.fft(np.cos((nz // 2) * Z) + 0 * X + 0 * Y)
    k_odd = -(nz // 2)     # GetWaveNums puts the oddball at -nz/2 (no sign flip for the tables, spectral.F90:845)
    assert np.abs(op.ddz_cmplx2cmplx(odd) - 1j * k_odd * odd).max() < 1e-12 * np.abs(odd).max()
    # shiftz: cells -> edges is a shift by -dz/2 applied in z-Fourier space; E2C undoes C2E
    a = sp.take_fft1d_z2z(fhat)
    b = sp.shiftz_C2E(a)
    fe = sp.ifft(sp.take_ifft1d_z2z(b))
    assert np.abs(fe - np.cos(2 * X) * np.sin(3 * Y) * np.sin(5 * (Z - d[2] / 2) + 0.3)).max() < 1e-12
    assert np.abs(sp.shiftz_E2C(b) - a).max() < 1e-12 * np.abs(a).max()


def test_poisson_and_dealias():
    nx, ny, nz = 32, 16, 12
    d, X, Y, Z = _grid(nx, ny, nz)
    op = OP.OpsPeriodic(nx, ny, nz, *d)
    ftrue = np.sin(6 * X) * np.cos(3 * Y) * np.sin(1 * Z)       # tests/test_PoissonPeriodic.F90:109-111
    rhs = -(36 + 9 + 1) * ftrue
    assert np.abs(op.SolvePoisson(rhs) - ftrue).max() < 1e-13
    # dealiasField: a mode inside the 2/3 box survives, one outside in ANY direction is removed (3-D mask, `>=` cut)
    keep = np.cos(3 * X) * np.cos(2 * Y) * np.cos(2 * Z)
    for kill in (np.cos(12 * X) * np.cos(2 * Y) + 0 * Z, np.cos(3 * X) * np.sin(6 * Y) + 0 * Z, np.cos(X) * np.cos(4 * Z) + 0 * Y):
        assert np.abs(op.dealiasField(keep + kill) - keep).max() < 1e-13
    # the cut is `>=`: |k| = (2/3)(n/2) exactly is removed (nz = 12: k3 = 4), one below survives (k3 = 3)
    edge = np.cos(3 * Z) + 0 * X + 0 * Y
    assert np.abs(op.dealiasField(edge) - edge).max() < 1e-13


@pytest.mark.parametrize("nz", [8, 16])
def test_real_fourier_collocation_operators(nz):
    """The REAL twins of the six fourierColl operators: equal to the complex ones on a field without oddball content, exact on
    resolved modes, and the oddball mode passes through every one of them untouched."""
    ny, nx = 6, 10
    dz = 2 * np.pi / nz
    ops = IG.Pade6stagg(nz, dz, scheme=2)
    zc = (np.arange(nz) + 0.5) * dz
    ze = np.arange(nz + 1) * dz
    amp = np.random.default_rng(1).standard_normal((1, ny, nx))
    k = nz // 2 - 1
    fC, fE = np.sin(k * zc + 0.4)[:, None, None] * amp, np.sin(k * ze + 0.4)[:, None, None] * amp
    dC, dE = k * np.cos(k * zc + 0.4)[:, None, None] * amp, k * np.cos(k * ze + 0.4)[:, None, None] * amp
    tol = 1e-12 * k * k
    assert np.abs(ops.ddz_E2C(fE) - dC).max() < tol and ops.ddz_E2C(fE).dtype == np.float64
    assert np.abs(ops.ddz_C2E(fC) - dE).max() < tol and ops.ddz_C2E(fC).shape == (nz + 1, ny, nx)
    assert np.abs(ops.interpz_E2C(fE) - fC).max() < tol
    assert np.abs(ops.interpz_C2E(fC) - fE).max() < tol
    assert np.abs(ops.d2dz2_C2C(fC) + k * k * fC).max() < tol
    assert np.abs(ops.d2dz2_E2E(fE) + k * k * fE).max() < tol
    for name, fin in (("ddz_E2C", fE), ("ddz_C2E", fC), ("interpz_E2C", fE), ("interpz_C2E", fC), ("d2dz2_C2C", fC), ("d2dz2_E2E", fE)):
        re = getattr(ops, name)(fin)
        cx = getattr(ops, name)(fin.astype(np.complex128))
        assert np.abs(cx.imag).max() < tol and np.abs(re - cx.real).max() < tol, name
    odd = ((-1.0) ** np.arange(nz))[:, None, None] * amp
    oddE = np.concatenate([odd, odd[:1]])
    for name, fin, edge_out in (("ddz_E2C", oddE, False), ("ddz_C2E", odd, True), ("interpz_E2C", oddE, False), ("interpz_C2E", odd, True),
                                ("d2dz2_C2C", odd, False), ("d2dz2_E2E", oddE, True)):
        out = getattr(ops, name)(fin)
        assert np.abs(out - (oddE if edge_out else odd)).max() < 1e-13, name
