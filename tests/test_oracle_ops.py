"""CPU: pin the oracle against the reference's own known-answer tests and against
independent dense cyclic solves.  The reference ships no golden vectors; these are its tests' oracles."""
import numpy as np
import pytest

from conftest import broadband


def _grid(n):
    d = 2 * np.pi / n
    x = np.arange(n) * d
    return d, x[None, None, :], x[None, :, None], x[:, None, None]


def test_cd10_analytic_32cubed(oracle):
    # tests/test_cd10.F90:51-81 — f = sin x + sin y + sin z on 32^3, d1 and d2 along every axis
    n = 32
    d, X, Y, Z = _grid(n)
    f = np.sin(X) + np.sin(Y) + np.sin(Z)
    for ax, c in enumerate((X, Y, Z)):
        assert np.abs(oracle.cd10(f, d, ax, 1) - np.cos(c)).max() < 1e-12
        assert np.abs(oracle.cd10(f, d, ax, 2) + np.sin(c)).max() < 1e-12


def test_cd06_analytic_256(oracle):
    # tests/test_cd06.F90:24-52 — 256 points per direction (thin slab here), 6th-order accuracy
    n = 256
    d = 2 * np.pi / n
    x = np.arange(n) * d
    f = np.broadcast_to(np.sin(x)[None, None, :], (2, 3, n)).copy()
    assert np.abs(oracle.cd06(f, d, 0) - np.cos(x)[None, None, :]).max() < 1e-12
    f = np.broadcast_to(np.sin(x)[None, :, None], (2, n, 3)).copy()
    assert np.abs(oracle.cd06(f, d, 1) - np.cos(x)[None, :, None]).max() < 1e-12
    f = np.broadcast_to(np.sin(x)[:, None, None], (n, 2, 3)).copy()
    assert np.abs(oracle.cd06(f, d, 2) - np.cos(x)[:, None, None]).max() < 1e-12


def _modified_wavenumber_check(op, n, k, symbol, is_deriv):
    """Single Fourier mode: every periodic compact operator returns T(k d) f or i k'(k d) f."""
    d = 2 * np.pi / n
    x = np.arange(n) * d
    w = k * d
    for ax in range(3):
        shape = [3, 3, 3]
        shape[2 - ax] = n
        idx = [None, None, None]
        idx[2 - ax] = slice(None)
        fc = np.broadcast_to(np.cos(k * x)[tuple(idx)], shape).copy()
        fs = np.broadcast_to(np.sin(k * x)[tuple(idx)], shape).copy()
        oc, os_ = op(fc, d, ax), op(fs, d, ax)
        if is_deriv == 1:  # d/dx cos = -k' sin
            kp = symbol(w) / d
            assert np.abs(oc + kp * fs).max() < 2e-12 * max(1.0, abs(kp))
            assert np.abs(os_ - kp * fc).max() < 2e-12 * max(1.0, abs(kp))
        else:
            T = symbol(w) / (d * d if is_deriv == 2 else 1.0)
            assert np.abs(oc - T * fc).max() < 2e-12 * max(1.0, abs(T))
            assert np.abs(os_ - T * fs).max() < 2e-12 * max(1.0, abs(T))


@pytest.mark.parametrize("n,k", [(32, 3), (64, 17), (40, 11)])
def test_modified_wavenumbers_and_transfer_functions(oracle, n, k):
    # CD10 first derivative: k'd = (2a sin w + 2b sin 2w + 2c sin 3w)/(1 + 2al cos w + 2be cos 2w), Lele (1992)
    a, b, c = (17 / 12) / 2, (101 / 150) / 4, (1 / 100) / 6
    _modified_wavenumber_check(lambda f, d, ax: oracle.cd10(f, d, ax, 1), n, k,
                               lambda w: (2 * a * np.sin(w) + 2 * b * np.sin(2 * w) + 2 * c * np.sin(3 * w)) /
                               (1 + 2 * 0.5 * np.cos(w) + 2 * 0.05 * np.cos(2 * w)), 1)
    a2, b2, c2 = 1065 / 1798, (1038 / 899) / 4, (79 / 1798) / 9
    _modified_wavenumber_check(lambda f, d, ax: oracle.cd10(f, d, ax, 2), n, k,
                               lambda w: (2 * a2 * (np.cos(w) - 1) + 2 * b2 * (np.cos(2 * w) - 1) + 2 * c2 * (np.cos(3 * w) - 1)) /
                               (1 + 2 * (334 / 899) * np.cos(w) + 2 * (43 / 1798) * np.cos(2 * w)), 2)
    a6, b6 = (14 / 9) / 2, (1 / 9) / 4
    _modified_wavenumber_check(lambda f, d, ax: oracle.cd06(f, d, ax), n, k,
                               lambda w: (2 * a6 * np.sin(w) + 2 * b6 * np.sin(2 * w)) / (1 + (2 / 3) * np.cos(w)), 1)
    # CF90 transfer function (tests/test_cf90.F90:108-116, tests/test_filters_parallel.F90:9-26)
    co = (9.9965e-1, 6.6652e-1, 1.6674e-1, 4.0e-5, -5.0e-6)
    _modified_wavenumber_check(lambda f, d, ax: oracle.cf90(f, ax), n, k,
                               lambda w: (co[0] + 2 * co[1] * np.cos(w) + 2 * co[2] * np.cos(2 * w) + 2 * co[3] * np.cos(3 * w) +
                                          2 * co[4] * np.cos(4 * w)) / (1 + 2 * 6.6624e-1 * np.cos(w) + 2 * 1.6688e-1 * np.cos(2 * w)), 0)
    g = (3565 / 10368, 3091 / 12960, 1997 / 25920, 149 / 12960, 107 / 103680)
    _modified_wavenumber_check(lambda f, d, ax: oracle.gaussian(f, ax), n, k,
                               lambda w: g[0] + 2 * g[1] * np.cos(w) + 2 * g[2] * np.cos(2 * w) + 2 * g[3] * np.cos(3 * w) +
                               2 * g[4] * np.cos(4 * w), 0)


def _circ(n, coef):
    A = np.zeros((n, n))
    for i in range(n):
        for o, c in coef.items():
            A[i, (i + o) % n] += c
    return A


@pytest.mark.parametrize("n", [8, 11, 32, 100])
def test_against_dense_cyclic_solve(oracle, n):
    rng = np.random.default_rng(n)
    d = 0.37
    g = rng.standard_normal((3, 4, n))
    a, b, c = (17 / 12) / 2 / d, (101 / 150) / 4 / d, (1 / 100) / 6 / d
    r = a * (np.roll(g, -1, 2) - np.roll(g, 1, 2)) + b * (np.roll(g, -2, 2) - np.roll(g, 2, 2)) + c * (np.roll(g, -3, 2) - np.roll(g, 3, 2))
    ref = np.linalg.solve(_circ(n, {0: 1, 1: .5, -1: .5, 2: .05, -2: .05}), r.reshape(-1, n).T).T.reshape(g.shape)
    assert np.abs(oracle.cd10(g, d, 0, 1) - ref).max() <= 1e-13 * np.abs(ref).max()
    if n >= 10:
        co = (9.9965e-1, 6.6652e-1, 1.6674e-1, 4.0e-5, -5.0e-6)
        r = co[0] * g + sum(co[m] * (np.roll(g, -m, 2) + np.roll(g, m, 2)) for m in range(1, 5))
        al, be = 6.6624e-1, 1.6688e-1
        ref = np.linalg.solve(_circ(n, {0: 1, 1: al, -1: al, 2: be, -2: be}), r.reshape(-1, n).T).T.reshape(g.shape)
        assert np.abs(oracle.cf90(g, 0) - ref).max() <= 1e-11 * np.abs(ref).max()  # cond(A) ~ 2e3


def test_axes_are_consistent(oracle):
    """x, y and z variants are the same line operator: transposing the field must commute with it."""
    f = broadband((12, 16, 20))
    for fn in (lambda a, ax: oracle.cd10(a, 0.1, ax, 1), lambda a, ax: oracle.cd10(a, 0.1, ax, 2),
               lambda a, ax: oracle.cd06(a, 0.1, ax), lambda a, ax: oracle.cf90(a, ax), lambda a, ax: oracle.gaussian(a, ax)):
        ox = fn(f, 0)
        oy = fn(np.ascontiguousarray(f.transpose(0, 2, 1)), 1).transpose(0, 2, 1)
        oz = fn(np.ascontiguousarray(f.transpose(2, 1, 0)), 2).transpose(2, 1, 0)
        assert np.abs(ox - oy).max() <= 1e-13 * np.abs(ox).max()
        assert np.abs(ox - oz).max() <= 1e-13 * np.abs(ox).max()


def test_init_error_codes(oracle):
    # 2 / 3 / 7 for illegal periodic n; n == 1 legal
    assert oracle.cd10_lu(5, 1)[0] == 2 and oracle.cd10_lu(8, 1)[0] == 0 and oracle.cd10_lu(1, 1)[0] == 0
    assert oracle.cd06_lu(4)[0] == 3 and oracle.cd06_lu(6)[0] == 0
    assert oracle.cf90_lu(9)[0] == 7 and oracle.cf90_lu(10)[0] == 0
    assert oracle.stagg_lu(4, 0)[0] == 21 and oracle.stagg_lu(5, 0)[0] == 0


def test_staggered_ops_analytic(oracle):
    # tests/test_PadeDer_periodic.F90:59-148 — 4x4x32, cells at (k-1/2)dz, edges at (k-1)dz, real + complex
    n = 32
    dz = 2 * np.pi / n
    zE = np.arange(n + 1) * dz
    zC = (np.arange(n) + 0.5) * dz
    ones = np.ones((1, 4, 4))
    fE, fC = np.cos(zE)[:, None, None] * ones, np.cos(zC)[:, None, None] * ones
    dE, dC = -np.sin(zE)[:, None, None] * ones, -np.sin(zC)[:, None, None] * ones
    tol = 5e-8  # 6th order at kdz = 2pi/32
    assert np.abs(oracle.stagg("ddz_E2C", fE, n, dz) - dC).max() < tol
    assert np.abs(oracle.stagg("ddz_C2E", fC, n, dz) - dE).max() < tol
    assert np.abs(oracle.stagg("interp_E2C", fE, n, dz) - fC).max() < tol
    assert np.abs(oracle.stagg("interp_C2E", fC, n, dz) - fE).max() < tol
    assert np.abs(oracle.stagg("d2dz2_C2C", fC, n, dz) + fC).max() < tol
    assert np.abs(oracle.stagg("d2dz2_E2E", fE, n, dz) + fE).max() < tol
    # complex specific = the real op on re and im separately (real LU)
    fc = fE + 1j * np.sin(zE)[:, None, None] * ones
    oc = oracle.stagg("ddz_E2C", fc, n, dz)
    assert np.array_equal(oc.real, oracle.stagg("ddz_E2C", np.ascontiguousarray(fc.real), n, dz))
    assert np.array_equal(oc.imag, oracle.stagg("ddz_E2C", np.ascontiguousarray(fc.imag), n, dz))


def test_staggered_modified_wavenumber(oracle):
    # GetKmod_CD06_stagg, tests/test_PoissonPeriodic.F90:12-25 / PadeDerOps.F90:1034-1053:
    # k'dz = (2a sin(w/2) + 2b/3... ) — derived here from the scheme itself: (2a sin(w/2) + 2b sin(3w/2))/(1+2al cos w)
    n, k = 32, 5
    dz = 2 * np.pi / n
    w = k * dz
    a, b, al = 63 / 62, (17 / 62) / 3, 9 / 62
    kp = (2 * a * np.sin(w / 2) + 2 * b * np.sin(3 * w / 2)) / (1 + 2 * al * np.cos(w)) / dz
    zE = np.arange(n + 1) * dz
    zC = (np.arange(n) + 0.5) * dz
    fE = np.cos(k * zE)[:, None, None] * np.ones((1, 2, 3))
    out = oracle.stagg("ddz_E2C", fE, n, dz)
    assert np.abs(out + kp * np.sin(k * zC)[:, None, None]).max() < 1e-12 * kp
