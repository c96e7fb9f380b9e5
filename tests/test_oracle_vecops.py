"""Pins the oracle's restatement of utilities/operators.F90 (gradient / curl / divergence) with analytic fields and
the discrete identities the tensor-product compact operators satisfy exactly (curl grad = 0, div curl = 0)."""
import numpy as np
import pytest


def _grid(n):
    d = 2 * np.pi / n
    x = np.arange(n) * d
    return d, x[None, None, :], x[None, :, None], x[:, None, None]


@pytest.mark.parametrize("method,tol", [("cd10", 1e-9), ("cd06", 2e-6)])
def test_gradient_analytic(oracle, method, tol):
    n = 48
    d, X, Y, Z = _grid(n)
    f = np.sin(X) * np.sin(2 * Y) * np.cos(Z) + 0 * (X + Y + Z)
    gx, gy, gz = oracle.gradient(f, d, d, d, method)
    assert np.abs(gx - np.cos(X) * np.sin(2 * Y) * np.cos(Z)).max() < tol
    assert np.abs(gy - 2 * np.sin(X) * np.cos(2 * Y) * np.cos(Z)).max() < 40 * tol
    assert np.abs(gz + np.sin(X) * np.sin(2 * Y) * np.sin(Z)).max() < tol


@pytest.mark.parametrize("method", ["cd10", "cd06"])
def test_discrete_identities(oracle, method):
    n = 32
    d, X, Y, Z = _grid(n)
    rng = np.random.default_rng(3)
    f = rng.standard_normal((n, n, n))
    gx, gy, gz = oracle.gradient(f, d, d, d, method)
    c = oracle.curl(gx, gy, gz, d, d, d, method)
    scale = np.abs(gx).max() / d
    assert np.abs(c).max() < 1e-12 * scale          # curl grad f = 0: the 1-D operators commute
    u, v, w = (rng.standard_normal((n, n, n)) for _ in range(3))
    cu = oracle.curl(u, v, w, d, d, d, method)
    div = oracle.divergence(cu[0], cu[1], cu[2], d, d, d, method)
    assert np.abs(div).max() < 1e-12 * np.abs(cu).max() / d   # div curl u = 0


def test_divergence_taylor_green(oracle):
    n = 40
    d, X, Y, Z = _grid(n)
    u = np.sin(X) * np.cos(Y) * np.cos(Z)
    v = -np.cos(X) * np.sin(Y) * np.cos(Z)
    w = np.zeros_like(u) + 0 * Z
    assert np.abs(oracle.divergence(u, v, w, d, d, d, "cd10")).max() < 1e-9


def test_curl_of_rotations_like_test_operators(oracle):
    """tests/test_operators.F90:67-140 checks curl on the three rigid rotations (u, v, w) = (y, -x, 0), (0, z, -y),
    (-z, 0, x) of a NON-periodic box (vorticity -2 along the rotation axis).  The periodic counterpart on the hot path:
    replace each coordinate by its sine, vorticity = -(cos a + cos b) along the axis, zero elsewhere."""
    n = 48
    d, X, Y, Z = _grid(n)
    zero = np.zeros((n, n, n))
    S = lambda c: np.sin(c) + zero
    cases = [((S(Y), -S(X), zero), 2, -(np.cos(X) + np.cos(Y))),
             ((zero, S(Z), -S(Y)), 0, -(np.cos(Y) + np.cos(Z))),
             ((-S(Z), zero, S(X)), 1, -(np.cos(Z) + np.cos(X)))]
    for (u, v, w), axis, exact in cases:
        c = oracle.curl(u, v, w, d, d, d, "cd10")
        for k in range(3):
            want = exact + zero if k == axis else zero
            assert np.abs(c[k] - want).max() < 1e-9, (axis, k)


def _T_cf90(w):
    co = (9.9965e-1, 6.6652e-1, 1.6674e-1, 4.0e-5, -5.0e-6)
    return (co[0] + 2 * co[1] * np.cos(w) + 2 * co[2] * np.cos(2 * w) + 2 * co[3] * np.cos(3 * w) + 2 * co[4] * np.cos(4 * w)) / \
        (1 + 2 * 6.6624e-1 * np.cos(w) + 2 * 1.6688e-1 * np.cos(2 * w))


def _T_gauss(w):
    g = (3565 / 10368, 3091 / 12960, 1997 / 25920, 149 / 12960, 107 / 103680)
    return g[0] + 2 * g[1] * np.cos(w) + 2 * g[2] * np.cos(2 * w) + 2 * g[3] * np.cos(3 * w) + 2 * g[4] * np.cos(4 * w)


@pytest.mark.parametrize("numtimes", [1, 3])
@pytest.mark.parametrize("methods", [("cf90", "cf90", "cf90"), ("gaussian", "cf90", "gaussian")])
def test_filter3d_single_mode_transfer(oracle, numtimes, methods):
    """filter3D (operators.F90:158-224) of one Fourier mode = product over the axes of (1-D transfer function)^numtimes — the
    3-D version of tests/test_cf90.F90:108-116 / tests/test_filters_parallel.F90:9-26."""
    nx, ny, nz = 32, 24, 40
    kx, ky, kz = 5, 3, 9
    x, y, z = (np.arange(n) * 2 * np.pi / n for n in (nx, ny, nz))
    f = np.cos(kx * x[None, None, :] + 0.3) * np.sin(ky * y[None, :, None] - 0.2) * np.cos(kz * z[:, None, None] + 1.1)
    T = 1.0
    for m, k, n in zip(methods, (kx, ky, kz), (nx, ny, nz)):
        T *= (_T_gauss if m == "gaussian" else _T_cf90)(2 * np.pi * k / n) ** numtimes
    out = oracle.filter3D(f, numtimes, methods)
    assert np.abs(out - T * f).max() < 2e-12
    assert 0.0 < T < 1.0


def test_filter3d_order_and_refilter(oracle):
    """numtimes = 2 equals two numtimes = 1 applications axis by axis (y twice, then x twice, then z twice), and constants pass."""
    rng = np.random.default_rng(11)
    f = rng.standard_normal((16, 20, 24))
    ref = f
    for axis in (1, 0, 2):
        for _ in range(2):
            ref = oracle.cf90(ref, axis)
    assert np.array_equal(oracle.filter3D(f, 2), ref)
    c = np.full((12, 16, 20), 2.5)
    assert np.abs(oracle.filter3D(c, 3) - 2.5).max() < 1e-11   # T(0) = 1 up to the rounding the nine pentadiagonal solves amplify
    assert np.array_equal(oracle.filter3D(f, 0), oracle.filter3D(f, 1))   # "do idx = 1, times2fil - 1": at least one pass


def test_filter3d_nonperiodic_z(oracle):
    """z_bc reaches the z filter only: symmetric walls (1, 1) in z reproduce the periodic filter on the even extension."""
    rng = np.random.default_rng(5)
    nz, ny, nx = 24, 16, 16
    f = rng.standard_normal((nz, ny, nx))
    out = oracle.filter3D(f, 1, periodic=(True, True, False), z_bc=(1, 1))
    ext = np.concatenate([f, f[-2:0:-1]], axis=0)    # even extension about both end points: 2 nz - 2 periodic points
    ref = oracle.filter3D(ext, 1)[:nz]
    assert np.abs(out - ref).max() < 1e-12
