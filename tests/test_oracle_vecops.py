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
