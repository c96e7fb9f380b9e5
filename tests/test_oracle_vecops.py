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
