import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def pdo():
    """The CUDA library through its Python mirror; GPU tests fail (not skip) if it is missing."""
    import padeops_b200
    padeops_b200.lib()
    return padeops_b200


def broadband(shape, seed=20240607):
    """Field F2 of SURVEY.md §8d: a few integer-wavenumber modes plus 1e-3 uniform noise; O(1) values."""
    import numpy as np
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    x = np.arange(nx) * (2 * np.pi / nx)
    y = np.arange(ny) * (2 * np.pi / ny)
    z = np.arange(nz) * (2 * np.pi / nz)
    f = np.zeros(shape)
    for m in range(1, 9):
        kx = int(rng.integers(1, max(2, nx // 4 + 1)))
        ky = int(rng.integers(1, max(2, ny // 4 + 1)))
        kz = int(rng.integers(1, max(2, nz // 4 + 1)))
        ph = rng.uniform(0, 2 * np.pi)
        f += (1.0 / m) * np.sin(kx * x[None, None, :] + ky * y[None, :, None] + kz * z[:, None, None] + ph)
    f += 1e-3 * rng.uniform(-1, 1, size=shape)
    return f
