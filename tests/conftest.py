import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "late_wrapper: the test that runs the not-yet-hardware-validated GPU tests in a subprocess")


def _is_late(item):
    """GPU tests written after a round's last GPU session carry xfail(strict=False) until their first hardware run."""
    return item.get_closest_marker("gpu") is not None and item.get_closest_marker("xfail") is not None and \
        item.get_closest_marker("late_wrapper") is None


def pytest_collection_modifyitems(config, items):
    """Isolation of unvalidated device code: a kernel that faults poisons the CUDA context of its process, and every test after
    it would fail with it.  So the late tests never run inside the main pytest process: there they are skipped, and
    tests/test_late_isolated.py runs them (and only them) in a child process with PDO_RUN_LATE=1."""
    if os.environ.get("PDO_RUN_LATE") == "1":
        keep, drop = [], []
        for it in items:
            (keep if _is_late(it) else drop).append(it)
        items[:] = keep
        config.hook.pytest_deselected(items=drop)
        return
    skip = pytest.mark.skip(reason="not yet validated on hardware: runs in the child process of tests/test_late_isolated.py")
    for it in items:
        if _is_late(it):
            it.add_marker(skip)


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
