"""CPU: the PRODUCT's non-periodic closures (csrc/nonperiodic.cuh: the __host__ __device__ per-point RHS and per-line
LU sweeps the CUDA kernels execute, plus csrc/nonperiodic.cu's coefficient / table builders) run on the host through a
test hook and compared with the oracle's statement-by-statement restatement of cd10.F90 / cf90.F90 — for all nine
(bc1, bcn) combinations, every axis, first and second derivative and the filter.  The product writes the symmetric /
antisymmetric rows as the interior stencil on the reflected line; the oracle writes them out like the Fortran: agreement
is expected to the last bit except where a table entry is rounded differently, hence 1e-14."""
import ctypes as C

import numpy as np
import pytest

KINDS = {"cd10_d1": 0, "cd10_d2": 1, "cf90": 2, "cd06_d1": 3, "gaussian": 4, "lstsq": 5}


def product_host(pdo, kind, f, dx, axis, bc1, bcn):
    nz, ny, nx = f.shape
    n, na, nb = {0: (nx, ny, nz), 1: (ny, nx, nz), 2: (nz, nx, ny)}[axis]
    out = np.empty_like(f)
    rc = pdo.lib().pdo_debug_np_line_host(KINDS[kind], n, float(dx), bc1, bcn, axis, C.c_void_p(f.ctypes.data), C.c_void_p(out.ctypes.data), na, nb)
    assert rc == 0, rc
    return out


def oracle_ref(oracle, kind, f, dx, axis, bc1, bcn):
    if kind == "cf90":
        return oracle.cf90_np(f, axis, bc1, bcn)
    if kind == "cd06_d1":
        return oracle.cd06_np(f, dx, axis)
    if kind == "gaussian":
        return oracle.gaussian_np(f, axis, bc1, bcn)
    if kind == "lstsq":
        return oracle.lstsq_np(f, axis)
    return oracle.cd10_np(f, dx, axis, 1 if kind == "cd10_d1" else 2, bc1, bcn)


@pytest.mark.parametrize("kind", sorted(KINDS))
@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("bc1", [0, 1, -1])
@pytest.mark.parametrize("bcn", [0, 1, -1])
def test_product_np_routines_match_oracle(pdo, oracle, kind, axis, bc1, bcn):
    if kind in ("cd06_d1", "lstsq") and (bc1, bcn) != (0, 0):
        pytest.skip("cd06 and lstsq have the one-sided closure only")
    shape = {0: (3, 4, 19), 1: (3, 19, 4), 2: (19, 3, 4)}[axis]
    rng = np.random.default_rng(1000 + 100 * axis + 10 * bc1 + bcn)
    f = rng.standard_normal(shape)
    dx = 0.37
    got = product_host(pdo, kind, f, dx, axis, bc1, bcn)
    ref = oracle_ref(oracle, kind, f, dx, axis, bc1, bcn)
    # cd06: the product carries the tridiagonal system through its pentadiagonal sweeps, the reference through Thomas: rounding differs
    tol = 1e-13 if kind == "cd06_d1" else 1e-14
    assert np.abs(got - ref).max() <= tol * np.abs(ref).max(), (kind, axis, bc1, bcn, np.abs(got - ref).max() / np.abs(ref).max())


def test_product_np_minimum_lengths_and_codes(pdo):
    f = np.zeros((1, 1, 7))
    out = np.empty_like(f)
    L = pdo.lib()
    assert L.pdo_debug_np_line_host(0, 7, 0.1, 0, 0, 0, C.c_void_p(f.ctypes.data), C.c_void_p(out.ctypes.data), 1, 1) == 2
    f = np.zeros((1, 1, 9)); out = np.empty_like(f)
    assert L.pdo_debug_np_line_host(2, 9, 0.1, 0, 0, 0, C.c_void_p(f.ctypes.data), C.c_void_p(out.ctypes.data), 1, 1) == 7
    f = np.zeros((1, 1, 16)); out = np.empty_like(f)
    assert L.pdo_debug_np_line_host(0, 16, 0.1, 2, 0, 0, C.c_void_p(f.ctypes.data), C.c_void_p(out.ctypes.data), 1, 1) == 324
