"""CPU: the PRODUCT's non-periodic staggered operators (csrc/stagg_np.cuh: the __host__ __device__ per-line routine the CUDA
kernel runs, and the host-side tridiagonal rows) executed on the host through test hooks, against the oracle's restatement of
cd06stagg%init_nonperiodic (oracle/stagg_np_oracle.py, pinned in tests/test_oracle_stagg_nonperiodic.py) — every operator,
every wall combination (even / odd / one-sided at each end), real and complex lines, shortest legal line included."""
import ctypes as C
import itertools

import numpy as np
import pytest

from oracle import stagg_np_oracle as SN

OPS = ["ddz_E2C", "ddz_C2E", "ddz_C2C", "ddz_E2E", "InterpZ_E2C", "InterpZ_C2E", "d2dz2_C2C", "d2dz2_E2E"]   # enum StaggNpOp order
TRI = ["TriD1_E2C", "TriD1_C2E", "TriD1_C2C", "TriD1_E2E", "TriInterp_E2C", "TriInterp_C2E", "TriD2_C2C", "TriD2_E2E"]
EDGE_IN = {0, 3, 4, 7}
EDGE_OUT = {1, 3, 5, 7}
# (isBotEven, isTopEven, isBotSided, isTopSided)
WALLS = [w for w in itertools.product([0, 1], [0, 1], [0, 1], [0, 1])]


@pytest.mark.parametrize("walls", WALLS)
@pytest.mark.parametrize("n", [5, 8, 33])
def test_tridiagonal_rows_match_oracle(pdo, n, walls):
    be, te, bs, ts = walls
    ref = SN.CD06StaggNP(n, 0.37, bool(te), bool(be), bool(ts), bool(bs))
    for op in range(8):
        m = n + 1 if op in EDGE_OUT else n
        rows = np.zeros((3, m))
        assert pdo.lib().pdo_debug_stagg_np_rows(op, n, be, te, bs, ts, C.c_void_p(rows.ctypes.data)) == 0
        ddn, dg, dup = getattr(ref, TRI[op])["rows"]
        assert np.array_equal(rows[0], ddn) and np.array_equal(rows[1], dg) and np.array_equal(rows[2], dup), OPS[op]


@pytest.mark.parametrize("walls", WALLS)
@pytest.mark.parametrize("n", [5, 6, 24])
@pytest.mark.parametrize("cplx", [False, True])
def test_line_routine_matches_oracle(pdo, n, walls, cplx):
    be, te, bs, ts = walls
    dx = 0.21
    ref = SN.CD06StaggNP(n, dx, bool(te), bool(be), bool(ts), bool(bs))
    rng = np.random.default_rng(n * 16 + be * 8 + te * 4 + bs * 2 + ts)
    n2, n1 = 3, 5
    for op, name in enumerate(OPS):
        rows_in = n + 1 if op in EDGE_IN else n
        rows_out = n + 1 if op in EDGE_OUT else n
        f = rng.standard_normal((rows_in, n2, n1))
        if cplx:
            f = f + 1j * rng.standard_normal((rows_in, n2, n1))
        want = getattr(ref, name)(f)
        got = np.zeros((rows_out, n2, n1), dtype=f.dtype)
        ncols = n1 * n2 * (2 if cplx else 1)       # a complex line is two interleaved real lines
        rc = pdo.lib().pdo_debug_stagg_np_host(op, n, dx, be, te, bs, ts, C.c_void_p(f.ctypes.data), C.c_void_p(got.ctypes.data), ncols)
        assert rc == 0
        assert want.shape == got.shape
        assert np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), (name, walls, np.abs(got - want).max())


def test_short_lines_are_refused(pdo):
    out = np.zeros(16)
    f = np.zeros(16)
    assert pdo.lib().pdo_debug_stagg_np_host(0, 4, 0.1, 1, 1, 0, 0, C.c_void_p(f.ctypes.data), C.c_void_p(out.ctypes.data), 1) == 21
