"""CPU: the PRODUCT's host-side factorisation (csrc/tables.cpp) through a host-only hook of the C ABI.  The per-chunk
algorithm of the CUDA kernels (banded.cu: chunk_interior / separator solve / chunk_finish) is re-enacted in numpy with the
product's own tables and must solve the cyclic system circ[b2 b1 1 b1 b2] x = r like a dense solver does — for every
matrix on the hot path, every chunk length, sparse and dense separator coupling.  No GPU, no oracle involved."""
import ctypes as C

import numpy as np
import pytest

MATS = {  # name: (bw, b1, b2)
    "cd10_d1": (2, 1.0 / 2.0, 1.0 / 20.0),          # derivatives/cd10.F90:16-17
    "cd10_d2": (2, 334.0 / 899.0, 43.0 / 1798.0),    # :23-24
    "cf90": (2, 6.6624e-1, 1.6688e-1),               # filters/cf90.F90:16-17
    "cd06_d1": (1, 1.0 / 3.0, 0.0),                  # derivatives/cd06.F90:14
    "stagg_d1": (1, 9.0 / 62.0, 0.0),                # derivatives/cd06stagg.F90:174-176
    "stagg_d2": (1, 2.0 / 11.0, 0.0),
    "stagg_interp": (1, 3.0 / 10.0, 0.0),
}


def tables(pdo, n, M, bw, b1, b2):
    from padeops_b200._lib import ChunkTables
    t = ChunkTables()
    rc = pdo.lib().pdo_debug_chunk_tables(n, M, bw, b1, b2, C.byref(t), C.sizeof(t))
    return rc, t


def chunk_solve(t, r):
    """x with A x = r, computed the way the kernels do it (one 'thread' per chunk)."""
    n, M, P, BW, W = t.n, t.M, t.P, t.BW, t.W
    mi = M - BW
    l1, l2, ginv, ug, bg = (np.array(a[:]) for a in (t.l1, t.l2, t.ginv, t.ug, t.bg))
    V = np.array([list(v) for v in t.V])
    U = np.array([list(u) for u in t.U])
    G = np.array([list(g) for g in t.G])
    z = np.zeros((P, mi))
    gA, gB = np.zeros((P, 2)), np.zeros((P, 2))
    for p in range(P):
        rr = r[p * M:(p + 1) * M]
        y = np.zeros(mi)
        for i in range(mi):
            y[i] = rr[i] - (l1[i] * y[i - 1] if i >= 1 else 0.0) - (l2[i] * y[i - 2] if (i >= 2 and BW == 2) else 0.0)
        for i in range(mi - 1, -1, -1):
            v = y[i] * ginv[i]
            if i + 1 < mi:
                v -= ug[i] * z[p, i + 1]
            if i + 2 < mi and BW == 2:
                v -= bg[i] * z[p, i + 2]
            z[p, i] = v
        if BW == 2:
            gA[p] = [rr[M - 2] - t.b2 * z[p, mi - 2] - t.b1 * z[p, mi - 1], rr[M - 1] - t.b2 * z[p, mi - 1]]
            gB[p] = [-t.b2 * z[p, 0], -t.b1 * z[p, 0] - t.b2 * z[p, 1]]
        else:
            gA[p, 0] = rr[M - 1] - t.b1 * z[p, mi - 1]
            gB[p, 0] = -t.b1 * z[p, 0]
    s = np.zeros((P, 2))
    for p in range(P):
        for d in range(2 * W + 1):
            q = (p - W + d) % P
            h = gA[q] + gB[(q + 1) % P]
            if BW == 2:
                s[p, 0] += G[d, 0] * h[0] + G[d, 1] * h[1]
                s[p, 1] += G[d, 2] * h[0] + G[d, 3] * h[1]
            else:
                s[p, 0] += G[d, 0] * h[0]
    x = np.zeros(n)
    for p in range(P):
        sp = s[(p - 1) % P]
        xi = z[p] - V[:mi, 0] * sp[0] - U[:mi, 0] * s[p, 0]
        if BW == 2:
            xi = xi - V[:mi, 1] * sp[1] - U[:mi, 1] * s[p, 1]
        x[p * M:p * M + mi] = xi
        x[p * M + mi:(p + 1) * M] = s[p, :BW]
    return x


def dense(n, b1, b2):
    A = np.eye(n)
    for i in range(n):
        A[i, (i + 1) % n] += b1
        A[i, (i - 1) % n] += b1
        A[i, (i + 2) % n] += b2
        A[i, (i - 2) % n] += b2
    return A


@pytest.mark.parametrize("name", sorted(MATS))
@pytest.mark.parametrize("n,M", [(64, 32), (96, 32), (256, 32), (1024, 32), (48, 16), (512, 16), (40, 8), (1024, 8)])
def test_chunk_factorisation_solves_the_cyclic_system(pdo, name, n, M):
    bw, b1, b2 = MATS[name]
    rc, t = tables(pdo, n, M, bw, b1, b2)
    if rc != 0:
        # legitimately not chunkable at this (n, M): the separator reach exceeds the table without being dense
        assert name == "cf90" and M < 32, (name, n, M)
        return
    assert (t.n, t.M, t.P, t.BW) == (n, M, n // M, bw)
    assert t.dense == int(2 * t.W + 1 >= t.P)
    rng = np.random.default_rng(n + M)
    r = rng.standard_normal(n)
    x = chunk_solve(t, r)
    ref = np.linalg.solve(dense(n, b1, b2), r)
    assert np.abs(x - ref).max() < 2e-14 * np.abs(ref).max() * max(1.0, np.linalg.cond(dense(n, b1, b2)) / 10)


def test_reach_of_the_separator_coupling(pdo):
    """The numbers DESIGN.md quotes: W = 2 (CD10), 1 (CD06), 6 (CF90) at M = 32; longer lines change nothing."""
    for name, want in (("cd10_d1", 2), ("cd06_d1", 1), ("cf90", 6)):
        for n in (1024, 2048, 8192):
            rc, t = tables(pdo, n, 32, *MATS[name])
            assert rc == 0 and t.W == want and not t.dense, (name, n, t.W)
    rc1, a = tables(pdo, 1024, 32, *MATS["cd10_d1"])
    rc2, b = tables(pdo, 8192, 32, *MATS["cd10_d1"])
    assert np.allclose(np.array([list(g) for g in a.G]), np.array([list(g) for g in b.G]), rtol=0, atol=1e-18)


def test_rejects_unchunkable(pdo):
    assert tables(pdo, 100, 32, 2, 0.5, 0.05)[0] == -1      # n % M != 0
    assert tables(pdo, 64, 4, 2, 0.5, 0.05)[0] == -1        # M - BW < BW + ...: chunk too short for a pentadiagonal block
