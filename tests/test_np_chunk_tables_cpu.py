"""CPU: the PRODUCT's chunk factorisation of the NON-CYCLIC boundary-row systems (csrc/np_chunk_tables.cpp) — groundwork
for the fast-path kernels of the non-periodic closures.  The per-chunk algorithm those kernels will run (three table sets:
first / mid / last chunk; position-dependent truncated separator inverse) is re-enacted in numpy with the product's own
tables and must solve the pentadiagonal system assembled from the product's own rows like a dense solver does."""
import ctypes as C

import numpy as np
import pytest

KINDS = {"cd10_d1": 0, "cd10_d2": 1, "cf90": 2}
SET_FIELDS = [("l1", 32), ("l2", 32), ("ginv", 32), ("ug", 32), ("bg", 32), ("V", 64), ("U", 64), ("cA", 3), ("cB", 3)]


def rows_of(pdo, kind, n, bc1, bcn):
    rows = np.zeros(5 * n)
    assert pdo.lib().pdo_debug_np_rows(KINDS[kind], n, bc1, bcn, C.c_void_p(rows.ctypes.data)) == 0
    return rows.reshape(5, n)


def tables_of(pdo, kind, n, M, bc1, bcn):
    nset = sum(k for _, k in SET_FIELDS)
    sets = np.zeros(3 * nset)
    G = np.zeros((n // M) * 33 * 4)
    meta = (C.c_int * 3)()
    rc = pdo.lib().pdo_debug_np_chunk_tables(KINDS[kind], n, M, bc1, bcn, C.c_void_p(sets.ctypes.data), C.c_void_p(G.ctypes.data), G.size, meta)
    if rc != 0:
        return rc, None
    P, W, per = meta[0], meta[1], meta[2]
    assert per == nset
    out = []
    for s in range(3):
        d, off = {}, s * nset
        for name, k in SET_FIELDS:
            d[name] = sets[off:off + k].copy()
            off += k
        d["V"], d["U"] = d["V"].reshape(32, 2), d["U"].reshape(32, 2)
        out.append(d)
    return 0, {"P": P, "W": W, "sets": out, "G": G[:P * (2 * W + 1) * 4].reshape(P, 2 * W + 1, 2, 2)}


def chunk_solve(t, n, M, r):
    P, W, mi = t["P"], t["W"], M - 2
    z = np.zeros((P, mi)); gA = np.zeros((P, 2)); gB = np.zeros((P + 1, 2))
    for p in range(P):
        s = t["sets"][0 if p == 0 else (2 if p == P - 1 else 1)]
        rr = r[p * M:(p + 1) * M]
        y = np.zeros(mi)
        for i in range(mi):
            y[i] = rr[i] - (s["l1"][i] * y[i - 1] if i >= 1 else 0.0) - (s["l2"][i] * y[i - 2] if i >= 2 else 0.0)
        for i in range(mi - 1, -1, -1):
            v = y[i] * s["ginv"][i]
            if i + 1 < mi:
                v -= s["ug"][i] * z[p, i + 1]
            if i + 2 < mi:
                v -= s["bg"][i] * z[p, i + 2]
            z[p, i] = v
        gA[p] = [rr[M - 2] - s["cA"][0] * z[p, mi - 2] - s["cA"][1] * z[p, mi - 1], rr[M - 1] - s["cA"][2] * z[p, mi - 1]]
        gB[p] = [-s["cB"][0] * z[p, 0], -s["cB"][1] * z[p, 0] - s["cB"][2] * z[p, 1]]
    h = gA + gB[1:]                      # gB_P = 0: nothing beyond the last chunk
    sep = np.zeros((P, 2))
    for p in range(P):
        for d in range(2 * W + 1):
            q = p - W + d
            if 0 <= q < P:
                sep[p] += t["G"][p, d] @ h[q]
    x = np.zeros(n)
    for p in range(P):
        s = t["sets"][0 if p == 0 else (2 if p == P - 1 else 1)]
        sp = sep[p - 1] if p > 0 else np.zeros(2)
        x[p * M:p * M + mi] = z[p] - s["V"][:mi] @ sp - s["U"][:mi] @ sep[p]
        x[p * M + mi:(p + 1) * M] = sep[p]
    return x


def dense(rows):
    bt, b, d, a, at = rows
    return np.diag(d) + np.diag(a[:-1], 1) + np.diag(at[:-2], 2) + np.diag(b[1:], -1) + np.diag(bt[2:], -2)


@pytest.mark.parametrize("kind", sorted(KINDS))
@pytest.mark.parametrize("bc1", [0, 1, -1])
@pytest.mark.parametrize("bcn", [0, 1, -1])
@pytest.mark.parametrize("n,M", [(64, 32), (256, 32), (96, 16)])
def test_noncyclic_chunk_factorisation_solves_the_boundary_row_system(pdo, kind, bc1, bcn, n, M):
    rows = rows_of(pdo, kind, n, bc1, bcn)
    rc, t = tables_of(pdo, kind, n, M, bc1, bcn)
    if rc != 0:
        assert kind == "cf90" and M < 32, (kind, n, M)      # CF90's separator reach at short chunks exceeds the table
        return
    r = np.random.default_rng(n + M + 3 * bc1 + bcn + 10).standard_normal(n)
    x = chunk_solve(t, n, M, r)
    A = dense(rows)
    ref = np.linalg.solve(A, r)
    assert np.abs(x - ref).max() < 5e-14 * np.abs(ref).max() * max(1.0, np.linalg.cond(A) / 10), (kind, bc1, bcn, n, M, t["W"])


def test_reach_matches_the_cyclic_case(pdo):
    for kind, want in (("cd10_d1", 2), ("cf90", 6)):
        rc, t = tables_of(pdo, kind, 1024, 32, 0, 0)
        assert rc == 0 and t["W"] == want, (kind, t["W"])
        # away from the ends the position-dependent inverse is the interior one: chunks 8 and 20 carry the same blocks
        assert np.allclose(t["G"][8], t["G"][20], rtol=0, atol=1e-17)
