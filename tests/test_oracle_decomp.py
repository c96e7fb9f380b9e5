"""CPU: the decomposition arithmetic and the simulated-rank transposes of the oracle, and the
product's own decomposition arithmetic (pure host code inside the CUDA library) against it."""
import numpy as np
import pytest


def test_distribute_matches_2decomp_rule(oracle):
    # 2D» decomp_2d.f90:715-716 comment: 17 points over 4 ranks → (4,4,4,5)
    st, en, sz = oracle.distribute(17, 4)
    assert sz == [4, 4, 4, 5] and st == [1, 5, 9, 13] and en == [4, 8, 12, 17]
    st, en, sz = oracle.distribute(257, 8)
    assert sz == [32] * 7 + [33]
    st, en, sz = oracle.distribute(513, 8)
    assert sz == [64] * 7 + [65]


@pytest.mark.parametrize("grid", [(1, 1), (2, 2), (1, 4), (4, 1), (2, 3), (3, 2)])
@pytest.mark.parametrize("cplx", [False, True])
def test_transposes_equal_global_redistribution(oracle, grid, cplx):
    """pack → ALLTOALLV → unpack must be exactly 'every rank ends up with its block of the global array'."""
    pr, pc = grid
    nx, ny, nz = 9, 8, 7  # uneven on purpose (ALLTOALLV is the normal case)
    rng = np.random.default_rng(1)
    G = rng.standard_normal((nz, ny, nx))
    if cplx:
        G = G + 1j * rng.standard_normal((nz, ny, nx))
    pens = {p: oracle.scatter_global(G, nx, ny, nz, pr, pc, p) for p in "xyz"}
    for d, (s, t) in enumerate((("x", "y"), ("y", "x"), ("y", "z"), ("z", "y"))):
        out = oracle.transpose(d, nx, ny, nz, pr, pc, pens[s])
        for a, b in zip(out, pens[t]):
            assert a.shape == b.shape and np.array_equal(a, b)


def test_product_decomp_arithmetic_matches_oracle(oracle, pdo):
    for (nx, ny, nz, pr, pc) in [(128, 128, 128, 2, 2), (257, 512, 512, 1, 8), (257, 512, 513, 2, 4), (17, 9, 5, 2, 2), (65, 64, 64, 4, 2)]:
        for r in range(pr * pc):
            assert pdo.decomp_info.for_rank(nx, ny, nz, pr, pc, r) == oracle.decomp_info(nx, ny, nz, pr, pc, r)


def test_product_decomp_rejects_bad_grid(pdo):
    with pytest.raises(pdo.PadeOpsError) as e:
        pdo.decomp_info.for_rank(4, 4, 4, 8, 1, 0)
    assert e.value.code == 6  # 2D» decomp_2d.f90:507-514


def test_wavenumbers(oracle):
    k = oracle.wavenums(8, 2 * np.pi / 8)
    assert np.allclose(k, [0, 1, 2, 3, -4, -3, -2, -1], atol=1e-14)
    k = oracle.wavenums(7, 2 * np.pi / 7)  # odd n: spacing 2pi/(n-1) per the reference's `dummy`
    assert abs(k[0]) < 1e-14 and k[1] > 0 and k[-1] < 0


def test_poisson_manufactured_solution(oracle):
    # tests/test_PoissonPeriodic.F90:109-118 — 64x32x16, (l,m,n) = (6,3,1)
    nx, ny, nz = 64, 32, 16
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    x, y, z = np.arange(nx) * dx, np.arange(ny) * dy, np.arange(nz) * dz
    ftrue = np.sin(6 * x)[None, None, :] * np.cos(3 * y)[None, :, None] * np.sin(1 * z)[:, None, None]
    rhs = -(36 + 9 + 1) * ftrue
    f = oracle.poisson_solve(rhs, dx, dy, dz)
    assert np.abs(f - ftrue).max() < 1e-13


def test_product_decomp_arithmetic_random_grids(oracle, pdo):
    """Index work is held bit-exact: the product's host-side decomposition (csrc/decomp.cu: fill_info, no GPU needed)
    against the oracle's restatement of 2DECOMP's distribute / partition on random uneven grids, plus the property that the
    pencils of all ranks tile the global box exactly once."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 6), st.integers(1, 6), st.integers(0, 40), st.integers(0, 40), st.integers(0, 40))
    def check(pr, pc, ax, ay, az):
        nx, ny, nz = max(pr, 1) + ax, max(pr, pc) + ay, max(pc, 1) + az
        cover = {p: np.zeros((nz, ny, nx), dtype=np.int32) for p in "xyz"}
        for r in range(pr * pc):
            got = pdo.decomp_info.for_rank(nx, ny, nz, pr, pc, r)
            assert got == oracle.decomp_info(nx, ny, nz, pr, pc, r)
            for p in "xyz":
                s, e, z = got[p + "st"], got[p + "en"], got[p + "sz"]
                assert all(e[i] - s[i] + 1 == z[i] for i in range(3))
                cover[p][s[2] - 1:e[2], s[1] - 1:e[1], s[0] - 1:e[0]] += 1
        for p in "xyz":
            assert (cover[p] == 1).all()
    check()
