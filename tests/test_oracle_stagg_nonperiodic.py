"""CPU: pins oracle/stagg_np_oracle.py (the non-periodic staggered CD06 operators of cd06stagg.F90 / STAGG_CD06_files) —
groundwork for SURVEY.md §8f rank 2.  Known answers: (i) the even / odd closures ARE the periodic staggered schemes applied
to the even / odd extension of the field, so they must reproduce the separately pinned periodic oracle (oracle.stagg,
oracle.cd06) on a 2n-cell periodic line; (ii) the one-sided closures differentiate / interpolate low-degree polynomials
exactly and converge on smooth functions; (iii) the Thomas factors solve the tridiagonal rows they were built from."""
import numpy as np
import pytest

from oracle.stagg_np_oracle import CD06StaggNP

n = 24
L = 1.7
dz = L / n
zC = (np.arange(n) + 0.5) * dz
zE = np.arange(n + 1) * dz
zC2 = (np.arange(2 * n) + 0.5) * dz        # the 2n-cell periodic extension (period 2L)
zE2 = np.arange(2 * n + 1) * dz


def col(v):
    return np.ascontiguousarray(v[:, None, None] * np.ones((1, 2, 3)))


FUN = {True: lambda z, m=3: np.cos(m * np.pi * z / L) + 0.3 * np.cos(5 * np.pi * z / L),      # even about both walls
       False: lambda z, m=2: np.sin(m * np.pi * z / L) - 0.2 * np.sin(7 * np.pi * z / L)}     # odd about both walls


@pytest.mark.parametrize("even", [True, False])
def test_even_odd_closures_equal_periodic_schemes_on_the_extension(oracle, even):
    st = CD06StaggNP(n, dz, isTopEven=even, isBotEven=even)
    f = FUN[even]
    fC, fE = col(f(zC)), col(f(zE))
    fC2, fE2 = col(f(zC2)), col(f(zE2))
    N = 2 * n
    tol = 2e-12
    ref = oracle.stagg("ddz_E2C", fE2, N, dz)
    assert np.abs(st.ddz_E2C(fE) - ref[:n]).max() < tol * np.abs(ref).max()
    ref = oracle.stagg("ddz_C2E", fC2, N, dz)
    assert np.abs(st.ddz_C2E(fC) - ref[:n + 1]).max() < tol * np.abs(ref).max()
    ref = oracle.stagg("interp_E2C", fE2, N, dz)
    assert np.abs(st.InterpZ_E2C(fE) - ref[:n]).max() < tol * np.abs(ref).max()
    ref = oracle.stagg("interp_C2E", fC2, N, dz)
    assert np.abs(st.InterpZ_C2E(fC) - ref[:n + 1]).max() < tol * np.abs(ref).max()
    ref = oracle.stagg("d2dz2_C2C", fC2, N, dz)
    assert np.abs(st.d2dz2_C2C(fC) - ref[:n]).max() < tol * np.abs(ref).max()
    ref = oracle.stagg("d2dz2_E2E", fE2, N, dz)
    assert np.abs(st.d2dz2_E2E(fE) - ref[:n + 1]).max() < tol * np.abs(ref).max()
    # collocated first derivatives: the periodic counterpart is cd06 (alpha = 1/3) on the extended cell / edge line
    ref = oracle.cd06(fC2, dz, 2)
    assert np.abs(st.ddz_C2C(fC) - ref[:n]).max() < tol * np.abs(ref).max()
    ref = oracle.cd06(np.ascontiguousarray(fE2[:N]), dz, 2)
    assert np.abs(st.ddz_E2E(fE) - ref[:n + 1]).max() < tol * np.abs(ref).max()


def test_mixed_walls(oracle):
    """even at the bottom, odd at the top: cos((m + 1/2) pi z / L) — against the analytic derivative / value (6th order)."""
    st = CD06StaggNP(48, L / 48, isTopEven=False, isBotEven=True)
    h = L / 48
    zc, ze = (np.arange(48) + 0.5) * h, np.arange(49) * h
    k = 1.5 * np.pi / L
    assert np.abs(st.ddz_E2C(col(np.cos(k * ze)))[:, 0, 0] + k * np.sin(k * zc)).max() < 1e-7
    assert np.abs(st.ddz_C2E(col(np.cos(k * zc)))[:, 0, 0] + k * np.sin(k * ze)).max() < 1e-7
    assert np.abs(st.InterpZ_E2C(col(np.cos(k * ze)))[:, 0, 0] - np.cos(k * zc)).max() < 1e-8
    assert np.abs(st.InterpZ_C2E(col(np.cos(k * zc)))[:, 0, 0] - np.cos(k * ze)).max() < 1e-8
    assert np.abs(st.d2dz2_C2C(col(np.cos(k * zc)))[:, 0, 0] + k * k * np.cos(k * zc)).max() < 1e-6
    assert np.abs(st.d2dz2_E2E(col(np.cos(k * ze)))[:, 0, 0] + k * k * np.cos(k * ze)).max() < 1e-6


def test_one_sided_closures_are_exact_on_cubics_and_converge(oracle):
    st = CD06StaggNP(n, dz, isTopEven=True, isBotEven=True, isTopSided=True, isBotSided=True)
    for k in range(4):
        d = k * zC ** (k - 1) if k > 0 else 0 * zC
        assert np.abs(st.ddz_E2C(col(zE ** k))[:, 0, 0] - d).max() < 1e-11, k
        dE = k * zE ** (k - 1) if k > 0 else 0 * zE
        assert np.abs(st.ddz_C2E(col(zC ** k))[:, 0, 0] - dE).max() < 1e-11, k
        assert np.abs(st.ddz_C2C(col(zC ** k))[:, 0, 0] - d).max() < 1e-11, k
    for k in range(3):      # the sided interpolation rows are 3-point (C2E) / 4-point (E2C): exact on quadratics
        assert np.abs(st.InterpZ_E2C(col(zE ** k))[:, 0, 0] - zC ** k).max() < 1e-12, k
        assert np.abs(st.InterpZ_C2E(col(zC ** k))[:, 0, 0] - zE ** k).max() < 1e-12, k
    errs = []
    for m in (24, 48):
        h = L / m
        zc, ze = (np.arange(m) + 0.5) * h, np.arange(m + 1) * h
        s2 = CD06StaggNP(m, h, True, True, True, True)
        errs.append(max(np.abs(s2.ddz_E2C(col(np.exp(ze)))[:, 0, 0] - np.exp(zc)).max(),
                        np.abs(s2.ddz_C2E(col(np.exp(zc)))[:, 0, 0] - np.exp(ze)).max(),
                        np.abs(s2.InterpZ_C2E(col(np.exp(zc)))[:, 0, 0] - np.exp(ze)).max()))
    assert errs[1] < errs[0] / 6          # at least ~3rd order globally with the low-order wall rows


@pytest.mark.parametrize("flags", [(True, True, False, False), (False, False, False, False), (True, False, True, False),
                                   (False, True, False, True), (True, True, True, True)])
def test_thomas_factors_solve_their_rows(flags):
    te, be, ts, bs = flags
    st = CD06StaggNP(12, 0.1, isTopEven=te, isBotEven=be, isTopSided=ts, isBotSided=bs)
    from oracle.stagg_np_oracle import _solve
    for name in ("TriD1_E2C", "TriD1_C2E", "TriD1_E2E", "TriD1_C2C", "TriD2_E2E", "TriD2_C2C", "TriInterp_E2C", "TriInterp_C2E"):
        T = getattr(st, name)
        ddn, dg, dup = T["rows"]
        m = dg.size
        A = np.diag(dg) + np.diag(dup[:-1], 1) + np.diag(ddn[1:], -1)
        rhs = np.random.default_rng(m).standard_normal((m, 1, 1))
        x = _solve(T, rhs)[:, 0, 0]
        assert np.abs(A @ x - rhs[:, 0, 0]).max() < 1e-12 * np.linalg.cond(A), (name, flags)


def test_complex_fields_take_the_same_path():
    st = CD06StaggNP(n, dz, isTopEven=True, isBotEven=False)
    fr, fi = col(np.cos(zE)), col(np.sin(2 * zE))
    got = st.ddz_E2C(fr + 1j * fi)
    assert np.allclose(got.real, st.ddz_E2C(fr), rtol=0, atol=1e-15) and np.allclose(got.imag, st.ddz_E2C(fi), rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        CD06StaggNP(4, 0.1, True, True)


def test_pade6stagg_wall_dispatch_of_the_oracle():
    """Pade6stagg with isPeriodic = .false. (PadeDerOps.F90:92-110, 185-205): (bot, top) picks one of derOO .. derSS; the second
    derivatives have no one-sided variant and return zero there; a field even about both walls differentiates like its
    periodic even extension."""
    from oracle import igrid_oracle as IG
    n, dz = 32, 1.0 / 32
    ops = IG.Pade6stagg(n, dz, scheme=1, isPeriodic=False)
    zc, ze = (np.arange(n) + 0.5) * dz, np.arange(n + 1) * dz
    fC = np.cos(3 * np.pi * zc)[:, None, None] * np.ones((1, 2, 3))      # even about z = 0 and z = 1
    dE = ops.ddz_C2E(fC, 1, 1)
    assert np.abs(dE + 3 * np.pi * np.sin(3 * np.pi * ze)[:, None, None]).max() < 1e-5
    assert not np.any(ops.d2dz2_C2C(fC, 0, 1)) and not np.any(ops.d2dz2_E2E(np.zeros((n + 1, 2, 3)) + 1.0, 1, 0))
    assert not np.any(ops.ddz_C2E(fC, 2, 1))
    errs = []
    for m in (32, 64):                                                   # no symmetry assumed: 3rd-order boundary rows
        o2 = IG.Pade6stagg(m, 1.0 / m, scheme=1, isPeriodic=False)
        zc2, ze2 = (np.arange(m) + 0.5) / m, np.arange(m + 1) / m
        one_sided = o2.ddz_C2E(np.cos(3 * np.pi * zc2)[:, None, None] * np.ones((1, 2, 3)), 0, 0)
        errs.append(np.abs(one_sided + 3 * np.pi * np.sin(3 * np.pi * ze2)[:, None, None]).max())
    assert errs[0] < 0.3 and errs[1] < errs[0] / 6.0
    with pytest.raises(AssertionError):
        IG.Pade6stagg(n, dz, scheme=2, isPeriodic=False)
