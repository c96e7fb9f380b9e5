"""CPU: pins the oracle's restatement of the NON-PERIODIC CD10 closures (cd10.F90:29-96, 429-707, 823-851, 1143-1262,
1636-1731; SURVEY.md §8f rank 2 — groundwork, the CUDA library still refuses periodic = .false.).  Known answers the
closures must satisfy by construction: every row of the one-sided scheme is at least 4th-order, so polynomials up to
degree 4 are differentiated exactly; the symmetric / antisymmetric closures are the interior 10th-order scheme applied to
the even / odd extension of f; and the LU sweeps must solve the pentadiagonal rows they were built from."""
import numpy as np
import pytest


def _lines(vals, axis, extra=(3, 2)):
    """A field whose lines along `axis` all equal `vals` (Fortran f(n1,n2,n3) = shape (n3,n2,n1))."""
    n = vals.size
    shape = {0: (extra[1], extra[0], n), 1: (extra[1], n, extra[0]), 2: (n, extra[1], extra[0])}[axis]
    idx = {0: (None, None, slice(None)), 1: (None, slice(None), None), 2: (slice(None), None, None)}[axis]
    return np.ascontiguousarray(np.broadcast_to(vals[idx], shape))


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("which", [1, 2])
def test_one_sided_closure_is_exact_for_quartics(oracle, axis, which):
    n = 41
    dx = 1.0 / (n - 1)
    x = np.arange(n) * dx
    for k in range(5):
        f = _lines(x ** k, axis)
        if which == 1:
            exact = k * x ** (k - 1) if k > 0 else 0 * x
        else:
            exact = k * (k - 1) * x ** (k - 2) if k > 1 else 0 * x
        got = oracle.cd10_np(f, dx, axis, which, 0, 0)
        assert np.abs(got - _lines(exact, axis)).max() < (5e-13 if which == 1 else 5e-10), (k, axis, which)
    # degree 5 is NOT exact (4th-order boundary rows): the closure really is in play
    got = oracle.cd10_np(_lines(x ** 5, axis), dx, axis, which, 0, 0)
    exact = 5 * x ** 4 if which == 1 else 20 * x ** 3
    assert 1e-8 < np.abs(got - _lines(exact, axis)).max() < 1e-1


@pytest.mark.parametrize("bc1,bcn,fn,d1,d2", [
    (1, 1, np.cos, lambda x: -np.sin(x), lambda x: -np.cos(x)),                      # even about 0 and pi
    (-1, -1, np.sin, np.cos, lambda x: -np.sin(x)),                                   # odd about 0 and pi
    (1, -1, lambda x: np.cos(x / 2), lambda x: -0.5 * np.sin(x / 2), lambda x: -0.25 * np.cos(x / 2)),   # even at 0, odd at pi
    (-1, 1, lambda x: np.sin(x / 2), lambda x: 0.5 * np.cos(x / 2), lambda x: -0.25 * np.sin(x / 2)),    # odd at 0, even at pi
])
def test_symmetry_closures_keep_tenth_order(oracle, bc1, bcn, fn, d1, d2):
    errs = {1: [], 2: []}
    for n in (17, 33):
        dx = np.pi / (n - 1)
        x = np.arange(n) * dx
        f = _lines(fn(x), 0)
        errs[1].append(np.abs(oracle.cd10_np(f, dx, 0, 1, bc1, bcn)[0, 0] - d1(x)).max())
        errs[2].append(np.abs(oracle.cd10_np(f, dx, 0, 2, bc1, bcn)[0, 0] - d2(x)).max())
    for which in (1, 2):
        assert errs[which][1] < 1e-10, (which, errs)
        order = np.log2(errs[which][0] / max(errs[which][1], 1e-300))
        assert order > 8.5 or errs[which][0] < 1e-12, (which, errs, order)     # 10th-order interior scheme, no boundary degradation


def test_symmetric_closure_equals_periodic_operator_on_the_even_extension(oracle):
    """bc = (1, 1) on [0, pi] with n points is the periodic scheme on the 2(n-1)-point even extension: same matrix, same
    right-hand side — the non-periodic restatement and the (separately pinned) periodic one must agree to rounding."""
    n = 21
    dx = np.pi / (n - 1)
    rng = np.random.default_rng(4)
    h = rng.standard_normal(n)
    ext = np.concatenate([h, h[-2:0:-1]])                 # even about both ends, period 2(n-1)
    per1 = oracle.cd10(_lines(ext, 0), dx, 0, 1)[0, 0]
    per2 = oracle.cd10(_lines(ext, 0), dx, 0, 2)[0, 0]
    np1 = oracle.cd10_np(_lines(h, 0), dx, 0, 1, 1, 1)[0, 0]
    np2 = oracle.cd10_np(_lines(h, 0), dx, 0, 2, 1, 1)[0, 0]
    assert np.abs(np1 - per1[:n]).max() < 1e-12 * np.abs(per1).max()
    assert np.abs(np2 - per2[:n]).max() < 1e-12 * np.abs(per2).max()
    odd = np.concatenate([h, -h[-2:0:-1]])
    odd[0] = odd[n - 1] = 0.0
    g = h.copy(); g[0] = g[-1] = 0.0
    per1 = oracle.cd10(_lines(odd, 0), dx, 0, 1)[0, 0]
    np1 = oracle.cd10_np(_lines(g, 0), dx, 0, 1, -1, -1)[0, 0]
    assert np.abs(np1 - per1[:n]).max() < 1e-12 * np.abs(per1).max()


@pytest.mark.parametrize("which", [1, 2])
@pytest.mark.parametrize("bc1", [0, 1, -1])
@pytest.mark.parametrize("bcn", [0, 1, -1])
def test_lu_sweeps_solve_the_rows_they_were_built_from(oracle, which, bc1, bcn):
    n = 24
    rc, P = oracle.cd10_np_penta(n, which, bc1, bcn)
    assert rc == 0
    bt, b, d, a, at = P[0], P[1], P[2], P[3], P[4]
    A = np.diag(d) + np.diag(a[:-1], 1) + np.diag(at[:-2], 2) + np.diag(b[1:], -1) + np.diag(bt[2:], -2)
    r = np.random.default_rng(n + which).standard_normal(n)
    x = oracle.cd10_np_solve_line(P, r)
    assert np.abs(A @ x - r).max() < 1e-12 * np.abs(r).max() * np.linalg.cond(A)
    assert np.abs(x - np.linalg.solve(A, r)).max() < 1e-12 * np.abs(x).max() * max(1.0, np.linalg.cond(A) / 10)


def test_short_lines_and_bad_codes_are_refused(oracle):
    assert oracle.cd10_np_penta(7, 1, 0, 0)[0] == 2
    assert oracle.cd10_np_penta(16, 1, 2, 0)[0] == 324      # cd10.F90:2044-2046


# ---- CF90 non-periodic (filters/cf90.F90:24-47, 276-418, 532-558, 672-801) ----
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_cf90_np_preserves_constants_and_keeps_the_end_points(oracle, axis):
    n = 32
    rng = np.random.default_rng(8)
    f = _lines(rng.standard_normal(n), axis)
    out = oracle.cf90_np(f, axis, 0, 0)
    first = {0: (Ellipsis, 0), 1: (slice(None), 0, slice(None)), 2: (0,)}[axis]
    last = {0: (Ellipsis, n - 1), 1: (slice(None), n - 1, slice(None)), 2: (n - 1,)}[axis]
    assert np.array_equal(out[first], f[first]) and np.array_equal(out[last], f[last])      # "first point is just identity"
    c = _lines(np.full(n, 2.5), axis)
    for bc in ((0, 0), (1, 1), (1, 0), (0, 1)):
        assert np.abs(oracle.cf90_np(c, axis, *bc) - 2.5).max() < 1e-12, bc
    # the highest representable mode is removed in the interior (what the filter is for), smooth modes pass
    x = np.arange(n)
    saw = oracle.cf90_np(_lines((-1.0) ** x, 0), 0, 1, 1)[0, 0]
    assert np.abs(saw[6:-6]).max() < 1e-3
    smooth = np.cos(np.pi * x / (n - 1))
    assert np.abs(oracle.cf90_np(_lines(smooth, 0), 0, 1, 1)[0, 0] - smooth).max() < 1e-6


def test_cf90_np_symmetry_closures_equal_the_periodic_filter_on_the_extension(oracle):
    n = 24
    rng = np.random.default_rng(5)
    h = rng.standard_normal(n)
    ext = np.concatenate([h, h[-2:0:-1]])
    per = oracle.cf90(_lines(ext, 0), 0)[0, 0]
    got = oracle.cf90_np(_lines(h, 0), 0, 1, 1)[0, 0]
    assert np.abs(got - per[:n]).max() < 1e-12 * np.abs(per).max()
    g = h.copy(); g[0] = g[-1] = 0.0
    odd = np.concatenate([g, -g[-2:0:-1]])
    per = oracle.cf90(_lines(odd, 0), 0)[0, 0]
    got = oracle.cf90_np(_lines(g, 0), 0, -1, -1)[0, 0]
    assert np.abs(got - per[:n]).max() < 1e-12 * np.abs(per).max()


@pytest.mark.parametrize("bc1", [0, 1, -1])
@pytest.mark.parametrize("bcn", [0, 1, -1])
def test_cf90_np_lu_solves_its_rows(oracle, bc1, bcn):
    n = 20
    rc, P = oracle.cf90_np_penta(n, bc1, bcn)
    assert rc == 0
    bt, b, d, a, at = P[0], P[1], P[2], P[3], P[4]
    A = np.diag(d) + np.diag(a[:-1], 1) + np.diag(at[:-2], 2) + np.diag(b[1:], -1) + np.diag(bt[2:], -2)
    r = np.random.default_rng(n).standard_normal(n)
    x = oracle.cd10_np_solve_line(P, r)
    assert np.abs(x - np.linalg.solve(A, r)).max() < 1e-12 * np.abs(x).max() * max(1.0, np.linalg.cond(A) / 10)
    assert oracle.cf90_np_penta(9, 0, 0)[0] == 7


# ---- CD06 non-periodic (derivatives/cd06.F90:27-58, 264-327, 432-449, 551-590): one-sided closure only ----
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_cd06_np_exact_for_quartics_and_sixth_order_inside(oracle, axis):
    n = 41
    dx = 1.0 / (n - 1)
    x = np.arange(n) * dx
    for k in range(5):
        exact = k * x ** (k - 1) if k > 0 else 0 * x
        assert np.abs(oracle.cd06_np(_lines(x ** k, axis), dx, axis) - _lines(exact, axis)).max() < 1e-12, (k, axis)
    err = []
    for m in (33, 65):
        h = np.pi / (m - 1)
        xx = np.arange(m) * h
        err.append(np.abs(oracle.cd06_np(_lines(np.sin(xx), 0), h, 0)[0, 0] - np.cos(xx)).max())
    assert err[1] < err[0] / 8       # at least third order globally (the boundary rows limit it), and converging


def test_cd06_np_thomas_solves_its_rows(oracle):
    n = 18
    rc, T = oracle.cd06_np_tri(n)
    assert rc == 0 and oracle.cd06_np_tri(5)[0] == 3
    a, b, c = T[3], T[4], T[5]
    A = np.diag(b) + np.diag(c[:-1], 1) + np.diag(a[1:], -1)
    # interior rows are the periodic scheme's (alpha = 1/3), the three boundary rows at each end mirror each other
    assert np.allclose(A[5, 4:7], [1 / 3, 1, 1 / 3]) and np.allclose(A[::-1, ::-1], A)
    # derivative of a random smooth-ish line through the oracle == dense solve of (rows, RHS implied by the operator on polynomials)
    x = np.linspace(0, 1, n)
    f = x ** 3
    got = oracle.cd06_np(_lines(f, 0), x[1] - x[0], 0)[0, 0]
    assert np.abs(got - 3 * x ** 2).max() < 1e-12


def test_gaussian_nonperiodic_closures(oracle):
    """gaussian%filter* with periodic = .false. (filters/gaussian.F90:22-46, 215-330): every boundary row sums to one (constants
    pass), the symmetric / antisymmetric closures are the periodic filter on the even / odd extension, and the one-sided rows
    b1..b4 are the published ones."""
    rng = np.random.default_rng(2)
    n = 20
    for bc1 in (0, 1, -1):
        for bcn in (0, 1, -1):
            c = oracle.gaussian_np(np.full((3, 4, n), 1.25), 0, bc1, bcn)
            if bc1 != -1 and bcn != -1:
                assert np.abs(c - 1.25).max() < 1e-15
    f = rng.standard_normal((2, 3, n))
    for bc in (1, -1):
        ext = np.concatenate([f, bc * f[..., -2:0:-1]], axis=-1)          # reflection about both end points: 2n - 2 periodic points
        ref = oracle.gaussian(ext, 0)[..., :n]
        got = oracle.gaussian_np(f, 0, bc, bc)
        assert np.abs(got - ref).max() < 1e-14
    got = oracle.gaussian_np(f, 0, 0, 0)
    assert np.allclose(got[..., 0], (5.0 / 6.0) * f[..., 0] + (1.0 / 6.0) * f[..., 1], rtol=0, atol=1e-15)
    assert np.allclose(got[..., n - 2], (2.0 / 3.0) * f[..., n - 2] + (1.0 / 6.0) * (f[..., n - 1] + f[..., n - 3]), rtol=0, atol=1e-15)
    assert np.allclose(got[..., 5:n - 4], oracle.gaussian(f, 0)[..., 5:n - 4], rtol=0, atol=1e-15)   # interior rows are the periodic ones
    for axis in (1, 2):                                                    # the other axes are the same operator on a moved axis
        g = rng.standard_normal((n, n, 5))
        a = oracle.gaussian_np(g, axis, 0, 1)
        b = np.moveaxis(oracle.gaussian_np(np.ascontiguousarray(np.moveaxis(g, {1: 1, 2: 0}[axis], 2)), 0, 0, 1), 2, {1: 1, 2: 0}[axis])
        assert np.array_equal(a, b)


def test_lstsq_filter(oracle):
    """lstsq%filter* (filters/lstsq.F90): the Gaussian filter's structure with the least-squares coefficients — transfer function
    0.5 + 0.6744132 cos w - 0.1744132 cos 3w (T(0) = 1, T(pi) = 0), the Gaussian's one-sided rows when non-periodic."""
    n = 32
    x = np.arange(n) * 2 * np.pi / n
    for k in (0, 3, 16):
        f = np.cos(k * x)[None, None, :] + np.zeros((2, 3, n))
        # the reference's coefficients are default-real literals widened to double: real(0.6744132, rkind) (lstsq.F90:16-18)
        b, dd = float(np.float32(0.6744132)), float(np.float32(-0.1744132))
        T = 0.5 + b * np.cos(2 * np.pi * k / n) + dd * np.cos(3 * 2 * np.pi * k / n)
        assert np.abs(oracle.lstsq(f, 0) - T * f).max() < 1e-14
    assert oracle.LSTSQ_COEFS[1] * 2 == 0.67441320419311523 and oracle.LSTSQ_COEFS[3] * 2 == -0.17441320419311523
    assert abs(0.5 + 0.6744132 - 0.1744132 - 1.0) < 1e-15 and abs(0.5 - 0.6744132 + 0.1744132) < 1e-15   # the decimal design values
    f = np.random.default_rng(4).standard_normal((2, 3, n))
    got = oracle.lstsq_np(f, 0)
    assert np.array_equal(got[..., :4], oracle.gaussian_np(f, 0, 0, 0)[..., :4]) and np.array_equal(got[..., -4:], oracle.gaussian_np(f, 0, 0, 0)[..., -4:])
    assert np.allclose(got[..., 4:n - 4], oracle.lstsq(f, 0)[..., 4:n - 4], rtol=0, atol=1e-15)
    g = np.random.default_rng(5).standard_normal((n, 12, 10))
    assert np.allclose(oracle.lstsq(g, 2), np.moveaxis(oracle.lstsq(np.ascontiguousarray(np.moveaxis(g, 0, 2)), 0), 2, 0), rtol=0, atol=0)
