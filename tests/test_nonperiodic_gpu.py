"""GPU: the non-periodic CD10 / CF90 closures (correctness path) through the C ABI against the
oracle: all nine (bc1, bcn) combinations, every axis, device and host arrays, and through the derivatives / filters
dispatch types with periodic = .false. (derivatives.F90:447-569 forwards bc1, bcn)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12
BCS = [(a, b) for a in (0, 1, -1) for b in (0, 1, -1)]


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(got, ref):
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_cd10_nonperiodic_all_boundary_codes(pdo, oracle, axis):
    shape = {0: (5, 6, 40), 1: (5, 40, 6), 2: (40, 5, 6)}[axis]
    n, dx = 40, 0.05
    f = np.random.default_rng(axis).standard_normal(shape)
    fd = _dev(f)
    h = pdo.cd10()
    assert h.init(n, dx, periodic_=False) == 0
    d1 = (h.dd1, h.dd2, h.dd3)[axis]
    d2 = (h.d2d1, h.d2d2, h.d2d3)[axis]
    for bc1, bcn in BCS:
        assert _rel(d1(fd, bc1_=bc1, bcn_=bcn).cpu().numpy(), oracle.cd10_np(f, dx, axis, 1, bc1, bcn)) < TOL, (bc1, bcn)
        assert _rel(d2(fd, bc1_=bc1, bcn_=bcn).cpu().numpy(), oracle.cd10_np(f, dx, axis, 2, bc1, bcn)) < TOL, (bc1, bcn)
    # host arrays (the unmodified-caller path)
    out = np.empty_like(f)
    d1(f, out, bc1_=1, bcn_=-1)
    assert _rel(out, oracle.cd10_np(f, dx, axis, 1, 1, -1)) < TOL
    with pytest.raises(pdo.PadeOpsError) as e:
        d1(fd, bc1_=2, bcn_=0)
    assert e.value.code == 324


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_cf90_nonperiodic_all_boundary_codes(pdo, oracle, axis):
    shape = {0: (4, 7, 33), 1: (4, 33, 7), 2: (33, 4, 7)}[axis]
    f = np.random.default_rng(10 + axis).standard_normal(shape)
    fd = _dev(f)
    h = pdo.cf90()
    assert h.init(33, periodic_=False) == 0
    fil = (h.filter1, h.filter2, h.filter3)[axis]
    for bc1, bcn in BCS:
        assert _rel(fil(fd, bc1_=bc1, bcn_=bcn).cpu().numpy(), oracle.cf90_np(f, axis, bc1, bcn)) < TOL, (bc1, bcn)


def test_dispatch_types_with_mixed_periodicity(pdo, oracle):
    """x periodic, y and z walls: derivatives%ddx/ddy/ddz and filters%filterx/y/z pick the right closure per axis."""
    nx, ny, nz = 32, 24, 16
    dx, dy, dz = 2 * np.pi / nx, 1.0 / (ny - 1), 1.0 / (nz - 1)
    f = np.random.default_rng(3).standard_normal((nz, ny, nx))
    fd = _dev(f)
    der = pdo.derivatives()
    der.init((nx, ny, nz), dx, dy, dz, True, False, False, "cd10", "cd10", "cd10")
    assert _rel(der.ddx(fd).cpu().numpy(), oracle.cd10(f, dx, 0, 1)) < TOL
    assert _rel(der.ddy(fd, None, 0, 0).cpu().numpy(), oracle.cd10_np(f, dy, 1, 1, 0, 0)) < TOL
    assert _rel(der.ddz(fd, None, 1, -1).cpu().numpy(), oracle.cd10_np(f, dz, 2, 1, 1, -1)) < TOL
    assert _rel(der.d2dy2(fd, None, -1, 1).cpu().numpy(), oracle.cd10_np(f, dy, 1, 2, -1, 1)) < TOL
    fil = pdo.filters()
    fil.init((nx, ny, nz), True, False, False, "cf90", "cf90", "cf90")
    assert _rel(fil.filterx(fd).cpu().numpy(), oracle.cf90(f, 0)) < TOL
    assert _rel(fil.filtery(fd, None, 0, 0).cpu().numpy(), oracle.cf90_np(f, 1, 0, 0)) < TOL
    assert _rel(fil.filterz(fd, None, 1, 1).cpu().numpy(), oracle.cf90_np(f, 2, 1, 1)) < TOL


def test_one_sided_closure_exact_on_quartics_on_the_gpu(pdo):
    n = 64
    dx = 1.0 / (n - 1)
    x = np.arange(n) * dx
    f = np.ascontiguousarray(np.broadcast_to((x ** 4)[None, None, :], (2, 3, n)))
    h = pdo.cd10()
    assert h.init(n, dx, periodic_=False) == 0
    assert np.abs(h.dd1(_dev(f)).cpu().numpy() - 4 * x ** 3).max() < 1e-11
    assert np.abs(h.d2d1(_dev(f)).cpu().numpy() - 12 * x ** 2).max() < 1e-8


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_cd06_nonperiodic_one_sided(pdo, oracle, axis):
    shape = {0: (3, 5, 24), 1: (3, 24, 5), 2: (24, 3, 5)}[axis]
    n, dx = 24, 0.1
    f = np.random.default_rng(20 + axis).standard_normal(shape)
    h = pdo.cd06()
    assert h.init(n, dx, periodic_=False) == 0
    got = (h.dd1, h.dd2, h.dd3)[axis](_dev(f)).cpu().numpy()
    assert _rel(got, oracle.cd06_np(f, dx, axis)) < TOL
    assert pdo.cd06().init(n, dx, periodic_=False, bc1_=1) == 1002


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_gaussian_nonperiodic_all_boundary_codes(pdo, oracle, axis):
    """gaussian%filter* with periodic = .false. (filters/gaussian.F90:215-330), also through the filters dispatch type and filter3D"""
    shape = {0: (5, 6, 40), 1: (5, 40, 6), 2: (40, 5, 6)}[axis]
    n = 40
    f = np.random.default_rng(10 + axis).standard_normal(shape)
    fd = _dev(f)
    h = pdo.gaussian()
    assert h.init(n, periodic_=False) == 0
    fil = (h.filter1, h.filter2, h.filter3)[axis]
    for bc1, bcn in BCS:
        assert _rel(fil(fd, bc1_=bc1, bcn_=bcn).cpu().numpy(), oracle.gaussian_np(f, axis, bc1, bcn)) < TOL, (bc1, bcn)
    if axis == 2:
        nz, ny, nx = 40, 16, 16
        g = np.random.default_rng(3).standard_normal((nz, ny, nx))
        gp = pdo.decomp_2d.init(nx, ny, nz, 1, 1)
        fl = pdo.filters()
        fl.init(gp, True, True, False, "cf90", "gaussian", "gaussian")
        ops = pdo.vector_ops()
        ops.init(gp, 0.1, 0.1, 0.1, "cd10")
        a = _dev(g)
        ops.filter3D(fl, a, 2, z_bc=(0, 1))
        assert _rel(a.cpu().numpy(), oracle.filter3D(g, 2, ("cf90", "gaussian", "gaussian"), (True, True, False), z_bc=(0, 1))) < TOL


def test_lstsq_filter_periodic_and_walls(pdo, oracle):
    """lstsq%filter1/2/3 (filters/lstsq.F90), alone and through the filters dispatch type (method "lstsq", filters.F90:111-118)"""
    n = 40
    for axis in (0, 1, 2):
        shape = {0: (5, 6, n), 1: (5, n, 6), 2: (n, 5, 6)}[axis]
        f = np.random.default_rng(20 + axis).standard_normal(shape)
        for periodic in (True, False):
            h = pdo.lstsq()
            assert h.init(n, periodic_=periodic) == 0
            got = (h.filter1, h.filter2, h.filter3)[axis](_dev(f)).cpu().numpy()
            assert _rel(got, oracle.lstsq(f, axis) if periodic else oracle.lstsq_np(f, axis)) < TOL, (axis, periodic)
    nz, ny, nx = 24, 16, 20
    g = np.random.default_rng(7).standard_normal((nz, ny, nx))
    gp = pdo.decomp_2d.init(nx, ny, nz, 1, 1)
    fl = pdo.filters()
    fl.init(gp, True, True, False, "lstsq", "cf90", "lstsq")
    assert _rel(fl.filterx(_dev(g)).cpu().numpy(), oracle.lstsq(g, 0)) < TOL
    assert _rel(fl.filterz(_dev(g)).cpu().numpy(), oracle.lstsq_np(g, 2)) < TOL
    with pytest.raises(pdo.PadeOpsError) as e:
        pdo.filters().init(gp, True, True, True, "spectral", "cf90", "cf90")
    assert e.value.code == 52


# ---------------------------------------------------------------------------------------------------------------------------------
# chunked fast path (csrc/np_chunk.cu): lines that are a multiple of 32 long take one fused pass; same bar against the oracle, and the
# two paths of the library against each other
# ---------------------------------------------------------------------------------------------------------------------------------
@pytest.fixture
def np_path(pdo):
    L = pdo.lib()
    yield lambda mode: L.pdo_debug_np_fast(mode)
    L.pdo_debug_np_fast(-1)


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("n", [64, 128, 256])
def test_chunked_fast_path_cd10_all_boundary_codes(pdo, oracle, np_path, axis, n):
    """ragged cross-sections (partial x tiles, lines that do not fill the last CTA) on purpose"""
    shape = {0: (5, 13, n), 1: (5, n, 37), 2: (n, 5, 21)}[axis]
    dx = 1.0 / (n - 1)
    f = np.random.default_rng(100 * axis + n).standard_normal(shape)
    fd = _dev(f)
    h = pdo.cd10()
    assert h.init(n, dx, periodic_=False) == 0
    d1 = (h.dd1, h.dd2, h.dd3)[axis]
    d2 = (h.d2d1, h.d2d2, h.d2d3)[axis]
    for bc1, bcn in BCS:
        np_path(1)
        g1, g2 = d1(fd, bc1_=bc1, bcn_=bcn).cpu().numpy(), d2(fd, bc1_=bc1, bcn_=bcn).cpu().numpy()
        np_path(0)
        s1, s2 = d1(fd, bc1_=bc1, bcn_=bcn).cpu().numpy(), d2(fd, bc1_=bc1, bcn_=bcn).cpu().numpy()
        r1, r2 = oracle.cd10_np(f, dx, axis, 1, bc1, bcn), oracle.cd10_np(f, dx, axis, 2, bc1, bcn)
        assert _rel(g1, r1) < TOL and _rel(g2, r2) < TOL, (bc1, bcn, _rel(g1, r1), _rel(g2, r2))
        assert _rel(s1, r1) < TOL and _rel(s2, r2) < TOL
        assert not np.array_equal(g1, s1) or n < 0      # the two paths associate differently: identical bits would mean one path ran twice


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("n", [128, 256])
def test_chunked_fast_path_cf90_and_cd06(pdo, oracle, np_path, axis, n):
    shape = {0: (3, 9, n), 1: (3, n, 41), 2: (n, 3, 19)}[axis]
    f = np.random.default_rng(7 * axis + n).standard_normal(shape)
    fd = _dev(f)
    cf = pdo.cf90()
    assert cf.init(n, periodic_=False) == 0
    fil = (cf.filter1, cf.filter2, cf.filter3)[axis]
    np_path(1)
    for bc1, bcn in BCS:
        assert _rel(fil(fd, bc1_=bc1, bcn_=bcn).cpu().numpy(), oracle.cf90_np(f, axis, bc1, bcn)) < TOL, (bc1, bcn)
    dx = 1.0 / (n - 1)
    c6 = pdo.cd06()
    assert c6.init(n, dx, periodic_=False) == 0
    d = (c6.dd1, c6.dd2, c6.dd3)[axis]
    assert _rel(d(fd).cpu().numpy(), oracle.cd06_np(f, dx, axis)) < TOL


def test_chunked_fast_path_exact_on_quartics_and_large_lines(pdo, np_path):
    """one-sided rows exact on x^4 (cd10.F90:33-77's closure is 4th-order at the wall), on 1024-point lines along every axis"""
    import torch
    n = 1024
    dx = 1.0 / (n - 1)
    x = np.arange(n) * dx
    np_path(1)
    h = pdo.cd10()
    assert h.init(n, dx, periodic_=False) == 0
    for axis, shape in ((0, (2, 40, n)), (1, (2, n, 48)), (2, (n, 2, 48))):
        sl = [None, None, None]
        sl[2 - axis] = slice(None)
        f = np.ascontiguousarray(np.broadcast_to((x ** 4)[tuple(sl)], shape))
        got = (h.dd1, h.dd2, h.dd3)[axis](torch.from_numpy(f).cuda()).cpu().numpy()
        assert np.abs(got - np.broadcast_to((4 * x ** 3)[tuple(sl)], shape)).max() < 2e-10, axis
