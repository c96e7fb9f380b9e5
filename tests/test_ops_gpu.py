"""GPU parity tests for the compact line operators: the CUDA path (through the C ABI) against the CPU
oracle on the same seeded inputs.  Bar: 1e-12 max error relative to max|ref| (north_star), double precision."""
import numpy as np
import pytest

from conftest import broadband

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _relerr(got, ref):
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)


def _shape_for(axis, n, na, nb):
    # Fortran f(n,na,nb)/f(na,n,nb)/f(na,nb,n) as C-order (n3,n2,n1)
    return {0: (nb, na, n), 1: (nb, n, na), 2: (n, nb, na)}[axis]


# n: 128/64 → chunk 32; 48 → 16; 40/24 → 8; 20/11/100/9 → generic any-n kernels
LINE_LENGTHS = [128, 64, 48, 40, 24, 20, 11, 100]


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("n", LINE_LENGTHS)
def test_cd10_matches_oracle(pdo, oracle, axis, n):
    dx = 2 * np.pi / n
    op = pdo.cd10()
    assert op.init(n, dx) == 0
    for (na, nb) in [(5, 3), (32, 6)]:
        f = broadband(_shape_for(axis, n, na, nb), seed=n + axis)
        fd = _dev(f)
        for which, fn in ((1, (op.dd1, op.dd2, op.dd3)[axis]), (2, (op.d2d1, op.d2d2, op.d2d3)[axis])):
            got = fn(fd).cpu().numpy()
            ref = oracle.cd10(f, dx, axis, which)
            assert _relerr(got, ref) < TOL, (axis, n, na, nb, which, _relerr(got, ref))


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("n", [128, 48, 40, 20, 7, 100])
def test_cd06_matches_oracle(pdo, oracle, axis, n):
    dx = 0.05
    op = pdo.cd06()
    assert op.init(n, dx) == 0
    f = broadband(_shape_for(axis, n, 7, 4), seed=n)
    got = (op.dd1, op.dd2, op.dd3)[axis](_dev(f)).cpu().numpy()
    assert _relerr(got, oracle.cd06(f, dx, axis)) < TOL


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("n", [128, 64, 48, 40, 20, 11, 100])
def test_filters_match_oracle(pdo, oracle, axis, n):
    f = broadband(_shape_for(axis, n, 9, 5), seed=3 * n)
    fd = _dev(f)
    cf = pdo.cf90()
    assert cf.init(n) == 0
    got = (cf.filter1, cf.filter2, cf.filter3)[axis](fd).cpu().numpy()
    # CF90's LHS is nearly singular at the Nyquist mode (cond ~ 2e3): same 1e-12 bar, still met
    assert _relerr(got, oracle.cf90(f, axis)) < TOL
    ga = pdo.gaussian()
    assert ga.init(n) == 0
    got = (ga.filter1, ga.filter2, ga.filter3)[axis](fd).cpu().numpy()
    assert _relerr(got, oracle.gaussian(f, axis)) < TOL


def test_chunked_and_generic_kernels_agree(pdo):
    """Two independent CUDA formulations (register chunk engine vs one-thread-per-line) on the same input."""
    import ctypes as C
    import torch
    from padeops_b200._lib import check, lib, ptr, stream_ptr
    n = 128
    op = pdo.cd10()
    assert op.init(n, 0.1) == 0
    for axis in range(3):
        f = _dev(broadband(_shape_for(axis, n, 6, 5), seed=axis))
        a = (op.dd1, op.dd2, op.dd3)[axis](f)
        b = torch.empty_like(f)
        check(lib().pdo_debug_cd10_generic(op._h, 1, axis, ptr(f), ptr(b), 6, 5, stream_ptr()))
        assert (a - b).abs().max().item() < 1e-12 * a.abs().max().item()


def test_host_pointer_dropin_path(pdo, oracle):
    """Host buffers in, host buffers out: the call a Fortran caller without device fields makes."""
    n = 64
    f = broadband((4, 6, n))
    op = pdo.cd10()
    assert op.init(n, 0.2) == 0
    out = np.empty_like(f)
    op.dd1(f, out)
    assert _relerr(out, oracle.cd10(f, 0.2, 0, 1)) < TOL


def test_degenerate_axis_and_bad_bc(pdo):
    import torch
    op = pdo.cd10()
    assert op.init(1, 0.1) == 0  # n == 1 is legal: derivative is zero (cd10.F90:2037-2040)
    f = torch.rand(3, 4, 1, dtype=torch.float64, device="cuda")
    assert op.dd1(f).abs().max().item() == 0.0
    cf = pdo.cf90()
    assert cf.init(1) == 0       # filter returns its input (cf90.F90:1028-1031)
    assert torch.equal(cf.filter1(f), f)
    op2 = pdo.cd10()
    assert op2.init(16, 0.1) == 0
    g = torch.rand(2, 2, 16, dtype=torch.float64, device="cuda")
    with pytest.raises(pdo.PadeOpsError) as e:
        op2.dd1(g, bc1_=2)
    assert e.value.code == 324   # cd10.F90:2044-2046


@pytest.mark.parametrize("n", [32, 64, 40, 21])
@pytest.mark.parametrize("cplx", [False, True])
def test_staggered_ops_match_oracle(pdo, oracle, n, cplx):
    dz = 2 * np.pi / n
    st = pdo.cd06stagg()
    st.init(n, dz)
    rng = np.random.default_rng(n)
    n2, n1 = 5, 9
    def field(planes):
        a = broadband((planes, n2, n1), seed=planes)
        if cplx:
            a = a + 1j * broadband((planes, n2, n1), seed=planes + 1)
        return a
    fC, fE = field(n), field(n + 1)
    fE_per = fE.copy()
    fE_per[n] = fE_per[0]  # the convention plane n+1 == plane 1 ...
    cases = [("ddz_E2C", st.ddz_E2C, fE_per), ("ddz_C2E", st.ddz_C2E, fC), ("interp_E2C", st.InterpZ_E2C, fE_per),
             ("interp_C2E", st.InterpZ_C2E, fC), ("d2dz2_C2C", st.d2dz2_C2C, fC), ("d2dz2_E2E", st.d2dz2_E2E, fE_per),
             # ... and WITHOUT it: the reference reads plane n+1 as stored, so must we
             ("ddz_E2C", st.ddz_E2C, fE), ("interp_E2C", st.InterpZ_E2C, fE), ("d2dz2_E2E", st.d2dz2_E2E, fE)]
    for name, fn, fin in cases:
        got = fn(_dev(fin)).cpu().numpy()
        ref = oracle.stagg(name, fin, n, dz)
        assert got.shape == ref.shape
        assert _relerr(got, ref) < TOL, (name, n, cplx, _relerr(got, ref))


def test_dispatch_types_on_a_pencil(pdo, oracle):
    """derivatives%ddx/ddy/ddz + filters%filterx/y/z with per-pencil sizes (derivatives.F90:447-569)."""
    class gp:
        xsz, ysz, zsz = (64, 6, 5), (6, 48, 5), (6, 5, 40)
    dx, dy, dz = 0.1, 0.2, 0.3
    der = pdo.derivatives()
    der.init(gp, dx, dy, dz, True, True, True, "cd10", "cd06", "cd10")
    fil = pdo.filters()
    fil.init(gp, True, True, True, "cf90", "gaussian", "cf90")
    fx, fy, fz = (broadband(tuple(reversed(s)), seed=i) for i, s in enumerate((gp.xsz, gp.ysz, gp.zsz)))
    assert _relerr(der.ddx(_dev(fx)).cpu().numpy(), oracle.cd10(fx, dx, 0, 1)) < TOL
    assert _relerr(der.ddy(_dev(fy)).cpu().numpy(), oracle.cd06(fy, dy, 1)) < TOL
    assert _relerr(der.ddz(_dev(fz)).cpu().numpy(), oracle.cd10(fz, dz, 2, 1)) < TOL
    assert _relerr(der.d2dx2(_dev(fx)).cpu().numpy(), oracle.cd10(fx, dx, 0, 2)) < TOL
    with pytest.raises(pdo.PadeOpsError):
        der.d2dy2(_dev(fy))  # "CD06 is incomplete right now" (derivatives.F90:525)
    assert _relerr(fil.filterx(_dev(fx)).cpu().numpy(), oracle.cf90(fx, 0)) < TOL
    assert _relerr(fil.filtery(_dev(fy)).cpu().numpy(), oracle.gaussian(fy, 1)) < TOL
    assert _relerr(fil.filterz(_dev(fz)).cpu().numpy(), oracle.cf90(fz, 2)) < TOL


def test_cfg1_shape_128cubed_sine_field(pdo, oracle):
    """BASELINE config 1 (tests/test_derivatives_parallel.F90): CD10 + CD06 on 128^3, f = sin x sin y cos z."""
    n = 128
    d = 2 * np.pi / n
    x = np.arange(n) * d
    f = np.sin(x)[None, None, :] * np.sin(x)[None, :, None] * np.cos(x)[:, None, None]
    fd = _dev(f)
    c10, c06 = pdo.cd10(), pdo.cd06()
    assert c10.init(n, d) == 0 and c06.init(n, d) == 0
    exact = [np.cos(x)[None, None, :] * np.sin(x)[None, :, None] * np.cos(x)[:, None, None],
             np.sin(x)[None, None, :] * np.cos(x)[None, :, None] * np.cos(x)[:, None, None],
             -np.sin(x)[None, None, :] * np.sin(x)[None, :, None] * np.sin(x)[:, None, None]]
    for ax in range(3):
        got = (c10.dd1, c10.dd2, c10.dd3)[ax](fd).cpu().numpy()
        assert _relerr(got, oracle.cd10(f, d, ax, 1)) < TOL
        assert np.abs(got - exact[ax]).max() < 1e-12  # 10th order on a k=1 mode: at round-off
        got = (c06.dd1, c06.dd2, c06.dd3)[ax](fd).cpu().numpy()
        assert _relerr(got, oracle.cd06(f, d, ax)) < TOL


@pytest.mark.parametrize("n", [512, 1024, 2048])
def test_full_size_lines_single_mode_identity(pdo, n):
    """Size-independent property at BASELINE line lengths: for f = cos(kx) every periodic compact operator
    returns the analytic symbol times f (tests/test_cf90.F90:108-116)."""
    import torch
    k = n // 5
    d = 2 * np.pi / n
    w = k * d
    x = torch.arange(n, dtype=torch.float64, device="cuda") * d
    c, s = torch.cos(k * x), torch.sin(k * x)
    a, b, cc = (17 / 12) / 2, (101 / 150) / 4, (1 / 100) / 6
    kp = (2 * a * np.sin(w) + 2 * b * np.sin(2 * w) + 2 * cc * np.sin(3 * w)) / (1 + np.cos(w) + 0.1 * np.cos(2 * w)) / d
    co = (9.9965e-1, 6.6652e-1, 1.6674e-1, 4.0e-5, -5.0e-6)
    T = (co[0] + 2 * sum(co[m] * np.cos(m * w) for m in range(1, 5))) / (1 + 2 * 6.6624e-1 * np.cos(w) + 2 * 1.6688e-1 * np.cos(2 * w))
    c10, cf = pdo.cd10(), pdo.cf90()
    assert c10.init(n, d) == 0 and cf.init(n) == 0
    for ax in range(3):
        shape = [24, 40, 24]
        shape[2 - ax] = n
        view = [1, 1, 1]
        view[2 - ax] = n
        fc = c.view(view).expand(shape).contiguous()
        fs = s.view(view).expand(shape).contiguous()
        got = (c10.dd1, c10.dd2, c10.dd3)[ax](fc)
        assert (got + kp * fs).abs().max().item() < 1e-12 * kp
        got = (cf.filter1, cf.filter2, cf.filter3)[ax](fc)
        assert (got - T * fc).abs().max().item() < 1e-12


def test_host_pointer_pipelined_path(pdo, oracle, monkeypatch):
    """The chunked full-duplex host path (capi_ops.cu: apply_host_pipelined) on every axis, with ragged last pieces,
    edge planes (staggered ops) and both split directions; thresholds shrunk so small fields take it."""
    monkeypatch.setenv("PDO_PIPE_MIN_MB", "0.001")
    monkeypatch.setenv("PDO_PIPE_CHUNK_MB", "0.05")
    n = 64
    d = 0.1
    op, cf = pdo.cd10(), pdo.cf90()
    assert op.init(n, d) == 0 and cf.init(n) == 0
    for axis, shape in [(0, (7, 33, n)), (1, (9, n, 40)), (1, (2, n, 300)), (2, (n, 13, 29))]:
        f = broadband(shape, seed=axis)
        out = np.empty_like(f)
        (op.dd1, op.dd2, op.dd3)[axis](f, out)
        assert _relerr(out, oracle.cd10(f, d, axis, 1)) < TOL, (axis, shape)
        (cf.filter1, cf.filter2, cf.filter3)[axis](f, out)
        assert _relerr(out, oracle.cf90(f, axis)) < TOL, (axis, shape)
    st = pdo.cd06stagg()
    st.init(n, d)
    fE = broadband((n + 1, 11, 37), seed=5) + 1j * broadband((n + 1, 11, 37), seed=6)
    fC = fE[:n].copy()
    got = np.empty((n, 11, 37), dtype=np.complex128)
    st.ddz_E2C(fE, got)
    assert _relerr(got, oracle.stagg("ddz_E2C", fE, n, d)) < TOL
    got = np.empty((n + 1, 11, 37), dtype=np.complex128)
    st.InterpZ_C2E(fC, got)
    assert _relerr(got, oracle.stagg("interp_C2E", fC, n, d)) < TOL
