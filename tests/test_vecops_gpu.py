"""operators.F90 drop-ins on one GPU against the oracle, and the z-slab distributed solve emulated on one GPU: the slabs
exchange halo planes and edge pieces through ordinary device buffers exactly as the GPUs of a z-group do through peer
memory (tests/mp_worker.py covers the real thing on >= 2 GPUs)."""
import ctypes as C

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


@pytest.mark.parametrize("method", ["cd10", "cd06"])
def test_vector_ops_single_rank(pdo, oracle, method):
    nx, ny, nz = 64, 48, 32
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    gp = pdo.decomp_2d.init(nx, ny, nz, 1, 1)
    ops = pdo.vector_ops()
    ops.init(gp, dx, dy, dz, method)
    assert ops.zmode == 0
    u, v, w = (broadband((nz, ny, nx), seed=s) for s in (1, 2, 3))
    ud, vd, wd = _dev(u), _dev(v), _dev(w)
    for got, ref in zip(ops.gradient(ud), oracle.gradient(u, dx, dy, dz, method)):
        assert _relerr(got.cpu().numpy(), ref) < TOL
    assert _relerr(ops.divergence(ud, vd, wd).cpu().numpy(), oracle.divergence(u, v, w, dx, dy, dz, method)) < TOL
    assert _relerr(ops.curl(ud, vd, wd).cpu().numpy(), oracle.curl(u, v, w, dx, dy, dz, method)) < TOL
    ops.destroy()


@pytest.mark.parametrize("which,nslabs,n,n1", [(0, 2, 512, 96), (0, 4, 1024, 64), (1, 2, 512, 130), (2, 2, 256, 64), (2, 4, 1024, 34),
                                              (0, 8, 2048, 32), (0, 8, 8192, 32)])
def test_zslab_emulated(pdo, oracle, which, nslabs, n, n1):
    """cd10 d1 / d2 and cd06 d1 along z with the line cut into `nslabs` slabs; n1 includes partial 32-column tiles; the
    8192-point line (256 chunks: the 8-GPU bench shape) exists only in z-slab mode, its whole-line reference runs on the
    any-n kernels."""
    import torch
    d = 2 * np.pi / n
    f = broadband((n, 1, n1), seed=n + n1)
    fd = _dev(f)
    out = torch.empty_like(fd)
    if which == 2:
        h = pdo.cd06(); assert h.init(n, d) == 0
        ref = oracle.cd06(f, d, 2)
    else:
        h = pdo.cd10(); assert h.init(n, d) == 0
        ref = oracle.cd10(f, d, 2, 1 if which == 0 else 2)
    rc = pdo.lib().pdo_debug_zslab_emulate(h._h, which, C.c_void_p(fd.data_ptr()), C.c_void_p(out.data_ptr()), n1, n, nslabs,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, pdo.lib().pdo_last_error()
    assert _relerr(out.cpu().numpy(), ref) < TOL
    # and it is the same per-chunk arithmetic as the single-GPU kernels: agreement with the whole-line solve at rounding level
    whole = (h.dd3 if which != 1 else h.d2d3)(fd)
    assert _relerr(out.cpu().numpy(), whole.cpu().numpy()) < 1e-14


def test_zslab_unsupported_fails_loudly(pdo):
    import torch
    h = pdo.cd10(); assert h.init(256, 0.1) == 0
    fd = torch.zeros((256, 1, 32), dtype=torch.float64, device="cuda")
    out = torch.empty_like(fd)
    # 4 slabs of 64 planes = 2 chunks each: fewer than the 2 x 3 edge chunks CD10 needs
    rc = pdo.lib().pdo_debug_zslab_emulate(h._h, 0, C.c_void_p(fd.data_ptr()), C.c_void_p(out.data_ptr()), 32, 256, 4, C.c_void_p(0))
    assert rc != 0


@pytest.mark.parametrize("numtimes", [1, 2, 3])
@pytest.mark.parametrize("methods", [("cf90", "cf90", "cf90"), ("gaussian", "cf90", "gaussian")])
def test_filter3d_single_rank(pdo, oracle, numtimes, methods):
    """filter3D (operators.F90:158-224) in place on a y-pencil field; every parity of the pass count lands in `arr`."""
    nx, ny, nz = 64, 48, 32
    d = 2 * np.pi / nx
    gp = pdo.decomp_2d.init(nx, ny, nz, 1, 1)
    ops = pdo.vector_ops()
    ops.init(gp, d, d, d, "cd10")
    fil = pdo.filters()
    fil.init(gp, True, True, True, *methods)
    f = broadband((nz, ny, nx), seed=7)
    fd = _dev(f)
    got = ops.filter3D(fil, fd, numtimes)
    assert got.data_ptr() == fd.data_ptr()
    assert _relerr(fd.cpu().numpy(), oracle.filter3D(f, numtimes, methods)) < TOL
    ops.destroy()


def test_filter3d_nonperiodic_z_and_mismatched_filters(pdo, oracle):
    nx, ny, nz = 32, 32, 40
    d = 2 * np.pi / nx
    gp = pdo.decomp_2d.init(nx, ny, nz, 1, 1)
    ops = pdo.vector_ops()
    ops.init(gp, d, d, d, "cd10")
    fil = pdo.filters()
    fil.init(gp, True, True, False, "cf90", "cf90", "cf90")
    f = broadband((nz, ny, nx), seed=9)
    fd = _dev(f)
    ops.filter3D(fil, fd, 2, z_bc=(1, -1))
    assert _relerr(fd.cpu().numpy(), oracle.filter3D(f, 2, periodic=(True, True, False), z_bc=(1, -1))) < TOL
    other = pdo.filters()
    other.init((nx, ny, nz + 8), True, True, True, "cf90", "cf90", "cf90")
    with pytest.raises(pdo.PadeOpsError) as e:
        ops.filter3D(other, fd)
    assert e.value.code == 234   # operators.F90:171-174
    ops.destroy()


def test_vector_ops_host_arrays(pdo, oracle):
    """the unmodified caller hands HOST arrays to gradient / divergence / curl / filter3D."""
    nx, ny, nz = 32, 24, 16
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    gp = pdo.decomp_2d.init(nx, ny, nz, 1, 1)
    ops = pdo.vector_ops()
    ops.init(gp, dx, dy, dz, "cd10")
    L = pdo.lib()
    u, v, w = (broadband((nz, ny, nx), seed=s) for s in (4, 5, 6))
    P = lambda a: C.c_void_p(a.ctypes.data)
    gx, gy, gz = (np.empty_like(u) for _ in range(3))
    assert L.pdo_operators_gradient(ops._h, P(u), P(gx), P(gy), P(gz), C.c_void_p(0)) == 0
    for got, ref in zip((gx, gy, gz), oracle.gradient(u, dx, dy, dz, "cd10")):
        assert _relerr(got, ref) < TOL
    div = np.empty_like(u)
    assert L.pdo_operators_divergence(ops._h, P(u), P(v), P(w), P(div), C.c_void_p(0)) == 0
    assert _relerr(div, oracle.divergence(u, v, w, dx, dy, dz, "cd10")) < TOL
    cu = np.empty((3,) + u.shape)
    assert L.pdo_operators_curl(ops._h, P(u), P(v), P(w), P(cu), C.c_void_p(0)) == 0
    assert _relerr(cu, oracle.curl(u, v, w, dx, dy, dz, "cd10")) < TOL
    fil = pdo.filters()
    fil.init(gp, True, True, True, "cf90", "cf90", "cf90")
    a = u.copy()
    assert L.pdo_operators_filter3d(ops._h, fil._h, P(a), 1, None, None, None, C.c_void_p(0)) == 0
    assert _relerr(a, oracle.filter3D(u, 1)) < TOL
    ops.destroy()
