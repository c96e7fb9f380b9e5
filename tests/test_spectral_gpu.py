"""GPU parity tests: pencil transposes (bit-exact), fft_3d passes and PoissonPeriodic against the oracle
(numpy pocketfft standing in for FFTW; tolerance 1e-12 relative to max|ref|)."""
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


@pytest.fixture(scope="module")
def comm(pdo):
    pdo.decomp_2d.comm_init()
    return pdo.decomp_2d


@pytest.mark.parametrize("cplx", [False, True])
def test_single_rank_transposes_are_bit_exact_copies(pdo, oracle, comm, cplx):
    from padeops_b200 import decomp as dc
    if comm.nproc != 1:
        pytest.skip("single-rank test")
    nx, ny, nz = 20, 12, 9
    gp = pdo.decomp_info(nx, ny, nz, 1, 1)
    assert {k: getattr(gp, k) for k in ("xsz", "ysz", "zsz")} == {k: oracle.decomp_info(nx, ny, nz, 1, 1, 0)[k] for k in ("xsz", "ysz", "zsz")}
    rng = np.random.default_rng(0)
    a = rng.standard_normal((nz, ny, nx))
    if cplx:
        a = a + 1j * rng.standard_normal((nz, ny, nx))
    for d, fn in enumerate((dc.transpose_x_to_y, dc.transpose_y_to_x, dc.transpose_y_to_z, dc.transpose_z_to_y)):
        ref = oracle.transpose(d, nx, ny, nz, 1, 1, [a])[0]
        got = fn(_dev(a), None, gp).cpu().numpy()
        assert np.array_equal(got, ref)
        # host-pointer (drop-in) path
        out = np.empty_like(a)
        fn(a, out, gp)
        assert np.array_equal(out, ref)


@pytest.mark.parametrize("shape", [(16, 32, 64), (12, 20, 48), (9, 10, 14), (64, 64, 64)])
def test_fft3d_passes_match_numpy(pdo, comm, shape):
    if comm.nproc != 1:
        pytest.skip("single-rank test")
    nz, ny, nx = shape
    f = broadband(shape, seed=nx)
    ft = pdo.fft_3d()
    assert ft.init(nx, ny, nz, "x", 0.1, 0.1, 0.1) == 0
    assert ft.get_complex_output_size() == (nx // 2 + 1, ny, nz)
    fd = _dev(f)
    # fft3_x2z: unnormalised r2c-x, c2c-y, c2c-z (fft_3d.F90:588-613)
    ref3 = np.fft.fft(np.fft.fft(np.fft.rfft(f, axis=2), axis=1), axis=0)
    got3 = ft.fft3_x2z(fd)
    assert _relerr(got3.cpu().numpy(), ref3) < TOL
    assert np.array_equal(fd.cpu().numpy(), f)  # intent(in) input untouched
    back = ft.ifft3_z2x(got3)
    assert _relerr(back.cpu().numpy(), f) < TOL
    assert _relerr(got3.cpu().numpy(), ref3) < TOL  # ifft3_z2x must not clobber its input either
    # fft2_x2y / ifft2_y2x (fft_3d.F90:645-663, 616-643)
    ref2 = np.fft.fft(np.fft.rfft(f, axis=2), axis=1)
    got2 = ft.fft2_x2y(fd)
    assert _relerr(got2.cpu().numpy(), ref2) < TOL
    assert _relerr(ft.ifft2_y2x(got2).cpu().numpy(), f) < TOL
    assert _relerr(got2.cpu().numpy(), ref2) < TOL
    if nx % 2 == 0:
        r2 = ref2.copy()
        r2[:, :, nx // 2] = 0
        refo = np.fft.irfft(np.fft.ifft(r2, axis=1), n=nx, axis=2)
        assert _relerr(ft.ifft2_y2x(got2, setOddBall=True).cpu().numpy(), refo) < TOL


# Power-of-two extents on a slab grid take the hand-written passes of csrc/fft2d.cu (Stockham radix-8 kernels; every other shape
# goes to cuFFT): one case per compiled transform length of the contiguous pass (nx = 16 ... 2048), of the strided pass as y
# (16 ... 1024) and as z (16 ... 1024), with plane / line counts that leave partial tiles.
_HW_SHAPES = ([(3, 16, nx) for nx in (16, 32, 64, 128, 256, 512, 1024, 2048)] + [(3, ny, 16 + 16 * (i % 2)) for i, ny in enumerate((32, 64, 128, 256, 512, 1024))]
              + [(nz, 16, 16) for nz in (16, 32, 64, 128, 256, 512, 1024)] + [(7, 64, 64), (5, 32, 512), (64, 128, 256)])


@pytest.mark.parametrize("shape", _HW_SHAPES)
def test_handwritten_fft_passes_match_numpy(pdo, comm, shape):
    if comm.nproc != 1:
        pytest.skip("single-rank test")
    nz, ny, nx = shape
    f = broadband(shape, seed=nx + ny)
    ft = pdo.fft_3d()
    assert ft.init(nx, ny, nz, "x", 0.1, 0.1, 0.1) == 0
    fd = _dev(f)
    ref2 = np.fft.fft(np.fft.rfft(f, axis=2), axis=1)
    got2 = ft.fft2_x2y(fd)
    assert _relerr(got2.cpu().numpy(), ref2) < TOL
    assert np.array_equal(fd.cpu().numpy(), f)
    back2 = ft.ifft2_y2x(got2)
    assert _relerr(back2.cpu().numpy(), f) < TOL
    assert _relerr(got2.cpu().numpy(), ref2) < TOL           # intent(in)
    # c2r ignores the imaginary parts of the x modes 0 and nx/2 after the y pass, as FFTW's c2r does; a non-Hermitian input pins that
    rng = np.random.default_rng(nx)
    g = ref2 + 0.1 * (rng.standard_normal(ref2.shape) + 1j * rng.standard_normal(ref2.shape))
    refg = np.fft.irfft(np.fft.ifft(g, axis=1), n=nx, axis=2)
    assert _relerr(ft.ifft2_y2x(_dev(g)).cpu().numpy(), refg) < TOL
    r2 = g.copy()
    r2[:, :, nx // 2] = 0
    refo = np.fft.irfft(np.fft.ifft(r2, axis=1), n=nx, axis=2)
    assert _relerr(ft.ifft2_y2x(_dev(g), setOddBall=True).cpu().numpy(), refo) < TOL
    ref3 = np.fft.fft(ref2, axis=0)
    got3 = ft.fft3_x2z(fd)
    assert _relerr(got3.cpu().numpy(), ref3) < TOL
    assert _relerr(ft.ifft3_z2x(got3).cpu().numpy(), f) < TOL
    assert _relerr(got3.cpu().numpy(), ref3) < TOL


def test_poisson_manufactured_solution(pdo, oracle, comm):
    # tests/test_PoissonPeriodic.F90:109-118: 64 x 32 x 16, (l,m,n) = (6,3,1)
    if comm.nproc != 1:
        pytest.skip("single-rank test")
    nx, ny, nz = 64, 32, 16
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    x, y, z = np.arange(nx) * dx, np.arange(ny) * dy, np.arange(nz) * dz
    ftrue = np.sin(6 * x)[None, None, :] * np.cos(3 * y)[None, :, None] * np.sin(z)[:, None, None]
    rhs = -(36 + 9 + 1) * ftrue
    for dir_id in (1, 2):
        po = pdo.PoissonPeriodic()
        po.init(dx, dy, dz, (nx, ny, nz), dir_id)
        import torch
        out = torch.empty((nz, ny, nx), dtype=torch.float64, device="cuda")
        po.poisson_solve(_dev(rhs), out)
        assert np.abs(out.cpu().numpy() - ftrue).max() < 1e-12
        assert _relerr(out.cpu().numpy(), oracle.poisson_solve(rhs, dx, dy, dz)) < TOL


def test_poisson_z_pencil_entry(pdo, oracle, comm):
    """dir_id = 3 (z-pencil in / out, PoissonPeriodic.F90:151-154): on one rank the three pencils coincide, so the result
    must equal the oracle's like dir_id = 1 does."""
    if comm.nproc != 1:
        pytest.skip("single-rank test")
    import torch
    nx, ny, nz = 32, 24, 16
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    rhs = broadband((nz, ny, nx), seed=5)
    rhs -= rhs.mean()
    po = pdo.PoissonPeriodic()
    po.init(dx, dy, dz, (nx, ny, nz), 3)
    out = torch.empty((nz, ny, nx), dtype=torch.float64, device="cuda")
    po.poisson_solve(_dev(rhs), out)
    assert _relerr(out.cpu().numpy(), oracle.poisson_solve(rhs, dx, dy, dz)) < TOL


@pytest.mark.parametrize("shape", [(32, 48, 64), (10, 12, 18), (128, 128, 128)])
def test_poisson_broadband_matches_oracle(pdo, oracle, comm, shape):
    if comm.nproc != 1:
        pytest.skip("single-rank test")
    nz, ny, nx = shape
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    rhs = broadband(shape, seed=7)
    po = pdo.PoissonPeriodic()
    po.init(dx, dy, dz, (nx, ny, nz), 1)
    ref = oracle.poisson_solve(rhs, dx, dy, dz)
    got = po.poisson_solve(_dev(rhs)).cpu().numpy()   # in-place specific
    assert _relerr(got, ref) < TOL
    host = rhs.copy()
    po.poisson_solve(host)                            # host-pointer drop-in path
    assert _relerr(host, ref) < TOL
    # property: Laplacian of the solution gives back the zero-mean part of rhs, to spectral accuracy
    k2 = (oracle.wavenums(nx, dx)[None, None, :nx // 2 + 1] ** 2 + oracle.wavenums(ny, dy)[None, :, None] ** 2 +
          oracle.wavenums(nz, dz)[:, None, None] ** 2)
    lap = np.fft.irfftn(-k2 * np.fft.rfftn(got), s=shape)
    assert np.abs(lap - (rhs - rhs.mean())).max() < 1e-10


def test_poisson_modified_wavenumbers(pdo, oracle, comm):
    """Get_ModKz-style callback: CD06-staggered modified wavenumber in z (tests/test_PoissonPeriodic.F90:12-25)."""
    if comm.nproc != 1:
        pytest.skip("single-rank test")
    nx, ny, nz = 32, 16, 24
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    kz = oracle.wavenums(nz, dz)
    w = kz * dz
    a, b, al = 63 / 62, (17 / 62) / 3, 9 / 62
    kzmod = (2 * a * np.sin(w / 2) + 2 * b * np.sin(3 * w / 2)) / (1 + 2 * al * np.cos(w)) / dz
    rhs = broadband((nz, ny, nx), seed=11)
    po = pdo.PoissonPeriodic()
    po.init(dx, dy, dz, (nx, ny, nz), 1, modkz=kzmod)
    got = po.poisson_solve(_dev(rhs)).cpu().numpy()
    h = np.fft.fft(np.fft.fft(np.fft.rfft(rhs, axis=2), axis=1), axis=0)
    h = oracle.poisson_multiply(h, oracle.wavenums(nx, dx)[:nx // 2 + 1], oracle.wavenums(ny, dy), kzmod, True)
    ref = np.fft.irfft(np.fft.ifft(np.fft.ifft(h, axis=0), axis=1), n=nx, axis=2)
    assert _relerr(got, ref) < TOL
