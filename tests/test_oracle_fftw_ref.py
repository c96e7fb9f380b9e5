"""(c) oracle pinning against REFERENCE-BUILT arithmetic: FFTW 3.3.5 compiled from the reference's own vendored tarball
(oracle/_ref, `make -C oracle ref`) and driven with the plan tuples of utilities/fft_3d.F90:256-306 (oracle/fftw_ref.py).

* where oracle/_ref exists (the build container; it also travels to the GPU box): the numpy/pocketfft stand-in used by oracle.py
  and igrid_oracle.py is checked against FFTW pass by pass, on PoissonPeriodic, and on a whole igrid substep with every FFT of
  the restatement swapped for FFTW;
* everywhere: the committed fixture tests/golden/fftw_ref_golden.npz (FFTW outputs, tests/golden/make_fftw_ref_golden.py) pins the
  oracle (CPU) and the CUDA path (cuFFT, `-m gpu`).
Bar: 1e-12 relative to max|ref| (north_star); measured 3e-16 ... 2e-15."""
import importlib.util
import os
import types

import numpy as np
import pytest

from oracle import fftw_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "fftw_ref_golden.npz")
TOL = 1e-12
need_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/lib/libfftw3.so not built (needs /root/reference: make -C oracle ref)")


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _maker():
    spec = importlib.util.spec_from_file_location("make_fftw_ref_golden", os.path.join(ROOT, "tests", "golden", "make_fftw_ref_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@need_ref
def test_ref_is_the_reference_tarball_version():
    assert R.version().startswith("fftw-3.3.5")     # dependencies/fftw-3.3.5.tar.gz


@need_ref
@pytest.mark.parametrize("shape", [(64, 32, 16), (16, 12, 10), (18, 9, 7), (33, 10, 6), (48, 48, 48)])
def test_pocketfft_stand_in_matches_fftw_pass_by_pass(oracle, shape):
    nx, ny, nz = shape
    f = np.random.default_rng(nx).standard_normal((nz, ny, nx))
    F = R.FFT3D(nx, ny, nz)
    h3 = F.fft3_x2z(f)
    assert _rel(np.fft.fft(np.fft.fft(np.fft.rfft(f, axis=2), axis=1), axis=0), h3) < TOL
    h2 = F.fft2_x2y(f)
    assert _rel(np.fft.fft(np.fft.rfft(f, axis=2), axis=1), h2) < TOL
    g = h3 * (1.0 + 0.25j)
    ref = F.ifft3_z2x(g)
    got = np.fft.irfft(np.fft.ifft(np.fft.ifft(g, axis=0), axis=1), n=nx, axis=2)
    assert _rel(got, ref) < TOL
    d = [2 * np.pi / m for m in shape]
    assert _rel(oracle.poisson_solve(f, *d), R.poisson_solve(f, *d, fft=F)) < TOL
    F.destroy()


@need_ref
def test_igrid_spectral_wrappers_match_fftw():
    """spectral%fft / ifft (spectral.F90:1413-1453 -> fft2_x2y / ifft2_y2x) of the igrid restatement"""
    from oracle import igrid_oracle as IG
    nx, ny, nz = 24, 16, 12
    sp = IG.Spectral(nx, ny, nz, 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz, True, 2.0 / 3.0, False)
    F = R.FFT3D(nx, ny, nz)
    f = np.random.default_rng(3).standard_normal((nz, ny, nx))
    h = sp.fft(f)
    assert _rel(h, F.fft2_x2y(f)) < TOL
    assert _rel(sp.ifft(h * (1 + 0.5j)), F.ifft2_y2x(h * (1 + 0.5j))) < TOL
    assert _rel(sp.ifft(h * (1 + 0.5j), True), F.ifft2_y2x(h * (1 + 0.5j), setOddBall=True)) < TOL


@need_ref
@pytest.mark.parametrize("vert,adv,tscheme", [(1, 1, 1), (2, 0, 2)])
def test_whole_igrid_step_on_fftw_arithmetic(vert, adv, tscheme):
    """The igrid restatement executes ~150 transforms per step; run it once on pocketfft and once with EVERY np.fft call of the
    module routed to FFTW 3.3.5: the two runs must agree far inside the parity bar, i.e. the oracle's answer does not depend on
    which FFT library stands behind it."""
    from oracle import igrid_oracle as IG

    class NP(types.ModuleType):
        def __getattr__(self, k):
            return getattr(np, k)
    shim = NP("numpy_with_fftw")
    shim.fft = R.npfft
    n = 16
    x = np.arange(n) * 2 * np.pi / n
    X, Y = x[None, None, :], x[None, :, None]
    zc, ze = (np.arange(n) + 0.5) * 2 * np.pi / n, np.arange(n + 1) * 2 * np.pi / n
    u = np.sin(X) * np.cos(Y) * np.cos(zc)[:, None, None]
    v = -np.cos(X) * np.sin(Y) * np.cos(zc)[:, None, None]
    w = 0.1 * np.sin(2 * X) * np.sin(Y) * np.sin(ze)[:, None, None]
    res = []
    for use_fftw in (False, True):
        old = IG.np
        IG.np = shim if use_fftw else np
        try:
            g = IG.IGrid(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, 100.0, u, v, w, TimeSteppingScheme=tscheme, AdvectionTerm=adv, NumericalSchemeVert=vert)
            for _ in range(2):
                g.timeAdvance(0.02)
            res.append((g.u.copy(), g.v.copy(), g.w.copy(), g.uhat.copy(), g.what.copy()))
        finally:
            IG.np = old
    for a, b in zip(*res):
        assert _rel(a, b) < TOL


def test_oracle_reproduces_the_fftw_golden_vectors(oracle, gold):
    """CPU, everywhere (the fixture travels): oracle.py's transforms against stored FFTW 3.3.5 outputs."""
    M = _maker()
    assert str(gold["fftw_version"]).startswith("fftw-3.3.5")
    for (nx, ny, nz) in M.SHAPES:
        f = M.inputs(nx, ny, nz)
        tag = f"{nx}x{ny}x{nz}"
        h3 = np.fft.fft(np.fft.fft(np.fft.rfft(f, axis=2), axis=1), axis=0)
        h2 = np.fft.fft(np.fft.rfft(f, axis=2), axis=1)
        assert _rel(h3, gold[f"fft3_x2z_{tag}"]) < TOL
        assert _rel(h2, gold[f"fft2_x2y_{tag}"]) < TOL
        g3 = gold[f"fft3_x2z_{tag}"] * (1.0 + 0.25j)
        assert _rel(np.fft.irfft(np.fft.ifft(np.fft.ifft(g3, axis=0), axis=1), n=nx, axis=2), gold[f"ifft3_z2x_{tag}"]) < TOL
        d = (2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz)
        assert _rel(oracle.poisson_solve(f, *d), gold[f"poisson_{tag}"]) < TOL


@need_ref
def test_golden_file_is_what_the_committed_script_generates(gold):
    M = _maker()
    nx, ny, nz = M.SHAPES[1]
    F = R.FFT3D(nx, ny, nz)
    # FFTW_MEASURE (the reference's planner flag) picks codelets by timing: a regeneration agrees to rounding, not bit for bit
    assert _rel(F.fft3_x2z(M.inputs(nx, ny, nz)), gold[f"fft3_x2z_{nx}x{ny}x{nz}"]) < 1e-14


@pytest.mark.gpu
def test_cuda_path_reproduces_the_fftw_golden_vectors(pdo, gold):
    """fft_3d / PoissonPeriodic on the GPU (cuFFT + the fused multiply) against FFTW 3.3.5's stored outputs."""
    import torch
    M = _maker()
    pdo.decomp_2d.comm_init()
    for (nx, ny, nz) in M.SHAPES:
        f = M.inputs(nx, ny, nz)
        tag = f"{nx}x{ny}x{nz}"
        fd = torch.from_numpy(f).cuda()
        ft = pdo.fft_3d()
        assert ft.init(nx, ny, nz, "x", 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz) == 0
        assert _rel(ft.fft3_x2z(fd).cpu().numpy(), gold[f"fft3_x2z_{tag}"]) < TOL
        assert _rel(ft.fft2_x2y(fd).cpu().numpy(), gold[f"fft2_x2y_{tag}"]) < TOL
        g3 = torch.from_numpy(gold[f"fft3_x2z_{tag}"] * (1.0 + 0.25j)).cuda()
        assert _rel(ft.ifft3_z2x(g3).cpu().numpy(), gold[f"ifft3_z2x_{tag}"]) < TOL
        g2 = torch.from_numpy(gold[f"fft2_x2y_{tag}"] * (1.0 + 0.25j)).cuda()
        assert _rel(ft.ifft2_y2x(g2, setOddBall=True).cpu().numpy(), gold[f"ifft2_y2x_{tag}"]) < TOL
        po = pdo.PoissonPeriodic()
        po.init(2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz, (nx, ny, nz), 1)
        assert _rel(po.poisson_solve(fd.clone()).cpu().numpy(), gold[f"poisson_{tag}"]) < TOL
        ft.destroy()
        po.destroy()
