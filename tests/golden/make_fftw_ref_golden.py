"""Generates tests/golden/fftw_ref_golden.npz: outputs of FFTW 3.3.5 — compiled from the reference's vendored tarball into
oracle/_ref (make -C oracle ref) and driven with fft_3d.F90's own plan tuples (oracle/fftw_ref.py) — for seeded inputs.
These are REFERENCE-ARITHMETIC vectors (the only ones this image can produce: the Fortran itself does not compile here); the
oracle (CPU, everywhere) and the CUDA path (GPU box, where /root/reference and possibly oracle/_ref are absent) are checked against them.
Run in the build container:  python tests/golden/make_fftw_ref_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SHAPES = [(64, 32, 16), (18, 9, 7), (24, 20, 12)]     # (nx, ny, nz); the first is tests/test_PoissonPeriodic.F90's grid


def inputs(nx, ny, nz):
    rng = np.random.default_rng(1000 * nx + 10 * ny + nz)
    return rng.standard_normal((nz, ny, nx))


def main():
    from oracle import fftw_ref as R
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    out = {"fftw_version": np.array(R.version())}
    for (nx, ny, nz) in SHAPES:
        f = inputs(nx, ny, nz)
        F = R.FFT3D(nx, ny, nz)
        tag = f"{nx}x{ny}x{nz}"
        h3 = F.fft3_x2z(f)
        h2 = F.fft2_x2y(f)
        out[f"fft3_x2z_{tag}"] = h3
        out[f"fft2_x2y_{tag}"] = h2
        out[f"ifft3_z2x_{tag}"] = F.ifft3_z2x(h3 * (1.0 + 0.25j))          # a non-Hermitian-looking spectrum: pins what c2r drops
        out[f"ifft2_y2x_{tag}"] = F.ifft2_y2x(h2 * (1.0 + 0.25j), setOddBall=True)
        out[f"poisson_{tag}"] = R.poisson_solve(f, 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz, fft=F)
        F.destroy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fftw_ref_golden.npz"), **out)
    print("wrote", len(out), "arrays;", out["fftw_version"])


if __name__ == "__main__":
    main()
