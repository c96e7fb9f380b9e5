#!/usr/bin/env python
"""Generates tests/golden/hotpath_golden.npz — small fixed input/output vectors for every operator on the hot path.

PROVENANCE: the reference (Fortran 2003 + MPI) cannot be compiled or imported in this image and ships no stored vectors
(SURVEY.md §8c), so these vectors are produced by the CPU oracle (oracle/, the line-by-line restatement of the reference
that tests/test_oracle_*.py pin against the reference's own analytic known-answer tests).  They are REGRESSION anchors:
they freeze the oracle's arithmetic at the commit that generated them, so that neither the oracle nor the CUDA path can
drift unnoticed, and they travel to the GPU box.  They are not independent evidence of parity with a Fortran build.

Usage: python tests/golden/make_golden.py   (rewrites the .npz next to this file; inputs are seeded)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def field(shape, seed):
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    x, y, z = (np.arange(n) * (2 * np.pi / n) for n in (nx, ny, nz))
    f = np.sin(x)[None, None, :] * np.cos(2 * y)[None, :, None] + np.sin(3 * z)[:, None, None] * np.cos(x)[None, None, :]
    return f + 0.05 * rng.standard_normal(shape)


def main():
    O.build()
    out = {}
    nx, ny, nz = 20, 16, 12          # CD10 needs n >= 8, CF90 n >= 10; all axes differ so an axis mix-up cannot hide
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    f = field((nz, ny, nx), 20240607)
    out["f"] = f
    for ax, d in enumerate((dx, dy, dz)):
        out[f"cd10_d1_ax{ax}"] = O.cd10(f, d, ax, 1)
        out[f"cd10_d2_ax{ax}"] = O.cd10(f, d, ax, 2)
        out[f"cd06_d1_ax{ax}"] = O.cd06(f, d, ax)
        out[f"cf90_ax{ax}"] = O.cf90(f, ax)
        out[f"gaussian_ax{ax}"] = O.gaussian(f, ax)
    # staggered operators along z (tests/test_PadeDer_periodic.F90 shape: thin in x, y), real and complex
    n = 32
    dzs = 2 * np.pi / n
    fC = field((n, 4, 4), 1)
    fE = field((n + 1, 4, 4), 2)
    out["stagg_fC"], out["stagg_fE"] = fC, fE
    for name, fin in (("ddz_E2C", fE), ("ddz_C2E", fC), ("interp_E2C", fE), ("interp_C2E", fC), ("d2dz2_C2C", fC), ("d2dz2_E2E", fE)):
        out["stagg_" + name] = O.stagg(name, fin, n, dzs)
        out["stagg_c_" + name] = O.stagg(name, fin + 1j * fin[::-1], n, dzs)
    # Poisson (tests/test_PoissonPeriodic.F90 shape, reduced) and the operators.F90 compositions
    out["poisson"] = O.poisson_solve(f, dx, dy, dz)
    u, v, w = field((nz, ny, nx), 3), field((nz, ny, nx), 4), field((nz, ny, nx), 5)
    out["u"], out["v"], out["w"] = u, v, w
    out["div_cd10"] = O.divergence(u, v, w, dx, dy, dz, "cd10")
    out["curl_cd10"] = O.curl(u, v, w, dx, dy, dz, "cd10")
    # transposes: rank pencils of a 2 x 2 grid with uneven sizes (17 x 9 x 11), x -> y -> z
    gx, gy, gz = 17, 9, 11
    G = np.arange(gx * gy * gz, dtype=np.float64).reshape(gz, gy, gx)
    for pen in "xyz":
        for r, a in enumerate(O.scatter_global(G, gx, gy, gz, 2, 2, pen)):
            out[f"pencil_{pen}_rank{r}"] = a
    # one TVD-RK3 step of the igrid periodic substep on 8^3
    from oracle import igrid_oracle as IG
    m = 8
    rng = np.random.default_rng(11)
    U, V = rng.standard_normal((m, m, m)), rng.standard_normal((m, m, m))
    W = rng.standard_normal((m + 1, m, m))
    W[m] = W[0]
    g = IG.IGrid(m, m, m, 2 * np.pi, 2 * np.pi, 2 * np.pi, 80.0, U, V, W, TimeSteppingScheme=1)
    g.timeAdvance(0.01)
    out["ig_U0"], out["ig_V0"], out["ig_W0"] = U, V, W
    out["ig_u1"], out["ig_v1"], out["ig_w1"] = g.u, g.v, g.w
    np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "hotpath_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
