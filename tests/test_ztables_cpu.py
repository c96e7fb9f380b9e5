"""CPU check of the z-Fourier path's host side: the PRODUCT's tables (through a host-only hook) against the oracle's, and the
kernels' algorithm for REAL arrays — two real columns transformed as one complex column, multiplied by the table whose oddball
entry is 1, transformed back — re-enacted in numpy with the product's tables against the oracle's r2c / c2r restatement
(spectral.F90:365-385, 507-526)."""
import ctypes as C

import numpy as np
import pytest

from oracle import igrid_oracle as IG

NAMES = ("k3_E2Cshift", "k3_C2Eshift", "E2Cshift", "C2Eshift", "mk3sq", "k3_C2Cder")


def _tables(pdo, nz, dz):
    t = np.zeros((2, 6, nz), dtype=np.complex128)
    rc = pdo.lib().pdo_debug_ztables(nz, float(dz), C.c_void_p(t.ctypes.data))
    assert rc == 0
    return t


@pytest.mark.parametrize("nz,Lz", [(8, 2 * np.pi), (32, 1.0), (50, 7.3)])
def test_product_tables_match_the_oracle(pdo, nz, Lz):
    dz = Lz / nz
    t = _tables(pdo, nz, dz)
    sp = IG.Spectral(8, 8, nz, 0.1, 0.1, dz, init_periodicInZ=True)
    for i, nm in enumerate(NAMES):
        ref = np.asarray(getattr(sp, nm), dtype=np.complex128)
        assert np.abs(t[0, i] - ref).max() <= 4e-16 * max(1.0, np.abs(ref).max()), nm
        twin = ref.copy()
        twin[nz // 2] = 1.0
        assert np.abs(t[1, i] - twin).max() <= 4e-16 * max(1.0, np.abs(ref).max()), nm
        # conjugate symmetry in k (what makes the pairing of real columns legal): t(nz - k) = conj(t(k))
        assert np.abs(t[1, i][1:] - np.conj(t[1, i][1:][::-1])).max() < 1e-13 * max(1.0, np.abs(ref).max()), nm


@pytest.mark.parametrize("P", [6, 7, 1])
@pytest.mark.parametrize("which", range(6))
def test_paired_real_columns_algorithm(pdo, P, which):
    nz, dz = 16, 2 * np.pi / 16
    t = _tables(pdo, nz, dz)[1, which]
    rng = np.random.default_rng(P * 10 + which)
    a = rng.standard_normal((nz, P))
    Pc = (P + 1) // 2
    w = np.zeros((nz, Pc), dtype=np.complex128)        # cudaMemcpy2D of P doubles per plane into a row of Pc complex numbers
    w.view(np.float64).reshape(nz, 2 * Pc)[:, :P] = a
    w = np.fft.fft(w, axis=0) * (t[:, None] * (1.0 / nz))
    w = np.fft.ifft(w, axis=0) * nz                    # unnormalised backward transform
    got = w.view(np.float64).reshape(nz, 2 * Pc)[:, :P]
    sp = IG.Spectral(8, 8, nz, 0.1, 0.1, dz, init_periodicInZ=True)
    ref = sp._real_z(a[:, None, :], getattr(sp, NAMES[which]))[:, 0, :]
    assert np.abs(got - ref).max() < 1e-13 * max(1.0, np.abs(ref).max())
