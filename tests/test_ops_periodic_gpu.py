"""GPU parity of Ops_Periodic (igrid_operators_periodic.F90:13-161), of the spectral type's z-Fourier in-place procedures and of
the REAL fourierColl procedures of Pade6stagg against the CPU oracle.  Bar: 1e-12 relative to max|ref| (north_star).

The oracle side is pinned in tests/test_oracle_ops_periodic.py."""
import numpy as np
import pytest

from conftest import broadband

pytestmark = [pytest.mark.gpu]
TOL = 1e-12


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(got, ref):
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)


def _cplx(shape, seed):
    return broadband(shape, seed) + 1j * broadband(shape, seed + 100)


@pytest.mark.parametrize("shape", [(32, 24, 16), (16, 16, 32), (15, 9, 8)])
def test_ops_periodic_matches_oracle(pdo, oracle, shape):
    """(15, 9, 8): an odd number of real columns per plane — the last column of the z-derivative is paired with zeros."""
    from oracle import ops_periodic_oracle as OP
    nx, ny, nz = shape
    d = [2 * np.pi / n for n in shape]
    ref = OP.OpsPeriodic(nx, ny, nz, *d)
    op = pdo.Ops_Periodic()
    op.init(nx, ny, nz, *d)
    f = broadband((nz, ny, nx), seed=sum(shape))
    fd = _dev(f)
    assert _rel(op.ddx(fd).cpu().numpy(), ref.ddx(f)) < TOL
    assert _rel(op.ddy(fd).cpu().numpy(), ref.ddy(f)) < TOL
    assert _rel(op.ddz(fd).cpu().numpy(), ref.ddz(f)) < TOL
    assert np.array_equal(fd.cpu().numpy(), f)                      # intent(in)
    c = _cplx((nz, ny, nx // 2 + 1), 3)
    assert _rel(op.ddz_cmplx2cmplx(_dev(c)).cpu().numpy(), ref.ddz_cmplx2cmplx(c)) < TOL
    p = op.SolvePoisson_oop(fd)
    assert _rel(p.cpu().numpy(), ref.SolvePoisson(f)) < TOL
    g = fd.clone()
    op.SolvePoisson_ip(g)
    assert _rel(g.cpu().numpy(), ref.SolvePoisson(f)) < TOL
    g = fd.clone()
    op.dealiasField(g)
    assert _rel(g.cpu().numpy(), ref.dealiasField(f)) < TOL
    # host arrays through the same entry points
    assert _rel(op.ddz(f.copy()), ref.ddz(f)) < TOL
    assert op.allocate3Dfield().shape == (nz, ny, nx)
    sp = op.link_spect()
    assert sp.physdecomp["xsz"] == (nx, ny, nz)
    op.destroy()


def test_oddball_mode_passes_through_the_real_z_derivative(pdo):
    nx, ny, nz = 8, 8, 12
    d = [2 * np.pi / n for n in (nx, ny, nz)]
    op = pdo.Ops_Periodic()
    op.init(nx, ny, nz, *d)
    z = np.arange(nz) * d[2]
    odd = np.cos((nz // 2) * z)[:, None, None] + np.zeros((nz, ny, nx))
    wave = np.sin(2 * z)[:, None, None] + np.zeros((nz, ny, nx))
    got = op.ddz(_dev(wave + 0.3 * odd)).cpu().numpy()
    assert np.abs(got - (2 * np.cos(2 * z)[:, None, None] + 0.3 * odd)).max() < 1e-13
    op.destroy()


@pytest.mark.parametrize("shape", [(16, 12, 16), (10, 6, 8)])
def test_spectral_z_fourier_procedures(pdo, oracle, shape):
    from oracle import igrid_oracle as IG
    nx, ny, nz = shape
    d = [2 * np.pi / n for n in shape]
    sp = pdo.spectral()
    sp.init("x", nx, ny, nz, *d, fixOddball=False, init_periodicInZ=True)
    ref = IG.Spectral(nx, ny, nz, *d, True, 2.0 / 3.0, False)
    a = broadband((nz, ny, nx), seed=2)
    assert _rel(sp.ddz_C2C_real_inplace(_dev(a)).cpu().numpy(), ref.ddz_C2C_real_inplace(a)) < TOL
    c = _cplx((nz, ny, nx // 2 + 1), 4)
    assert _rel(sp.ddz_C2C_complex_inplace(_dev(c)).cpu().numpy(), ref.ddz_C2C_complex_inplace(c)) < TOL
    assert _rel(sp.shiftz_E2C(_dev(c)).cpu().numpy(), ref.shiftz_E2C(c)) < TOL
    assert _rel(sp.shiftz_C2E(_dev(c)).cpu().numpy(), ref.shiftz_C2E(c)) < TOL
    sp.destroy()


@pytest.mark.parametrize("shape", [(16, 12, 16), (9, 5, 8)])
def test_pade6stagg_real_fourier_collocation(pdo, oracle, shape):
    from oracle import igrid_oracle as IG
    nx, ny, nz = shape
    d = [2 * np.pi / n for n in shape]
    spC = pdo.spectral()
    spC.init("x", nx, ny, nz, *d, fixOddball=False, init_periodicInZ=True)
    der = pdo.Pade6stagg()
    der.init(spC.physdecomp, spC.spectdecomp, dz=d[2], scheme=2, isPeriodic=True, spectC=spC)
    rops = IG.Pade6stagg(nz, d[2], scheme=2)
    fC, fE = broadband((nz, ny, nx), 4), broadband((nz + 1, ny, nx), 5)
    fE[nz] = fE[0]
    for name in ("ddz_E2C", "interpz_E2C", "d2dz2_E2E"):
        got = getattr(der, name)(_dev(fE)).cpu().numpy()
        ref = getattr(rops, name)(fE)
        assert got.shape == ref.shape and _rel(got, ref) < TOL, name
    for name in ("ddz_C2E", "interpz_C2E", "d2dz2_C2C"):
        got = getattr(der, name)(_dev(fC)).cpu().numpy()
        ref = getattr(rops, name)(fC)
        assert got.shape == ref.shape and _rel(got, ref) < TOL, name
    der.destroy()
    spC.destroy()


def test_field_files_round_trip(pdo, tmp_path):
    """WriteField3D / ReadField3D (igrid_operators_periodic.F90:162-205): Run<rid>_<label>_t<tidx>.out, flat global Fortran order"""
    nx, ny, nz = 16, 12, 8
    d = [2 * np.pi / n for n in (nx, ny, nz)]
    op = pdo.Ops_Periodic()
    op.init(nx, ny, nz, *d, InputDir=str(tmp_path), OutputDir=str(tmp_path))
    f = broadband((nz, ny, nx), seed=1)
    op.WriteField3D(_dev(f), "uVel", 12, 3)
    assert (tmp_path / "Run03_uVel_t000012.out").read_bytes() == f.tobytes()
    back = op.ReadField3D(op.allocate3Dfield(), "uVel", 12, 3)
    assert np.array_equal(back.cpu().numpy(), f)
    with pytest.raises(pdo.PadeOpsError) as e:
        op.ReadField3D(op.allocate3Dfield(), "vVel", 12, 3)
    assert e.value.code == 321
    op.destroy()
