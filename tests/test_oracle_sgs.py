"""Pins the oracle's restatement of the eddy-viscosity SGS models of the periodic igrid (sgsmod_igrid.F90:156-268,
sgs_models/{smagorinsky, sigma, AMD, eddyViscosity}.F90) with closed-form values of the three kernels, their defining
properties (sigma and AMD vanish for pure shear and solid rotation, sigma for isotropic expansion) and the sign of the
resolved-scale dissipation."""
import numpy as np
import pytest

from oracle import igrid_oracle as IG


def _model(mid, Csgs, n=8, scheme=1):
    d = 2 * np.pi / n
    C = IG.Spectral(n, n, n, d, d, d, init_periodicInZ=True)
    E = IG.Spectral(n, n, n + 1, d, d, d)
    return IG.SGS(C, E, IG.Pade6stagg(n, d, scheme), SGSModelID=mid, Csgs=Csgs), d


def _grad(M, shape=(2, 2, 2)):
    """nine constant gradient fields from the matrix M[i][j] = du_i/dx_j"""
    return [np.full(shape, float(M[i][j])) for i in range(3) for j in range(3)]


def test_kernels_closed_forms():
    sm, d = _model(0, 0.17)
    delta = (1.5 * d * 1.5 * d * 1.5 * d) ** (1.0 / 3.0)
    assert abs(sm.cmodel_global - (0.17 * delta) ** 2) < 1e-18
    g = _grad([[0, 0, 2.0], [0, 0, 0], [0, 0, 0]])                    # pure shear du/dz = 2: |S| = sqrt(2 S_ij S_ij) = 2
    assert np.allclose(sm.kernel(g, sm.get_Sij(g)), 2.0, rtol=0, atol=1e-15)
    g = _grad([[1.0, 0, 0], [0, -3.0, 0], [0, 0, 2.0]])
    assert np.allclose(sm.kernel(g, sm.get_Sij(g)), np.sqrt(2.0 * (1 + 9 + 4)), rtol=1e-15)
    sg, _ = _model(1, 1.5)
    g = _grad([[3.0, 0, 0], [0, 2.0, 0], [0, 0, 1.0]])                # singular values 3, 2, 1: sigma3 (s1 - s2)(s2 - s3) / s1^2 = 1/9
    assert np.allclose(sg.kernel(g, sg.get_Sij(g)), 1.0 / 9.0, rtol=1e-9)
    am, d = _model(2, 1.67)
    cx = 1.67 * d * np.sqrt(1.0 / 12.0)
    cz = 1.67 * d / np.sqrt(10.0)                                      # cd06: Poincare constant 1/sqrt(10) (PadeDerOps.F90:1025-1026)
    assert abs(am.camd_x - cx) < 1e-15 and abs(am.camd_z - cz) < 1e-15 and am.cmodel_global == 1.0
    a, b, c = 1.0, 1.0, -2.0
    g = _grad([[a, 0, 0], [0, b, 0], [0, 0, c]])
    num = (a * cx) ** 2 * a + (b * cx) ** 2 * b + (c * cz) ** 2 * c
    assert np.allclose(am.kernel(g, am.get_Sij(g)), max(-num / (a * a + b * b + c * c), 0.0), rtol=1e-14)
    amf, _ = _model(2, 1.67, scheme=2)
    assert abs(amf.camd_z - 1.67 * d / np.sqrt(12.0)) < 1e-15          # fourierColl: 1/sqrt(12)


@pytest.mark.parametrize("mid", [1, 2])
def test_sigma_and_amd_vanish_where_they_should(mid):
    m, _ = _model(mid, 1.0)
    for M in ([[0, 0, 2.0], [0, 0, 0], [0, 0, 0]],                    # pure shear
              [[0, -1.5, 0], [1.5, 0, 0], [0, 0, 0]]):                # solid-body rotation
        g = _grad(M)
        assert np.abs(m.kernel(g, m.get_Sij(g))).max() < 1e-12
    if mid == 1:
        g = _grad([[0.7, 0, 0], [0, 0.7, 0], [0, 0, 0.7]])            # isotropic expansion
        assert np.abs(m.kernel(g, m.get_Sij(g))).max() < 1e-7


@pytest.mark.parametrize("mid,Csgs", [(0, 0.17), (1, 1.5), (2, 1.67)])
@pytest.mark.parametrize("explicitE", [False, True])
def test_sgs_term_drains_resolved_energy(mid, Csgs, explicitE):
    n = 16
    L = 2 * np.pi
    rng = np.random.default_rng(0)
    u, v = 0.3 * rng.standard_normal((n, n, n)), 0.3 * rng.standard_normal((n, n, n))
    w = 0.3 * rng.standard_normal((n + 1, n, n))
    w[n] = w[0]
    g = IG.IGrid(n, n, n, L, L, L, 1e4, u, v, w, SGS_=dict(SGSModelID=mid, Csgs=Csgs, explicitCalcEdgeEddyViscosity=explicitE))
    d = g.duidxj
    dC = [d["dudx"], d["dudy"], d["dudzC"], d["dvdx"], d["dvdy"], d["dvdzC"], d["dwdxC"], d["dwdyC"], d["dwdz"]]
    dE = [d["dudxE"], d["dudyE"], d["dudz"], d["dvdxE"], d["dvdyE"], d["dvdz"], d["dwdx"], d["dwdy"], d["dwdzE"]]
    z = np.zeros_like(g.uhat)
    fu, fv, fw = g.sgsmodel.getRHS_SGS(z, z, np.zeros_like(g.what), dC, dE)
    assert g.sgsmodel.nu_sgs_C.min() >= 0.0 and g.sgsmodel.nu_sgs_E.min() >= 0.0
    P = (g.u * g.spectC.ifft(fu)).mean() + (g.v * g.spectC.ifft(fv)).mean() + (g.w[:n] * g.spectE.ifft(fw)[:n]).mean()
    assert P < 0.0                                                     # - tau_ij S_ij <= 0 for nu >= 0
    # the time step with the model loses more energy than without
    g0 = IG.IGrid(n, n, n, L, L, L, 1e4, u, v, w)
    g.timeAdvance(0.01)
    g0.timeAdvance(0.01)
    e = lambda q: (q.u ** 2 + q.v ** 2 + q.wC ** 2).mean()
    assert e(g) < e(g0)


@pytest.mark.parametrize("mid,Csgs", [(0, 0.17), (1, 1.5), (2, 1.67)])
def test_product_point_kernels_match_the_oracle(mid, Csgs):
    """The PRODUCT's pointwise SGS kernels (csrc/sgs_kernels.cuh, the __host__ __device__ code the CUDA kernels run) through a
    host-only hook, against the oracle on random velocity-gradient tensors, degenerate ones included."""
    import ctypes as C
    import padeops_b200 as pdo
    m, _ = _model(mid, Csgs)
    rng = np.random.default_rng(mid)
    cases = [rng.standard_normal(9) * s for s in (1.0, 1e-3, 50.0) for _ in range(40)]
    cases += [np.zeros(9), np.array([0, 0, 2.0, 0, 0, 0, 0, 0, 0]), np.array([0, -1.5, 0, 1.5, 0, 0, 0, 0, 0]), np.array([0.7, 0, 0, 0, 0.7, 0, 0, 0, 0.7])]
    cx, cy, cz = (m.camd_x, m.camd_y, m.camd_z) if mid == 2 else (0.0, 0.0, 0.0)
    for d in cases:
        g = [np.full((1,), v) for v in d]
        S = m.get_Sij(g)
        want = m.cmodel_global * m.kernel(g, S)[0]
        nu, S6 = C.c_double(0.0), np.zeros(6)
        dd = np.ascontiguousarray(d, dtype=np.float64)
        assert pdo.lib().pdo_debug_sgs_point(mid, m.cmodel_global, cx, cy, cz, C.c_void_p(dd.ctypes.data), C.byref(nu), C.c_void_p(S6.ctypes.data)) == 0
        assert np.allclose(S6, [s[0] for s in S], rtol=0, atol=0)
        assert abs(nu.value - want) <= 1e-12 * max(abs(want), 1e-3 * np.abs(d).max() ** (1 if mid == 0 else 1)), (mid, d, nu.value, want)
