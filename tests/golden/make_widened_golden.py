#!/usr/bin/env python
"""Generates tests/golden/widened_golden.npz — small fixed input/output vectors for the rows widened beyond the periodic hot path
(filter3D, Ops_Periodic, the staggered wall operators, the Fourier-collocation operators, the wall-bounded
projection, the igrid run-deck variants).  Same provenance and purpose as make_golden.py: oracle-generated REGRESSION anchors
(the reference cannot run here), frozen at the commit that generated them, travelling to the GPU box.

Usage: python tests/golden/make_widened_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import igrid_oracle as IG  # noqa: E402
from oracle import ops_periodic_oracle as OP  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import stagg_np_oracle as SN  # noqa: E402

STAGG_OPS = ("ddz_E2C", "ddz_C2E", "ddz_C2C", "ddz_E2E", "InterpZ_E2C", "InterpZ_C2E", "d2dz2_C2C", "d2dz2_E2E")
STAGG_EDGE_IN = {"ddz_E2C", "ddz_E2E", "InterpZ_E2C", "d2dz2_E2E"}
WALLS = {"EO": (True, False, False, False), "SS": (True, True, True, True), "OS": (False, False, False, True)}   # isTopEven, isBotEven, isTopSided, isBotSided
PADE_OPS = ("ddz_E2C", "ddz_C2E", "interpz_E2C", "interpz_C2E", "d2dz2_C2C", "d2dz2_E2E")
PADE_EDGE_IN = {"ddz_E2C", "interpz_E2C", "d2dz2_E2E"}


def rnd(shape, seed, cplx=False):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(shape)
    return a + 1j * rng.standard_normal(shape) if cplx else a


def igrid_cases():
    m = 8
    U, V = 0.3 * rnd((m, m, m), 21), 0.3 * rnd((m, m, m), 22)
    W = 0.3 * rnd((m + 1, m, m), 23)
    W[m] = W[0]
    L = (2 * np.pi,) * 3
    deck = dict(TimeSteppingScheme=2, AdvectionTerm=0, NumericalSchemeVert=2, HITForcing_=dict(kmin=1.0, kmax=2.5, Nwaves=6, EpsAmplitude=0.05, RandSeedToAdd=0),
                SGS_=dict(SGSModelID=2, Csgs=1.67, explicitCalcEdgeEddyViscosity=False))
    Ww = W.copy()
    Ww[0] = 0.0
    Ww[m] = 0.0
    return {"deck": ((U, V, W), (*L, 1.0e10), deck), "slip": ((U, V, Ww), (*L[:2], 2.0, 100.0), dict(TimeSteppingScheme=1, PeriodicInZ=False, topWall=2, botWall=2)),
            "noslip": ((U, V, Ww), (*L[:2], 2.0, 100.0), dict(TimeSteppingScheme=1, PeriodicInZ=False, topWall=1, botWall=1)),      # Stokes pressure on (default)
            "noslip_nostokes": ((U, V, Ww), (*L[:2], 2.0, 100.0), dict(TimeSteppingScheme=1, PeriodicInZ=False, topWall=1, botWall=2,
                                                                      ComputeStokesPressure=False))}


def main():
    O.build()
    out = {}
    nx, ny, nz = 20, 16, 12
    f = rnd((nz, ny, nx), 1)
    out["f3d_in"] = f
    out["f3d_cf90_x2"] = O.filter3D(f, 2)
    out["f3d_mixed_x1"] = O.filter3D(f, 1, ("gaussian", "cf90", "gaussian"))
    # Ops_Periodic
    mx, my, mz = 16, 12, 8
    d = [2 * np.pi / n for n in (mx, my, mz)]
    op = OP.OpsPeriodic(mx, my, mz, *d)
    g = rnd((mz, my, mx), 2)
    c = rnd((mz, my, mx // 2 + 1), 3, True)
    out["opp_in"], out["opp_cin"] = g, c
    out["opp_ddx"], out["opp_ddy"], out["opp_ddz"] = op.ddx(g), op.ddy(g), op.ddz(g)
    out["opp_ddz_c2c"], out["opp_poisson"], out["opp_dealias"] = op.ddz_cmplx2cmplx(c), op.SolvePoisson(g), op.dealiasField(g)
    # staggered wall operators, n = 16, real and complex
    n, dz = 16, 1.0 / 16
    fC, fE = rnd((n, 3, 3), 4), rnd((n + 1, 3, 3), 5)
    cC, cE = fC + 1j * fC[::-1], fE + 1j * fE[::-1]
    out["snp_fC"], out["snp_fE"], out["snp_cC"], out["snp_cE"] = fC, fE, cC, cE
    for tag, (te, be, ts, bs) in WALLS.items():
        ref = SN.CD06StaggNP(n, dz, te, be, ts, bs)
        for name in STAGG_OPS:
            edge = name in STAGG_EDGE_IN
            out[f"snp_{tag}_{name}"] = getattr(ref, name)(fE if edge else fC)
            out[f"snp_{tag}_c_{name}"] = getattr(ref, name)(cE if edge else cC)
    # Fourier collocation in z, real and complex
    pf = IG.Pade6stagg(n, 2 * np.pi / n, scheme=2)
    fE2 = fE.copy()
    fE2[n] = fE2[0]
    cE2 = fE2 + 1j * np.roll(fE2, 3, axis=0)
    cE2[n] = cE2[0]
    out["four_fE"], out["four_cE"] = fE2, cE2
    for name in PADE_OPS:
        edge = name in PADE_EDGE_IN
        out[f"four_{name}"] = getattr(pf, name)(fE2 if edge else fC)
        out[f"four_c_{name}"] = getattr(pf, name)(cE2 if edge else cC)
    # wall-bounded projection
    px, py, pz = 8, 6, 10
    pd = [2 * np.pi / px, 2 * np.pi / py, 1.0 / pz]
    spC, spE = IG.Spectral(px, py, pz, *pd), IG.Spectral(px, py, pz + 1, *pd)
    P = IG.PadePoisson(*pd, spC, spE, IG.Pade6stagg(pz, pd[2], 1, isPeriodic=False), PeriodicInZ=False)
    uh, vh, wh = rnd((pz, py, px // 2 + 1), 6, True), rnd((pz, py, px // 2 + 1), 7, True), rnd((pz + 1, py, px // 2 + 1), 8, True)
    wh[0] = 0
    wh[pz] = 0
    out["wp_u"], out["wp_v"], out["wp_w"] = uh, vh, wh
    out["wp_u1"], out["wp_v1"], out["wp_w1"] = P.PressureProjection(uh, vh, wh)
    Ps = IG.PadePoisson(*pd, spC, spE, IG.Pade6stagg(pz, pd[2], 1, isPeriodic=False), PeriodicInZ=False, computeStokesPressure=True, Lz=1.0)
    whs = rnd((pz + 1, py, px // 2 + 1), 9, True)        # nonzero on the walls
    out["wps_w"] = whs
    out["wps_u1"], out["wps_v1"], out["wps_w1"] = Ps.PressureProjection(uh, vh, whs)
    # igrid variants, one step on 8^3
    for tag, ((U, V, W), (Lx, Ly, Lz, Re), kw) in igrid_cases().items():
        m = U.shape[0]
        gsim = IG.IGrid(m, m, m, Lx, Ly, Lz, Re, U, V, W, **kw)
        gsim.timeAdvance(0.005)
        out[f"ig_{tag}_U0"], out[f"ig_{tag}_V0"], out[f"ig_{tag}_W0"] = U, V, W
        out[f"ig_{tag}_u1"], out[f"ig_{tag}_v1"], out[f"ig_{tag}_w1"] = gsim.u, gsim.v, gsim.w
    # HIT forcing draw for (tidStart 3, RandSeedToAdd 1), first two steps
    hf = IG.HITForcing(spC, kmin=2.0, kmax=6.0, Nwaves=12, tidStart=3, RandSeedToAdd=1)
    for step in range(2):
        hf.pick_random_wavenumbers()
        out[f"hit_waves_step{step}"] = np.stack([hf.wave_x, hf.wave_y, hf.wave_z])
        hf.update_seeds()
    np.savez_compressed(os.path.join(HERE, "widened_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "widened_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
