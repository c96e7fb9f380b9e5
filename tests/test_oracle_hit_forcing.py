"""Pins the oracle's restatement of HIT_shell_forcing (incompressible/forcingIsotropic.F90:45-314): seed arithmetic, the
sample -> integer-wavenumber map, and the forcing itself — a forced mode receives, in physical space, the plane wave
(normfact Eps / den / Nwaves / (nx ny nz)) conj(u_hat) e^{i k.x} projected the way the 2-D real transform projects it; the energy
injection rate sum(u . f) per forced (non-degenerate) mode is Eps / Nwaves x (|u_hat|^2 / den) by construction."""
import numpy as np
import pytest

from oracle import igrid_oracle as IG


def _setup(nx=16, ny=12, nz=16, **kw):
    d = [2 * np.pi / n for n in (nx, ny, nz)]
    sp = IG.Spectral(nx, ny, nz, *d, init_periodicInZ=True)
    return sp, IG.HITForcing(sp, **kw), d


def test_seed_arithmetic_and_splitmix():
    sp, f, _ = _setup(tidStart=5, RandSeedToAdd=2)
    assert (f.seed0, f.seed1, f.seed2, f.seed3) == (7 + 2223345, 7 + 2223345 + 1423246, 7 + 2223345 + 8723446, 7 + 2223345 + 3423444)
    f.update_seeds()
    assert f.seed0 == 7 + 2 * 2223345 and f.seed1 == f.seed0 + 1423246
    # SplitMix64 known answers (seed 0: the published first outputs of the reference implementation)
    mask = (1 << 64) - 1
    first = [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    u = IG.splitmix64_uniform(0, 3)
    assert [int(x * 9007199254740992.0) for x in u] == [v >> 11 for v in first]
    assert np.all((u >= 0) & (u < 1)) and mask


def test_wavenumbers_from_samples_and_shell():
    sp, f, _ = _setup(kmin=4.0, kmax=5.0, Nwaves=200)
    f.wavenumbers_from_samples(np.array([4.5, 5.0]), np.array([0.0, -1.0]), np.array([0.0, 1.0]))
    assert f.wave_x.tolist() == [5, 1] or f.wave_x.tolist() == [5, 0]      # ceiling(|4.5|) = 5; sqrt(1 - 1) = 0 -> ceiling(0) = 0
    assert f.wave_x.tolist() == [5, 0] and f.wave_y.tolist() == [0, 0] and f.wave_z.tolist() == [0, 5]
    f.pick_random_wavenumbers()
    k = np.sqrt(f.wave_x ** 2.0 + f.wave_y ** 2.0 + f.wave_z ** 2.0)
    assert np.all(k >= 4.0 - 1e-12) and np.all(k <= 5.0 + np.sqrt(3.0))   # ceiling moves each component up by < 1
    assert f.wave_x.min() >= 0 and len(set(zip(f.wave_x, f.wave_y, f.wave_z))) > 20


def test_forcing_of_a_single_mode_is_the_scaled_conjugate_wave():
    nx, ny, nz = 16, 12, 16
    sp, f, d = _setup(nx, ny, nz, Nwaves=1, EpsAmplitude=0.3)
    rng = np.random.default_rng(2)
    u, v = rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz, ny, nx))
    w = rng.standard_normal((nz + 1, ny, nx)); w[nz] = w[0]
    spE = IG.Spectral(nx, ny, nz + 1, *d)
    uh, vh, wh = sp.fft(u), sp.fft(v), spE.fft(w)
    kx, ky, kz = 2, 3, 4
    f.set_wavenumbers([kx], [ky], [kz])
    zero = np.zeros_like(uh)
    zeroE = np.zeros_like(wh)
    ur, vr, wr = f.getRHS_HITforcing(zero, zero, zeroE, uh, vh, wh, False)
    U = np.fft.fft(uh, axis=0)[kz, ky, kx]
    V = np.fft.fft(vh, axis=0)[kz, ky, kx]
    W = (np.fft.fft(wh[:nz], axis=0) * sp.E2Cshift[:, None, None])[kz, ky, kx]
    den = abs(U) ** 2 + abs(V) ** 2 + abs(W) ** 2 + 1e-14
    fac = (nx * ny * nz) ** 2 * 0.3 / den
    z = np.arange(nz)
    wave = np.exp(2j * np.pi * kz * z / nz) / nz
    exp_u = np.zeros_like(uh); exp_u[:, ky, kx] = fac * np.conj(U) * wave
    assert np.abs(ur - exp_u).max() < 1e-12 * np.abs(exp_u).max()
    exp_v = np.zeros_like(uh); exp_v[:, ky, kx] = fac * np.conj(V) * wave
    assert np.abs(vr - exp_v).max() < 1e-12 * np.abs(exp_v).max()
    exp_w = np.zeros_like(wh)
    exp_w[:nz, ky, kx] = fac * np.conj(W) * sp.C2Eshift[kz] * wave
    exp_w[nz] = exp_w[0]
    assert np.abs(wr - exp_w).max() < 1e-12 * np.abs(exp_w).max()
    # duplicates add up; modes outside the half-spectrum are skipped
    f.Nwaves = 3
    f.set_wavenumbers([kx, kx, nx], [ky, ky, 1], [kz, kz, 1])
    ur3, _, _ = f.getRHS_HITforcing(zero, zero, zeroE, uh, vh, wh, False)
    assert np.abs(ur3 - (2.0 / 3.0) * exp_u).max() < 1e-12 * np.abs(exp_u).max()


@pytest.mark.parametrize("scheme", [1, 2])
def test_forced_taylor_green_gains_energy_and_stays_solenoidal(scheme):
    """igrid with useHITForcing: the forcing enters populate_rhs after the viscous term (igrid.F90:1907-1910), wavenumbers are
    redrawn once per time step; the projected field stays divergence-free and, at high Re, its energy grows."""
    n = 16
    L = 2 * np.pi
    rng = np.random.default_rng(7)          # energy in every mode of the forced shell (an empty mode has den = 1e-14: unbounded forcing)
    u, v = 0.3 * rng.standard_normal((n, n, n)), 0.3 * rng.standard_normal((n, n, n))
    w = 0.3 * rng.standard_normal((n + 1, n, n))
    w[n] = w[0]
    hit = dict(kmin=1.0, kmax=2.5, Nwaves=12, EpsAmplitude=0.5, RandSeedToAdd=3)
    g = IG.IGrid(n, n, n, L, L, L, 1.0e4, u, v, w, TimeSteppingScheme=scheme, HITForcing_=hit)
    g0 = IG.IGrid(n, n, n, L, L, L, 1.0e4, u, v, w, TimeSteppingScheme=scheme)
    e_start = (g.u ** 2 + g.v ** 2 + g.wC ** 2).mean()
    waves = []
    for _ in range(3):
        g.timeAdvance(0.01)
        g0.timeAdvance(0.01)
        waves.append(tuple(g.hitforce.wave_x))
    assert len(set(waves)) == 3                                       # a new draw every step, none inside the RK stages
    e_forced, e_free = (g.u ** 2 + g.v ** 2 + g.wC ** 2).mean(), (g0.u ** 2 + g0.v ** 2 + g0.wC ** 2).mean()
    assert e_forced > e_free and e_forced > e_start
    rate = 0.5 * (e_forced - e_free) / 0.03                           # kinetic energy per unit time: of the order of EpsAmplitude
    assert 0.5 * 0.5 < rate < 6.0 * 0.5
    _, _, _, div = g.poiss.DivergenceCheck(g.uhat, g.vhat, g.what)
    assert np.abs(div).max() < 1e-11


@pytest.mark.parametrize("tid,add,updates", [(0, 0, 0), (5, 2, 3), (1000, 17, 1200)])
def test_product_draw_matches_the_oracle(tid, add, updates):
    """The PRODUCT's host side of the forcing (seed arithmetic, SplitMix64, the sample -> wavenumber map) through a host-only
    hook, against the oracle: the two must pick the same modes at every step of a long run."""
    import ctypes as C
    import padeops_b200 as pdo
    n = 40
    sp, f, _ = _setup(kmin=3.0, kmax=7.5, Nwaves=n, tidStart=tid, RandSeedToAdd=add)
    for _ in range(updates):
        f.update_seeds()
    f.pick_random_wavenumbers()
    seeds = (C.c_longlong * 4)()
    w = [(C.c_int * n)() for _ in range(3)]
    assert pdo.lib().pdo_debug_hit_draw(3.0, 7.5, n, tid, add, updates, seeds, *w) == 0
    assert list(seeds) == [f.seed0, f.seed1, f.seed2, f.seed3]
    assert list(w[0]) == f.wave_x.tolist() and list(w[1]) == f.wave_y.tolist() and list(w[2]) == f.wave_z.tolist()


def test_sparse_dft_algorithm_of_the_kernels_equals_the_fft_formulation():
    """csrc/igrid.cu evaluates the forcing without whole-field transforms: (U, V, Wraw) = direct DFT of the forced columns
    (partial sums per z-slab, summed over ranks), then every wave adds its plane wave to the right-hand sides.  Re-enacted here
    in numpy, two z-slabs with the uneven cell / edge plane split of a 1 x 2 grid, against the oracle's FFT formulation."""
    nx, ny, nz = 12, 10, 16
    sp, f, d = _setup(nx, ny, nz, Nwaves=6, EpsAmplitude=0.2)
    spE = IG.Spectral(nx, ny, nz + 1, *d)
    rng = np.random.default_rng(4)
    u, v = rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz, ny, nx))
    w = rng.standard_normal((nz + 1, ny, nx)); w[nz] = w[0]
    uh, vh, wh = sp.fft(u), sp.fft(v), spE.fft(w)
    waves = ([2, 3, 2, 0, 6, 40], [1, 9, 1, 0, 3, 1], [4, 0, 4, 7, 15, 1])      # a duplicate, ky in the upper half, one mode off the grid
    f.set_wavenumbers(*waves)
    r0 = [rng.standard_normal(uh.shape) + 1j * rng.standard_normal(uh.shape) for _ in range(2)] + \
         [rng.standard_normal(wh.shape) + 1j * rng.standard_normal(wh.shape)]
    want = f.getRHS_HITforcing(r0[0], r0[1], r0[2], uh, vh, wh, False)
    # --- the kernels' algorithm ---
    nxh = nx // 2 + 1
    slabsC = [(0, 8), (8, 16)]
    slabsE = [(0, 8), (8, 17)]
    n = len(waves[0])
    part = np.zeros((n, 3), dtype=np.complex128)
    for (c0, c1), (e0, e1) in zip(slabsC, slabsE):                  # hit_reduce_kernel on each rank, then the allreduce
        for i, (kx, ky, kz) in enumerate(zip(*waves)):
            if not (0 <= kx < nxh and 0 <= ky < ny and 0 <= kz < nz):
                continue
            zc = np.arange(c0, c1)
            ph = np.exp(-2j * np.pi * ((kz * zc) % nz) / nz)
            part[i, 0] += (uh[c0:c1, ky, kx] * ph).sum()
            part[i, 1] += (vh[c0:c1, ky, kx] * ph).sum()
            ze = np.arange(e0, min(e1, nz))
            part[i, 2] += (wh[e0:min(e1, nz), ky, kx] * np.exp(-2j * np.pi * ((kz * ze) % nz) / nz)).sum()
    got = [a.copy() for a in r0]
    normfact = (nx * ny * nz) ** 2.0
    for i, (kx, ky, kz) in enumerate(zip(*waves)):                  # hit_apply_kernel: waves one after another, every plane
        if not (0 <= kx < nxh and 0 <= ky < ny and 0 <= kz < nz):
            continue
        U, V, W = part[i, 0], part[i, 1], part[i, 2] * sp.E2Cshift[kz]
        den = abs(U) ** 2 + abs(V) ** 2 + abs(W) ** 2 + 1e-14
        fac = normfact * 0.2 / den / n
        zc = np.arange(nz)
        ph = np.exp(2j * np.pi * ((kz * zc) % nz) / nz)
        got[0][:, ky, kx] += fac * np.conj(U) / nz * ph
        got[1][:, ky, kx] += fac * np.conj(V) / nz * ph
        ze = np.arange(nz + 1)
        ze[nz] = 0
        got[2][:, ky, kx] += fac * np.conj(W) * sp.C2Eshift[kz] / nz * np.exp(2j * np.pi * ((kz * ze) % nz) / nz)
    for a, b in zip(got, want):
        assert np.abs(a - b).max() < 1e-12 * np.abs(b).max()


def test_injected_draw_is_kept_through_the_next_new_step_only():
    sp, f, _ = _setup(kmin=1.0, kmax=3.0, Nwaves=4, tidStart=0)
    g = IG.HITForcing(sp, kmin=1.0, kmax=3.0, Nwaves=4, tidStart=0)
    z = np.zeros((sp.nz, sp.ny, sp.nxh), dtype=complex)
    zE = np.zeros((sp.nz + 1, sp.ny, sp.nxh), dtype=complex)
    f.set_wavenumbers([1, 2, 1, 0], [0, 1, 1, 2], [1, 1, 2, 1])
    f.getRHS_HITforcing(z, z, zE, z + 1.0, z + 1.0, zE + 1.0, True)
    g.getRHS_HITforcing(z, z, zE, z + 1.0, z + 1.0, zE + 1.0, True)
    assert f.wave_x.tolist() == [1, 2, 1, 0]                      # kept ...
    assert (f.seed0, f.seed1) == (g.seed0, g.seed1)               # ... while the seeds advance like in an undisturbed run
    f.getRHS_HITforcing(z, z, zE, z + 1.0, z + 1.0, zE + 1.0, True)
    g.getRHS_HITforcing(z, z, zE, z + 1.0, z + 1.0, zE + 1.0, True)
    assert f.wave_x.tolist() == g.wave_x.tolist() and f.wave_z.tolist() == g.wave_z.tolist()    # the next step draws again, in step
