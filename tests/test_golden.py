"""The committed fixtures of tests/golden/ (provenance in make_golden.py: oracle-generated regression anchors, the
reference cannot run here).  CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA path through the C ABI
reproduces them within north_star's 1e-12 (transposes: bit-exact)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_golden.npz")
TOL = 1e-12


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _spacing(f):
    nz, ny, nx = f.shape
    return 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_oracle_reproduces_golden_bitwise(oracle, gold):
    f = gold["f"]
    d = _spacing(f)
    for ax in range(3):
        assert np.array_equal(oracle.cd10(f, d[ax], ax, 1), gold[f"cd10_d1_ax{ax}"])
        assert np.array_equal(oracle.cd10(f, d[ax], ax, 2), gold[f"cd10_d2_ax{ax}"])
        assert np.array_equal(oracle.cd06(f, d[ax], ax), gold[f"cd06_d1_ax{ax}"])
        assert np.array_equal(oracle.cf90(f, ax), gold[f"cf90_ax{ax}"])
        assert np.array_equal(oracle.gaussian(f, ax), gold[f"gaussian_ax{ax}"])
    n = gold["stagg_fC"].shape[0]
    for name, fin in (("ddz_E2C", "stagg_fE"), ("ddz_C2E", "stagg_fC"), ("interp_E2C", "stagg_fE"), ("interp_C2E", "stagg_fC"),
                      ("d2dz2_C2C", "stagg_fC"), ("d2dz2_E2E", "stagg_fE")):
        assert np.array_equal(oracle.stagg(name, gold[fin], n, 2 * np.pi / n), gold["stagg_" + name])
    assert _rel(oracle.poisson_solve(f, *d), gold["poisson"]) < 1e-14      # pocketfft may differ in the last bit across numpy builds
    assert np.array_equal(oracle.divergence(gold["u"], gold["v"], gold["w"], *d, "cd10"), gold["div_cd10"])
    G = np.arange(17 * 9 * 11, dtype=np.float64).reshape(11, 9, 17)
    for pen in "xyz":
        for r, a in enumerate(oracle.scatter_global(G, 17, 9, 11, 2, 2, pen)):
            assert np.array_equal(a, gold[f"pencil_{pen}_rank{r}"])
    # transposes of the golden pencils land exactly on the golden pencils of the next orientation
    for d_, (s, t) in enumerate((("x", "y"), ("y", "x"), ("y", "z"), ("z", "y"))):
        outs = oracle.transpose(d_, 17, 9, 11, 2, 2, [gold[f"pencil_{s}_rank{r}"] for r in range(4)])
        for r in range(4):
            assert np.array_equal(outs[r], gold[f"pencil_{t}_rank{r}"])


def test_igrid_oracle_reproduces_golden(oracle, gold):
    from oracle import igrid_oracle as IG
    m = gold["ig_U0"].shape[0]
    g = IG.IGrid(m, m, m, 2 * np.pi, 2 * np.pi, 2 * np.pi, 80.0, gold["ig_U0"], gold["ig_V0"], gold["ig_W0"], TimeSteppingScheme=1)
    g.timeAdvance(0.01)
    for nm in ("u", "v", "w"):
        assert _rel(getattr(g, nm), gold[f"ig_{nm}1"]) < 1e-13


@pytest.mark.gpu
def test_cuda_reproduces_golden(pdo, gold):
    import torch
    f = gold["f"]
    d = _spacing(f)
    nz, ny, nx = f.shape
    fd = torch.from_numpy(f).cuda()
    c10, c06, cf, ga = [pdo.cd10() for _ in range(3)], [pdo.cd06() for _ in range(3)], [pdo.cf90() for _ in range(3)], [pdo.gaussian() for _ in range(3)]
    for ax, n in enumerate((nx, ny, nz)):
        assert c10[ax].init(n, d[ax]) == 0 and c06[ax].init(n, d[ax]) == 0 and cf[ax].init(n) == 0 and ga[ax].init(n) == 0
        dd = (c10[ax].dd1, c10[ax].dd2, c10[ax].dd3)[ax]
        d2 = (c10[ax].d2d1, c10[ax].d2d2, c10[ax].d2d3)[ax]
        assert _rel(dd(fd).cpu().numpy(), gold[f"cd10_d1_ax{ax}"]) < TOL
        assert _rel(d2(fd).cpu().numpy(), gold[f"cd10_d2_ax{ax}"]) < TOL
        assert _rel((c06[ax].dd1, c06[ax].dd2, c06[ax].dd3)[ax](fd).cpu().numpy(), gold[f"cd06_d1_ax{ax}"]) < TOL
        assert _rel((cf[ax].filter1, cf[ax].filter2, cf[ax].filter3)[ax](fd).cpu().numpy(), gold[f"cf90_ax{ax}"]) < TOL
        assert _rel((ga[ax].filter1, ga[ax].filter2, ga[ax].filter3)[ax](fd).cpu().numpy(), gold[f"gaussian_ax{ax}"]) < TOL
    n = gold["stagg_fC"].shape[0]
    st = pdo.cd06stagg()
    st.init(n, 2 * np.pi / n)
    for name, fn, fin in (("ddz_E2C", st.ddz_E2C, "stagg_fE"), ("ddz_C2E", st.ddz_C2E, "stagg_fC"), ("interp_E2C", st.InterpZ_E2C, "stagg_fE"),
                          ("interp_C2E", st.InterpZ_C2E, "stagg_fC"), ("d2dz2_C2C", st.d2dz2_C2C, "stagg_fC"), ("d2dz2_E2E", st.d2dz2_E2E, "stagg_fE")):
        assert _rel(fn(torch.from_numpy(gold[fin]).cuda()).cpu().numpy(), gold["stagg_" + name]) < TOL
        cin = gold[fin] + 1j * gold[fin][::-1]
        assert _rel(fn(torch.from_numpy(np.ascontiguousarray(cin)).cuda()).cpu().numpy(), gold["stagg_c_" + name]) < TOL
    po = pdo.PoissonPeriodic()
    po.init(d[0], d[1], d[2], (nx, ny, nz), 1)
    assert _rel(po.poisson_solve(fd.clone()).cpu().numpy(), gold["poisson"]) < TOL
    gp = pdo.decomp_2d.init(nx, ny, nz, 1, 1)
    ops = pdo.vector_ops()
    ops.init(gp, d[0], d[1], d[2], "cd10")
    u, v, w = (torch.from_numpy(gold[k]).cuda() for k in "uvw")
    assert _rel(ops.divergence(u, v, w).cpu().numpy(), gold["div_cd10"]) < TOL
    assert _rel(ops.curl(u, v, w).cpu().numpy(), gold["curl_cd10"]) < TOL
    m = gold["ig_U0"].shape[0]
    g = pdo.igrid()
    g.init(m, m, m, 2 * np.pi, 2 * np.pi, 2 * np.pi, 80.0, gold["ig_U0"], gold["ig_V0"], gold["ig_W0"], TimeSteppingScheme=1)
    g.timeAdvance(0.01)
    for nm in ("u", "v", "w"):
        assert _rel(g.get(nm), gold[f"ig_{nm}1"]) < TOL


# ---- the rows widened beyond the periodic hot path (tests/golden/make_widened_golden.py) ----
WIDE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "widened_golden.npz")


@pytest.fixture(scope="module")
def wide():
    return np.load(WIDE)


def _wide_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_widened_golden", os.path.join(os.path.dirname(WIDE), "make_widened_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_reproduces_widened_golden(oracle, wide):
    from oracle import igrid_oracle as IG
    from oracle import ops_periodic_oracle as OP
    from oracle import stagg_np_oracle as SN
    M = _wide_module()
    f = wide["f3d_in"]
    assert np.array_equal(oracle.filter3D(f, 2), wide["f3d_cf90_x2"])
    assert np.array_equal(oracle.filter3D(f, 1, ("gaussian", "cf90", "gaussian")), wide["f3d_mixed_x1"])
    g, c = wide["opp_in"], wide["opp_cin"]
    mz, my, mx = g.shape
    op = OP.OpsPeriodic(mx, my, mz, 2 * np.pi / mx, 2 * np.pi / my, 2 * np.pi / mz)
    for nm, val in (("ddx", op.ddx(g)), ("ddy", op.ddy(g)), ("ddz", op.ddz(g)), ("ddz_c2c", op.ddz_cmplx2cmplx(c)), ("poisson", op.SolvePoisson(g)),
                    ("dealias", op.dealiasField(g))):
        assert _rel(val, wide["opp_" + nm]) < 1e-13, nm          # pocketfft: last-bit differences across numpy builds
    n = wide["snp_fC"].shape[0]
    for tag, (te, be, ts, bs) in M.WALLS.items():
        ref = SN.CD06StaggNP(n, 1.0 / n, te, be, ts, bs)
        for name in M.STAGG_OPS:
            edge = name in M.STAGG_EDGE_IN
            assert np.array_equal(getattr(ref, name)(wide["snp_fE" if edge else "snp_fC"]), wide[f"snp_{tag}_{name}"])
            assert np.array_equal(getattr(ref, name)(wide["snp_cE" if edge else "snp_cC"]), wide[f"snp_{tag}_c_{name}"])
    pf = IG.Pade6stagg(n, 2 * np.pi / n, scheme=2)
    for name in M.PADE_OPS:
        edge = name in M.PADE_EDGE_IN
        assert _rel(getattr(pf, name)(wide["four_fE"] if edge else wide["snp_fC"]), wide[f"four_{name}"]) < 1e-13
        assert _rel(getattr(pf, name)(wide["four_cE"] if edge else wide["snp_cC"]), wide[f"four_c_{name}"]) < 1e-13
    uh, vh, wh = wide["wp_u"], wide["wp_v"], wide["wp_w"]
    pz, py, nxh = uh.shape
    px = 2 * (nxh - 1)
    pd = [2 * np.pi / px, 2 * np.pi / py, 1.0 / pz]
    P = IG.PadePoisson(*pd, IG.Spectral(px, py, pz, *pd), IG.Spectral(px, py, pz + 1, *pd), IG.Pade6stagg(pz, pd[2], 1, isPeriodic=False), PeriodicInZ=False)
    for a, nm in zip(P.PressureProjection(uh, vh, wh), ("wp_u1", "wp_v1", "wp_w1")):
        assert _rel(a, wide[nm]) < 1e-13
    Ps = IG.PadePoisson(*pd, IG.Spectral(px, py, pz, *pd), IG.Spectral(px, py, pz + 1, *pd), IG.Pade6stagg(pz, pd[2], 1, isPeriodic=False),
                        PeriodicInZ=False, computeStokesPressure=True, Lz=1.0)
    for a, nm in zip(Ps.PressureProjection(uh, vh, wide["wps_w"]), ("wps_u1", "wps_v1", "wps_w1")):
        assert _rel(a, wide[nm]) < 1e-13
    for tag, ((U, V, W), (Lx, Ly, Lz, Re), kw) in M.igrid_cases().items():
        assert np.array_equal(U, wide[f"ig_{tag}_U0"]) and np.array_equal(W, wide[f"ig_{tag}_W0"])
        m = U.shape[0]
        sim = IG.IGrid(m, m, m, Lx, Ly, Lz, Re, U, V, W, **kw)
        sim.timeAdvance(0.005)
        for nm in ("u", "v", "w"):
            assert _rel(getattr(sim, nm), wide[f"ig_{tag}_{nm}1"]) < 1e-12, (tag, nm)
    hf = IG.HITForcing(IG.Spectral(px, py, pz, *pd), kmin=2.0, kmax=6.0, Nwaves=12, tidStart=3, RandSeedToAdd=1)
    for step in range(2):
        hf.pick_random_wavenumbers()
        assert np.array_equal(np.stack([hf.wave_x, hf.wave_y, hf.wave_z]), wide[f"hit_waves_step{step}"])
        hf.update_seeds()


def test_product_host_code_reproduces_widened_golden(pdo, wide):
    """No GPU needed: the staggered wall operators' host-device line routine and the forcing's draw, run on the host through the
    test hooks, against the committed vectors."""
    import ctypes as C
    M = _wide_module()
    n = wide["snp_fC"].shape[0]
    for tag, (te, be, ts, bs) in M.WALLS.items():
        for op, name in enumerate(M.STAGG_OPS):
            for pre, key in (("", "snp_f"), ("c_", "snp_c")):
                fin = np.ascontiguousarray(wide[key + ("E" if name in M.STAGG_EDGE_IN else "C")])
                want = wide[f"snp_{tag}_{pre}{name}"]
                got = np.zeros_like(want)
                ncols = fin.shape[1] * fin.shape[2] * (2 if pre else 1)
                rc = pdo.lib().pdo_debug_stagg_np_host(op, n, 1.0 / n, int(be), int(te), int(bs), int(ts), C.c_void_p(fin.ctypes.data),
                                                       C.c_void_p(got.ctypes.data), ncols)
                assert rc == 0 and _rel(got, want) < 1e-13, (tag, name, pre)
    seeds = (C.c_longlong * 4)()
    for step in range(2):
        w = [(C.c_int * 12)() for _ in range(3)]
        assert pdo.lib().pdo_debug_hit_draw(2.0, 6.0, 12, 3, 1, step, seeds, *w) == 0
        assert np.array_equal(np.array([list(a) for a in w]), wide[f"hit_waves_step{step}"])


@pytest.mark.gpu
def test_cuda_reproduces_widened_golden(pdo, wide):
    import torch
    M = _wide_module()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    f = wide["f3d_in"]
    nz, ny, nx = f.shape
    d = _spacing(f)
    gp = pdo.decomp_2d.init(nx, ny, nz, 1, 1)
    ops = pdo.vector_ops()
    ops.init(gp, *d, "cd10")
    for methods, numtimes, key in ((("cf90",) * 3, 2, "f3d_cf90_x2"), (("gaussian", "cf90", "gaussian"), 1, "f3d_mixed_x1")):
        fil = pdo.filters()
        fil.init(gp, True, True, True, *methods)
        a = dev(f)
        ops.filter3D(fil, a, numtimes)
        assert _rel(a.cpu().numpy(), wide[key]) < TOL
    g, c = wide["opp_in"], wide["opp_cin"]
    mz, my, mx = g.shape
    op = pdo.Ops_Periodic()
    op.init(mx, my, mz, 2 * np.pi / mx, 2 * np.pi / my, 2 * np.pi / mz)
    assert _rel(op.ddx(dev(g)).cpu().numpy(), wide["opp_ddx"]) < TOL
    assert _rel(op.ddy(dev(g)).cpu().numpy(), wide["opp_ddy"]) < TOL
    assert _rel(op.ddz(dev(g)).cpu().numpy(), wide["opp_ddz"]) < TOL
    assert _rel(op.ddz_cmplx2cmplx(dev(c)).cpu().numpy(), wide["opp_ddz_c2c"]) < TOL
    assert _rel(op.SolvePoisson_oop(dev(g)).cpu().numpy(), wide["opp_poisson"]) < TOL
    assert _rel(op.dealiasField(dev(g)).cpu().numpy(), wide["opp_dealias"]) < TOL
    n = wide["snp_fC"].shape[0]
    for tag, (te, be, ts, bs) in M.WALLS.items():
        st = pdo.cd06stagg()
        st.init(n, 1.0 / n, isTopEven=te, isBotEven=be, isTopSided=ts, isBotSided=bs)
        for name in M.STAGG_OPS:
            edge = name in M.STAGG_EDGE_IN
            assert _rel(getattr(st, name)(dev(wide["snp_fE" if edge else "snp_fC"])).cpu().numpy(), wide[f"snp_{tag}_{name}"]) < TOL
            assert _rel(getattr(st, name)(dev(wide["snp_cE" if edge else "snp_cC"])).cpu().numpy(), wide[f"snp_{tag}_c_{name}"]) < TOL
    for tag, ((U, V, W), (Lx, Ly, Lz, Re), kw) in M.igrid_cases().items():
        m = U.shape[0]
        kw = dict(kw)
        hit, sgs = kw.pop("HITForcing_", None), kw.pop("SGS_", None)
        sim = pdo.igrid()
        sim.init(m, m, m, Lx, Ly, Lz, Re, U, V, W, computeAllGradients=sgs is not None, **kw)
        if sgs:
            sim.enableSGS(**sgs)
        if hit:
            sim.enableHITForcing(**hit)
        sim.timeAdvance(0.005)
        for nm in ("u", "v", "w"):
            assert _rel(sim.get(nm), wide[f"ig_{tag}_{nm}1"]) < 1e-10, (tag, nm)
