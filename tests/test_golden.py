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
