"""Every strided / contiguous kernel variant against the CPU oracle at the BASELINE line lengths (512, 1024,
2048 → clusters of 1, 2, 4 CTAs), including ragged x-extents (partial tiles, odd n1 → 8-byte cp.async path)
and the staggered edge planes.  The default dispatch picks one variant per shape; here each is forced."""
import numpy as np
import pytest

from conftest import broadband

pytestmark = pytest.mark.gpu
TOL = 1e-12

MODES = {"auto": 0, "t512": 1, "t256": 2, "cluster": 3, "cluster4": 4, "cpipe": 5, "pipe1": 6, "stma": 7, "ctma64": 8, "ctma32": 9, "ctma32s": 10, "cpipe_t": 11}


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _relerr(got, ref):
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.fixture
def variant(pdo):
    L = pdo.lib()
    yield lambda mode, xth=-1: L.pdo_debug_set_variant(MODES[mode], xth)
    L.pdo_debug_set_variant(-1, -1)


@pytest.mark.parametrize("mode", ["cpipe", "pipe1", "cluster", "t512", "stma", "ctma64", "ctma32", "ctma32s", "cpipe_t"])
@pytest.mark.parametrize("n,n1", [(512, 96), (512, 33), (1024, 64), (1024, 45), (2048, 40), (1024, 46), (256, 90)])
def test_strided_variants_cd10_cf90(pdo, oracle, variant, mode, n, n1):
    """axis 1 (f(n1, n, n3)) and axis 2 (f(na, nb, n)) with the forced variant; enough tiles (>= 148) for the persistent ones."""
    if mode == "stma" and (n1 % 2 or n > 1024):
        pytest.skip("tensor-map rows must be 16-byte multiples; three 2048-row tiles do not fit in shared memory")
    if mode == "cpipe_t" and (n1 % 2 or n == 256):
        pytest.skip("tensor-map rows must be 16-byte multiples; cpipe needs lines of at least 16 chunks")
    if mode.startswith("ctma") and (n1 % 2 or (mode != "ctma32" and n == 2048) or n == 256):
        pytest.skip("tensor-map rows must be 16-byte multiples; 2048-point lines need clusters of 16 at 64 columns; "
                    "CF90's separator reach covers the whole 256-point line (dense tables)")
    variant(mode)
    d = 2 * np.pi / n
    c10, cf, c06 = pdo.cd10(), pdo.cf90(), pdo.cd06()
    assert c10.init(n, d) == 0 and cf.init(n) == 0 and c06.init(n, d) == 0
    n3 = 160 if n1 < 64 else 80
    f = broadband((n3, n, n1), seed=n + n1)
    fd = _dev(f)
    assert _relerr(c10.dd2(fd).cpu().numpy(), oracle.cd10(f, d, 1, 1)) < TOL
    assert _relerr(c10.d2d2(fd).cpu().numpy(), oracle.cd10(f, d, 1, 2)) < TOL
    narrow = mode in ("ctma64", "ctma32s")   # 4 chunks per CTA: CF90's separator reach (7 chunks) does not fit, must fail loudly
    if narrow:
        with pytest.raises(pdo.PadeOpsError):
            cf.filter2(fd)
    else:
        assert _relerr(cf.filter2(fd).cpu().numpy(), oracle.cf90(f, 1)) < TOL
    assert _relerr(c06.dd2(fd).cpu().numpy(), oracle.cd06(f, d, 1)) < TOL
    # axis 2: same memory seen as f(na, nb, n) with na*nb = n1*n3'
    g = broadband((n, 48 if n < 2048 else 24, 100), seed=n)
    gd = _dev(g)
    assert _relerr(c10.dd3(gd).cpu().numpy(), oracle.cd10(g, d, 2, 1)) < TOL
    if not narrow:
        assert _relerr(cf.filter3(gd).cpu().numpy(), oracle.cf90(g, 2)) < TOL
    ga = pdo.gaussian()
    assert ga.init(n) == 0
    assert _relerr(ga.filter3(gd).cpu().numpy(), oracle.gaussian(g, 2)) < TOL
    if mode in ("stma", "ctma64", "ctma32", "ctma32s", "cpipe_t"):   # the TMA variants never fall back: they run or the call fails
        assert pdo.lib().pdo_debug_last_variant() == MODES[mode]


@pytest.mark.parametrize("mode", ["cpipe", "pipe1", "stma"])
@pytest.mark.parametrize("n", [512, 1024])
@pytest.mark.parametrize("cplx", [False, True])
def test_staggered_edge_planes_variants(pdo, oracle, variant, mode, n, cplx):
    if mode == "stma" and not cplx:
        pytest.skip("odd n1: rows are not 16-byte multiples (the complex case, 2*n1 doubles per row, is covered)")
    variant(mode)
    dz = 2 * np.pi / n
    st = pdo.cd06stagg()
    st.init(n, dz)
    n2, n1 = 70, 77  # odd n1: the real case takes the 8-byte cp.async path, the complex one the 16-byte path

    def field(planes):
        a = broadband((planes, n2, n1), seed=planes)
        if cplx:
            a = a + 1j * broadband((planes, n2, n1), seed=planes + 1)
        return a
    fC, fE = field(n), field(n + 1)
    for name, fn, fin in [("ddz_E2C", st.ddz_E2C, fE), ("ddz_C2E", st.ddz_C2E, fC), ("interp_E2C", st.InterpZ_E2C, fE),
                          ("interp_C2E", st.InterpZ_C2E, fC), ("d2dz2_C2C", st.d2dz2_C2C, fC), ("d2dz2_E2E", st.d2dz2_E2E, fE)]:
        got = fn(_dev(fin)).cpu().numpy()
        ref = oracle.stagg(name, fin, n, dz)
        assert got.shape == ref.shape
        assert _relerr(got, ref) < TOL, (name, n, cplx, mode, _relerr(got, ref))


@pytest.mark.parametrize("xth", [128, 256, 1000, 1016])
@pytest.mark.parametrize("n", [512, 1024, 2048, 96])
def test_contiguous_variants(pdo, oracle, variant, xth, n):
    """128 / 256: register-staged x kernel by CTA size; 1000 / 1016: the TMA (bulk-copy) pipeline on the operator's own
    chunk length / on 16-point chunks.  A variant that does not cover a shape must fail loudly, never fall back."""
    if xth == 1016 and n % 32 != 0:
        pytest.skip("the M=16 alternate tables exist only next to M=32 ones")
    variant("auto", xth)
    d = 2 * np.pi / n
    c10, cf = pdo.cd10(), pdo.cf90()
    assert c10.init(n, d) == 0 and cf.init(n) == 0
    f = broadband((7, 31, n), seed=n)  # 217 lines: ragged last tile
    fd = _dev(f)
    assert _relerr(c10.dd1(fd).cpu().numpy(), oracle.cd10(f, d, 0, 1)) < TOL
    assert _relerr(c10.d2d1(fd).cpu().numpy(), oracle.cd10(f, d, 0, 2)) < TOL
    assert _relerr(cf.filter1(fd).cpu().numpy(), oracle.cf90(f, 0)) < TOL
    assert pdo.lib().pdo_debug_last_variant() == xth
    c06, ga = pdo.cd06(), pdo.gaussian()
    assert c06.init(n, d) == 0 and ga.init(n) == 0
    assert _relerr(c06.dd1(fd).cpu().numpy(), oracle.cd06(f, d, 0)) < TOL
    assert _relerr(ga.filter1(fd).cpu().numpy(), oracle.gaussian(f, 0)) < TOL


@pytest.mark.parametrize("xth", [1000, 1016])
def test_contiguous_tma_many_tiles(pdo, oracle, variant, xth):
    """Enough lines that every persistent CTA cycles its three tile buffers several times (148 CTAs x 8 lines x >3)."""
    variant("auto", xth)
    n = 1024
    d = 2 * np.pi / n
    c10 = pdo.cd10()
    assert c10.init(n, d) == 0
    f = broadband((5, 1203, n), seed=3)   # 6015 lines: 752 tiles of 8 -> 5+ per CTA, ragged tail
    got = c10.dd1(_dev(f)).cpu().numpy()
    assert _relerr(got, oracle.cd10(f, d, 0, 1)) < TOL
