"""Full-field parity at the BASELINE configurations, through the DEFAULT dispatch (no forced variant, no plan): the kernels these
tests exercise are the kernels bench.py times.
  config 2   CD06 dd1 / dd2 / dd3 + CF90 filter1 / filter2 / filter3 on a 512^3 double periodic field (2^27 points each)
  headline   CD10 ddx / ddy (+ d2) on a 1024 x 1024 x 64 slab and ddz (+ d2) on 1024 x 64 x 1024: 1024-point lines, 2^26 points
Bar: 1e-12 relative to max|ref| against the oracle (north_star); the variant the library reports is pinned to the deterministic
table of banded.cu: default_variant, the same codes bench.py prints in roofline.per_kernel."""
import numpy as np
import pytest

from conftest import broadband

pytestmark = pytest.mark.gpu
TOL = 1e-12
XTMA, PIPE1, CPIPE, STMA, CTMA32, CPIPE_T = 1000, 6, 5, 7, 9, 11


def _rel(got, ref):
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)


def _field(shape, seed):
    """broadband() costs minutes at 2^27 points: a few product modes + noise do the same job (O(1), no symmetry a wrong index
    could hide behind)"""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    x, y, z = (np.arange(m) * (2 * np.pi / m) for m in (nx, ny, nz))
    f = np.sin(3 * x + 0.3)[None, None, :] * np.cos(5 * y + 0.1)[None, :, None] * np.sin(2 * z + 0.7)[:, None, None]
    f += 0.5 * np.cos(11 * x)[None, None, :] * np.sin(7 * z)[:, None, None] + 0.25 * np.sin(9 * y + 1.0)[None, :, None]
    f += 1e-3 * rng.uniform(-1, 1, size=shape)
    return f


def test_config2_cd06_cf90_512_cubed(pdo, oracle):
    import torch
    n = 512
    d = 2 * np.pi / n
    f = _field((n, n, n), 2)
    fd = torch.from_numpy(f).cuda()
    out = torch.empty_like(fd)
    c06, cf = pdo.cd06(), pdo.cf90()
    assert c06.init(n, d) == 0 and cf.init(n) == 0
    L = pdo.lib()
    L.pdo_debug_set_variant(-1, -1)
    want06 = {0: XTMA, 1: CTMA32, 2: CTMA32}
    want90 = {0: XTMA, 1: STMA, 2: CPIPE_T}
    for ax, fn in enumerate((c06.dd1, c06.dd2, c06.dd3)):
        fn(fd, out)
        assert L.pdo_debug_last_variant() == want06[ax], ("cd06", ax, L.pdo_debug_last_variant())
        assert _rel(out.cpu().numpy(), oracle.cd06(f, d, ax)) < TOL, ("cd06", ax)
    for ax, fn in enumerate((cf.filter1, cf.filter2, cf.filter3)):
        fn(fd, out)
        assert L.pdo_debug_last_variant() == want90[ax], ("cf90", ax, L.pdo_debug_last_variant())
        assert _rel(out.cpu().numpy(), oracle.cf90(f, ax)) < TOL, ("cf90", ax)


@pytest.mark.parametrize("which", [1, 2])
def test_headline_cd10_1024_point_lines(pdo, oracle, which):
    import torch
    n = 1024
    d = 2 * np.pi / n
    c10 = pdo.cd10()
    assert c10.init(n, d) == 0
    L = pdo.lib()
    L.pdo_debug_set_variant(-1, -1)
    fx = _field((64, n, n), 3)            # x and y lines of 1024 points
    fxd = torch.from_numpy(fx).cuda()
    out = torch.empty_like(fxd)
    for ax, want in ((0, XTMA), (1, PIPE1)):
        fn = ((c10.dd1, c10.dd2, c10.dd3) if which == 1 else (c10.d2d1, c10.d2d2, c10.d2d3))[ax]
        fn(fxd, out)
        assert L.pdo_debug_last_variant() == want, (ax, L.pdo_debug_last_variant())
        assert _rel(out.cpu().numpy(), oracle.cd10(fx, d, ax, which)) < TOL, (ax, which)
    fz = _field((n, 64, n), 4)            # z lines of 1024 points, row stride 512 KB ... use the far-row shape of the bench:
    fzd = torch.from_numpy(fz).cuda()
    outz = torch.empty_like(fzd)
    fn = c10.dd3 if which == 1 else c10.d2d3
    fn(fzd, outz)
    assert L.pdo_debug_last_variant() in (PIPE1, CPIPE), L.pdo_debug_last_variant()
    assert _rel(outz.cpu().numpy(), oracle.cd10(fz, d, 2, which)) < TOL
    del fzd, outz
    fz2 = _field((n, 128, n), 5)          # n1 = 128 * 1024 doubles = 1 MB rows: the megabyte-stride branch (cpipe), as at 1024^3
    fz2d = torch.from_numpy(fz2).cuda()
    out2 = torch.empty_like(fz2d)
    fn(fz2d, out2)
    assert L.pdo_debug_last_variant() == CPIPE, L.pdo_debug_last_variant()
    assert _rel(out2.cpu().numpy(), oracle.cd10(fz2, d, 2, which)) < TOL


def test_explicit_plan_is_stored_and_changes_nothing(pdo, oracle):
    """pdo_cd10_plan times the candidates once, at set-up; calls on that shape then run the stored variant and agree with the
    default dispatch to rounding."""
    import torch
    n = 512
    d = 2 * np.pi / n
    f = _field((96, n, 384), 6)           # 2^24+ points, so the planner's size threshold is met
    fd = torch.from_numpy(f).cuda()
    c10 = pdo.cd10()
    assert c10.init(n, d) == 0
    L = pdo.lib()
    L.pdo_debug_set_variant(-1, -1)
    base = c10.dd2(fd).cpu().numpy()
    v_default = L.pdo_debug_last_variant()
    v1, v2 = c10.plan(1, 384, 96)
    assert v1 != 0 and v2 != 0
    got = c10.dd2(fd).cpu().numpy()
    assert L.pdo_debug_last_variant() == v1
    assert _rel(got, base) < 1e-14 and _rel(got, oracle.cd10(f, d, 1, 1)) < TOL
    other = pdo.cd10()                    # a plan lives in its handle only
    assert other.init(n, d) == 0
    other.dd2(fd)
    assert L.pdo_debug_last_variant() == v_default
