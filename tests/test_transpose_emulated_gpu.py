"""a16 / a17 on ONE GPU: a p_row x p_col process grid is emulated with one device buffer pair per simulated rank, and the
library runs the very geometry (2DECOMP's counts, displacements, pack / unpack boxes; 2D» decomp_2d.f90:739-794,
transpose_x_to_y.f90:332-513, transpose_y_to_z.f90:342-431) and the very kernels of the multi-GPU transposes on them — the fused
store kernel (box_push_kernel), the copy-engine plane, the TMA bulk-push kernel (bulk_push_kernel) and the NCCL path's
pack / unpack kernels (cudaMemcpyAsync standing in for the grouped ncclSend / ncclRecv).  Only the epoch-flag handshakes and
the NVLink hop are absent (tests/mp_worker.py covers those on 2 / 4 / 8 GPUs).  Bar: bit-exact against the oracle's
simulated-rank ALLTOALLV."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIRS = [(0, "x", "y"), (1, "y", "x"), (2, "y", "z"), (3, "z", "y")]
PATHS = {1: "sm store kernel", 2: "copy engines", 3: "TMA bulk push", 4: "pack / exchange / unpack"}


def _run(pdo, O, nx, ny, nz, pr, pc, cplx, path, seed=5):
    import torch
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((nz, ny, nx))
    if cplx:
        G = G + 1j * rng.standard_normal((nz, ny, nx))
    R = pr * pc
    pens = {p: O.scatter_global(G, nx, ny, nz, pr, pc, p) for p in "xyz"}
    L = pdo.lib()
    for d, s, t in DIRS:
        ref = O.transpose(d, nx, ny, nz, pr, pc, pens[s])
        src = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in pens[s]]
        dst = [torch.full(tuple(a.shape), float("nan"), dtype=src[0].dtype, device="cuda") for a in pens[t]]
        sp = (C.c_void_p * R)(*[a.data_ptr() for a in src])
        dp = (C.c_void_p * R)(*[a.data_ptr() for a in dst])
        rc = L.pdo_debug_transpose_emulate(nx, ny, nz, pr, pc, d, 2 if cplx else 1, path, sp, dp, None)
        assert rc == 0, L.pdo_last_error()
        for r in range(R):
            got = dst[r].cpu().numpy()
            assert got.shape == ref[r].shape
            assert np.array_equal(got, ref[r]), (PATHS[path], f"{s}->{t}", (pr, pc), (nx, ny, nz), cplx, r)
            assert np.array_equal(ref[r], pens[t][r])      # the oracle's own consistency: a transpose of the scattered global field


@pytest.mark.parametrize("path", sorted(PATHS))
@pytest.mark.parametrize("grid", [(1, 4), (4, 1), (2, 2), (2, 4), (3, 2)])
@pytest.mark.parametrize("cplx", [False, True])
def test_emulated_grid_transposes_are_bit_exact(pdo, oracle, grid, cplx, path):
    pr, pc = grid
    for (nx, ny, nz) in [(16, 12, 8), (17, 9, 11), (33, 16, 10)]:      # even, uneven (extras to the last ranks), spectral-like nx/2+1
        if min(nx, ny) < pr or min(ny, nz) < pc:
            continue
        _run(pdo, oracle, nx, ny, nz, pr, pc, cplx, path)


@pytest.mark.parametrize("path", [1, 3, 4])
@pytest.mark.parametrize("grid", [(1, 8), (2, 4)])
def test_emulated_transposes_many_chunks(pdo, oracle, grid, path):
    """Boxes large enough for the persistent kernels to rotate their stage rings many times (bulk push: 32 KB pieces, 6 stages)."""
    pr, pc = grid
    _run(pdo, oracle, 256, 64, 72, pr, pc, False, path, seed=11)
    _run(pdo, oracle, 129, 40, 48, pr, pc, True, path, seed=12)
