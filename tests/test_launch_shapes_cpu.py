"""CPU: launch-shape logic of the cluster + TMA strided kernel (banded.cu: ctma_chunks_per_cta) through a host-only hook.
Every shape it returns must be launchable on sm_100a: portable cluster size, 64..512 threads, one tensor-map box per CTA,
halo reach inside the neighbouring CTA, three (or, when the ping-pong exchange arrays are large, two) tile buffers + exchange
arrays within the 227 KB opt-in shared memory."""
import ctypes as C

import pytest

CAP = 227 * 1024 - 256


def cfg(pdo, P, XT, HB, HW, BW, pc_max):
    smem = C.c_longlong(0)
    pc = pdo.lib().pdo_debug_ctma_config(P, XT, HB, HW, BW, pc_max, C.byref(smem))
    return pc, smem.value


@pytest.mark.parametrize("XT", [32, 64])
@pytest.mark.parametrize("pc_max", [4, 16])
@pytest.mark.parametrize("HB,HW,BW", [(3, 3, 2), (2, 2, 1), (4, 7, 2), (4, 1, 0)])
def test_every_returned_shape_is_launchable(pdo, XT, pc_max, HB, HW, BW):
    for P in range(1, 129):
        pc, smem = cfg(pdo, P, XT, HB, HW, BW, pc_max)
        if pc == 0:
            continue
        assert P % pc == 0 and 1 <= P // pc <= 8
        assert 64 <= XT * pc <= 512 and pc * 32 <= 256 and pc <= pc_max
        assert BW == 0 or HW <= pc
        assert 0 < smem <= CAP
        want = [8 * (nb * (pc * 32 + 2 * HB) * XT + 4 * max(BW, 1) * (pc + 2 * HW) * XT) for nb in (3, 2)]
        assert smem == (want[0] if want[0] <= CAP else want[1])


def test_bench_shapes_are_covered(pdo):
    # 1024-point lines (32 chunks): CD10 gets 4 chunks per CTA in clusters of 8 (two CTAs per SM) and 8 x 4 (one per SM)
    assert cfg(pdo, 32, 32, 3, 3, 2, 4)[0] == 4
    assert cfg(pdo, 32, 32, 3, 3, 2, 16)[0] == 8
    # CF90's reach of 7 chunks needs 8 chunks per CTA; 64-column tiles cannot hold it
    assert cfg(pdo, 32, 32, 4, 7, 2, 4)[0] == 0 and cfg(pdo, 32, 32, 4, 7, 2, 16)[0] == 8 and cfg(pdo, 32, 64, 4, 7, 2, 16)[0] == 0
    # the 8-GPU bench's z-slab (1024 planes per GPU) and the 2048-point y lines of a 2048^3 run
    assert cfg(pdo, 32, 32, 3, 3, 2, 16)[1] == 8 * (3 * 262 * 32 + 4 * 2 * 14 * 32)   # CD10: three 262-row tile buffers + ping-pong exchange
    assert cfg(pdo, 32, 32, 4, 7, 2, 16)[1] == 8 * (2 * 264 * 32 + 4 * 2 * 22 * 32)   # CF90: its 7-chunk reach leaves room for two
    assert cfg(pdo, 64, 32, 3, 3, 2, 16)[0] == 8                   # cluster of 8
    assert cfg(pdo, 6, 32, 3, 3, 2, 16)[0] == 0                    # 6 chunks: no power-of-two split that holds the halo
