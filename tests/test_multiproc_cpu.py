"""CPU, world_size 2 over gloo: the host-side rendezvous logic of the N>1 path (NCCL-id broadcast, per-rank
decomposition arithmetic tiling the global domain, bench.py's reference arm under torchrun)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
import padeops_b200 as pdo
from padeops_b200 import _lib
# 1) per-rank pencils from the product's decomposition arithmetic must tile the global box exactly once
nx, ny, nz, pr, pc = 17, 10, 9, 1, world
info = pdo.decomp_info.for_rank(nx, ny, nz, pr, pc, rank)
infos = [None] * world
dist.all_gather_object(infos, info)
for pen in "xyz":
    cover = np.zeros((nz, ny, nx), dtype=int)
    for d in infos:
        st, en = d[pen + "st"], d[pen + "en"]
        cover[st[2]-1:en[2], st[1]-1:en[1], st[0]-1:en[0]] += 1
    assert (cover == 1).all(), pen
# 1b) decomp_2d_write_one's per-rank writer with REAL concurrency: rank 0 creates / truncates, everyone else writes its own sub-box of
#     the same file at the same time; the file must be the global array, and every rank reads its pencil back (io_test.f90)
import ctypes as C
L = pdo.lib()
i3 = lambda v: (C.c_int * 3)(*[int(x) for x in v])
G = np.arange(1, nx * ny * nz + 1, dtype=np.float64).reshape(nz, ny, nx)
fn = os.path.join(os.environ["PDO_TMPDIR"], "u.dat").encode()
for pen in "xyz":
    sz, st = info[pen + "sz"], [a - 1 for a in info[pen + "st"]]
    blk = np.ascontiguousarray(G[st[2]:st[2] + sz[2], st[1]:st[1] + sz[1], st[0]:st[0] + sz[0]])
    if rank == 0:
        assert L.pdo_io_write_block(fn, i3((nx, ny, nz)), i3(sz), i3(st), 1, C.c_void_p(blk.ctypes.data), 1) == 0
    dist.barrier()
    if rank != 0:
        assert L.pdo_io_write_block(fn, i3((nx, ny, nz)), i3(sz), i3(st), 1, C.c_void_p(blk.ctypes.data), 0) == 0
    dist.barrier()
    assert open(fn, "rb").read() == G.tobytes(), pen
    back = np.zeros_like(blk)
    assert L.pdo_io_read_block(fn, i3((nx, ny, nz)), i3(sz), i3(st), 1, C.c_void_p(back.ctypes.data)) == 0
    assert np.array_equal(back, blk)
    dist.barrier()
# 2) the NCCL-id broadcast reaches pdo_comm_init on every rank; without a GPU it must refuse (no CPU fallback)
try:
    pdo.decomp_2d.comm_init()
    outcome = "inited"
except pdo.PadeOpsError as e:
    outcome = "err%%d" %% e.code
outs = [None] * world
dist.all_gather_object(outs, outcome)
if rank == 0:
    print("GLOO_WORKER", outs, flush=True)
dist.barrier()
dist.destroy_process_group()
''' % ROOT


def test_world2_gloo_host_logic(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(w)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, PDO_TMPDIR=str(tmp_path)))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import torch
    want = "inited" if torch.cuda.is_available() else "err1004"
    assert f"GLOO_WORKER ['{want}', '{want}']" in r.stdout, r.stdout[-2000:]


def test_reference_arm_under_torchrun_prints_once():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"]
    env = dict(os.environ, PDO_BENCH_CPU_N="128")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["cpu_baseline"]["kind"] == "port"
    assert lines[0]["value"] > 0 and lines[0]["e2e"]["h2d_bytes_per_step"] == 0


def test_committed_bench_lines_meet_the_contract():
    """The bench lines kept under profiles/ (written by bench.py on the B200 boxes) and a fresh reference-arm line satisfy the
    driver's JSON contract: tools/check_bench_line.py."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from check_bench_line import check
    for n in (1, 2, 4):
        fn = os.path.join(ROOT, "profiles", f"r01_bench_n{n}.json")
        line = [json.loads(l) for l in open(fn) if l.startswith("{")][0]
        assert check(line) == [], (fn, check(line))
        assert line["n_gpus"] == n
    full = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_n1.json")))
    assert full["e2e"]["h2d_bytes_per_step"] > 0 and full["cpu_baseline"]["kind"] == "port" and full["roofline"]["traffic"]
