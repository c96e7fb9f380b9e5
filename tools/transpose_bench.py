#!/usr/bin/env python
"""Pencil-transpose throughput (torchrun, one rank per GPU): GB/s sent per GPU and fraction of the NVLink 5
all-to-all roofline (900 GB/s per direction per GPU), SURVEY.md 8d.  Bit-exactness is checked in tests/mp_worker.py.
Usage: torchrun --nproc-per-node P tools/transpose_bench.py [n] [--grids 1xP,2x4]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import padeops_b200 as pdo
from padeops_b200 import decomp as dc

NVLINK = 900.0


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pdo.decomp_2d.comm_init()
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if args else 1024
    grids = [(1, world)] + ([(2, world // 2)] if world >= 4 else []) + ([(world, 1)] if world > 1 else [])
    for a in sys.argv[1:]:
        if a.startswith("--grids="):
            grids = [tuple(int(v) for v in g.split("x")) for g in a.split("=")[1].split(",")]
    for (pr, pc) in grids:
        for cplx in (False, True):
            nx = n // 2 + 1 if cplx else n
            gp = pdo.decomp_info(nx, n, n, pr, pc)
            dt = torch.complex128 if cplx else torch.float64
            bufs = {p: torch.zeros(tuple(reversed(getattr(gp, p + "sz"))), dtype=dt, device="cuda") for p in "xyz"}
            bufs["x"].real.uniform_() if cplx else bufs["x"].uniform_()
            if "--nccl" not in sys.argv:
                for p in "xyz":
                    pdo.decomp_2d.register(bufs[p])
            for name, fn, s, d, p in (("x_to_y", dc.transpose_x_to_y, "x", "y", pr), ("y_to_x", dc.transpose_y_to_x, "y", "x", pr),
                                      ("y_to_z", dc.transpose_y_to_z, "y", "z", pc), ("z_to_y", dc.transpose_z_to_y, "z", "y", pc)):
                for _ in range(3):
                    fn(bufs[s], bufs[d], gp)
                torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record()
                for _ in range(reps):
                    fn(bufs[s], bufs[d], gp)
                e1.record(); torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = t.item()
                L = bufs[s].numel() * bufs[s].element_size()
                sent = L * (p - 1) / p
                if rank == 0:
                    print(json.dumps({"op": "transpose_" + name, "complex": cplx, "global": [nx, n, n], "grid": [pr, pc], "n_gpus": world, "path": "nccl" if "--nccl" in sys.argv else "p2p",
                                      "local_MiB": round(L / 2**20, 1), "sent_MiB_per_gpu": round(sent / 2**20, 1), "ms": round(ms, 4),
                                      "link_GBps_per_gpu": round(sent / ms / 1e6, 1), "nvlink_frac": round(sent / ms / 1e6 / NVLINK, 3),
                                      "hbm_GBps_local": round(2 * L / ms / 1e6, 1)}), flush=True)
            for p in "xyz":
                pdo.decomp_2d.deregister(bufs[p])
            del bufs
            gp.destroy()
    dist.barrier()
    pdo.decomp_2d.finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
