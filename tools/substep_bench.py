#!/usr/bin/env python
"""ms per igrid time step / RK substep (BASELINE.json metric iii) on synthetic Taylor-Green + broadband fields.
Usage: python tools/substep_bench.py [n] [scheme] [steps] [variant]   (torchrun for multi-GPU; grid 1 x N)
variant: "base" (skew-symmetric, CD06 in z, viscous: the BASELINE workload), "hit" (the authors' HIT_Periodic deck: rotational form,
Fourier collocation in z, AMD model Csgs = 1.67, shell forcing kmin 4 kmax 5 Nwaves 20 Eps 0.05, Re = 1e10), "slip" (slip walls).
The "hit" variant is a TIMING workload only: the shell forcing scales every forced mode by 1 / its own energy, which this synthetic
start (Taylor-Green + two shell modes + 1e-3 noise) keeps tiny, so the field grows without bound within a few steps (the CPU oracle
does the same on the same start) and max_div is meaningless there; parity of that deck is tests/test_igrid_gpu.py's job."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import padeops_b200 as pdo


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    scheme = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    variant = sys.argv[4] if len(sys.argv) > 4 else "base"
    dbg = lambda m: (print(f"[rank {rank}] {m}", file=sys.stderr, flush=True) if os.environ.get("PDO_BENCH_TRACE") else None)
    dbg("process group up")
    pdo.decomp_2d.comm_init()
    dbg("comm_init done")
    info = pdo.decomp_info.for_rank(n, n, n, 1, world, rank)
    infoE = pdo.decomp_info.for_rank(n, n, n + 1, 1, world, rank)
    d = 2 * np.pi / n
    x = torch.arange(n, device="cuda", dtype=torch.float64) * d
    def zc(inf, edge):
        k = torch.arange(inf["xst"][2] - 1, inf["xen"][2], device="cuda", dtype=torch.float64)
        return (k * d) if edge else ((k + 0.5) * d)
    X, Y = x[None, None, :], x[None, :, None]
    ZC, ZE = zc(info, False)[:, None, None], zc(infoE, True)[:, None, None]
    u = (torch.sin(X) * torch.cos(Y) * torch.cos(ZC)).contiguous()
    v = (-torch.cos(X) * torch.sin(Y) * torch.cos(ZC)).contiguous()
    w = (0.1 * torch.sin(2 * X) * torch.sin(Y) * torch.sin(ZE)).contiguous()
    g = pdo.igrid()
    if variant == "hit":
        # the shell forcing scales with 1 / (energy in 4 <= |k| <= 5): Taylor-Green alone has none there
        # (the forced wavevectors are drawn at random inside the shell and each is scaled by 1 / its own energy: every mode needs some)
        gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
        u = (u + 0.05 * torch.sin(4 * Y) * torch.cos(2 * ZC) * torch.cos(X) + 1e-3 * torch.randn(u.shape, generator=gen, device="cuda", dtype=torch.float64)).contiguous()
        v = (v + 0.05 * torch.sin(3 * X) * torch.cos(3 * ZC) * torch.cos(Y) + 1e-3 * torch.randn(v.shape, generator=gen, device="cuda", dtype=torch.float64)).contiguous()
        w = (w + 1e-3 * torch.randn(w.shape, generator=gen, device="cuda", dtype=torch.float64)).contiguous()
        g.init(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, 1.0e10, u, v, w, TimeSteppingScheme=scheme, prow=1, pcol=world, AdvectionTerm=0,
               NumericalSchemeVert=2, computeAllGradients=True)
        g.enableSGS(SGSModelID=2, Csgs=1.67)
        g.enableHITForcing(kmin=4.0, kmax=5.0, Nwaves=20, EpsAmplitude=0.05)
        label = "HIT_Periodic deck: rotational, Fourier z, AMD, shell forcing"
    elif variant == "slip":
        w = (0.1 * torch.sin(2 * X) * torch.sin(Y) * torch.sin(ZE / 2)).contiguous()     # w = 0 on both walls of [0, 2 pi]
        g.init(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, 1600.0, u, v, w, TimeSteppingScheme=scheme, prow=1, pcol=world, PeriodicInZ=False)
        label = "slip walls, skew-symmetric, CD06 z, viscous"
    else:
        g.init(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, 1600.0, u, v, w, TimeSteppingScheme=scheme, prow=1, pcol=world)
        label = "skew-symmetric, CD06 z, viscous"
    dbg("igrid init done")
    dt = 0.2 * d
    g.timeAdvance(dt)
    torch.cuda.synchronize()
    dbg("first step done")
    L = pdo.lib()
    l0 = L.pdo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.timeAdvance(dt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    nsub = 3 if scheme == 1 else 5
    max_div = g.maxDivergence()   # collective: every rank calls it
    if rank == 0:
        print(json.dumps({"workload": f"igrid {n}^3, {label}, {'TVD-RK3' if scheme == 1 else 'SSP-RK45'}",
                          "n_gpus": world, "ms_per_step": round(ms, 3), "ms_per_substep": round(ms / nsub, 3),
                          "launches_per_substep": (L.pdo_launch_count() - l0) // (steps * nsub),
                          "Mpoints_per_s_per_substep": round(n ** 3 / (ms / nsub) / 1e3, 1), "max_div": max_div}), flush=True)


if __name__ == "__main__":
    main()
