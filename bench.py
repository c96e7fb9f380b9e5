#!/usr/bin/env python
"""bench.py — headline benchmark of the PadeOps operator hot path on B200.

Metric (BASELINE.json): Gpoints/s per compact derivative, with the HBM-roofline fraction.
Workload at N=1: CD10 ddx + ddy + ddz on a 1024^3 double-precision periodic field (the north_star target
config; one "step" = the three derivative calls, 3 * 2^30 points).  At N>1 (torchrun, one rank per GPU)
the global field is 1024 x 1024 x (1024 N), decomposed 1 x N like 2DECOMP would: every rank owns 2^30
points (weak scaling); ddx / ddy are pencil-local, ddz goes y->z transpose, derivative, z->y transpose
over NCCL, exactly the choreography of tests/test_derivatives_parallel.F90:94-126.

`value`       device-resident throughput (inputs already in HBM), CUDA-event timed, max over ranks.
`e2e`         same metric through the C ABI with HOST (pinned) buffers: H2D + kernel + D2H per call; next to it the box's own
              full-duplex pinned-copy ceiling (`pcie_ceiling_GBps_per_gpu_each_way`), which is what bounds this leg.
`roofline`    the slowest of the three legs: 16 B/pt algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json.
`cpu_baseline` the oracle port (flat-MPI emulation: one worker per host core, each owning a pencil) on the SAME 1024^3 field,
              1 warm-up + 2 timed passes.  `--impl reference` runs that arm alone with the given --steps / --warmup.
`substep`     BASELINE metric (iii): ms per igrid RK substep at 512^3 (strong scaling); flat copy in `substep_ms`.
`transposes`, `poisson`, `cd10_2048`   BASELINE configs 3 and 5 at this N: the four pencil transposes of the 1024^3 field (real +
              complex, grids 1xN and 2xN/2) with their NVLink fraction, PoissonPeriodic on 1024^3 (strong), CD10 ddx/ddy/ddz/d2dx2 on
              2048^3 (strong; skipped where two 64 GiB/N fields do not fit).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: the contract is ONE JSON line

METRIC = "Gpoints/s per CD10 derivative (ddx+ddy+ddz), double precision"
UNIT = "Gpoints/s"
BYTES_PER_POINT = 16.0  # read f once + write df once


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement run the way the reference runs — flat MPI, one worker per core,
# each worker owning the pencil a 2DECOMP rank would own (transposes between pencils are not timed,
# which favours the CPU).
# ------------------------------------------------------------------------------------------------
def cpu_arm(n_global, steps, warmup, cores=None, budget_s=None):
    """CD10 ddx + ddy + ddz of an n^3 field on the host cores.  Worker i owns the pencil a 2DECOMP rank of a 1 x C grid would own:
    (n, n, n/C) for x and y, (n, n/C, n) for z; input and output arrays are allocated and touched before the timed region
    (the reference's callers own both).  budget_s caps the timed steps by the rate measured during warm-up."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.build()
    cores = cores or os.cpu_count() or 1
    n = n_global
    d = 2 * np.pi / n
    nloc = [n // cores + (1 if i >= cores - n % cores else 0) for i in range(cores)]
    nloc = [nl for nl in nloc if nl > 0]
    W = len(nloc)
    fxy, oxy, fz, oz = [None] * W, [None] * W, [None] * W, [None] * W

    def make(i):   # the z-pencil block reuses the memory of the x/y one (values are immaterial to the timing): 16 n^3 bytes in all
        rng = np.random.default_rng(20240607 + i)
        fxy[i] = np.empty((nloc[i], n, n))
        rng.random(out=fxy[i].reshape(-1))
        oxy[i] = np.zeros((nloc[i], n, n))
        fz[i] = fxy[i].reshape(n, nloc[i], n)
        oz[i] = oxy[i].reshape(n, nloc[i], n)

    O.cd10(np.zeros((1, 1, n)), d, 0, 1)  # builds LU once (untimed, like init)

    def work(i):
        O.cd10(fxy[i], d, 0, 1, out=oxy[i])
        O.cd10(fxy[i], d, 1, 1, out=oxy[i])
        O.cd10(fz[i], d, 2, 1, out=oz[i])

    pts = 3.0 * n ** 3
    times, wt = [], []
    with ThreadPoolExecutor(max_workers=W) as ex:
        list(ex.map(make, range(W)))
        for it in range(max(1, warmup)):
            t0 = time.perf_counter()
            list(ex.map(work, range(W)))
            wt.append(time.perf_counter() - t0)
        if budget_s is not None:
            steps = max(1, min(steps, int(budget_s / max(wt[-1], 1e-9))))
        for it in range(steps):
            t0 = time.perf_counter()
            list(ex.map(work, range(W)))
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return {"value": pts / t / 1e9, "unit": UNIT, "cores": W, "kind": "port",
            "sample": f"CD10 ddx+ddy+ddz on the full {n}^3 field split 1x{W} (one worker per core, own pencil each; arrays caller-owned "
                      f"and pre-touched); {len(times)} timed passes after {max(1, warmup)} warm-up, {t*1e3:.1f} ms/pass; "
                      f"oracle/padeops_oracle.c compiled -O3 -march=native"}, t, len(times)


def run_reference(args):
    """The reference arm: the CPU restatement of the reference's own path (the Fortran + MPI original cannot be built in this image,
    DESIGN.md 3) on all host cores, SAME config as our arm (cd10 ddx+ddy+ddz on n^3, n = --n = 1024), same steps / warm-up; the
    timed steps are capped only if they would run past ~3 minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncpu = int(os.environ.get("PDO_BENCH_CPU_N", str(args.n)))
    warm = max(1, min(args.warmup, 3))
    cb, t, used = cpu_arm(ncpu, args.steps, warm, budget_s=170.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": used,
            "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cd10 ddx+ddy+ddz, periodic, {ncpu}x{ncpu}x{ncpu} field (CPU arm: the same field our 1-GPU arm runs)",
                       "n": ncpu, "global": [ncpu, ncpu, ncpu]},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML (SM clock + clock-event reasons) every ~2 ms from a thread while the timed region runs."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, dev):
        self.sm, self.mask, self.mx, self.err = [], 0, None, None
        self._stop = threading.Event()
        try:
            import pynvml as nv
            nv.nvmlInit()
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; map through it when set
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[dev]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else dev
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:  # noqa
            self.err = f"nvml unavailable: {e}"
            self.t = None
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception as e:  # noqa
                self.err = str(e)
                return
            time.sleep(0.002)

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err]}
        self._stop.set()
        self.t.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(v for k, v in self.BAD.items() if self.mask & k), "samples": len(sm)}


def substep_leg(n, world, rank, steps=2):
    """BASELINE.json metric (iii): ms per igrid RK substep.  Periodic box n^3 (strong scaling: the SAME global grid at every
    N, slabs 1 x N), Taylor-Green + a z-dependent perturbation, skew-symmetric advection, CD06 staggered z operators, viscous,
    TVD-RK3, fixed dt (the configuration of tests/test_igrid_gpu.py, which pins it against the oracle)."""
    import numpy as np
    import torch
    import padeops_b200 as pdo
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
    g.init(n, n, n, 2 * np.pi, 2 * np.pi, 2 * np.pi, 1600.0, u, v, w, TimeSteppingScheme=1, prow=1, pcol=world)
    dt = 0.2 * d
    g.timeAdvance(dt)
    torch.cuda.synchronize()
    L = pdo.lib()
    l0 = L.pdo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        g.timeAdvance(dt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    out = {"workload": f"igrid periodic {n}^3 (global, strong scaling), grid 1x{world}, skew-symmetric, CD06 z, viscous, TVD-RK3",
           "ms_per_substep": ms / 3.0, "ms_per_step": ms, "unit": "ms", "scaling": "strong",
           "launches_per_substep": int((L.pdo_launch_count() - l0) // (steps * 3)),
           "Mpoints_per_s": n ** 3 / (ms / 3.0) / 1e3, "max_divergence": float(g.maxDivergence())}
    g.destroy() if hasattr(g, "destroy") else None
    return out


NVLINK_GBS = 900.0   # NVLink 5, per direction per GPU


def _ev_time(fn, st, reps, warm, world):
    """median-free device timing: warm-up, barrier, reps back to back between two events on the launching stream; max over ranks"""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps):
        fn()
    b.record(st)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms


def transposes_leg(n, world, rank):
    """BASELINE config 3, first half: the four pencil transposes of an n^3 real field and of its (n/2+1) x n x n complex spectrum on
    grids 1 x N and 2 x N/2; bytes sent per GPU = L (p-1)/p against 900 GB/s per direction.  Destinations are
    registered (peer-writable), so the fused NVLink path runs; bit-exactness is tests/mp_worker.py's and the emulated-grid test's job."""
    import torch
    import padeops_b200 as pdo
    from padeops_b200 import decomp as dc
    st = torch.cuda.current_stream()
    grids = [(1, world)] + ([(2, world // 2)] if world >= 4 else [])
    rows = []
    for (pr, pc) in grids:
        for cplx in (False, True):
            nx = n // 2 + 1 if cplx else n
            gp = pdo.decomp_info(nx, n, n, pr, pc)
            dt = torch.complex128 if cplx else torch.float64
            bufs = {p: torch.zeros(tuple(reversed(getattr(gp, p + "sz"))), dtype=dt, device="cuda") for p in "xyz"}
            (bufs["x"].real if cplx else bufs["x"]).uniform_()
            for p in "xyz":
                pdo.decomp_2d.register(bufs[p])
            for name, fn, s_, d_, p in (("x_to_y", dc.transpose_x_to_y, "x", "y", pr), ("y_to_x", dc.transpose_y_to_x, "y", "x", pr),
                                        ("y_to_z", dc.transpose_y_to_z, "y", "z", pc), ("z_to_y", dc.transpose_z_to_y, "z", "y", pc)):
                if p == 1:
                    continue   # one rank in that sub-communicator: the pencils coincide, nothing crosses a link
                ms = _ev_time(lambda: fn(bufs[s_], bufs[d_], gp), st, 8, 3, world)
                L = bufs[s_].numel() * bufs[s_].element_size()
                sent = L * (p - 1) / p
                rows.append({"op": name, "complex": cplx, "grid": [pr, pc], "ms": round(ms, 4), "sent_MB_per_gpu": round(sent / 1e6, 1),
                             "link_GBps_per_gpu": round(sent / ms / 1e6, 1), "nvlink_frac": round(sent / ms / 1e6 / NVLINK_GBS, 3)})
            for p in "xyz":
                pdo.decomp_2d.deregister(bufs[p])
            del bufs
            gp.destroy()
            torch.cuda.empty_cache()
    worst = min(rows, key=lambda r: r["nvlink_frac"]) if rows else None
    return {"workload": f"pencil transposes of a {n}^3 real field / its {n//2+1}x{n}x{n} complex spectrum, {world} GPUs", "peak_GBps": NVLINK_GBS,
            "rows": rows, "min_nvlink_frac": worst["nvlink_frac"] if worst else None, "plane": os.environ.get("PDO_P2P_MODE", "auto")}


def poisson_leg(n, world, rank):
    """BASELINE config 3, second half: PoissonPeriodic%poisson_solve (dir_id = 1) on the global n^3 grid, slabs 1 x N (strong scaling):
    fft3_x2z + multiply + ifft3_z2x = 176 algorithmic bytes per real point."""
    import numpy as np
    import torch
    import padeops_b200 as pdo
    st = torch.cuda.current_stream()
    d = 2 * np.pi / n
    po = pdo.PoissonPeriodic()
    po.init(d, d, d, (n, n, n), 1, p_row=1, p_col=world)
    info = pdo.decomp_info.for_rank(n, n, n, 1, world, rank)
    rhs = torch.rand(tuple(reversed(info["xsz"])), dtype=torch.float64, device="cuda")
    out = torch.empty_like(rhs)
    ms = _ev_time(lambda: po.poisson_solve(rhs, out), st, 5, 2, world)
    po.destroy()
    peak, _ = peaks()
    gb = 176.0 * n ** 3 / world / 1e9
    return {"workload": f"PoissonPeriodic poisson_solve, {n}^3 global, grid 1x{world} (strong scaling)", "ms": round(ms, 4),
            "Gpoints_per_s": round(n ** 3 / ms / 1e6, 2), "algorithmic_GB_per_gpu": round(gb, 3),
            "hbm_frac_per_gpu": round(gb / (ms * 1e-3) / peak, 3)}


def cd10_2048_leg(world, rank):
    """BASELINE config 5: CD10 ddx / ddy / ddz + d2dx2 on a 2048^3 field (64 GiB), y-pencil, slabs 1 x N, strong scaling.  x and y are
    pencil-local; z goes through the distributed z-slab solve (pdo_operators_ddz).  Skipped when 2 fields + scratch do not fit."""
    import numpy as np
    import torch
    import padeops_b200 as pdo
    n = 2048
    st = torch.cuda.current_stream()
    need = 2 * 8 * n ** 3 / world * 1.12 + (2 << 30)
    free, _tot = torch.cuda.mem_get_info()
    ok = free > need
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    if not ok:
        return {"skipped": f"needs {need/2**30:.0f} GiB per GPU at {world} GPU(s), {free/2**30:.0f} GiB free"}
    d = 2 * np.pi / n
    gp = pdo.decomp_2d.init(n, n, n, 1, world)
    f = torch.empty(tuple(reversed(gp.ysz)), dtype=torch.float64, device="cuda")
    f.uniform_()
    df = torch.empty_like(f)
    npts = f.numel()
    der = pdo.derivatives()
    der.init(gp, d, d, d, True, True, True, "cd10", "cd10", "cd10")
    ops = None
    if world > 1:
        ops = pdo.vector_ops()
        ops.init(gp, d, d, d, "cd10", allow_zslab=True)
    peak, _ = peaks()
    out = {"workload": f"cd10 ddx / ddy / ddz / d2dx2, 2048^3 global (64 GiB per field), grid 1x{world}, {npts} points per GPU (strong scaling)",
           "z_path": "local" if ops is None else ("z-slab distributed solve" if ops.zmode == 1 else "transposes"), "ops": {}}
    calls = [("ddx", lambda: der.ddx(f, df)), ("ddy", lambda: der.ddy(f, df)),
             ("ddz", (lambda: der.ddz(f, df)) if ops is None else (lambda: ops.ddz(f, df))), ("d2dx2", lambda: der.d2dx2(f, df))]
    tot = 0.0
    for nm, fn in calls:
        ms = _ev_time(fn, st, 4, 2, world)
        tot += ms
        out["ops"][nm] = {"ms": round(ms, 4), "Gpoints_per_s": round(n ** 3 / ms / 1e6, 1),
                          "hbm_frac_per_gpu": round(BYTES_PER_POINT * npts / (ms * 1e-3) / 1e9 / peak, 3)}
    out["ms_all_four"] = round(tot, 4)
    out["Gpoints_per_s_all_four"] = round(4.0 * n ** 3 / tot / 1e6, 1)
    if ops is not None:
        ops.destroy()
    der.destroy()
    del f, df
    torch.cuda.empty_cache()
    return out


def pcie_ceiling(nbytes, world):
    """What the host gives a plain pinned full-duplex copy of the e2e leg's size (H2D and D2H at once, all ranks at once): the
    ceiling of any host-pointer path on this box."""
    import torch
    try:
        h_in = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
        h_out = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
        d_in = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
        d_out = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    except RuntimeError:
        return None
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def go():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    go()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 2
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
    return nbytes / dt / 1e9




def run_ours(args):
    import numpy as np
    import torch
    import padeops_b200 as pdo
    from padeops_b200 import decomp as dc
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = pdo.lib()
    n = args.n
    # global field: N * n^3 points; 2DECOMP grid p_row x p_col
    # slab grids 1 x N: x- and y-pencils coincide, only ddz needs transposes (2DECOMP's own auto-tuner, best_2d_grid,
    # picks the grid by timing; 1 x N is what it converges to on an all-to-all fabric)
    grids = {1: (1, 1), 2: (1, 2), 4: (1, 4), 8: (1, 8)}
    # weak scaling along z: every rank holds the same n x n x n block at every N (the box grows in z), so the per-GPU
    # kernels are identical across N and the scaling run isolates what the decomposition costs
    mult = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 1, 4), 8: (1, 1, 8)}
    assert world in grids, "bench.py supports 1, 2, 4 or 8 GPUs"
    p_row, p_col = grids[world]
    nx, ny, nz = (n * m for m in mult[world])
    dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
    gp = pdo.decomp_2d.init(nx, ny, nz, p_row, p_col)
    npts_rank = gp.ysz[0] * gp.ysz[1] * gp.ysz[2]
    der = pdo.derivatives()
    der.init(gp, dx, dy, dz, True, True, True, "cd10", "cd10", "cd10")

    def pencil(sz):
        return torch.empty(tuple(reversed(sz)), dtype=torch.float64, device="cuda")

    g = torch.Generator(device="cuda").manual_seed(20240607 + rank)
    f = torch.rand(tuple(reversed(gp.ysz)), dtype=torch.float64, device="cuda", generator=g)   # the field lives in the y-pencil
    df = torch.empty_like(f)
    tin = tout = None
    if world > 1:   # transpose destination: peer-writable, so the fused NVLink path is taken (collective, same order on all ranks)
        pdo.decomp_2d.register(df)
    st = torch.cuda.current_stream()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]

    # N > 1: the field lives in the y-pencil and the three derivatives go through the operators.F90-level entry points
    # (pdo_operators_ddx/ddy/ddz, what operators.F90:gradient does per axis).  Along z the library distributes the compact
    # solve over the z-group (halo planes + reduced-system edge pieces over NVLink) instead of transposing the field;
    # --transposes forces the reference's choreography (transpose, differentiate, transpose back) for comparison.
    ops = None
    if world > 1:
        ops = pdo.vector_ops()
        ops.init(gp, dx, dy, dz, "cd10", allow_zslab=not args.transposes)

    gx = gy = None
    if ops is not None:
        gx, gy = torch.empty_like(f), torch.empty_like(f)

    def step(i=None):
        e = ev[i] if i is not None else None
        if ops is not None:
            # one gradient call = ddx + ddy + ddz of the field into three outputs; the z exchange overlaps the x / y kernels
            ops.gradient(f, gx, gy, df)
            return
        # tests/test_derivatives_parallel.F90:94-126: transpose to the pencil where the axis is local, differentiate, transpose back
        if e: e[0].record(st)
        if p_row > 1:
            a, b = tin.view(tuple(reversed(gp.xsz))), tout.view(tuple(reversed(gp.xsz)))
            dc.transpose_y_to_x(f, a, gp)
            der.ddx(a, b)
            dc.transpose_x_to_y(b, df, gp)
        else:
            der.ddx(f, df)            # one rank in the row communicator: x- and y-pencils coincide
        if e: e[1].record(st)
        der.ddy(f, df)
        if e: e[2].record(st)
        if p_col > 1:
            a, b = tin.view(tuple(reversed(gp.zsz))), tout.view(tuple(reversed(gp.zsz)))
            dc.transpose_y_to_z(f, a, gp)
            der.ddz(a, b)
            dc.transpose_z_to_y(b, df, gp)
        else:
            der.ddz(f, df)
        if e: e[3].record(st)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    VAR = {128: "chunk_x_kernel<128 thr>", 256: "chunk_x_kernel<256 thr>", 1000: "chunk_x_tma_kernel<M=32, 256 thr>",
           1016: "chunk_x_tma_kernel<M=16, 512 thr>", 1: "chunk_strided_kernel<512>", 2: "chunk_strided_kernel<256>",
           3: "chunk_strided_cluster_kernel", 5: "chunk_strided_cpipe_kernel", 6: "chunk_strided_pipe_kernel",
           7: "chunk_strided_tma_kernel", 8: "chunk_strided_ctma_kernel<XT=64>", 9: "chunk_strided_ctma_kernel<XT=32>",
           10: "chunk_strided_ctma_kernel<XT=32, 2 CTA/SM>", 11: "chunk_strided_cpipe_kernel<tma loads>"}
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    kern_of = []   # which kernel variant the planner settled on, per axis
    for ax, fn in enumerate((der.ddx, der.ddy, der.ddz)):
        if (ax == 0 and p_row > 1) or (ax == 2 and p_col > 1):
            kern_of.append("z-slab distributed solve: halo push + edge pass + chunk_strided_ctma_kernel" if (ops is not None and ops.zmode == 1)
                           else "transposes + kernel")
            continue
        fn(f, df)
        kern_of.append(VAR.get(L.pdo_debug_last_variant(), "?"))
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = L.pdo_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_region:   # ncu --profile-from-start off: capture exactly the timed launches
        torch.cuda.cudart().cudaProfilerStart()
    t0.record(st)
    for i in range(args.steps):
        step(i)
    t1.record(st)
    barrier()
    if args.profile_region:
        torch.cuda.cudart().cudaProfilerStop()
    launches = L.pdo_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms = t0.elapsed_time(t1)
    if ops is None:
        per = [sum(e[j].elapsed_time(e[j + 1]) for e in ev) / args.steps for j in range(3)]  # ddx, ddy, ddz
    else:   # per-axis times measured one call at a time, outside the timed region (inside it they overlap)
        per = []
        for fn in (ops.ddx, ops.ddy, ops.ddz):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for _ in range(5):
                fn(f, df)
            b.record(st)
            barrier()
            per.append(a.elapsed_time(b) / 5)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_step = ms / args.steps
    pts_step_rank = 3.0 * npts_rank
    value = pts_step_rank * world / (ms_step * 1e-3) / 1e9

    # ---- e2e: the C-ABI call with HOST buffers (H2D + kernel + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        def pinned(m):   # 2 x 8 m^3 bytes of page-locked host memory per rank; None if the host refuses
            try:
                a = torch.rand((m, m, m), dtype=torch.float64).pin_memory()
                return a, torch.empty_like(a).pin_memory()
            except RuntimeError:
                return None
        ne = args.e2e_n or n
        bufs = pinned(ne)
        ok = bufs is not None
        if world > 1:   # the decision must be the same on every rank (the leg ends in a collective)
            import torch.distributed as dist
            t = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = bool(t.item())
        if not ok:
            bufs = None
            ne = min(ne, 512)
            bufs = pinned(ne)
        fh, oh = bufs
        ce = pdo.cd10()
        assert ce.init(ne, 2 * np.pi / ne) == 0
        def e2e_step():
            ce.dd1(fh, oh); ce.dd2(fh, oh); ce.dd3(fh, oh)
        e2e_step()
        barrier()
        k = max(1, min(args.steps, 3))
        tt = time.perf_counter()
        for _ in range(k):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - tt) / k
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = t.item()
        e2e = {"value": 3.0 * ne ** 3 * world / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 3 * 8 * ne ** 3,
               "d2h_bytes_per_step": 3 * 8 * ne ** 3, "n": ne, "ms_per_step": dt * 1e3,
               "pcie_GBps_per_gpu_each_way": 3 * 8 * ne ** 3 / dt / 1e9,
               "note": "pdo_cd10_dd1/dd2/dd3 called with pinned HOST pointers; the library stages H2D/D2H (chunked, full duplex)"}
        del fh, oh, bufs
        # the box's own ceiling for this leg: a plain pinned full-duplex copy of one field each way, all ranks at once
        try:
            ceil = pcie_ceiling(8 * ne ** 3, world)
            if ceil:
                e2e["pcie_ceiling_GBps_per_gpu_each_way"] = ceil
                e2e["frac_of_pcie_ceiling"] = e2e["pcie_GBps_per_gpu_each_way"] / ceil
        except Exception as ex:  # noqa
            e2e["pcie_ceiling_error"] = str(ex)[:120]
        torch.cuda.empty_cache()

    # ---- igrid RK substep (metric iii), guarded: a failure or a stall here must not cost the headline line ----
    sub = None
    sub_holder = {}
    if not args.no_substep:
        torch.cuda.synchronize()
        if world > 1:
            pdo.decomp_2d.deregister(df)
        if ops is not None:
            ops.destroy()
        del f, df, tin, tout, gx, gy
        torch.cuda.empty_cache()

        def _guard(key, fn):
            try:
                sub_holder[key] = fn()
            except Exception as ex:  # noqa
                sub_holder[key] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
            torch.cuda.empty_cache()

        def _run_sub():
            torch.cuda.set_device(local)   # the current device is per thread
            _guard("v", lambda: substep_leg(args.substep_n, world, rank))
            if not args.no_legs:
                # BASELINE configs 3 and 5 (collective legs: every rank runs them in the same order)
                if world > 1:
                    _guard("transposes", lambda: transposes_leg(args.n, world, rank))
                _guard("poisson", lambda: poisson_leg(args.n, world, rank))
                _guard("cd10_2048", lambda: cd10_2048_leg(world, rank))
            sub_holder["done"] = True
        th = threading.Thread(target=_run_sub, daemon=True)
        th.start()
        th.join(timeout=float(args.substep_timeout + (0 if args.no_legs else args.legs_timeout)))
        sub = sub_holder.get("v", {"error": f"substep leg did not finish within {args.substep_timeout} s"})
    if rank != 0:
        if not args.no_substep and "done" not in sub_holder:
            os._exit(0)
        return
    peak, peak_src = peaks()
    names = [f"cd10 dd{a} ({k})" for a, k in zip("xyz", kern_of)]
    per_k = {}
    for nm, t in zip(names, per):
        per_k[nm] = {"ms": t, "GBps": BYTES_PER_POINT * npts_rank / (t * 1e-3) / 1e9}
    # the genuinely slowest leg, also when it is the distributed z solve (halo exchange + edge pass + kernel): its 16 B/pt are
    # still the algorithmic bytes, so the fraction says what the decomposition costs
    dom = max(range(3), key=lambda j: per[j])
    ach = BYTES_PER_POINT * npts_rank / (per[dom] * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_POINT * npts_rank, "per_kernel": per_k}
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (tools/ncu_summary.py)
    import glob
    base = names[dom].split("(")[1].rstrip(")").split("<")[0]
    for tr in sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json"))):
        try:
            v = json.load(open(tr)).get(base)
            if v is not None:
                roof["traffic"] = v
                roof["traffic_source"] = os.path.relpath(tr, ROOT)
        except Exception:
            pass
    cb = None
    if not args.no_cpu:
        cb, _, _ = cpu_arm(n, 2, 1)      # bounded sample: the full field, 1 warm-up + 2 timed passes (~10 s on 16 cores)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"cd10 ddx+ddy+ddz, periodic, {nx}x{ny}x{nz} field, 2DECOMP grid {p_row}x{p_col}, "
                                   f"{npts_rank} points per GPU" + ("" if world == 1 else (
                                       "; y-pencil field through pdo_operators_gradient (operators.F90:gradient), z via the distributed z-slab solve (no transposes), its exchange overlapped with the x / y kernels"
                                       if ops.zmode == 1 else "; y-pencil field through pdo_operators_gradient, off-pencil axes by transposes")),
                       "n": n, "global": [nx, ny, nz], "grid": [p_row, p_col],
                       "l2": "inputs (8 GiB per field at n=1024) larger than the 126 MB L2, no flush needed"},
            "roofline": roof, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "substep": sub}
    if isinstance(sub, dict) and "ms_per_substep" in sub:   # metric (iii) as flat top-level keys too
        line["substep_ms"] = sub["ms_per_substep"]
        line["substep_launches"] = sub.get("launches_per_substep")
    for key in ("transposes", "poisson", "cd10_2048"):      # BASELINE configs 3 and 5
        if key in sub_holder:
            line[key] = sub_holder[key]
    if isinstance(line.get("transposes"), dict) and line["transposes"].get("min_nvlink_frac") is not None:
        line["transposes_min_nvlink_frac"] = line["transposes"]["min_nvlink_frac"]
    if isinstance(line.get("poisson"), dict) and "ms" in line["poisson"]:
        line["poisson_ms"] = line["poisson"]["ms"]
    if isinstance(line.get("cd10_2048"), dict) and "ms_all_four" in line["cd10_2048"]:
        line["cd10_2048_ms"] = line["cd10_2048"]["ms_all_four"]
    print(json.dumps(line), flush=True)
    if not args.no_substep and "done" not in sub_holder:
        os._exit(0)   # a guarded leg is still stuck in a collective: leave without joining it


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1024, help="points per direction per GPU")
    ap.add_argument("--e2e-n", type=int, default=0, dest="e2e_n", help="field size of the host-pointer (e2e) leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--transposes", action="store_true", help="N > 1: force the reference's transpose choreography along z")
    ap.add_argument("--no-substep", action="store_true", dest="no_substep")
    ap.add_argument("--substep-n", type=int, default=512, dest="substep_n", help="global grid of the igrid substep leg")
    ap.add_argument("--substep-timeout", type=int, default=150, dest="substep_timeout")
    ap.add_argument("--no-legs", action="store_true", dest="no_legs", help="skip the config-3 / config-5 legs (transposes, Poisson, 2048^3)")
    ap.add_argument("--legs-timeout", type=int, default=240, dest="legs_timeout")
    ap.add_argument("--profile-region", action="store_true", dest="profile_region",
                    help="cudaProfilerStart/Stop around the timed region (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
