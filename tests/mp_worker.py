"""Worker for the multi-GPU tests (launched by torchrun; one rank per GPU).  Checks, per rank, bit-exact
transposes (real + complex, even + uneven sizes) against the oracle's simulated-rank ALLTOALLV, the
distributed CD10 derivative choreography, and the pencil-decomposed FFT / Poisson solve."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import padeops_b200 as pdo
    from padeops_b200 import decomp as dc
    from oracle import oracle as O
    pdo.decomp_2d.comm_init()
    grids = [(1, world), (world, 1)] + ([(2, world // 2)] if world >= 4 else [])
    nfail = 0
    if os.environ.get("PDO_MP_LATE") == "1":
        return late(pdo, O, rank, world, grids)

    def ok(cond, what):
        nonlocal nfail
        if not cond:
            nfail += 1
            print(f"[rank {rank}] FAIL {what}", flush=True)

    for (pr, pc) in grids:
        for (nx, ny, nz) in [(16, 12, 8), (17, 9, 11), (33, 16, 10)]:
            if min(nx, ny) < pr or min(ny, nz) < pc:
                continue
            gp = pdo.decomp_info(nx, ny, nz, pr, pc)
            for cplx in (False, True):
                rng = np.random.default_rng(5)
                G = rng.standard_normal((nz, ny, nx))
                if cplx:
                    G = G + 1j * rng.standard_normal((nz, ny, nx))
                pens = {p: O.scatter_global(G, nx, ny, nz, pr, pc, p) for p in "xyz"}
                for d, (s, t, fn) in enumerate((("x", "y", dc.transpose_x_to_y), ("y", "x", dc.transpose_y_to_x),
                                                ("y", "z", dc.transpose_y_to_z), ("z", "y", dc.transpose_z_to_y))):
                    ref = O.transpose(d, nx, ny, nz, pr, pc, pens[s])[rank]
                    got = fn(torch.from_numpy(pens[s][rank]).cuda(), None, gp).cpu().numpy()
                    ok(got.shape == ref.shape and np.array_equal(got, ref), f"transpose {s}->{t} grid {pr}x{pc} {nx}x{ny}x{nz} cplx={cplx}")
                    ok(np.array_equal(ref, pens[t][rank]), "oracle self-consistency")
    # the same transposes into REGISTERED destinations: fused pack + NVLink store + unpack path (decomp.cu: box_push_kernel)
    p2p = bool(pdo.lib().pdo_comm_p2p_enabled())
    print(f"[rank {rank}] p2p enabled: {p2p}", flush=True)
    for (pr, pc) in grids:
        for (nx, ny, nz) in [(16, 12, 8), (17, 9, 11), (33, 16, 10), (128, 96, 80)]:
            if min(nx, ny) < pr or min(ny, nz) < pc:
                continue
            gp = pdo.decomp_info(nx, ny, nz, pr, pc)
            for cplx in (False, True):
                rng = np.random.default_rng(7)
                G = rng.standard_normal((nz, ny, nx))
                if cplx:
                    G = G + 1j * rng.standard_normal((nz, ny, nx))
                pens = {p: O.scatter_global(G, nx, ny, nz, pr, pc, p) for p in "xyz"}
                dt = torch.complex128 if cplx else torch.float64
                dsts = {p: pdo.decomp_2d.register(torch.zeros(tuple(reversed(getattr(gp, p + "sz"))), dtype=dt, device="cuda")) for p in "xyz"}
                for rep in range(2):  # twice: epochs / flag reuse
                    for s_, t_, fn in (("x", "y", dc.transpose_x_to_y), ("y", "x", dc.transpose_y_to_x), ("y", "z", dc.transpose_y_to_z),
                                       ("z", "y", dc.transpose_z_to_y)):
                        dsts[t_].zero_()
                        fn(torch.from_numpy(pens[s_][rank]).cuda(), dsts[t_], gp)
                        ok(np.array_equal(dsts[t_].cpu().numpy(), pens[t_][rank]), f"p2p transpose {s_}->{t_} grid {pr}x{pc} {nx}x{ny}x{nz} cplx={cplx}")
                for t_ in dsts.values():
                    pdo.decomp_2d.deregister(t_)  # before the tensors are freed
            gp.destroy()
    # distributed derivative choreography (tests/test_derivatives_parallel.F90:94-126) on 64^3
    n = 64
    d = 2 * np.pi / n
    pr, pc = (1, world)
    gp = pdo.decomp_info(n, n, n, pr, pc)
    der = pdo.derivatives()
    der.init(gp, d, d, d, True, True, True, "cd10", "cd10", "cd10")
    x = np.arange(n) * d
    G = np.sin(x)[None, None, :] * np.sin(x)[None, :, None] * np.cos(x)[:, None, None] + 0.1 * np.random.default_rng(1).standard_normal((n, n, n))
    fy = torch.from_numpy(O.scatter_global(G, n, n, n, pr, pc, "y")[rank]).cuda()
    for ax, (to, back, pen) in enumerate(((dc.transpose_y_to_x, dc.transpose_x_to_y, "x"), (None, None, "y"), (dc.transpose_y_to_z, dc.transpose_z_to_y, "z"))):
        ref = O.scatter_global(O.cd10(G, d, ax, 1), n, n, n, pr, pc, "y")[rank]
        if to is None:
            got = der.ddy(fy)
        else:
            a = to(fy, None, gp)
            b = (der.ddx, None, der.ddz)[ax](a)
            got = back(b, None, gp)
        err = np.abs(got.cpu().numpy() - ref).max() / np.abs(ref).max()
        ok(err < 1e-12, f"distributed cd10 axis {ax}: rel err {err:.2e}")
    # operators.F90 drop-ins on decomposed fields; on 1 x world slabs of 256 planes per GPU
    # the z-derivative takes the distributed z-slab solve, on world x 1 it is local, and allow_zslab=False is the
    # reference's transpose choreography
    for (pr, pc) in grids:
        for method in ("cd10", "cd06"):
            nx, ny = 32, 32
            nz = 256 * pc if pc > 1 else 64         # 8 chunks per slab (CD10 needs 2 x 3 edge chunks and a power-of-two split)
            if nx < pr or ny < pr:
                continue
            dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
            rng = np.random.default_rng(21)
            U, V, W = (rng.standard_normal((nz, ny, nx)) for _ in range(3))
            gp = pdo.decomp_info(nx, ny, nz, pr, pc)
            loc = [torch.from_numpy(O.scatter_global(A, nx, ny, nz, pr, pc, "y")[rank]).cuda() for A in (U, V, W)]
            refs = {"grad": O.gradient(U, dx, dy, dz, method), "div": O.divergence(U, V, W, dx, dy, dz, method),
                    "curl": O.curl(U, V, W, dx, dy, dz, method)}
            sl = lambda A: O.scatter_global(A, nx, ny, nz, pr, pc, "y")[rank]
            for allow in (True, False):
                ops = pdo.vector_ops()
                ops.init(gp, dx, dy, dz, method, allow_zslab=allow)
                want_mode = 0 if pc == 1 else (1 if (allow and p2p) else 2)
                ok(ops.zmode == want_mode, f"vector_ops zmode {ops.zmode} != {want_mode} grid {pr}x{pc} {method} allow={allow}")
                for rep in range(3):   # parity buffers and epochs of the z-slab exchange
                    g = ops.gradient(loc[0])
                    for c in range(3):
                        e = np.abs(g[c].cpu().numpy() - sl(refs["grad"][c])).max() / np.abs(refs["grad"][c]).max()
                        ok(e < 1e-12, f"gradient[{c}] grid {pr}x{pc} {method} zmode {ops.zmode} rep {rep}: {e:.2e}")
                dv = ops.divergence(*loc).cpu().numpy()
                ok(np.abs(dv - sl(refs["div"])).max() < 1e-12 * np.abs(refs["div"]).max(), f"divergence grid {pr}x{pc} {method} zmode {ops.zmode}")
                cu = ops.curl(*loc).cpu().numpy()
                for c in range(3):
                    ok(np.abs(cu[c] - sl(refs["curl"][c])).max() < 1e-12 * np.abs(refs["curl"]).max(), f"curl[{c}] grid {pr}x{pc} {method} zmode {ops.zmode}")
                ops.destroy()
            gp.destroy()
    # pencil-decomposed FFT and Poisson
    for (pr, pc) in grids:
        for (nx, ny, nz) in [(32, 16, 24), (18, 12, 10)]:
            if min(nx // 2 + 1, ny) < pr or min(ny, nz) < pc:
                continue
            dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
            rng = np.random.default_rng(9)
            G = rng.standard_normal((nz, ny, nx))
            ft = pdo.fft_3d()
            rc = ft.init(nx, ny, nz, "x", dx, dy, dz, p_row=pr, p_col=pc)
            ok(rc == 0, "fft init")
            fx = torch.from_numpy(O.scatter_global(G, nx, ny, nz, pr, pc, "x")[rank]).cuda()
            H = np.fft.fft(np.fft.fft(np.fft.rfft(G, axis=2), axis=1), axis=0)
            refz = O.scatter_global(H, nx // 2 + 1, ny, nz, pr, pc, "z")[rank]
            got = ft.fft3_x2z(fx)
            ok(np.abs(got.cpu().numpy() - refz).max() < 1e-12 * np.abs(H).max(), f"fft3_x2z grid {pr}x{pc} {nx}x{ny}x{nz}")
            back = ft.ifft3_z2x(got)
            ok(np.abs(back.cpu().numpy() - fx.cpu().numpy()).max() < 1e-12 * np.abs(G).max(), f"ifft3_z2x grid {pr}x{pc}")
            H2 = np.fft.fft(np.fft.rfft(G, axis=2), axis=1)
            refy = O.scatter_global(H2, nx // 2 + 1, ny, nz, pr, pc, "y")[rank]
            got2 = ft.fft2_x2y(fx)
            ok(np.abs(got2.cpu().numpy() - refy).max() < 1e-12 * np.abs(H2).max(), f"fft2_x2y grid {pr}x{pc}")
            ok(np.abs(ft.ifft2_y2x(got2).cpu().numpy() - fx.cpu().numpy()).max() < 1e-12 * np.abs(G).max(), f"ifft2_y2x grid {pr}x{pc}")
            ref = O.poisson_solve(G, dx, dy, dz)
            for dir_id, pen in ((1, "x"), (2, "y")):
                po = pdo.PoissonPeriodic()
                po.init(dx, dy, dz, (nx, ny, nz), dir_id, p_row=pr, p_col=pc)
                rin = torch.from_numpy(O.scatter_global(G, nx, ny, nz, pr, pc, pen)[rank]).cuda()
                out = torch.empty_like(rin)
                po.poisson_solve(rin, out)
                rr = O.scatter_global(ref, nx, ny, nz, pr, pc, pen)[rank]
                ok(np.abs(out.cpu().numpy() - rr).max() < 1e-12 * np.abs(ref).max(), f"poisson dir {dir_id} grid {pr}x{pc} {nx}x{ny}x{nz}")
    # igrid periodic substep on decomposed fields: one TVD-RK3 step against the single-rank oracle
    from oracle import igrid_oracle as IG
    nx, ny, nz = 16, 16, 16
    rng = np.random.default_rng(11)
    U, V = rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz, ny, nx))
    W = rng.standard_normal((nz + 1, ny, nx))
    W[nz] = W[0]
    Lbox = (2 * np.pi,) * 3
    ref = IG.IGrid(nx, ny, nz, *Lbox, 80.0, U, V, W, TimeSteppingScheme=1)
    ref.timeAdvance(0.01)
    for (pr, pc) in grids:
        if min(nx // 2 + 1, ny) < pr or min(ny, nz) < pc:
            continue
        loc = [O.scatter_global(A, nx, ny, n3, pr, pc, "x")[rank] for A, n3 in ((U, nz), (V, nz), (W, nz + 1))]
        g = pdo.igrid()
        g.init(nx, ny, nz, *Lbox, 80.0, *loc, TimeSteppingScheme=1, prow=pr, pcol=pc)
        g.timeAdvance(0.01)
        for nm, n3 in (("u", nz), ("v", nz), ("w", nz + 1), ("wC", nz)):
            rr = O.scatter_global(getattr(ref, nm), nx, ny, n3, pr, pc, "x")[rank]
            got = g.get(nm)
            ok(got.shape == rr.shape and np.abs(got - rr).max() < 5e-12 * np.abs(getattr(ref, nm)).max(), f"igrid {nm} grid {pr}x{pc}")
        ok(g.maxDivergence() < 1e-11, f"igrid divergence grid {pr}x{pc}")
        g.destroy()
    # reductions (utilities/reductions.F90)
    ok(pdo.decomp_2d.p_maxval(float(rank)) == float(world - 1), "p_maxval")
    ok(pdo.decomp_2d.p_sum(1.0) == float(world), "p_sum")
    t = torch.tensor([nfail], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("MP_WORKER_RESULT", "PASS" if t.item() == 0 else f"FAIL({t.item()})", flush=True)
    dist.barrier()
    pdo.decomp_2d.finalize()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


def late(pdo, O, rank, world, grids):
    """Sections written after the round's last multi-GPU session (own result line, own non-strict test): Poisson with z-pencil
    input, filter3D on decomposed fields, the igrid substep's rotational / Fourier-z variants."""
    nfail = 0

    def ok(cond, what):
        nonlocal nfail
        if not cond:
            nfail += 1
            print(f"[rank {rank}] FAIL {what}", flush=True)

    for (pr, pc) in grids:
        nx, ny, nz = 32, 16, 24
        if min(nx // 2 + 1, ny) < pr or min(ny, nz) < pc:
            continue
        dx, dy, dz = 2 * np.pi / nx, 2 * np.pi / ny, 2 * np.pi / nz
        G = np.random.default_rng(9).standard_normal((nz, ny, nx))
        ref = O.poisson_solve(G, dx, dy, dz)
        po = pdo.PoissonPeriodic()
        po.init(dx, dy, dz, (nx, ny, nz), 3, p_row=pr, p_col=pc)
        rin = torch.from_numpy(O.scatter_global(G, nx, ny, nz, pr, pc, "z")[rank]).cuda()
        out = torch.empty_like(rin)
        po.poisson_solve(rin, out)
        rr = O.scatter_global(ref, nx, ny, nz, pr, pc, "z")[rank]
        ok(np.abs(out.cpu().numpy() - rr).max() < 1e-12 * np.abs(ref).max(), f"poisson dir 3 grid {pr}x{pc}")
    # filter3D (operators.F90:158-224) on decomposed y-pencil fields, 1 / 2 passes, CF90 and mixed methods
    for (pr, pc) in grids:
        nx, ny, nz = 48, 32, 40
        if min(nx, ny) < pr or min(ny, nz) < pc:
            continue
        d = 2 * np.pi / nx
        F = np.random.default_rng(31).standard_normal((nz, ny, nx))
        gp = pdo.decomp_info(nx, ny, nz, pr, pc)
        ops = pdo.vector_ops()
        ops.init(gp, d, d, d, "cd10")
        for methods in (("cf90", "cf90", "cf90"), ("gaussian", "cf90", "gaussian")):
            fil = pdo.filters()
            fil.init(gp, True, True, True, *methods)
            for numtimes in (1, 2):
                a = torch.from_numpy(O.scatter_global(F, nx, ny, nz, pr, pc, "y")[rank]).cuda()
                ops.filter3D(fil, a, numtimes)
                ref = O.filter3D(F, numtimes, methods)
                rr = O.scatter_global(ref, nx, ny, nz, pr, pc, "y")[rank]
                ok(np.abs(a.cpu().numpy() - rr).max() < 1e-12 * np.abs(ref).max(), f"filter3D grid {pr}x{pc} {methods} x{numtimes}")
        ops.destroy()
        gp.destroy()
    # igrid substep variants on decomposed fields
    from oracle import igrid_oracle as IG
    nx, ny, nz = 16, 16, 16
    rng = np.random.default_rng(11)
    U, V = rng.standard_normal((nz, ny, nx)), rng.standard_normal((nz, ny, nx))
    W = rng.standard_normal((nz + 1, ny, nx))
    W[nz] = W[0]
    Lbox = (2 * np.pi,) * 3
    for adv, vert in ((0, 1), (1, 2), (0, 2)):
        ref = IG.IGrid(nx, ny, nz, *Lbox, 80.0, U, V, W, TimeSteppingScheme=1, AdvectionTerm=adv, NumericalSchemeVert=vert)
        ref.timeAdvance(0.01)
        for (pr, pc) in grids:
            if min(nx // 2 + 1, ny) < pr or min(ny, nz) < pc:
                continue
            loc = [O.scatter_global(A, nx, ny, n3, pr, pc, "x")[rank] for A, n3 in ((U, nz), (V, nz), (W, nz + 1))]
            g = pdo.igrid()
            g.init(nx, ny, nz, *Lbox, 80.0, *loc, TimeSteppingScheme=1, prow=pr, pcol=pc, AdvectionTerm=adv, NumericalSchemeVert=vert)
            g.timeAdvance(0.01)
            for nm, n3 in (("u", nz), ("v", nz), ("w", nz + 1)):
                rr = O.scatter_global(getattr(ref, nm), nx, ny, n3, pr, pc, "x")[rank]
                got = g.get(nm)
                ok(got.shape == rr.shape and np.abs(got - rr).max() < 5e-12 * np.abs(getattr(ref, nm)).max(),
                   f"igrid adv={adv} vert={vert} {nm} grid {pr}x{pc}")
            g.destroy()
    # wall-bounded igrid (slip walls) on decomposed fields
    nx, ny, nz, Lz = 16, 16, 16, 2.0
    xx = np.arange(nx) * 2 * np.pi / nx
    zc, ze = (np.arange(nz) + 0.5) * Lz / nz, np.arange(nz + 1) * Lz / nz
    X, Y = xx[None, None, :], xx[None, :, None]
    U = np.sin(X) * np.cos(Y) * np.cos(np.pi * zc / Lz)[:, None, None]
    V = -np.cos(X) * np.sin(Y) * np.cos(np.pi * zc / Lz)[:, None, None]
    W = 0.25 * np.sin(X) * np.sin(2 * Y) * np.sin(2 * np.pi * ze / Lz)[:, None, None]
    box = (2 * np.pi, 2 * np.pi, Lz)
    ref = IG.IGrid(nx, ny, nz, *box, 100.0, U, V, W, TimeSteppingScheme=1, PeriodicInZ=False, topWall=2, botWall=2)
    ref.timeAdvance(0.005)
    for (pr, pc) in grids:
        if min(nx // 2 + 1, ny) < pr or min(ny, nz) < pc:
            continue
        loc = [O.scatter_global(A, nx, ny, n3, pr, pc, "x")[rank] for A, n3 in ((U, nz), (V, nz), (W, nz + 1))]
        g = pdo.igrid()
        g.init(nx, ny, nz, *box, 100.0, *loc, TimeSteppingScheme=1, prow=pr, pcol=pc, PeriodicInZ=False, topWall=2, botWall=2)
        g.timeAdvance(0.005)
        for nm, n3 in (("u", nz), ("v", nz), ("w", nz + 1)):
            rr = O.scatter_global(getattr(ref, nm), nx, ny, n3, pr, pc, "x")[rank]
            ok(np.abs(g.get(nm) - rr).max() < 5e-12 * np.abs(ref.u).max(), f"wall-bounded igrid {nm} grid {pr}x{pc}")
        g.destroy()
    t = torch.tensor([nfail], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("MP_WORKER_LATE", "PASS" if t.item() == 0 else f"FAIL({t.item()})", flush=True)
    dist.barrier()
    pdo.decomp_2d.finalize()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
