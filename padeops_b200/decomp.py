"""Mirror of the 2DECOMP&FFT module interface used on the hot path (2D» decomp_2d.f90, transpose_*.f90).

decomp_2d.init(nx,ny,nz,p_row,p_col) ~ decomp_2d_init; decomp_info ~ TYPE(DECOMP_INFO) from
decomp_info_init; transpose_x_to_y(src, dst, decomp) etc. are generic over real / complex like the
Fortran interfaces.  One process per GPU; the NCCL unique id is broadcast with torch.distributed
(the Fortran host would use MPI_Bcast) — plumbing only, no data-path collective goes through torch.
"""
import ctypes as C

from ._lib import DecompInfo, check, lib, ptr, stream_ptr


class decomp_info:
    """TYPE(DECOMP_INFO): xst/xen/xsz, yst/yen/ysz, zst/zen/zsz (1-based starts, like the Fortran)."""

    def __init__(self, nx, ny, nz, p_row=0, p_col=0):
        self._h = C.c_void_p(None)
        check(lib().pdo_decomp_init(C.byref(self._h), int(nx), int(ny), int(nz), int(p_row), int(p_col)))
        info = DecompInfo()
        check(lib().pdo_decomp_get_info(self._h, C.byref(info)))
        for nm, _ in DecompInfo._fields_:
            setattr(self, nm, tuple(getattr(info, nm)))
        self.nx, self.ny, self.nz = nx, ny, nz

    def destroy(self):
        if self._h:
            lib().pdo_decomp_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    @staticmethod
    def for_rank(nx, ny, nz, p_row, p_col, rank):
        """Pure arithmetic: what `rank` of a p_row x p_col grid owns (no communicator, no GPU)."""
        info = DecompInfo()
        check(lib().pdo_decomp_info_for(int(nx), int(ny), int(nz), int(p_row), int(p_col), int(rank), C.byref(info)))
        return {nm: tuple(getattr(info, nm)) for nm, _ in DecompInfo._fields_}


class decomp_2d:
    """Module-level state of decomp_2d: nrank, nproc and the main decomposition."""
    nrank = 0
    nproc = 1
    main = None
    _inited = False

    @classmethod
    def comm_init(cls):
        """Replaces MPI_Init + communicator creation.  Reads RANK / WORLD_SIZE; for WORLD_SIZE > 1 uses
        torch.distributed (already initialised by the caller) only to broadcast the 128-byte NCCL id."""
        import os
        if cls._inited:
            return
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        uid = C.create_string_buffer(128)
        if world > 1:
            import torch
            import torch.distributed as dist
            assert dist.is_initialized(), "initialise torch.distributed first (it only carries the NCCL id)"
            if rank == 0:
                check(lib().pdo_comm_unique_id(uid))
            t = torch.tensor(list(uid.raw), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=0)
            uid = C.create_string_buffer(bytes(t.cpu().tolist()), 128)
        check(lib().pdo_comm_init(rank, world, uid))
        cls.nrank, cls.nproc, cls._inited = rank, world, True

    @classmethod
    def init(cls, nx, ny, nz, p_row=0, p_col=0):
        """decomp_2d_init(nx,ny,nz,p_row,p_col) (2D» decomp_2d.f90:297-455)."""
        cls.comm_init()
        cls.main = decomp_info(nx, ny, nz, p_row, p_col)
        return cls.main

    @classmethod
    def finalize(cls):
        lib().pdo_comm_finalize()
        cls._inited = False
        cls.main = None

    @staticmethod
    def register(t):
        """Collective: make tensor `t` (device) a legal destination of the fused NVLink transposes."""
        check(lib().pdo_comm_register_buffer(C.c_void_p(t.data_ptr()), t.numel() * t.element_size()))
        return t

    @staticmethod
    def deregister(t):
        """Local: must be called before a registered tensor is freed."""
        check(lib().pdo_comm_deregister_buffer(C.c_void_p(t.data_ptr())))

    @staticmethod
    def p_maxval(x):
        out = C.c_double(0.0)
        check(lib().pdo_p_maxval(float(x), C.byref(out)))
        return out.value

    @staticmethod
    def p_sum(x):
        out = C.c_double(0.0)
        check(lib().pdo_p_sum(float(x), C.byref(out)))
        return out.value


def _transpose(name, src, dst, decomp, src_pen, dst_pen, stream):
    decomp = decomp or decomp_2d.main
    w = 2 if "complex" in str(src.dtype) else 1
    assert tuple(src.shape) == tuple(reversed(getattr(decomp, src_pen + "sz"))), (tuple(src.shape), getattr(decomp, src_pen + "sz"))
    if dst is None:
        shape = tuple(reversed(getattr(decomp, dst_pen + "sz")))
        dst = src.new_empty(shape) if hasattr(src, "new_empty") else __import__("numpy").empty(shape, dtype=src.dtype)
    check(getattr(lib(), f"pdo_transpose_{name}")(decomp._h, ptr(src), ptr(dst), w, stream_ptr(stream)))
    return dst


def transpose_x_to_y(src, dst=None, decomp=None, stream=None):
    return _transpose("x_to_y", src, dst, decomp, "x", "y", stream)


def transpose_y_to_x(src, dst=None, decomp=None, stream=None):
    return _transpose("y_to_x", src, dst, decomp, "y", "x", stream)


def transpose_y_to_z(src, dst=None, decomp=None, stream=None):
    return _transpose("y_to_z", src, dst, decomp, "y", "z", stream)


def transpose_z_to_y(src, dst=None, decomp=None, stream=None):
    return _transpose("z_to_y", src, dst, decomp, "z", "y", stream)


# ---- decomp_2d_io (2D» io_write_one.f90, io_read_one.f90): one distributed array <-> the flat global Fortran-order file ----
def decomp_2d_write_one(ipencil, var, filename, opt_decomp=None):
    decomp = opt_decomp or decomp_2d.main
    pen = "xyz"[int(ipencil) - 1]
    assert tuple(var.shape) == tuple(reversed(getattr(decomp, pen + "sz"))), (tuple(var.shape), getattr(decomp, pen + "sz"))
    w = 2 if "complex" in str(var.dtype) else 1
    check(lib().pdo_decomp_write_one(decomp._h, int(ipencil), ptr(var), w, str(filename).encode()))


def decomp_2d_read_one(ipencil, var, filename, opt_decomp=None):
    decomp = opt_decomp or decomp_2d.main
    pen = "xyz"[int(ipencil) - 1]
    assert tuple(var.shape) == tuple(reversed(getattr(decomp, pen + "sz"))), (tuple(var.shape), getattr(decomp, pen + "sz"))
    w = 2 if "complex" in str(var.dtype) else 1
    check(lib().pdo_decomp_read_one(decomp._h, int(ipencil), ptr(var), w, str(filename).encode()))
    return var
