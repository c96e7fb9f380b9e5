"""padeops_b200 — B200-native (sm_100a CUDA + NCCL) implementation of the PadeOps operator hot path.

The product is `lib/libpadeops_b200.so` (C ABI in include/padeops_b200.h).  This package is the
Python-side mirror of the reference's Fortran module interfaces on top of that ABI, used by the
tests and the bench harness; PyTorch only supplies device memory, streams and the rendezvous.
There is no CPU fallback: importing works anywhere, computing requires the CUDA library and a GPU.
"""
from ._lib import PadeOpsError, build_library, lib, library_path  # noqa: F401
from .operators import (cd06, cd06stagg, cd10, cf90, derivatives, filters, gaussian, lstsq)  # noqa: F401
from .decomp import decomp_2d, decomp_2d_read_one, decomp_2d_write_one, decomp_info  # noqa: F401
from .spectral import PoissonPeriodic, fft_3d  # noqa: F401
from .igrid import HIT_shell_forcing, Ops_Periodic, Pade6stagg, igrid, padepoisson, spectral  # noqa: F401
from .vecops import vector_ops  # noqa: F401
