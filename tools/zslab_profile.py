#!/usr/bin/env python
"""The z-slab distributed solve's kernels (halo push, edge pass, fused solve) on ONE GPU through the emulation hook, at the
bench shape (1024 planes per slab, 1024 x 1024 columns), for an ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/zslab_profile.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import padeops_b200 as pdo

nslabs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n1 = int(sys.argv[3]) if len(sys.argv) > 3 else 1024 * 1024
n = nslabs * nl
h = pdo.cd10()
assert h.init(n, 2 * np.pi / n) == 0
f = torch.rand((n, 1, n1), dtype=torch.float64, device="cuda")
o = torch.empty_like(f)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for rep in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rc = pdo.lib().pdo_debug_zslab_emulate(h._h, 0, C.c_void_p(f.data_ptr()), C.c_void_p(o.data_ptr()), n1, n, nslabs, st)
    b.record(); torch.cuda.synchronize()
    assert rc == 0
    print(f"emulated {nslabs} slabs x {nl} planes x {n1} columns: {a.elapsed_time(b):.3f} ms (includes buffer alloc/free)")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
h.dd3(f, o); torch.cuda.synchronize()
a.record(); h.dd3(f, o); b.record(); torch.cuda.synchronize()
print(f"whole-line dd3 on the same field: {a.elapsed_time(b):.3f} ms")
