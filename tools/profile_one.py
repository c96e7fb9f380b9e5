#!/usr/bin/env python
"""Run one operator a few times on a modest field so ncu can capture it.
Usage: python tools/profile_one.py <op> <axis> [n] [na] [nb] [reps]   (op: cd10|cd10d2|cd06|cf90|gaussian)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import padeops_b200 as pdo

op, axis = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
na = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
nb = int(sys.argv[5]) if len(sys.argv) > 5 else 128
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 4
shape = {0: (nb, na, n), 1: (nb, n, na), 2: (n, nb, na)}[axis]
f = torch.rand(shape, dtype=torch.float64, device="cuda")
o = torch.empty_like(f)
d = 2 * np.pi / n
if op in ("cd10", "cd10d2"):
    h = pdo.cd10(); assert h.init(n, d) == 0
    fn = ((h.dd1, h.dd2, h.dd3) if op == "cd10" else (h.d2d1, h.d2d2, h.d2d3))[axis]
elif op == "cd06":
    h = pdo.cd06(); assert h.init(n, d) == 0
    fn = (h.dd1, h.dd2, h.dd3)[axis]
elif op == "cf90":
    h = pdo.cf90(); assert h.init(n) == 0
    fn = (h.filter1, h.filter2, h.filter3)[axis]
else:
    h = pdo.gaussian(); assert h.init(n) == 0
    fn = (h.filter1, h.filter2, h.filter3)[axis]
for _ in range(reps):
    fn(f, o)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); fn(f, o); b.record(); torch.cuda.synchronize()
pts = n * na * nb
print(f"{op} axis {axis} shape {shape}: {a.elapsed_time(b):.4f} ms, {16*pts/a.elapsed_time(b)/1e6:.1f} GB/s")
