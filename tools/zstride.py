#!/usr/bin/env python
"""ddz (solve axis outermost) throughput vs row stride n1 = na*nb at fixed line length: separates TLB reach effects
from kernel structure.  Usage: python tools/zstride.py [n] [op]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import padeops_b200 as pdo

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
opn = sys.argv[2] if len(sys.argv) > 2 else "cd10"
h = {"cd10": pdo.cd10, "cd06": pdo.cd06}.get(opn, pdo.cd10)()
assert h.init(n, 2 * np.pi / n) == 0
for lg in range(12, 21):
    n1 = 1 << lg
    f = torch.rand((n, 1, n1), dtype=torch.float64, device="cuda")
    o = torch.empty_like(f)
    for _ in range(3):
        h.dd3(f, o)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, min(50, (1 << 22) // n1))
    a.record()
    for _ in range(reps):
        h.dd3(f, o)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(json.dumps({"op": opn + ".dd3", "n": n, "n1": n1, "stride_bytes": 8 * n1, "MB": 8 * n * n1 / 2**20, "ms": round(ms, 4),
                      "GBps": round(16 * n * n1 / ms / 1e6, 1)}), flush=True)
    del f, o
