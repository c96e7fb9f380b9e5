#!/usr/bin/env python
"""Non-periodic closures: chunked fast path (np_chunk.cu) vs the one-thread-per-line sweeps, per operator and axis.
Usage: python tools/npsweep.py [n ...]   -> one JSON line per (op, axis, n, path); 16 B per point against the measured HBM peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import padeops_b200 as pdo
from opsweep import timeit, PEAK


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [512]
    L = pdo.lib()
    for n in sizes:
        dx = 1.0 / (n - 1)
        f = torch.rand((n, n, n), dtype=torch.float64, device="cuda")
        o = torch.empty_like(f)
        c10, cf, c6 = pdo.cd10(), pdo.cf90(), pdo.cd06()
        assert c10.init(n, dx, periodic_=False) == 0 and cf.init(n, periodic_=False) == 0 and c6.init(n, dx, periodic_=False) == 0
        ops = [("cd10.dd", (c10.dd1, c10.dd2, c10.dd3), (0, 0)), ("cd10.dd[+1,-1]", (c10.dd1, c10.dd2, c10.dd3), (1, -1)),
               ("cd10.d2d", (c10.d2d1, c10.d2d2, c10.d2d3), (0, 0)), ("cf90.filter", (cf.filter1, cf.filter2, cf.filter3), (0, 0)),
               ("cd06.dd", (c6.dd1, c6.dd2, c6.dd3), (0, 0))]
        for name, fns, (b1, bn) in ops:
            for ax, fn in enumerate(fns):
                for mode, tag in ((1, "chunked"), (0, "sweeps")):
                    L.pdo_debug_np_fast(mode)
                    med, best = timeit(lambda: fn(f, o, bc1_=b1, bcn_=bn), reps=5 if mode else 3, warm=2 if mode else 1)
                    print(json.dumps({"op": f"{name}{ax+1}", "n": n, "path": tag, "ms": round(med, 4), "GBps": round(16 * n ** 3 / med / 1e6, 1),
                                      "frac": round(16 * n ** 3 / med / 1e6 / PEAK, 3)}), flush=True)
        L.pdo_debug_np_fast(-1)
        del f, o


if __name__ == "__main__":
    main()
