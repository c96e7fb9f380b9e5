#!/usr/bin/env python
"""Per-operator / per-axis device throughput sweep (CUDA events, device-resident fields).
Usage: python tools/opsweep.py [n ...]   → one JSON line per (op, axis, n)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import padeops_b200 as pdo

PEAK = 6551.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [512, 1024]
    for n in sizes:
        d = 2 * np.pi / n
        f = torch.rand((n, n, n), dtype=torch.float64, device="cuda")
        o = torch.empty_like(f)
        c10, c06, cf, ga = pdo.cd10(), pdo.cd06(), pdo.cf90(), pdo.gaussian()
        assert c10.init(n, d) == 0 and c06.init(n, d) == 0 and cf.init(n) == 0 and ga.init(n) == 0
        ops = [("cd10.dd", (c10.dd1, c10.dd2, c10.dd3)), ("cd10.d2d", (c10.d2d1, c10.d2d2, c10.d2d3)),
               ("cd06.dd", (c06.dd1, c06.dd2, c06.dd3)), ("cf90.filter", (cf.filter1, cf.filter2, cf.filter3)),
               ("gaussian.filter", (ga.filter1, ga.filter2, ga.filter3))]
        # reference points: a plain device copy and torch's FFT-free elementwise pass on the same field
        med, best = timeit(lambda: o.copy_(f))
        print(json.dumps({"op": "torch copy_", "n": n, "ms": med, "GBps": 16 * n ** 3 / med / 1e6, "frac": 16 * n ** 3 / med / 1e6 / PEAK}), flush=True)
        for name, fns in ops:
            for ax, fn in enumerate(fns):
                med, best = timeit(lambda: fn(f, o))
                print(json.dumps({"op": f"{name}{ax+1}", "n": n, "ms": round(med, 4), "ms_min": round(best, 4),
                                  "Gpts": round(n ** 3 / med / 1e6, 1), "GBps": round(16 * n ** 3 / med / 1e6, 1),
                                  "frac": round(16 * n ** 3 / med / 1e6 / PEAK, 3)}), flush=True)
        del f, o


if __name__ == "__main__":
    main()
