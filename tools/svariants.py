#!/usr/bin/env python
"""Times every strided-axis (y, z) kernel variant per operator: 6 pipe1, 5 cpipe, 3 cluster/streaming, 1 t512, 7 TMA pipeline.
Usage: python tools/svariants.py [n ...]   → one JSON line per (op, axis, n, variant)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import padeops_b200 as pdo
from opsweep import timeit, PEAK

NAMES = {6: "pipe1", 5: "cpipe", 3: "cluster", 1: "t512", 7: "stma", 8: "ctma64", 9: "ctma32", 10: "ctma32s", 11: "cpipe_t"}


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [512, 1024]
    L = pdo.lib()
    for n in sizes:
        d = 2 * np.pi / n
        f = torch.rand((n, n, n), dtype=torch.float64, device="cuda")
        o = torch.empty_like(f)
        ref = torch.empty_like(f)
        c10, c06, cf, ga = pdo.cd10(), pdo.cd06(), pdo.cf90(), pdo.gaussian()
        assert c10.init(n, d) == 0 and c06.init(n, d) == 0 and cf.init(n) == 0 and ga.init(n) == 0
        for name, fn in [("cd10.dd2", c10.dd2), ("cd10.dd3", c10.dd3), ("cd10.d2d2", c10.d2d2), ("cd10.d2d3", c10.d2d3),
                         ("cd06.dd2", c06.dd2), ("cd06.dd3", c06.dd3), ("cf90.filter2", cf.filter2), ("cf90.filter3", cf.filter3),
                         ("gaussian.filter2", ga.filter2), ("gaussian.filter3", ga.filter3)]:
            L.pdo_debug_set_variant(1, -1)
            fn(f, ref)
            for mode in (6, 5, 11, 9, 10):
                L.pdo_debug_set_variant(mode, -1)
                try:
                    fn(f, o)
                except Exception as e:
                    print(json.dumps({"op": name, "n": n, "variant": NAMES[mode], "error": str(e)[:80]}), flush=True)
                    continue
                ran = L.pdo_debug_last_variant()
                err = float((o - ref).abs().max() / ref.abs().max())
                med, best = timeit(lambda: fn(f, o), reps=6, warm=2)
                print(json.dumps({"op": name, "n": n, "variant": NAMES[mode], "ran": ran, "ms": round(med, 4), "ms_min": round(best, 4),
                                  "GBps": round(16 * n ** 3 / med / 1e6, 1), "frac": round(16 * n ** 3 / med / 1e6 / PEAK, 3),
                                  "maxrel_vs_t512": err}), flush=True)
        L.pdo_debug_set_variant(-1, -1)
        del f, o, ref


if __name__ == "__main__":
    main()
