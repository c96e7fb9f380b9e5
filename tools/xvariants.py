#!/usr/bin/env python
"""Times every contiguous-axis kernel variant (128 / 256 register-staged, 1000 / 1016 TMA pipeline) per operator.
Usage: python tools/xvariants.py [n ...]   → one JSON line per (op, n, variant)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import padeops_b200 as pdo
from opsweep import timeit, PEAK


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
        for name, fn in [("cd10.dd1", c10.dd1), ("cd10.d2d1", c10.d2d1), ("cd06.dd1", c06.dd1), ("cf90.filter1", cf.filter1),
                         ("gaussian.filter1", ga.filter1)]:
            L.pdo_debug_set_variant(-1, 128)
            fn(f, ref)
            for xth in (128, 256, 1000, 1016):
                L.pdo_debug_set_variant(-1, xth)
                try:
                    fn(f, o)
                except Exception as e:  # variant does not cover this shape
                    print(json.dumps({"op": name, "n": n, "variant": xth, "error": str(e)[:80]}), flush=True)
                    continue
                err = float((o - ref).abs().max() / ref.abs().max())
                med, best = timeit(lambda: fn(f, o))
                print(json.dumps({"op": name, "n": n, "variant": xth, "ms": round(med, 4), "ms_min": round(best, 4),
                                  "GBps": round(16 * n ** 3 / med / 1e6, 1), "frac": round(16 * n ** 3 / med / 1e6 / PEAK, 3),
                                  "maxrel_vs_128": err}), flush=True)
        L.pdo_debug_set_variant(-1, -1)
        del f, o, ref


if __name__ == "__main__":
    main()
