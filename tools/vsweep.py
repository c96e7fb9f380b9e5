#!/usr/bin/env python
"""Variant sweep on arbitrary pencils: times the forced kernel variants of one operator along one axis of an (nx, ny, nz) field.
Usage: python tools/vsweep.py SPEC [SPEC ...]      SPEC = op:axis:nx,ny,nz:variant[,variant...]
       op in cd10.dd cd10.d2d cd06.dd cf90.filter gaussian.filter; axis in 1 2 3; variants by name (strided) or number (x axis)
e.g.   python tools/vsweep.py cd10.dd:3:1024,1024,1024:cpipe,ctma32,cpipe_t cd10.dd:1:2048,2048,256:1000,1016,128
One JSON line per (spec, variant): median / min ms, GB/s (16 B per point) and fraction of the measured HBM peak, max relative
difference against the t512 (or 128-thread) variant."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import padeops_b200 as pdo
from opsweep import timeit, PEAK

MODES = {"auto": 0, "t512": 1, "t256": 2, "cluster": 3, "cluster4": 4, "cpipe": 5, "pipe1": 6, "stma": 7, "ctma64": 8, "ctma32": 9,
         "ctma32s": 10, "cpipe_t": 11}


def main():
    L = pdo.lib()
    for spec in sys.argv[1:]:
        opn, ax, shp, vs = spec.split(":")
        ax = int(ax)
        nx, ny, nz = (int(v) for v in shp.split(","))
        n = (nx, ny, nz)[ax - 1]
        fam, meth = opn.split(".")
        obj = getattr(pdo, fam)()
        rc = obj.init(n, 2 * np.pi / n) if fam in ("cd10", "cd06") else obj.init(n)
        assert rc == 0
        fn = getattr(obj, f"{meth}{ax}")
        f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda")
        o = torch.empty_like(f)
        ref = torch.empty_like(f)
        if ax == 1:
            L.pdo_debug_set_variant(-1, 128)
        else:
            L.pdo_debug_set_variant(1, -1)
        try:
            fn(f, ref)
        except Exception:
            ref = None
        for v in vs.split(","):
            if ax == 1:
                L.pdo_debug_set_variant(-1, int(v))
            else:
                L.pdo_debug_set_variant(MODES[v], -1)
            rec = {"op": f"{opn}{ax}", "shape": [nx, ny, nz], "variant": v}
            try:
                fn(f, o)
            except Exception as e:
                rec["error"] = str(e)[:100]
                print(json.dumps(rec), flush=True)
                continue
            rec["ran"] = L.pdo_debug_last_variant()
            if ref is not None:
                rec["maxrel_vs_ref"] = float((o - ref).abs().max() / ref.abs().max())
            med, best = timeit(lambda: fn(f, o), reps=8, warm=2)
            pts = nx * ny * nz
            rec.update({"ms": round(med, 4), "ms_min": round(best, 4), "GBps": round(16 * pts / med / 1e6, 1),
                        "frac": round(16 * pts / med / 1e6 / PEAK, 3), "frac_best": round(16 * pts / best / 1e6 / PEAK, 3)})
            print(json.dumps(rec), flush=True)
        L.pdo_debug_set_variant(-1, -1)
        del f, o, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
