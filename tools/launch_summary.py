#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel.  Usage: launch_summary.py file.csv [out.json]"""
import collections
import csv
import json
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        ms = v / 1e6 if u.startswith("ns") else (v / 1e3 if u.startswith("us") else v)
        k = re.sub(r"\(.*", "", r[ki])[:110]
        agg[k][0] += 1
        agg[k][1] += ms
    tot = sum(v[1] for v in agg.values())
    out = {"total_ms": round(tot, 3), "kernels": [{"kernel": k, "launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / tot, 4),
                                                   "ms_per_launch": round(v[1] / v[0], 4)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    for k in out["kernels"][:30]:
        print(f"{k['launches']:5d} {k['ms']:9.2f} {k['share']:6.3f} {k['ms_per_launch']:8.4f}  {k['kernel']}")
    print("total ms", out["total_ms"])
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
