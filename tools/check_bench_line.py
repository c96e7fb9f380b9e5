#!/usr/bin/env python
"""Validates a bench.py JSON line against the driver's contract (keys, types, internal consistency).
Usage: python tools/check_bench_line.py <file with one JSON line> ...   (exit 1 on the first violation)"""
import json
import sys

REQ = {"metric": str, "value": (int, float), "unit": str, "n_gpus": int, "steps": int, "warmup": int, "ms_per_step": (int, float),
       "higher_is_better": bool, "scaling": str, "dtype": str, "data": str, "config": dict}


def check(line, reference=False):
    errs = []
    for k, t in REQ.items():
        if k not in line:
            errs.append(f"missing {k}")
        elif not isinstance(line[k], t):
            errs.append(f"{k}: {type(line[k]).__name__}")
    if "vs_baseline" not in line:
        errs.append("missing vs_baseline")
    if "workload" not in line.get("config", {}):
        errs.append("config.workload missing")
    if line.get("warmup", 0) < (0 if reference else 3):
        errs.append("warmup < 3")
    e2e = line.get("e2e")
    if e2e is not None and (not isinstance(e2e, dict) or not all(k in e2e for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"))):
        errs.append("e2e incomplete")    # None = leg skipped with --no-e2e
    cb = line.get("cpu_baseline")
    if cb is not None and (not isinstance(cb, dict) or not all(k in cb for k in ("value", "unit", "cores", "kind", "sample"))):
        errs.append("cpu_baseline incomplete")    # None = leg skipped with --no-cpu
    if reference:
        if line.get("impl") != "reference":
            errs.append("impl != reference")
        return errs
    roof = line.get("roofline")
    if not isinstance(roof, dict) or not all(k in roof for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")):
        errs.append("roofline incomplete")
    else:
        if abs(roof["frac"] - roof["achieved"] / roof["peak"]) > 1e-9:
            errs.append("roofline.frac != achieved / peak")
        if roof["bound"] not in ("hbm", "tensor"):
            errs.append("roofline.bound")
    if not isinstance(line.get("gpu_launches"), int) or line["gpu_launches"] <= 0:
        errs.append("gpu_launches")
    clk = line.get("clocks")
    if not isinstance(clk, dict) or not all(k in clk for k in ("sm_mhz", "sm_max_mhz", "reasons")):
        errs.append("clocks incomplete")
    elif any(r in clk["reasons"] for r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")):
        errs.append("throttled run")
    # value must be what ms_per_step implies: 3 derivatives x points per GPU x GPUs
    pts = line["config"].get("global")
    if pts:
        npts = pts[0] * pts[1] * pts[2]
        implied = 3.0 * npts / (line["ms_per_step"] * 1e-3) / 1e9
        if abs(implied - line["value"]) > 1e-6 * line["value"]:
            errs.append(f"value {line['value']} != 3 * points / time = {implied}")
    return errs


def main():
    bad = 0
    for fn in sys.argv[1:]:
        for raw in open(fn):
            if not raw.startswith("{"):
                continue
            line = json.loads(raw)
            errs = check(line, reference=line.get("impl") == "reference")
            print(fn, "OK" if not errs else errs)
            bad += bool(errs)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
