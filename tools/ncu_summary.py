#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into the per-kernel summary kept under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/<tag>_raw.csv profiles/<name>.json [profiles/traffic_rNN.json]
The optional third file maps kernel base name -> DRAM bytes per launch (read+write), which bench.py reports as
roofline.traffic."""
import csv
import json
import re
import sys

KEEP = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct_of_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "regs_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "smsp__inst_executed.sum": "warp_insts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "dram__cycles_elapsed.avg.per_second": "dram_clock",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out, traffic = [], {}
    for r in data:
        name = r[hdr.index("Kernel Name")]
        base = re.sub(r"^void\s+", "", name).split("<")[0].split("(")[0]
        d = {"kernel": name.split("(")[0], "base": base}
        for k, short in KEEP.items():
            hits = [i for i, h in enumerate(hdr) if h == k or h.endswith("." + k)]
            if not hits:
                continue
            i = hits[0]
            v = num(r[i])
            if v is None:
                continue
            u = units[i]
            if u in SCALE and short.startswith("dram_"):
                v *= SCALE[u]
                u = "byte"
            d[short] = v
            d[short + "_unit"] = u
        stalls = {}
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio$", h)
            if m and num(r[i]) and num(r[i]) >= 0.2:
                stalls[m.group(1)] = round(num(r[i]), 2)
        d["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        if "dram_read" in d and "dram_write" in d:
            d["dram_total_bytes"] = d["dram_read"] + d["dram_write"]
            traffic[base] = d["dram_total_bytes"]
            t = d.get("time")
            if t:
                sec = t * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(d.get("time_unit", "ns"), 1e-9)
                d["dram_GBps_under_ncu"] = d["dram_total_bytes"] / sec / 1e9
        out.append(d)
    json.dump({"source": sys.argv[1], "note": "ncu --set full --clock-control none; per-launch times are cold-cache and serialised",
               "kernels": out}, open(sys.argv[2], "w"), indent=1)
    if len(sys.argv) > 3:
        json.dump(traffic, open(sys.argv[3], "w"), indent=1)
    for d in out:
        print(d["base"], d.get("time"), d.get("time_unit"), "dram", d.get("dram_total_bytes"), d["stall_warps_per_issue"])


if __name__ == "__main__":
    main()
