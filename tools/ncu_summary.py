#!/usr/bin/env python
"""Summarise `ncu --set full` reports (gpurun_out/<tag>_{enc,dec,senc,sdec}.ncu-rep) into one JSON file:
duration, DRAM bytes, issue utilisation, stall reasons per issue, instruction count of the captured launch.

  python tools/ncu_summary.py <tag> profiles/<tag>_ncu_full_summary.json [key=path.ncu-rep ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active")


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    res = {"kernel": vals[hdr.index("Kernel Name")]}
    for i, name in enumerate(hdr):
        if name in KEEP or (name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio")):
            try:
                res[name] = {"value": float(vals[i].replace(",", "")), "unit": units[i]}
            except ValueError:
                pass
    return res


def main():
    tag, dst = sys.argv[1], sys.argv[2]
    out = {}
    for key in ("enc", "dec", "senc", "sdec"):
        p = os.path.join("gpurun_out", f"{tag}_{key}.ncu-rep")
        if os.path.exists(p):
            out[key] = summarise(p)
    for extra in sys.argv[3:]:
        key, p = extra.split("=", 1)
        out[key] = summarise(p)
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in out.items():
        rd, wr = v.get("dram__bytes_read.sum", {}), v.get("dram__bytes_write.sum", {})
        print(k, v["kernel"][:50], v.get("gpu__time_duration.sum"), "dram", rd, wr)


if __name__ == "__main__":
    main()
