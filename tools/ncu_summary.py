#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): one block of key metrics per captured launch.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/rNN_name.txt]"""
import csv
import re
import subprocess
import sys

PAT = re.compile(r"^(Kernel Name|Grid Size|Block Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                 r"launch__registers_per_thread|launch__occupancy_limit_registers|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                 r"sm__pipe_(fma|fmaheavy|alu|fp64)_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__inst_executed_pipe_(fma|alu|lsu|uniform|xu)\.avg\.pct_of_peak_sustained_active|"
                 r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__inst_executed\.sum|sm__cycles_elapsed\.max|"
                 r"smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio|smsp__average_warp_latency_per_inst_issued\.ratio|lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct)$")


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        stalls = []
        for i, h in enumerate(hdr):
            if not PAT.match(h):
                continue
            if "issue_stalled" in h:
                try:
                    stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
                continue
            print(f"{h} = {r[i]} {units[i]}")
        stalls.sort(reverse=True)
        print("top stall reasons (warps stalled per issue-active cycle):", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))
        print("---")


if __name__ == "__main__":
    main()
