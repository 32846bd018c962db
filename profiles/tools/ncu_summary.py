#!/usr/bin/env python
"""One-screen summary of an ncu report: python profiles/tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_global_loads"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("Kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print(f"  {k} [{u.get(k, '')}] = {d[k]}")


if __name__ == "__main__":
    main()
