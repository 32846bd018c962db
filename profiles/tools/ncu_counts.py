#!/usr/bin/env python
"""gpurun_out/counts_<kernel>.csv (profiles/tools/ncu_counts.sh) -> profiles/kernel_counts.json, read by bench.py:
per-launch instruction / byte counts of each hot kernel on its bench workload."""
import csv, glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = {}
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "counts_*.csv"))):
    name = os.path.basename(path)[len("counts_"):-4]
    m = {}
    kern = None
    for r in csv.reader(open(path)):
        if len(r) >= 15 and r[0].isdigit():
            kern = r[4]
            try:
                m[r[12]] = float(r[14].replace(",", ""))
            except ValueError:
                pass
    if not m:
        continue
    out[name] = {
        "kernel_name": kern,
        "warp_inst": m.get("smsp__inst_executed.sum"),
        "thread_inst": m.get("smsp__thread_inst_executed.sum"),
        "fp64_warp_inst": m.get("smsp__inst_executed_pipe_fp64.sum"),
        "fp64_thread_inst": m.get("smsp__thread_inst_executed_pipe_fp64_pred_on.sum"),
        "alu_warp_inst": m.get("smsp__inst_executed_pipe_alu.sum"),
        "fma_warp_inst": m.get("smsp__inst_executed_pipe_fma.sum"),
        "lsu_warp_inst": m.get("smsp__inst_executed_pipe_lsu.sum"),
        "l2_bytes": m.get("lts__t_bytes.sum"),
        "l1_bytes": m.get("l1tex__t_bytes.sum"),
        "smem_wavefronts": m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "dram_bytes": (m.get("dram__bytes_read.sum") or 0) + (m.get("dram__bytes_write.sum") or 0),
        "ncu_duration_ms": (m.get("gpu__time_duration.sum") or 0) / 1e6,
        "ncu_issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "ncu_fp64_pipe_active_pct": m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "source": f"profiles/{tag}/counts_{name}.csv (ncu --metrics pass, profiles/tools/ncu_counts.sh)",
    }
# the single-query nearest scan (256 MiB stream): device time of the kernel alone from the bench's ncu launch list
ll = sorted(glob.glob(os.path.join(ROOT, "profiles", tag, "bench_launches_*.csv")))
if ll:
    d = [float(r[-1]) for r in csv.reader(open(ll[-1])) if len(r) > 14 and "nearest_tile_kernel<1>" in r[4] and r[-2] == "ns"]
    if d:
        d.sort()
        out["nearest_tile_kernel_single_query"] = {"ncu_duration_us": d[len(d) // 2] / 1e3, "launches": len(d),
                                                   "source": f"profiles/{tag}/{os.path.basename(ll[-1])} (gpu__time_duration, median)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "kernel_counts.json"), "w"), indent=1)
for k, v in out.items():
    print(k, {a: b for a, b in v.items() if a != "source"})
