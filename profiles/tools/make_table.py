#!/usr/bin/env python
"""Markdown measurement table from a bench.py JSON line: python profiles/tools/make_table.py profiles/r1/bench_r1_final.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]; e = d["e2e"]; c = d.get("cpu_baseline", {}); s = d.get("secondary", {})
def f(v, u=""): return f"{v:,.0f}{u}" if v >= 100 else f"{v:.3g}{u}"
rows = [
 ("RRT expansions/s, inputs resident (`value`)", f(d["value"]), f"{d['ms_per_step']:.1f} ms per step of {d['expansions_per_step']:,} expansions, {d['n_gpus']} GPU"),
 ("RRT expansions/s, host buffers in and out (`e2e`)", f(e["value"]), f"{e['ms_per_step']:.1f} ms; H2D {e['h2d_bytes_per_step']/1e6:.0f} MB, D2H {e['d2h_bytes_per_step']/1e6:.0f} MB per step; {e.get('api','')}"),
 ("fused kernel, algorithmic GB/s (`roofline`)", f(r["achieved"]), f"{r['frac']:.3f} of the measured HBM copy peak {r['peak']:.0f} GB/s; DRAM traffic per launch (ncu): {r['traffic']}"),
]
if c:
    rows.append(("CPU port, all host threads (`cpu_baseline`)", f(c["value"]), f"{c['cores']} threads; single thread {f(c.get('single_core_value', 0))}; {c['sample']}"))
for k, v in s.items():
    extra = []
    if "ms" in v: extra.append(f"{v['ms']:.3f} ms")
    if "roofline" in v: extra.append(f"{v['roofline']['achieved']:.0f} GB/s algorithmic = {v['roofline']['frac']:.3f} of HBM peak")
    if "cpu_baseline" in v: extra.append(f"CPU port {f(v['cpu_baseline']['value'])} {v['cpu_baseline']['unit']} on {v['cpu_baseline']['cores']} threads")
    for kk in ("expanded", "los_checks_per_sec", "visible_fraction", "pixel_tests_per_sec_upper", "queries"):
        if kk in v: extra.append(f"{kk} {f(v[kk])}")
    rows.append((f"{k}: {v['metric']}", f(v["value"]) + " " + v.get("unit", ""), "; ".join(extra)))
print("| Quantity | Value | Notes |\n|---|---|---|")
for a, b, cc in rows: print(f"| {a} | {b} | {cc} |")
print(f"\nClocks during the timed region: {d['clocks']}")
