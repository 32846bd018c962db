#!/usr/bin/env python
"""Aggregate an ncu report's source page per CUDA source line.

    python profiles/tools/ncu_hotspots.py report.ncu-rep [top_n]

Prints, per file and per line, the share of executed warp instructions and of stall samples, plus
the average active threads per instruction (divergence)."""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file = None
    hdr = None
    lines = collections.OrderedDict()
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 10 or r[0] == "":
            continue
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)  # first "Source" column is the CUDA source
        try:
            inst = int(d["Instructions Executed"]); tinst = int(d["Thread Instructions Executed"]); smp = int(d["# Samples"])
        except ValueError:
            continue
        key = (cur_file, int(d["Line No"]))
        a = lines.setdefault(key, [0, 0, 0, d["Source"].strip()])
        a[0] += inst; a[1] += tinst; a[2] += smp
    ti = sum(a[0] for a in lines.values()) or 1
    ts = sum(a[2] for a in lines.values()) or 1
    tt = sum(a[1] for a in lines.values())
    print(f"total warp inst {ti}  thread inst {tt}  avg active {tt / ti:.2f}  samples {ts}")
    files = collections.Counter(); fs = collections.Counter()
    for (f, _), a in lines.items():
        files[f] += a[0]; fs[f] += a[2]
    for f, v in files.most_common():
        print(f"  {f:28s} inst {100 * v / ti:5.1f}%  samples {100 * fs[f] / ts:5.1f}%")
    for (f, ln), a in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        act = a[1] / a[0] if a[0] else 0
        print(f"{100 * a[0] / ti:5.2f}% s={100 * a[2] / ts:5.2f}% act={act:4.1f} {f}:{ln}  {a[3][:110]}")


if __name__ == "__main__":
    main()
