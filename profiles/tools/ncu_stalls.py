#!/usr/bin/env python
"""Warp-stall breakdown and opcode mix of an ncu report: python profiles/tools/ncu_stalls.py report.ncu-rep"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; tot = collections.Counter(); ops = collections.Counter(); opi = collections.Counter()
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    d = {}
    for k, v in zip(hdr, r): d.setdefault(k, v)
    for k, v in d.items():
        if k.startswith('stall_') and '(Not Issued)' not in k:
            try: tot[k] += int(v)
            except ValueError: pass
    src = d['Source'].split()
    op = src[1] if src and src[0].startswith('@') else (src[0] if src else '?')
    op = op.split('.')[0]
    try: ops[op] += int(d['# Samples']); opi[op] += int(d['Instructions Executed'])
    except ValueError: pass
s = sum(tot.values()) or 1
print("stall reasons (share of warp samples):")
for k, v in tot.most_common(10): print(f"  {100*v/s:6.2f}% {k}")
si = sum(opi.values()) or 1; ss = sum(ops.values()) or 1
print("opcode mix (share of executed warp instructions / of samples):")
for k, v in opi.most_common(16): print(f"  {100*v/si:6.2f}% {100*ops[k]/ss:6.2f}% {k}")
