#!/usr/bin/env python
"""Static SASS size and executed instructions per source function (ncu source page, needs -lineinfo):
   python profiles/tools/ncu_funcs.py report.ncu-rep [source_dir]"""
import collections, csv, os, re, subprocess, sys
rep = sys.argv[1]
srcdir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "theta_rrt_b200", "csrc")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the report embeds the source it was built from (--import-source on): rebuild function boundaries from it
files = collections.defaultdict(dict)
cur = None; hdr = None; curline = None
static = collections.Counter(); dyn = collections.Counter(); smp = collections.Counter()
fn_re = re.compile(r'^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__global__|__device__|__host__|TL_FN|TL_ENTRY)[^;]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(')
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "":
        curline = int(r[0]); files[cur][curline] = r[1]; continue
    try: inst = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")])
    except ValueError: inst = s = 0
    static[(cur, curline)] += 1; dyn[(cur, curline)] += inst; smp[(cur, curline)] += s
def func_of(f, line):
    path = os.path.join(srcdir, f)
    if not os.path.exists(path): return f
    if f not in func_of.cache:
        names = []; last = f + ":<top>"
        for i, t in enumerate(open(path), 1):
            m = fn_re.match(t)
            if m and not t.strip().endswith(";"): last = m.group(1)
            names.append(last)
        func_of.cache[f] = names
    names = func_of.cache[f]
    return names[line - 1] if 0 < line <= len(names) else f
func_of.cache = {}
agg = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), n in static.items():
    a = agg[(f, func_of(f, l))]; a[0] += n; a[1] += dyn[(f, l)]; a[2] += smp[(f, l)]
tot = sum(a[0] for a in agg.values()); td = sum(a[1] for a in agg.values()) or 1; ts = sum(a[2] for a in agg.values()) or 1
print(f"static {tot} instr = {tot*16/1024:.1f} KB; executed {td/1e9:.2f} G warp-instr")
hot = 0
for (f, fn), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[0]*16/1024:7.1f} KB  dyn {100*a[1]/td:5.1f}%  samples {100*a[2]/ts:5.1f}%  {f}:{fn}")
