"""LOS microbenchmark on the cfg-4 grid (8192^2, 2^20 random rays): python profiles/tools/mb_los.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from theta_rrt_b200 import OccupancyGrid, Planner
dev = torch.device("cuda:0")
big = bench.synthetic_map(8192, 0.1, 8, 42)
pl = Planner(OccupancyGrid(big, device=dev))
seg = bench.make_segments(big, 1 << 20, 7)
d_seg = torch.from_numpy(seg).to(dev)
out = torch.empty(len(seg), dtype=torch.uint8, device=dev)
ref = None
for lanes in (1, 2, 4, 8, 16, 32):
    os.environ["TRRT_LOS_LANES"] = str(lanes)
    for _ in range(3): pl.los(d_seg, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): pl.los(d_seg, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    res = out.cpu().numpy()
    if ref is None: ref = res.copy()
    print(f"lanes {lanes:2d}: {ms*1e3:8.1f} us  {len(seg)/ms/1e6:8.2f} G checks/s  same as lanes=1: {bool((res == ref).all())}", flush=True)
