"""Theta* batch (8 192 map2 queries, the bench's) against the number of resident searches: python profiles/tools/mb_theta_slots.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from theta_rrt_b200 import OccupancyGrid, Planner
dev = torch.device("cuda:0")
m2 = bench.load_maps()["map2"]
pt = Planner(OccupancyGrid(m2, device=dev))
cells = np.argwhere(m2)
rq = np.random.default_rng(5)
a, b = cells[rq.integers(len(cells), size=8192)], cells[rq.integers(len(cells), size=8192)]
sg = torch.from_numpy(np.stack([a[:, 1], a[:, 0], b[:, 1], b[:, 0]], 1).astype(np.int32)).to(dev)
sms = torch.cuda.get_device_properties(0).multi_processor_count
for wps in (16, 20, 24, 28, 32, 0):
    ns = sms * wps
    fn = lambda: pt.theta(sg, path_cap=64, n_slots=ns)
    for _ in range(3):
        r = fn()
    torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record()
    for _ in range(3):
        r = fn()
    y.record(); torch.cuda.synchronize()
    ms = x.elapsed_time(y) / 3
    print(f"warps per SM {wps or 'default':>7}: {ms:7.2f} ms  {float(r.expanded.sum()) / ms / 1e3:6.1f} M expansions/s", flush=True)
