"""Nearest-scan microbenchmark: python profiles/tools/mb_nearest.py  (CUDA events, 20 reps after 5 warm-ups)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from theta_rrt_b200 import OccupancyGrid, Planner

dev = torch.device("cuda:0")
pl = Planner(OccupancyGrid(np.ones((8, 8), bool), device=dev))
rng = np.random.default_rng(3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n_nodes, nq in ((1 << 24, 1), (1 << 24, 8), (1 << 22, 1), (1 << 20, 1), (1 << 20, 64), (1 << 20, 4096), (5001, 4096)):
    x = torch.from_numpy(rng.uniform(0, 8191, n_nodes)).to(dev)
    y = torch.from_numpy(rng.uniform(0, 8191, n_nodes)).to(dev)
    q = torch.from_numpy(rng.integers(0, 8192, size=(nq, 2)).astype(np.int32)).to(dev)
    for _ in range(5):
        pl.nearest(x, y, q)
    ts = []
    for _ in range(20):
        if n_nodes * 16 < (200 << 20):
            flush.zero_()  # tree smaller than L2: evict it so the scan reads HBM
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pl.nearest(x, y, q); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    print(f"nodes {n_nodes:9d} queries {nq:5d}: {ms*1e3:9.1f} us  stream {16.0*n_nodes/ms/1e6:8.1f} GB/s  algorithmic {16.0*n_nodes*nq/ms/1e6:10.1f} GB/s", flush=True)
