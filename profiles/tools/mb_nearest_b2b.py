import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from theta_rrt_b200 import OccupancyGrid, Planner
dev = torch.device("cuda:0")
pl = Planner(OccupancyGrid(np.ones((8, 8), bool), device=dev))
rng = np.random.default_rng(3)
for n_nodes, nq in ((1 << 24, 1), (1 << 25, 1), (1 << 24, 4)):
    x = torch.from_numpy(rng.uniform(0, 8191, n_nodes)).to(dev)
    y = torch.from_numpy(rng.uniform(0, 8191, n_nodes)).to(dev)
    q = torch.from_numpy(rng.integers(0, 8192, size=(nq, 2)).astype(np.int32)).to(dev)
    for _ in range(5):
        pl.nearest(x, y, q)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): pl.nearest(x, y, q)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)/50
    print(f"nodes {n_nodes} q {nq}: {ms*1e3:.1f} us/call back-to-back -> {16.0*n_nodes/ms/1e6:.1f} GB/s", flush=True)
