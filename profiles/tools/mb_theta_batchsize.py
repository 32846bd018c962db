import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from theta_rrt_b200 import OccupancyGrid, Planner
dev = torch.device("cuda:0")
m2 = bench.load_maps()["map2"]
pt = Planner(OccupancyGrid(m2, device=dev))
cells = np.argwhere(m2)
for nq in (8192, 32768, 131072):
    rq = np.random.default_rng(5)
    a, b = cells[rq.integers(len(cells), size=nq)], cells[rq.integers(len(cells), size=nq)]
    sg = torch.from_numpy(np.stack([a[:, 1], a[:, 0], b[:, 1], b[:, 0]], 1).astype(np.int32)).to(dev)
    fn = lambda: pt.theta(sg, path_cap=64)
    for _ in range(2): r = fn()
    torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record(); r = fn(); y.record(); torch.cuda.synchronize()
    ms = x.elapsed_time(y)
    print(f"queries {nq:7d}: {ms:8.2f} ms  {float(r.expanded.sum()) / ms / 1e3:6.1f} M expansions/s  longest search {int(r.expanded.max())} expansions", flush=True)
