"""Theta* microbenchmark (map2): single reference query and a 2048-query batch, per lane count."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from theta_rrt_b200 import OccupancyGrid, Planner
dev = torch.device("cuda:0")
m2 = bench.load_maps()["map2"]
pt = Planner(OccupancyGrid(m2, device=dev))
def timed(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
one = torch.tensor([[280, 0, 8, 280]], dtype=torch.int32, device=dev)
cells = np.argwhere(m2); rq = np.random.default_rng(5); nqt = 2048
a, b = cells[rq.integers(len(cells), size=nqt)], cells[rq.integers(len(cells), size=nqt)]
sg = torch.from_numpy(np.stack([a[:, 1], a[:, 0], b[:, 1], b[:, 0]], 1).astype(np.int32)).to(dev)
for lanes in (8, 16, 32):
    ms1 = timed(lambda: pt.theta(one, lanes=lanes))
    r = pt.theta(one, lanes=lanes).host()
    msb = timed(lambda: pt.theta(sg, path_cap=64, lanes=lanes))
    rb = pt.theta(sg, path_cap=64, lanes=lanes).host()
    print(f"lanes {lanes:2d}: single query {ms1:7.2f} ms ({int(r['expanded'][0])} expanded, cost {float(r['cost'][0]):.6f}); "
          f"batch {nqt}: {msb:7.2f} ms = {float(rb['expanded'].sum())/msb/1e3:6.1f} M expansions/s, {float(rb['n_los'].sum())/msb/1e3:6.1f} M LOS/s", flush=True)
