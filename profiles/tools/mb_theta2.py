"""Theta* batch-size / residency sweep on map2."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from theta_rrt_b200 import OccupancyGrid, Planner
dev = torch.device("cuda:0")
m2 = bench.load_maps()["map2"]
cells = np.argwhere(m2); rq = np.random.default_rng(5)
for wps in (0,):
    pt = Planner(OccupancyGrid(m2, device=dev))
    for nqt in (2048, 4096, 8192):
        a, b = cells[rq.integers(len(cells), size=nqt)], cells[rq.integers(len(cells), size=nqt)]
        sg = torch.from_numpy(np.stack([a[:, 1], a[:, 0], b[:, 1], b[:, 0]], 1).astype(np.int32)).to(dev)
        for _ in range(2): r = pt.theta(sg, path_cap=64, lanes=32)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): r = pt.theta(sg, path_cap=64, lanes=32)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"auto queries {nqt}: {ms:7.2f} ms  {float(r.expanded.sum())/ms/1e3:6.1f} M expansions/s  slots {r.extra['n_slots']} workspace {r.extra['workspace_bytes']/1e9:.1f} GB", flush=True)
    del pt
    torch.cuda.empty_cache()
