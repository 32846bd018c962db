"""Small fixed workloads for ncu captures (one launch of each hot kernel after one warm-up launch).

    python profiles/prof_workload.py rrt [nq] [K] [lanes]
    python profiles/prof_workload.py nearest | los | theta
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from theta_rrt_b200 import OccupancyGrid, Params, Planner  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "rrt"
    dev = torch.device("cuda:0")
    maps = bench.load_maps()
    if what == "rrt":
        nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
        K = int(sys.argv[3]) if len(sys.argv) > 3 else 2001
        lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
        free = maps["map1"]
        starts, goals, sxy, sth = bench.make_rrt_workload(free, nq, K)
        p = Planner(OccupancyGrid(free, device=dev), Params(tol_xy=0.0, K=K))
        d = [torch.from_numpy(a).to(dev) for a in (starts, goals, sxy, sth)]
        for _ in range(2):
            r = p.rrt(*d, K=K, lanes=lanes)
        torch.cuda.synchronize()
        print("rrt done", int(r.iters.sum()), int(r.n_nodes.sum()))
    elif what == "nearest":
        rng = np.random.default_rng(3)
        p = Planner(OccupancyGrid(np.ones((8, 8), bool), device=dev))
        n = 1 << 20
        x = torch.from_numpy(rng.uniform(0, 8191, n)).to(dev)
        y = torch.from_numpy(rng.uniform(0, 8191, n)).to(dev)
        q = torch.from_numpy(rng.integers(0, 8192, size=(4096, 2)).astype(np.int32)).to(dev)
        for _ in range(2):
            p.nearest(x, y, q)
        p.nearest(x, y, q[:1].contiguous())
        torch.cuda.synchronize()
    elif what == "los":
        big = bench.synthetic_map(8192, 0.1, 8, 42)
        p = Planner(OccupancyGrid(big, device=dev))
        seg = torch.from_numpy(bench.make_segments(big, 1 << 20, 7)).to(dev)
        for _ in range(2):
            p.los(seg)
        torch.cuda.synchronize()
    elif what == "theta1":
        p = Planner(OccupancyGrid(maps["map2"], device=dev))
        one = torch.tensor([[280, 0, 8, 280]], dtype=torch.int32, device=dev)
        for _ in range(2):
            r = p.theta(one, lanes=32)
        torch.cuda.synchronize()
        print("theta1 done", int(r.expanded.sum()))
    elif what == "theta":
        m2 = maps["map2"]
        p = Planner(OccupancyGrid(m2, device=dev))
        cells = np.argwhere(m2)
        rq = np.random.default_rng(5)
        nqt = int(sys.argv[2]) if len(sys.argv) > 2 else 8192  # the bench batch
        a, b = cells[rq.integers(len(cells), size=nqt)], cells[rq.integers(len(cells), size=nqt)]
        sg = torch.from_numpy(np.stack([a[:, 1], a[:, 0], b[:, 1], b[:, 0]], 1).astype(np.int32)).to(dev)
        for _ in range(2):
            r = p.theta(sg, path_cap=64)
        torch.cuda.synchronize()
        print("theta done", int(r.expanded.sum()))


if __name__ == "__main__":
    main()
