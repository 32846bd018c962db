"""cfg-5 RRT half (K = 1001, 64 random 256^2 maps) through one build of the library:
    python profiles/tools/mb_cfg5.py [so path] [queries]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from theta_rrt_b200 import _lib
if len(sys.argv) > 1 and sys.argv[1] != "-":
    _lib.SO_PATH = os.path.abspath(sys.argv[1])
    import ctypes
    probe = ctypes.CDLL(_lib.SO_PATH)
    _lib.SIGNATURES = {k: v for k, v in _lib.SIGNATURES.items() if hasattr(probe, k)}
import bench
from theta_rrt_b200 import OccupancyGrid, Params, Planner
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
c5 = bench.make_cfg5(0, 1, nq5=nq)
dev = torch.device("cuda:0")
p5 = Planner(OccupancyGrid(c5["maps"], device=dev), Params(tol_xy=0.0, K=c5["K"]))
d5 = [torch.from_numpy(v).to(dev) for v in (c5["starts"], c5["goals"], c5["sxy"], c5["sth"])]
dm = torch.from_numpy(c5["mid_r"]).to(dev)
for want_u in (False, True):
    for _ in range(2):
        r = p5.rrt(*d5, K=c5["K"], map_id=dm, want_u=want_u)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        r = p5.rrt(*d5, K=c5["K"], map_id=dm, want_u=want_u)
    b.record(); torch.cuda.synchronize()
    print(f"{os.path.basename(_lib.SO_PATH):24s} cfg5 rrt {nq} queries want_u={want_u}: {a.elapsed_time(b) / 3:8.2f} ms  iters {int(r.iters.sum())} nodes {int(r.n_nodes.sum())}", flush=True)
