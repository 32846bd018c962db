import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from theta_rrt_b200 import OccupancyGrid, Params, Planner
dev = torch.device("cuda:0")
free = bench.load_maps()["map1"]
K = 5001; nq = 4096
starts, goals, sxy, sth = bench.make_rrt_workload(free, nq, K)
p = Planner(OccupancyGrid(free, device=dev), Params(tol_xy=0.0, K=K))
d = [torch.from_numpy(a).to(dev) for a in (starts, goals, sxy, sth)]
r = p.rrt(*d, K=K, schedule=2)
torch.cuda.synchronize()
print(int(r.iters.sum()))
