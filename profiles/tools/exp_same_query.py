import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from theta_rrt_b200 import OccupancyGrid, Params, Planner
dev = torch.device("cuda:0")
free = bench.load_maps()["map1"]
K = 5001; nq = 4096
starts, goals, sxy, sth = bench.make_rrt_workload(free, 64, K)
p = Planner(OccupancyGrid(free, device=dev), Params(tol_xy=0.0, K=K))
r = p.rrt(starts, goals, sxy, sth, K=K).host()
full = [q for q in range(64) if r["iters"][q] == K - 1]
print("full-length queries among first 64:", len(full), "n_nodes", [int(r["n_nodes"][q]) for q in full[:8]])
def run(sel, label):
    idx = np.array(sel)
    d = [torch.from_numpy(np.ascontiguousarray(a[idx])).to(dev) for a in (starts, goals, sxy, sth)]
    for _ in range(3): res = p.rrt(*d, K=K)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): res = p.rrt(*d, K=K)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    it = int(res.iters.sum())
    print(f"{label}: {ms:.2f} ms, {it/ms/1e3:.1f} M expansions/s", flush=True)
run([full[0]] * nq, "4096 copies of one full-length query")
run([full[i % len(full)] for i in range(nq)], "4096 = cycling full-length queries")
run([i % 64 for i in range(nq)], "4096 = cycling first 64 queries (27% early)")
