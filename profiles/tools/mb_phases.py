"""How long would one window of 4096 queries x 32 lanes take phase by phase (standalone step kernels)?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from theta_rrt_b200 import OccupancyGrid, Params, Planner
dev = torch.device("cuda:0")
free = bench.load_maps()["map1"]
p = Planner(OccupancyGrid(free, device=dev), Params(tol_xy=0.0))
rng = np.random.default_rng(1)
n = 4096 * 32
cells = np.argwhere(free)
o = cells[rng.integers(len(cells), size=n)] + rng.uniform(0, 1, (n, 2))
inp = np.stack([o[:, 1], o[:, 0], rng.uniform(-180, 180, n), rng.integers(0, 100, n).astype(float), rng.integers(0, 100, n).astype(float), rng.uniform(-180, 180, n)], 1)
d_in = torch.from_numpy(inp).to(dev)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
out, straight = p.steer(d_in)
print(f"steer_batch  {n} items: {timed(lambda: p.steer(d_in)):8.1f} us")
so = out.cpu().numpy(); st = straight.cpu().numpy().astype(bool)
rows = np.stack([inp[:, 0], inp[:, 1], so[:, 0], so[:, 1], so[:, 3], so[:, 4], so[:, 5], so[:, 6], st.astype(float)], 1)
rows[st, 5:8] = 0.0
d_rows = torch.from_numpy(rows).to(dev)
for lanes in (1, 8):
    print(f"arc_batch lanes={lanes} {n} items: {timed(lambda: p.arc_blocked(d_rows, lanes=lanes)):8.1f} us")
dr = np.stack([inp[:, 0], inp[:, 1], inp[:, 2], so[:, 3], so[:, 4], so[:, 5], so[:, 6], so[:, 7] / 3], 1)[~st]
d_dr = torch.from_numpy(dr).to(dev)
print(f"drive_batch  {len(dr)} items: {timed(lambda: p.drive(d_dr)):8.1f} us")
seg = np.stack([inp[:, 0], inp[:, 1], so[:, 0], so[:, 1]], 1).astype(np.int32)
d_seg = torch.from_numpy(seg).to(dev)
print(f"los_batch    {n} items: {timed(lambda: p.los(d_seg)):8.1f} us")
# scan: 4096 queries each against its own tree of 1237 nodes ~ one tree of 1237 nodes, 131072 query points
x = torch.from_numpy(rng.uniform(0, 99, 1237)).to(dev); y = torch.from_numpy(rng.uniform(0, 99, 1237)).to(dev)
q = torch.from_numpy(rng.integers(0, 100, size=(n, 2)).astype(np.int32)).to(dev)
print(f"nearest      {n} points x 1237 nodes: {timed(lambda: p.nearest(x, y, q)):8.1f} us")
