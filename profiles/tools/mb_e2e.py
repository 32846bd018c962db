"""Where does the host-buffer pipeline (Planner.rrt_host) spend its time?  python profiles/tools/mb_e2e.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from theta_rrt_b200 import OccupancyGrid, Params, Planner
nq, K = 4096, 5001
dev = torch.device("cuda:0")
free = bench.load_maps()["map1"]
cache = f"/tmp/wl_{nq}_{K}.npz"
if os.path.exists(cache):
    z = np.load(cache); w = [z[k] for k in ("a", "b", "c", "d")]
else:
    w = bench.make_rrt_workload(free, nq, K); np.savez(cache, a=w[0], b=w[1], c=w[2], d=w[3])
p = Planner(OccupancyGrid(free, device=dev), Params(tol_xy=0.0, K=K))
h_in = [torch.from_numpy(a).pin_memory() for a in w]
d_in = [t.to(dev) for t in h_in]

def timed(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    p.host_sync()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t) * 1e3 / n

keep = {}
print("resident, dense           %.2f ms (wall %.2f)" % timed(lambda: keep.__setitem__("r", p.rrt(*d_in, K=K))))
print("resident, dense, no u     %.2f ms (wall %.2f)" % timed(lambda: keep.__setitem__("r", p.rrt(*d_in, K=K, want_u=False))))
pk = {}
def f():
    if "p" in pk:
        pk["p"]["pack_total"].zero_()  # first row to use
    r = p.rrt(*d_in, K=K, pack=True, packed=pk.get("p"))
    pk["p"] = {k: getattr(r, k) for k in ("pack_x", "pack_y", "pack_theta", "pack_parent", "pack_u", "row_start", "pack_total")}
print("resident, packed          %.2f ms (wall %.2f)" % timed(f))
names = {"node_x": ((nq, K), torch.float64), "node_y": ((nq, K), torch.float64), "node_theta": ((nq, K), torch.float64),
         "parent": ((nq, K), torch.int32), "u": ((nq, K, 5), torch.float64), "n_nodes": ((nq,), torch.int32),
         "sol": ((nq,), torch.int32), "status": ((nq,), torch.int32), "iters": ((nq,), torch.int32), "row_start": ((nq,), torch.int64)}
h_out = {k: torch.empty(s, dtype=dt).pin_memory() for k, (s, dt) in names.items()}
dense = {k: v for k, v in h_out.items() if k != "row_start"}
for chunks in (4, 16):
    for vro, u in ((False, True), (True, True), (True, False), (False, False)):
        out = dict(h_out if vro else dense)
        if not u: out.pop("u")
        ms, wall = timed(lambda: p.rrt_host(*h_in, out=out, K=K, chunks=chunks, wait=False, valid_rows_only=vro))
        ms2, wall2 = timed(lambda: (p.rrt_host(*h_in, out=out, K=K, chunks=chunks, wait=True, valid_rows_only=vro), torch.cuda.synchronize()), n=3, warm=1)
        print(f"rrt_host chunks={chunks:2d} packed={int(vro)} u={int(u)}: streamed {ms:7.2f} ms   serial {ms2:7.2f} ms", flush=True)
h16 = [h_in[0], h_in[1], h_in[2].to(torch.int16).pin_memory(), h_in[3]]
out = dict(h_out); out.pop("u")
print("int16 xy, packed, no u, chunks=16: streamed %.2f ms" % timed(lambda: p.rrt_host(*h16, out=out, K=K, chunks=16, wait=False, valid_rows_only=True))[0])
# raw copies
big = torch.empty(int(0.62e9), dtype=torch.uint8, device=dev); hb = torch.empty(int(0.62e9), dtype=torch.uint8).pin_memory()
print("D2H 0.62 GB single copy   %.2f ms" % timed(lambda: hb.copy_(big, non_blocking=True))[0])
print("H2D 0.33 GB (inputs)      %.2f ms" % timed(lambda: [d.copy_(h, non_blocking=True) for d, h in zip(d_in, h_in)])[0])
