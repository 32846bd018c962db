"""cfg 5 on one GPU: the RRT batch and the Theta* batch back to back against both issued at once on two streams.
    python profiles/tools/mb_cfg5_streams.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from theta_rrt_b200 import OccupancyGrid, Params, Planner
dev = torch.device("cuda:0")
c5 = bench.make_cfg5(0, 1)
p5 = Planner(OccupancyGrid(c5["maps"], device=dev), Params(tol_xy=0.0, K=c5["K"]))
d5 = [torch.from_numpy(v).to(dev) for v in (c5["starts"], c5["goals"], c5["sxy"], c5["sth"])]
sg5 = torch.from_numpy(c5["sg"]).to(dev)
dm_r, dm_t = torch.from_numpy(c5["mid_r"]).to(dev), torch.from_numpy(c5["mid_t"]).to(dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
sms = torch.cuda.get_device_properties(0).multi_processor_count
keep = {}


def serial():
    keep["r"] = p5.rrt(*d5, K=c5["K"], map_id=dm_r, want_u=False)
    keep["t"] = p5.theta(sg5, map_id=dm_t, path_cap=64)


def both(theta_first, wps):
    def f():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        def th():
            with torch.cuda.stream(s1):
                keep["t"] = p5.theta(sg5, map_id=dm_t, path_cap=64, n_slots=sms * wps if wps else 0)
        def rr():
            with torch.cuda.stream(s2):
                keep["r"] = p5.rrt(*d5, K=c5["K"], map_id=dm_r, want_u=False)
        (th(), rr()) if theta_first else (rr(), th())
        cur.wait_stream(s1); cur.wait_stream(s2)
    return f


def timed(fn, n=3, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print(f"back to back: {timed(serial):7.2f} ms", flush=True)
ref = (keep["r"].n_nodes.clone(), keep["t"].expanded.clone())
for tf in (True, False):
    for wps in (0, 16, 12, 8):
        ms = timed(both(tf, wps))
        ok = torch.equal(ref[0], keep["r"].n_nodes) and torch.equal(ref[1], keep["t"].expanded)
        print(f"two streams, {'Theta* first' if tf else 'RRT first   '}, Theta* warps per SM {wps or 'default':>7}: {ms:7.2f} ms  same results {ok}", flush=True)
