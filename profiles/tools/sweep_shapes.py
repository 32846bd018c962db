"""CTA-shape sweep of the fused RRT kernel: builds one library per (threads, blocks/SM) and times cfg 3 with each.
    python profiles/tools/sweep_shapes.py build     (here: nvcc)
    python profiles/tools/sweep_shapes.py           (GPU box)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SHAPES = [(384, 2), (768, 1), (512, 1), (640, 1), (256, 2)]
def so(t, b): return os.path.join(ROOT, "profiles", "tools", f"_libthetarrt_{t}x{b}.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    from theta_rrt_b200 import build as B
    procs = [subprocess.Popen([B.nvcc_path(), *B.NVCC_FLAGS, f"-DTRRT_SPEC_THREADS={t}", f"-DTRRT_SPEC_BLOCKS_PER_SM={b}", "-o", so(t, b),
                               os.path.join(B.CSRC, "thetarrt.cu")]) for t, b in SHAPES]
    sys.exit(max(p.wait() for p in procs))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import numpy as np, torch
    from theta_rrt_b200 import _lib
    _lib.SO_PATH = sys.argv[2]
    import bench
    from theta_rrt_b200 import OccupancyGrid, Params, Planner
    dev = torch.device("cuda:0")
    free = bench.load_maps()["map1"]
    z = np.load("/tmp/wl.npz")
    p = Planner(OccupancyGrid(free, device=dev), Params(tol_xy=0.0, K=5001))
    d = [torch.from_numpy(z[k]).to(dev) for k in ("starts", "goals", "sxy", "sth")]
    for _ in range(3): p.rrt(*d, K=5001)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): r = p.rrt(*d, K=5001)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"{os.path.basename(sys.argv[2]):32s} {ms:7.2f} ms  {int(r.iters.sum()) / ms / 1e3:7.1f} M expansions/s", flush=True)
    sys.exit(0)
import numpy as np
import bench
free = bench.load_maps()["map1"]
starts, goals, sxy, sth = bench.make_rrt_workload(free, 4096, 5001)
np.savez("/tmp/wl.npz", starts=starts, goals=goals, sxy=sxy, sth=sth)
for t, b in SHAPES + SHAPES:
    if os.path.exists(so(t, b)):
        subprocess.call([sys.executable, os.path.abspath(__file__), "one", so(t, b)])
