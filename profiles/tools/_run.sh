timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_main_chain.py -x -q -k "theta or cfg5 or main or dropin" 2>&1 | tail -3
python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
import bench
from theta_rrt_b200 import OccupancyGrid, Planner
maps = bench.load_maps(); m2 = maps["map2"]; dev = torch.device("cuda:0")
pt = Planner(OccupancyGrid(m2, device=dev))
one = torch.tensor([[280, 0, 8, 280]], dtype=torch.int32, device=dev)
def timed(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for strips in (False, True):
    print("strips", strips, "theta single query ms", timed(lambda: pt.theta(one, lanes=32, strips=strips)))
cells = np.argwhere(m2); rq = np.random.default_rng(5); nqt = 8192
a, b = cells[rq.integers(len(cells), size=nqt)], cells[rq.integers(len(cells), size=nqt)]
sg = torch.from_numpy(np.stack([a[:, 1], a[:, 0], b[:, 1], b[:, 0]], 1).astype(np.int32)).to(dev)
for strips in (False, True):
    ms = timed(lambda: pt.theta(sg, path_cap=64, strips=strips))
    r = pt.theta(sg, path_cap=64, strips=strips)
    print("strips", strips, "theta batch ms", ms, "M exp/s", float(r.expanded.sum()) / ms / 1e3, "nlos", int(r.n_los.sum()))
PY
