"""Times the fused RRT kernel (cfg 3, resident inputs) for several builds of libthetarrt.so:
    python profiles/tools/variant_bench.py build NAME [nvcc flags...]     -> profiles/tools/_variants/NAME.so
    python profiles/tools/variant_bench.py run [NAME ...]                  (default: every built variant)
Each variant runs in its own process (3 warm-ups, 5 timed launches, CUDA events) and reports a checksum of the trees, which
must be the same for all of them."""
import glob, os, subprocess, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "profiles", "tools", "_variants")


def build(name, flags):
    from theta_rrt_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    so = os.path.join(VDIR, name + ".so")
    cmd = [B.nvcc_path(), *B.NVCC_FLAGS, "-Xptxas", "-v", *flags, "-o", so, os.path.join(B.CSRC, "thetarrt.cu")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode:
        print(out.stderr[-3000:]); sys.exit(1)
    lines = out.stderr.splitlines()
    for i, l in enumerate(lines):
        if "rrt_kernel_specILi32" in l and "Compiling" in l:
            print(name, "|", lines[i + 2].strip(), "|", lines[i + 3].strip())
            break


def one(so, nq, K):
    import numpy as np, torch
    from theta_rrt_b200 import _lib
    _lib.SO_PATH = so
    import ctypes
    probe = ctypes.CDLL(so)  # older builds lack the newer entry points: bind only what the variant exports
    _lib.SIGNATURES = {k: v for k, v in _lib.SIGNATURES.items() if hasattr(probe, k)}
    import bench
    from theta_rrt_b200 import OccupancyGrid, Params, Planner
    dev = torch.device("cuda:0")
    free = bench.load_maps()["map1"]
    cache = f"/tmp/wl_{nq}_{K}.npz"
    if os.path.exists(cache):
        z = np.load(cache); w = [z[k] for k in ("a", "b", "c", "d")]
    else:
        w = bench.make_rrt_workload(free, nq, K)
        np.savez(cache, a=w[0], b=w[1], c=w[2], d=w[3])
    p = Planner(OccupancyGrid(free, device=dev), Params(tol_xy=0.0, K=K))
    d = [torch.from_numpy(a).to(dev) for a in w]
    for _ in range(3):
        r = p.rrt(*d, K=K)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record(); r = p.rrt(*d, K=K); b.record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    h = r.host()
    n = h["n_nodes"]
    crc = zlib.crc32(n.tobytes())
    for q in range(0, nq, 37):
        crc = zlib.crc32(h["parent"][q, :n[q]].tobytes(), crc)
        crc = zlib.crc32(h["node_x"][q, :n[q]].tobytes(), crc)
        crc = zlib.crc32(h["u"][q, 1:n[q]].tobytes(), crc)
    print(f"{os.path.basename(so):28s} ms min {min(ms):7.2f} mean {sum(ms) / len(ms):7.2f}  iters {int(h['iters'].sum())} crc {crc:08x}", flush=True)


def one_los(so):
    """cfg-4 rays (2^20 on the 8192^2 grid) through the strip kernel of one build."""
    import numpy as np, torch
    from theta_rrt_b200 import _lib
    _lib.SO_PATH = so
    import bench
    from theta_rrt_b200 import OccupancyGrid, Planner
    dev = torch.device("cuda:0")
    big = bench.synthetic_map(8192, 0.1, 8, 42)
    pl = Planner(OccupancyGrid(big, device=dev))
    seg = torch.from_numpy(bench.make_segments(big, 1 << 20, 7)).to(dev)
    out = torch.empty(seg.shape[0], dtype=torch.uint8, device=dev)
    for _ in range(3):
        pl.los(seg, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        pl.los(seg, out=out)
    b.record(); torch.cuda.synchronize()
    print(f"{os.path.basename(so):28s} los {a.elapsed_time(b) / 20 * 1e3:7.1f} us  visible {int(out.sum())} crc {zlib.crc32(out.cpu().numpy().tobytes()):08x}", flush=True)


def one_theta(so):
    """Theta* on map2: the single query of main.py:57 and the bench batch of 8 192 random free-cell queries."""
    import numpy as np, torch
    from theta_rrt_b200 import _lib
    _lib.SO_PATH = so
    import bench
    from theta_rrt_b200 import OccupancyGrid, Planner
    dev = torch.device("cuda:0")
    m2 = bench.load_maps()["map2"]
    pt = Planner(OccupancyGrid(m2, device=dev))
    one_q = torch.tensor([[280, 0, 8, 280]], dtype=torch.int32, device=dev)
    cells = np.argwhere(m2)
    rq = np.random.default_rng(5)
    a, b = cells[rq.integers(len(cells), size=8192)], cells[rq.integers(len(cells), size=8192)]
    sg = torch.from_numpy(np.stack([a[:, 1], a[:, 0], b[:, 1], b[:, 0]], 1).astype(np.int32)).to(dev)

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x.record()
        for _ in range(n):
            fn()
        y.record(); torch.cuda.synchronize()
        return x.elapsed_time(y) / n
    ms1 = timed(lambda: pt.theta(one_q, lanes=32), 3)
    msb = timed(lambda: pt.theta(sg, path_cap=64), 3)
    r = pt.theta(sg, path_cap=64).host()
    crc = 0
    for k in ("status", "expanded", "path_len", "cost", "n_los", "pushes", "path"):
        if k in r and r[k] is not None:
            crc = zlib.crc32(np.ascontiguousarray(r[k]).tobytes(), crc)
    print(f"{os.path.basename(so):28s} theta single {ms1:7.2f} ms  batch {msb:7.2f} ms  {r['expanded'].sum() / msb / 1e3:6.1f} M exp/s  crc {crc:08x}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "one_los":
        one_los(sys.argv[2])
    elif sys.argv[1] == "one_theta":
        one_theta(sys.argv[2])
    elif sys.argv[1] == "theta":
        for n in sys.argv[2:]:
            subprocess.run([sys.executable, os.path.abspath(__file__), "one_theta", os.path.join(VDIR, n + ".so")])
    elif sys.argv[1] == "los":
        for n in sys.argv[2:]:
            subprocess.run([sys.executable, os.path.abspath(__file__), "one_los", os.path.join(VDIR, n + ".so")])
    elif sys.argv[1] == "build":
        build(sys.argv[2], sys.argv[3:])
    elif sys.argv[1] == "one":
        one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    else:
        names = sys.argv[2:]
        nq, K = 4096, 5001
        sos = [os.path.join(VDIR, n + ".so") for n in names] if names else sorted(glob.glob(os.path.join(VDIR, "*.so")))
        for so in sos:
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", so, str(nq), str(K)])
