#!/usr/bin/env python
"""bench.py -- theta-rrt planning inner loop on B200: RRT expansions/s (+ LOS checks/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (BASELINE.json configs[2], SURVEY.md 8d cfg 3): batched RRT, 4096 independent
queries per GPU on map1 (100x100), K=5001 (5000 expansions each), tol_xy=0 so every query runs all
iterations, start/goal uniform over free cells (default_rng(1234)), per-query sample stream equal to
np.random.seed(q) + 5000 x rand_conf(goal_q).  One "step" = one fused-kernel pass over the batch.

The single JSON line carries
  value      RRT expansions/s with all inputs resident in HBM (CUDA events on the launch stream)
  e2e        the same metric through the public host-buffer API (pinned H2D of the sample streams,
             kernel, D2H of the trees) inside the timed region
  roofline   the fused rrt kernel: algorithmic bytes (16 B per scanned (query,node) pair + 16 B per
             sample + tree writes) / kernel time, against the measured HBM copy peak
  secondary  cfg-4 microbenchmarks (LOS checks/s over an 8192^2 grid, nearest-node scan over a
             2^20-node tree, each with its own roofline) and Theta* on map2
  cpu_baseline  the C oracle (a port of the reference) on the host cores, bounded sample

--impl reference times the reference's algorithm on the host CPU (the C oracle port, all host
threads): the reference itself is pure Python living outside this repository, so it cannot run on
the GPU box.  Multi-GPU (torchrun, one rank per GPU): queries are sharded by rank, no data-path
collective; a gather of the per-query summaries over NCCL closes each step ("scaling": "weak").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NQ_PER_GPU = 4096
K_RRT = 5001
MAP_SEED = 1234
WORKLOAD = ("cfg3: batched RRT, %d independent queries per GPU on map1.png (100x100), K=%d "
            "(5000 expansions each), tol_xy=0, seeded rand_conf streams")


# --------------------------------------------------------------------------- workload
def load_maps():
    z = np.load(os.path.join(ROOT, "tests", "golden", "maps.npz"))
    return {k: z[k].astype(bool) for k in z.files}


def random_queries(free, nq, seed, offset=0):
    rng = np.random.default_rng(seed)
    cells = np.argwhere(free)
    n_all = offset + nq
    a = cells[rng.integers(len(cells), size=n_all)]
    b = cells[rng.integers(len(cells), size=n_all)]
    hs, hg = rng.uniform(-180, 180, n_all), rng.uniform(-180, 180, n_all)
    starts = np.stack([a[:, 1], a[:, 0], hs], axis=1).astype(np.float64)[offset:]
    goals = np.stack([b[:, 1], b[:, 0], hg], axis=1).astype(np.float64)[offset:]
    return starts, goals


def make_rrt_workload(free, nq, K, first_query=0):
    from theta_rrt_b200 import samples
    starts, goals = random_queries(free, nq, MAP_SEED, offset=first_query)
    sxy = np.empty((nq, K - 1, 2), np.int32)
    sth = np.empty((nq, K - 1), np.float64)
    for q in range(nq):
        g = ((goals[q, 0], goals[q, 1]), goals[q, 2])
        sxy[q], sth[q] = samples.make_stream(g, K - 1, first_query + q, free.shape)
    return starts, goals, sxy, sth


def synthetic_map(n, p, block, seed):
    rng = np.random.default_rng(seed)
    nb = (n + block - 1) // block
    coarse = rng.random((nb, nb)) >= p
    return np.kron(coarse, np.ones((block, block), bool))[:n, :n]


def make_segments(free, n, seed, maxlen=512):
    """cfg 4 raycast input: n segments from a random free cell, length uniform [1, maxlen] px, uniform direction."""
    rng = np.random.default_rng(seed)
    side = free.shape[0]
    cells = np.argwhere(free)
    c = cells[rng.integers(len(cells), size=n)]
    length = rng.uniform(1, maxlen, n)
    ang = rng.uniform(0, 2 * np.pi, n)
    x0, y0 = c[:, 1], c[:, 0]
    x1 = np.clip((x0 + length * np.cos(ang)).astype(np.int64), 0, side - 1)
    y1 = np.clip((y0 + length * np.sin(ang)).astype(np.int64), 0, side - 1)
    return np.stack([x0, y0, x1, y1], axis=1).astype(np.int32)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if t0 is not None and (ts < t0 - 0.1 or ts > t1 + 0.3):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_traffic(kernel):
    """Per-launch DRAM traffic of a kernel from the committed ncu capture summary, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            rec = json.load(open(p)).get(kernel)
            return rec["bytes"] if isinstance(rec, dict) else rec
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------- reference arm (CPU)
def run_reference_arm(args, rank, world):
    """The reference's algorithm on the host CPU: C oracle port, all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import c_oracle as O
    free = load_maps()["map1"]
    cores = os.cpu_count() or 1
    nq = max(4 * cores, 32)  # ~0.2 s of one core per query, four queries per thread: a step is about a second of wall
    starts, goals, sxy, sth = make_rrt_workload(free, nq, K_RRT)
    P = O.Params(tol_xy=0.0)
    O.lib()
    for _ in range(max(args.warmup, 1)):
        O.rrt_batch(free, starts[:cores], goals[:cores], sxy[:cores], sth[:cores], K_RRT, P, threads=cores, want_nodes=False)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        r = O.rrt_batch(free, starts, goals, sxy, sth, K_RRT, P, threads=cores, want_nodes=True)
        done += int(r["iters"].sum())  # executed loop iterations (queries the reference would abort stop early)
    dt = time.perf_counter() - t0
    value = done / dt
    sample = f"{nq} of the {NQ_PER_GPU} queries per step, K={K_RRT}, {cores} threads"
    line = {"impl": "reference", "metric": "rrt_expansions_per_sec", "value": value, "unit": "expansions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % (NQ_PER_GPU, K_RRT), "queries_per_gpu": NQ_PER_GPU, "K": K_RRT,
                       "sample_queries_per_step": nq, "host_threads": cores},
            "cpu_baseline": {"value": value, "unit": "expansions/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "expansions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference is pure Python outside the repo; this arm is its C restatement (oracle/) on the host"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def time_steps(torch, fn, steps, warmup, dist_on):
    """W warm-up calls, then `steps` timed calls with per-step CUDA events on the current stream."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
        torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.time()
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t1 = time.time()
    if dist_on:
        dist.barrier()
        torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    return ms, t0, t1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=NQ_PER_GPU, help="queries per GPU (default: the cfg-3 size)")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per query of the fused kernel (0 = auto)")
    ap.add_argument("--schedule", type=int, default=0, help="0 = speculative window (default), 1 = cooperative")
    ap.add_argument("--valid-rows-d2h", action="store_true",
                    help="e2e: pack the tree rows that exist on the device and fetch only those (Planner.rrt_host "
                         "valid_rows_only); the default from 4 ranks on, where the host link is the limit")
    ap.add_argument("--dense-d2h", action="store_true", help="e2e: always copy the full [q][K] tree arrays")
    ap.add_argument("--chunks", type=int, default=16, help="pieces of the e2e host-buffer pipeline (Planner.rrt_host)")
    ap.add_argument("--skip-secondary", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from theta_rrt_b200 import OccupancyGrid, Params, Planner, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    # one process per GPU: stay on the GPU's NUMA node before any pinned buffer exists (not at N = 1, where the CPU
    # baseline legs of this process use every host core)
    numa = shard.bind_host_to_device(local_rank) if (dist_on and not os.environ.get("TRRT_NO_NUMA_BIND")) else None
    if dist_on:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    maps = load_maps()
    free = maps["map1"]
    nq = args.queries
    K = K_RRT
    total_q = nq * world
    lo, hi = shard.shard_range(total_q, rank, world)  # contiguous block of query ids of this rank
    starts, goals, sxy, sth = make_rrt_workload(free, hi - lo, K, first_query=lo)
    planner = Planner(OccupancyGrid(free, device=dev), Params(tol_xy=0.0, K=K))

    # ---- device-resident inputs (value) and pinned host inputs (e2e)
    h_in = [torch.from_numpy(a).pin_memory() for a in (starts, goals, sxy, sth)]
    d_in = [t.to(dev) for t in h_in]
    torch.cuda.synchronize()

    # one untimed instrumented run: counters for the algorithmic-byte accounting
    r0 = planner.rrt(*d_in, K=K, counters=True, lanes=args.lanes, schedule=args.schedule)
    torch.cuda.synchronize()
    counters = r0.counters.sum(dim=0).cpu().numpy().astype(np.int64)
    n_nodes_total = int(r0.n_nodes.sum().item())
    iters_total = int(r0.iters.sum().item())
    # tol_xy=0 disables the goal test, but a query still ends early when the reference itself would raise
    # (quirk Q7: rrt.py:170-171 calls drive() on a straight-line steer -> TypeError); such queries are
    # reported with status 4 and only their executed iterations count as work
    status_bad = int((r0.status > 1).sum().item())
    del r0

    launches = {"n": 0}
    keep = {}

    def step_resident():
        keep["res"] = planner.rrt(*d_in, K=K, lanes=args.lanes, schedule=args.schedule)
        launches["n"] += 1
        if dist_on:  # optional gather of the per-query summaries
            rec = torch.stack([keep["res"].n_nodes, keep["res"].sol, keep["res"].status], dim=1)
            keep["gathered"] = shard.gather_records(rec, total_q)

    clocks = ClockSampler(local_rank)
    clocks.start()
    ms, t0, t1 = time_steps(torch, step_resident, args.steps, args.warmup, dist_on)
    clk = clocks.stop(t0, t1)
    launches_timed = args.steps
    total_ms = float(sum(ms))
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    it_all = torch.tensor([iters_total, n_nodes_total, status_bad], dtype=torch.int64, device=dev)
    if dist_on:
        dist.all_reduce(it_all, op=dist.ReduceOp.SUM)
    expansions_per_step_all = int(it_all[0].item())  # loop iterations actually executed (rrt.py:141)
    value = expansions_per_step_all * args.steps / (total_ms_max / 1e3)

    # ---- roofline of the fused kernel (this rank's launch)
    alg_bytes = 16 * int(counters[0]) + 16 * iters_total + (28 + 40) * n_nodes_total
    ms_kernel = total_ms / args.steps
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (ms_kernel / 1e3) / 1e9
    roofline = {"kernel": "rrt_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": profile_traffic("rrt_kernel"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "pairs_scanned_per_launch": int(counters[0]),
                "note": "16 B per scanned (query,node) pair + 16 B per sample + 68 B per inserted node; trees of "
                        "<= 80 KB per query are L2/L1 resident, so frac may exceed what DRAM traffic alone would give"}

    # ---- e2e through the host-buffer API: pinned H2D of the inputs, kernel, D2H of the trees
    nql = hi - lo
    h_out = {k: torch.empty(s, dtype=dt).pin_memory() for k, (s, dt) in {
        "node_x": ((nql, K), torch.float64), "node_y": ((nql, K), torch.float64), "node_theta": ((nql, K), torch.float64),
        "parent": ((nql, K), torch.int32), "u": ((nql, K, 5), torch.float64), "n_nodes": ((nql,), torch.int32),
        "sol": ((nql,), torch.int32), "status": ((nql,), torch.int32), "iters": ((nql,), torch.int32)}.items()}
    h2d = sum(t.numel() * t.element_size() for t in h_in)
    d2h_dense = sum(t.numel() * t.element_size() for t in h_out.values())
    # measured on this pool's boxes: dense 51 ms per step on one GPU against 68 ms packed; 4 GPUs 89 ms dense, 75 ms packed
    vro = args.valid_rows_d2h or (world >= 4 and not args.dense_d2h)
    # valid_rows_only: the rows that exist (68 B per node) are packed on the device and fetched with linear copies; the
    # per-query scalars and the row index travel whole
    if vro:
        h_out["row_start"] = torch.empty(nql, dtype=torch.int64).pin_memory()
    d2h = (n_nodes_total * 68 + sum(h_out[k].numel() * h_out[k].element_size() for k in ("n_nodes", "sol", "status", "iters", "row_start"))
           if vro else d2h_dense)

    def step_e2e():
        # public host-buffer API: pinned inputs in, pinned trees out, transfers of one piece overlap the others' kernels
        planner.rrt_host(*h_in, out=h_out, K=K, chunks=args.chunks, lanes=args.lanes, schedule=args.schedule, valid_rows_only=vro)

    ms_e2e, _, _ = time_steps(torch, step_e2e, args.steps, 1, dist_on)  # every step waited for before the next starts
    serial_ms = float(sum(ms_e2e)) / args.steps

    # the same steps as a stream of batches: piece c of step i+1 queues behind piece c of step i, so the copies of one
    # step also overlap the planning of the next; the timed region still contains every step's H2D and D2H
    def step_e2e_streamed():
        planner.rrt_host(*h_in, out=h_out, K=K, chunks=args.chunks, wait=False, lanes=args.lanes, schedule=args.schedule,
                         valid_rows_only=vro)

    step_e2e_streamed(); planner.host_sync(); torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
        torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(args.steps):
        step_e2e_streamed()
    planner.host_sync()
    eb.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
        torch.cuda.synchronize()
    te = torch.tensor([float(ea.elapsed_time(eb))], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = expansions_per_step_all * args.steps / (float(te.item()) / 1e3)
    launches_e2e = args.steps * args.chunks * (2 if vro else 1)  # fused kernel (+ pack kernel) per piece
    assert int(h_out["n_nodes"].sum()) == n_nodes_total  # the host really received this step's trees
    if vro:  # ... and the rows themselves: the last row of every tree against the resident result
        last_dev = (keep["res"].n_nodes.long() - 1).clamp(min=0)
        last_host = h_out["row_start"] + last_dev.cpu()
        rows = torch.arange(nql, device=dev)
        assert torch.equal(h_out["node_x"].view(-1)[last_host], keep["res"].node_x[rows, last_dev].cpu())
        assert torch.equal(h_out["parent"].view(-1)[last_host], keep["res"].parent[rows, last_dev].cpu())
    keep.clear()
    del h_out
    torch.cuda.empty_cache()

    line = {"metric": "rrt_expansions_per_sec", "value": value, "unit": "expansions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % (nq, K),
                       "queries_per_gpu": nq, "K": K, "lanes_per_query": args.lanes or 32,
                       "schedule": "speculative window" if args.schedule == 0 else "cooperative",
                       "parallelism": "query-sharded x%d, no data-path collective" % world,
                       "l2_policy": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2" %
                                    ((h2d + d2h) / 1e9)},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "expansions/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": float(te.item()) / args.steps,
                    "api": "Planner.rrt_host(pinned inputs, pinned outputs, chunks=%d, wait=False%s) per step, "
                           "host_sync() after the last step" % (args.chunks, ", valid_rows_only=True" if vro else ""),
                    "d2h_dense_bytes_per_step": int(d2h_dense),
                    "serial_ms_per_step": serial_ms, "host_numa_binding": numa,
                    "note": "steps are streamed: the transfers of a step overlap the planning of its neighbours; "
                            "serial_ms_per_step is the same call with every step completed before the next starts"},
            "gpu_launches": launches_timed, "gpu_launches_e2e": launches_e2e,
            "roofline": roofline,
            "expansions_per_step": expansions_per_step_all,
            "accepted_nodes_per_sec": (int(it_all[1].item()) - total_q) * args.steps / (total_ms_max / 1e3),
            "queries_ended_by_reference_TypeError": int(it_all[2].item())}

    if rank == 0 and not args.skip_secondary:
        line["secondary"] = secondary_benchmarks(torch, dev, maps, peak, peak_src, args)
    if rank == 0 and world == 1 and not args.skip_cpu:
        line["cpu_baseline"] = cpu_baseline(free, starts, goals, sxy, sth)
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def secondary_benchmarks(torch, dev, maps, peak, peak_src, args):
    """cfg 4 microbenchmarks + Theta*; each timed with CUDA events after warm-up."""
    from theta_rrt_b200 import OccupancyGrid, Planner
    out = {}
    steps = max(args.steps, 5)

    def timed(fn, n=steps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    # ---- cfg 4: 8192^2 grid, 2^20 segments, 2^20-node tree, 4096 queries
    big = synthetic_map(8192, 0.1, 8, 42)
    pl = Planner(OccupancyGrid(big, device=dev))
    seg = make_segments(big, 1 << 20, 7)
    d_seg = torch.from_numpy(seg).to(dev)
    d_out = torch.empty(len(seg), dtype=torch.uint8, device=dev)
    ms_rows = timed(lambda: pl.los(d_seg, out=d_out, layout="rows"))
    vis_rows = d_out.cpu().numpy().astype(bool)
    ms = timed(lambda: pl.los(d_seg, out=d_out))
    vis = d_out.cpu().numpy().astype(bool)
    assert np.array_equal(vis, vis_rows), "los: strip and row layouts disagree"
    px = np.maximum(np.abs(seg[:, 2] - seg[:, 0]), np.abs(seg[:, 3] - seg[:, 1])) + 1
    los_bytes = 4.0 * float(px[vis].sum()) + 17.0 * len(seg)
    out["los_cfg4"] = {"metric": "los_checks_per_sec", "value": len(seg) / (ms / 1e3), "unit": "checks/s", "ms": ms,
                       "segments": len(seg), "grid": "8192x8192: strip copy (32 MiB, L2 resident) of the bit-packed rows (8 MiB)",
                       "visible_fraction": float(vis.mean()),
                       "pixel_tests_per_sec_upper": float(px.sum()) / (ms / 1e3),
                       "rows_layout_ms": ms_rows, "rows_layout_checks_per_sec": len(seg) / (ms_rows / 1e3),
                       "roofline": {"kernel": "los_tiled_kernel", "bound": "hbm",
                                    "achieved": los_bytes / (ms / 1e3) / 1e9,
                                    "peak": peak, "unit": "GB/s",
                                    "frac": los_bytes / (ms / 1e3) / 1e9 / peak,
                                    "traffic": profile_traffic("los_tiled_kernel"), "peak_source": peak_src,
                                    "note": "4 B word per pixel test of fully walked (visible) rays + 16 B segment in + "
                                            "1 B out; blocked rays stop early so their tests are not counted; the kernel "
                                            "is bound by the integer pipes (ncu), not by HBM"}}
    rng = np.random.default_rng(3)
    n_nodes = 1 << 20
    x = torch.from_numpy(rng.uniform(0, 8191, n_nodes)).to(dev)
    y = torch.from_numpy(rng.uniform(0, 8191, n_nodes)).to(dev)
    qxy = torch.from_numpy(rng.integers(0, 8192, size=(4096, 2)).astype(np.int32)).to(dev)
    ms = timed(lambda: pl.nearest(x, y, qxy))
    ach = 16.0 * n_nodes * 4096 / (ms / 1e3) / 1e9
    out["nearest_cfg4"] = {"metric": "nearest_queries_per_sec", "value": 4096 / (ms / 1e3), "unit": "queries/s",
                           "ms": ms, "nodes": n_nodes, "queries": 4096,
                           "roofline": {"kernel": "nearest_tile_kernel", "bound": "hbm", "achieved": ach, "peak": peak,
                                        "unit": "GB/s", "frac": ach / peak, "traffic": profile_traffic("nearest_tile_kernel"),
                                        "peak_source": peak_src,
                                        "note": "algorithmic 16 B per (query,node); 8 queries share each node load "
                                                "(register tiling), so DRAM traffic is far below the algorithmic bytes "
                                                "and frac > 1 is expected: the kernel is fp64-issue bound"}}
    # single-query scans over a tree larger than L2 (2^24 nodes = 256 MiB): the pure HBM-streaming case
    n_big = 1 << 24
    xb = torch.from_numpy(rng.uniform(0, 8191, n_big)).to(dev)
    yb = torch.from_numpy(rng.uniform(0, 8191, n_big)).to(dev)
    q1 = qxy[:1].contiguous()
    ms = timed(lambda: pl.nearest(xb, yb, q1))
    ach = 16.0 * n_big / (ms / 1e3) / 1e9
    out["nearest_single_query_hbm"] = {"metric": "nearest_scan_GBps", "value": ach, "unit": "GB/s", "ms": ms,
                                       "nodes": n_big, "queries": 1,
                                       "roofline": {"kernel": "nearest_tile_kernel", "bound": "hbm", "achieved": ach,
                                                    "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                                    "peak_source": peak_src,
                                                    "note": "one query, 256 MiB SoA tree (> L2): every byte comes from HBM"}}
    del xb, yb, x, y
    if not args.skip_cpu:
        # CPU port (oracle) on a bounded sample of the same rays, all host threads
        from oracle import c_oracle as O
        cores = os.cpu_count() or 1
        O.lib()
        ns = 1 << 17
        O.lineofsight_batch(big, seg[:4096], threads=cores)
        t = time.perf_counter()
        cpu_vis = O.lineofsight_batch(big, seg[:ns], threads=cores)
        dt = time.perf_counter() - t
        assert np.array_equal(cpu_vis, vis[:ns])  # same booleans as the kernel
        out["los_cfg4"]["cpu_baseline"] = {"value": ns / dt, "unit": "checks/s", "cores": cores, "kind": "port",
                                           "sample": f"first {ns} of the {len(seg)} rays on {cores} threads"}
    # ---- Theta* on map2: the reference's single query and a batch of random free-cell queries
    m2 = maps["map2"]
    pt = Planner(OccupancyGrid(m2, device=dev))
    one = torch.tensor([[280, 0, 8, 280]], dtype=torch.int32, device=dev)
    ms1 = timed(lambda: pt.theta(one, lanes=32), n=3, warm=1)
    r = pt.theta(one, lanes=32).host()
    out["theta_cfg2"] = {"metric": "theta_single_query_ms", "value": ms1, "unit": "ms", "expanded": int(r["expanded"][0]),
                         "los_checks": int(r["n_los"][0]), "cost": float(r["cost"][0]),
                         "expansions_per_sec": int(r["expanded"][0]) / (ms1 / 1e3)}
    cells = np.argwhere(m2)
    rq = np.random.default_rng(5)
    nqt = 8192
    a, b = cells[rq.integers(len(cells), size=nqt)], cells[rq.integers(len(cells), size=nqt)]
    sg = torch.from_numpy(np.stack([a[:, 1], a[:, 0], b[:, 1], b[:, 0]], 1).astype(np.int32)).to(dev)
    msb = timed(lambda: pt.theta(sg, path_cap=64), n=3, warm=1)
    rb = pt.theta(sg, path_cap=64).host()
    out["theta_batch_map2"] = {"metric": "theta_expansions_per_sec", "value": float(rb["expanded"].sum()) / (msb / 1e3),
                               "unit": "expansions/s", "ms": msb, "queries": nqt,
                               "los_checks_per_sec": float(rb["n_los"].sum()) / (msb / 1e3),
                               "found": int((rb["status"] == 0).sum())}
    # ---- cfg 5, the share of one GPU of an 8-GPU box: 4096 RRT queries (K = 1001) + 4096 Theta* queries, each bound
    # to one of 64 random 256x256 maps (4x4-block Bernoulli obstacles, p = 0.15, default_rng(7 + map))
    from theta_rrt_b200 import Params, samples
    n_maps5, side5, nq5, K5 = 64, 256, 4096, 1001
    maps5 = np.stack([synthetic_map(side5, 0.15, 4, 7 + m) for m in range(n_maps5)])
    p5 = Planner(OccupancyGrid(maps5, device=dev), Params(tol_xy=0.0, K=K5))
    r5 = np.random.default_rng(77)
    mid_r = r5.integers(0, n_maps5, nq5).astype(np.int32)
    mid_t = r5.integers(0, n_maps5, nq5).astype(np.int32)
    free_cells = [np.argwhere(m) for m in maps5]
    pick = lambda mids: np.stack([free_cells[m][r5.integers(len(free_cells[m]))] for m in mids])
    a5, b5 = pick(mid_r), pick(mid_r)
    starts5 = np.stack([a5[:, 1], a5[:, 0], r5.uniform(-180, 180, nq5)], 1).astype(np.float64)
    goals5 = np.stack([b5[:, 1], b5[:, 0], r5.uniform(-180, 180, nq5)], 1).astype(np.float64)
    sxy5 = np.empty((nq5, K5 - 1, 2), np.int32); sth5 = np.empty((nq5, K5 - 1))
    for q in range(nq5):
        sxy5[q], sth5[q] = samples.make_stream(((goals5[q, 0], goals5[q, 1]), goals5[q, 2]), K5 - 1, 500000 + q, (side5, side5))
    ta, tb = pick(mid_t), pick(mid_t)
    sg5 = torch.from_numpy(np.stack([ta[:, 1], ta[:, 0], tb[:, 1], tb[:, 0]], 1).astype(np.int32)).to(dev)
    d5 = [torch.from_numpy(v).to(dev) for v in (starts5, goals5, sxy5, sth5)]
    dm_r, dm_t = torch.from_numpy(mid_r).to(dev), torch.from_numpy(mid_t).to(dev)
    res5 = {}

    def step5():
        res5["rrt"] = p5.rrt(*d5, K=K5, map_id=dm_r)
        res5["theta"] = p5.theta(sg5, map_id=dm_t, path_cap=64)
    ms5 = timed(step5, n=3, warm=1)
    it5, ex5 = int(res5["rrt"].iters.sum()), int(res5["theta"].expanded.sum())
    out["cfg5_mixed_share_of_one_gpu"] = {
        "metric": "mixed_queries_per_sec", "value": 2 * nq5 / (ms5 / 1e3), "unit": "queries/s", "ms": ms5,
        "rrt_queries": nq5, "K": K5, "theta_queries": nq5, "maps": f"{n_maps5} x {side5}x{side5}",
        "rrt_expansions": it5, "theta_expansions": ex5, "expansions_per_sec": (it5 + ex5) / (ms5 / 1e3),
        "theta_found": int((res5["theta"].status == 0).sum()),
        "note": "1/8 of BASELINE cfg 5 (65536 queries over 8 GPUs), the RRT batch and the Theta* batch back to back"}
    if not args.skip_cpu:
        from oracle import c_oracle as O
        cores = os.cpu_count() or 1
        nsmp = min(nqt, 4 * cores)
        sgh = sg[:nsmp].cpu().numpy()
        t = time.perf_counter()
        rc = O.astar_batch(m2, sgh, thetastar=True, threads=cores)
        dt = time.perf_counter() - t
        assert np.array_equal(rc["expanded"], rb["expanded"][:nsmp])  # same searches as the kernel
        out["theta_batch_map2"]["cpu_baseline"] = {"value": float(rc["expanded"].sum()) / dt, "unit": "expansions/s",
                                                   "cores": cores, "kind": "port",
                                                   "sample": f"first {nsmp} of the {nqt} queries on {cores} threads"}
    return out


def cpu_baseline(free, starts, goals, sxy, sth):
    """The C oracle (a port of the reference's algorithm) on the host: one core and all cores, bounded sample."""
    from oracle import c_oracle as O
    cores = os.cpu_count() or 1
    P = O.Params(tol_xy=0.0)
    K = sth.shape[1] + 1
    O.lib()
    n1 = min(8, len(starts))
    t = time.perf_counter()
    r = O.rrt_batch(free, starts[:n1], goals[:n1], sxy[:n1], sth[:n1], K, P, threads=1, want_nodes=False)
    one = int(r["iters"].sum()) / (time.perf_counter() - t)
    nall = min(len(starts), max(4 * cores, 32))
    t = time.perf_counter()
    r = O.rrt_batch(free, starts[:nall], goals[:nall], sxy[:nall], sth[:nall], K, P, threads=cores, want_nodes=False)
    allc = int(r["iters"].sum()) / (time.perf_counter() - t)
    return {"value": allc, "unit": "expansions/s", "cores": cores, "kind": "port",
            "sample": f"first {nall} of the step's queries on {cores} threads (single thread: first {n1} queries)",
            "single_core_value": one,
            "python_reference_note": "the unmodified Python reference ran at 214 expansions/s on one core in the build "
                                     "container (BASELINE.md); it is not present on the GPU box"}


if __name__ == "__main__":
    main()
