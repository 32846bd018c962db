"""Times the UNMODIFIED Python reference (rrt.rrt, rrt.py:130-206) on host cores: one process, and one process per
core.  TEST / BASELINE INFRASTRUCTURE ONLY -- used by bench.py's CPU legs.

The reference files come from /root/reference (build container) or from the git-ignored copy oracle/_ref made by
oracle/make_ref.py (GPU box).  Each worker process imports the reference once (matplotlib stubbed, `search` first),
sets builtins.imarray and the parameters of main.py:15-32, replaces rrt.rand_conf by the injected sample stream and
times rrt.rrt(start, goal) on its queries.  A query the reference aborts with its own TypeError (quirk Q7) counts the
iterations it executed.
"""
from __future__ import annotations

import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_dir():
    for d in (os.environ.get("THETA_RRT_REFERENCE", "/root/reference"), os.path.join(HERE, "_ref")):
        if os.path.isfile(os.path.join(d, "search.py")):
            return d
    return None


def _worker(args):
    free, jobs, K, tol_xy = args
    os.environ.setdefault("THETA_RRT_REFERENCE", reference_dir() or "")
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import live_reference as L
    L.REFERENCE_DIR = reference_dir()
    search, rrt, _ = L.load()
    L.set_map(free)
    done, t0 = 0, time.perf_counter()
    for start, goal, sxy, sth in jobs:
        L.set_params(K=K, tol_xy=tol_xy)
        count = [0]
        it = iter(zip(sxy.tolist(), sth.tolist()))

        def rand_conf(goal_):
            count[0] += 1
            (x, y), th = next(it)
            return ((x, y), th)
        orig = rrt.rand_conf
        rrt.rand_conf = rand_conf
        try:
            with L.quiet():
                try:
                    rrt.rrt(((start[0], start[1]), start[2]), ((goal[0], goal[1]), goal[2]), debug=False)
                    done += count[0]
                except TypeError:  # Q7: the iteration that raises is not counted
                    done += count[0] - 1
        finally:
            rrt.rand_conf = orig
            L.set_params()
    return done, time.perf_counter() - t0


def time_reference(free, starts, goals, sxy, sth, K, procs, per_proc=1, tol_xy=0.0):
    """Expansions/s of the Python reference with `procs` worker processes, `per_proc` queries each (the first K-1
    samples of each stream).  Workers are plain subprocesses of this file (job in a temporary .npz, result on stdout).
    Returns (rate, iterations, wall seconds) or None when the reference files are absent."""
    if reference_dir() is None:
        return None
    import json
    import subprocess
    import tempfile
    import numpy as np
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        procs_ = []
        for p in range(procs):
            qs = slice(p * per_proc, (p + 1) * per_proc)
            path = os.path.join(tmp, f"job{p}.npz")
            np.savez(path, free=np.asarray(free, bool), starts=np.asarray(starts)[qs], goals=np.asarray(goals)[qs],
                     sxy=np.asarray(sxy)[qs, :K - 1], sth=np.asarray(sth)[qs, :K - 1])
            procs_.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "worker", path, str(K), repr(float(tol_xy))],
                                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        res = []
        for pr in procs_:
            out, err = pr.communicate()
            if pr.returncode != 0:
                raise RuntimeError("python reference worker failed: " + err[-2000:])
            res.append(json.loads(out.strip().splitlines()[-1]))
    wall = time.perf_counter() - t0
    iters = sum(r["iters"] for r in res)
    busy = max(r["seconds"] for r in res)  # the slowest worker's time inside rrt.rrt (interpreter start and imports excluded)
    return iters / busy, iters, wall


if __name__ == "__main__" and len(sys.argv) >= 5 and sys.argv[1] == "worker":
    import json
    import numpy as np
    z = np.load(sys.argv[2])
    jobs = [(z["starts"][i], z["goals"][i], z["sxy"][i], z["sth"][i]) for i in range(len(z["starts"]))]
    it, sec = _worker((z["free"], jobs, int(sys.argv[3]), float(sys.argv[4])))
    print(json.dumps({"iters": it, "seconds": sec}))
