"""Generates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Outputs (committed):
    maps.npz            map1 / map2 / blank as np.array(Image.open(p).convert('1'))  (main.py:38-42)
    theta_kat.json      search.astar known answers (SURVEY.md Appendix B + main.py:57)
    los_kat.npz         search.lineofsight booleans for random segments on map1 / map2
    circle_kat.json     search.getCircle pixel sets
    rrt_kat.npz         rrt.rrt runs with injected seeded streams: parents, nearest indices,
                        LOS booleans, node coordinates, steer / drive / getArc records
    stream_kat.npz      rand_conf streams for fixed seeds
The float arrays are what THIS container's numpy/scipy/glibc produced; tests compare
them with a tolerance (discrete arrays are compared exactly).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import live_reference as L  # noqa: E402

REF = "/root/reference"


def main():
    search, rrt, _ = L.load()
    maps = {n: L.load_png(os.path.join(REF, n + ".png")) for n in ("map1", "map2", "blank")}
    np.savez_compressed(os.path.join(HERE, "maps.npz"), **{k: v.astype(np.uint8) for k, v in maps.items()})

    # ---------------- Theta* / A* known answers
    kat = []
    cases = [("map2", (280, 0), (8, 280), True), ("map2", (8, 280), (280, 0), True), ("map2", (280, 0), (8, 280), False),
             ("map1", (5, 5), (90, 50), True), ("map1", (5, 5), (90, 50), False), ("map1", (2, 97), (97, 2), True),
             ("map1", (20, 20), (50, 50), True), ("map1", (50, 50), (80, 80), True), ("map1", (5, 5), (150, 50), True),
             ("map1", (5, 5), (5, 5), True), ("map1", (-1, 5), (5, 5), True), ("blank", (0, 0), (299, 150), True),
             ("blank", (299, 299), (0, 0), False), ("map1", (5, 5), (6, 6), True)]
    # a blocked endpoint on map1
    blocked = np.argwhere(~maps["map1"])
    by, bx = (int(v) for v in blocked[len(blocked) // 2])
    cases.append(("map1", (5, 5), (bx, by), True))
    for name, s, g, th in cases:
        L.set_map(maps[name])
        r = L.run_astar(s, g, thetastar=th)
        kat.append(dict(map=name, start=s, goal=g, thetastar=th, path=r["path"], expanded=r["expanded"], cost=r["cost"],
                        n_los=len(r["los"]), los_true=int(sum(r["los"])),
                        los_hex=np.packbits(np.array(r["los"], np.uint8)).tobytes().hex(), stdout=r["stdout"]))
    json.dump(kat, open(os.path.join(HERE, "theta_kat.json"), "w"), indent=0)

    # ---------------- LOS booleans
    rng = np.random.default_rng(2024)
    los = {}
    for name in ("map1", "map2"):
        L.set_map(maps[name])
        n = maps[name].shape[0]
        seg = rng.integers(-3, n + 3, size=(3000, 4))
        seg[:200, 2:] = seg[:200, :2] + rng.integers(-2, 3, size=(200, 2))  # short / degenerate rays
        out = np.array([search.lineofsight((a, b), (c, d)) for a, b, c, d in seg], bool)
        los[name + "_seg"] = seg.astype(np.int32)
        los[name + "_los"] = out
    np.savez_compressed(os.path.join(HERE, "los_kat.npz"), **los)

    # ---------------- getCircle pixel sets
    circ = []
    L.set_map(maps["map1"])
    for _ in range(120):
        r = float(rng.choice([0, 1, 2, 2.33, 3, 5, 7.9, 12, 30, 64, 99.5, 150, 1000.7, 41014.2]))
        cx, cy = (float(v) for v in rng.uniform(-1.2 * r - 3, 100 + 1.2 * r + 3, 2))
        px = search.getCircle((cx, cy), r)
        circ.append(dict(center=(cx, cy), r=r, pixels=sorted(set((int(a), int(b)) for a, b in px))))
    json.dump(circ, open(os.path.join(HERE, "circle_kat.json"), "w"))

    # ---------------- sample streams
    st = {}
    for i, (goal, seed, name) in enumerate([(((90, 50), 90.0), 0, "map1"), (((3, 97), -170.5), 3, "map1"),
                                            (((73, 38), 135.0), 11, "map2")]):
        L.set_map(maps[name])
        s = L.make_stream(goal, 400, seed)
        st[f"s{i}_goal"] = np.array([goal[0][0], goal[0][1], goal[1]], float)
        st[f"s{i}_seed"] = np.array([seed])
        st[f"s{i}_shape"] = np.array(maps[name].shape)
        st[f"s{i}_xy"] = np.array([p[0] for p in s], np.int32)
        st[f"s{i}_th"] = np.array([p[1] for p in s], float)
    np.savez_compressed(os.path.join(HERE, "stream_kat.npz"), **st)

    # ---------------- RRT runs
    rk = {}
    runs = [("map1", ((5, 5), 0.0), ((90, 50), 90.0), 0, 300, 10),       # BASELINE cfg 1
            ("map1", ((5, 5), 0.0), ((90, 50), 90.0), 0, 1501, 0),
            ("map2", ((280, 0), 169.6), ((73, 38), 135.0), 0, 300, 10),
            ("blank", ((10, 10), 45.0), ((250, 200), -30.0), 5, 400, 10)]
    free1 = np.argwhere(maps["map1"])
    rq = np.random.default_rng(1234)
    for q in range(6):  # cfg-3 style random queries, shortened
        a = free1[rq.integers(len(free1))]
        b = free1[rq.integers(len(free1))]
        runs.append(("map1", ((int(a[1]), int(a[0])), float(rq.uniform(-180, 180))),
                     ((int(b[1]), int(b[0])), float(rq.uniform(-180, 180))), q, 801, 0))
    for i, (name, start, goal, seed, K, tol) in enumerate(runs):
        L.set_map(maps[name])
        s = L.make_stream(goal, K - 1, seed)
        r = L.run_rrt(start, goal, s, K=K, tol_xy=tol)
        p = f"r{i}_"
        rk[p + "map"] = np.array(name)
        rk[p + "start"] = np.array([start[0][0], start[0][1], start[1]], float)
        rk[p + "goal"] = np.array([goal[0][0], goal[0][1], goal[1]], float)
        rk[p + "K"] = np.array([K])
        rk[p + "tol_xy"] = np.array([tol], float)
        rk[p + "sxy"] = np.array([q[0] for q in s], np.int32)
        rk[p + "sth"] = np.array([q[1] for q in s], float)
        rk[p + "raised"] = np.array(r["raised"] or "")
        if r["raised"]:
            rk[p + "iterations"] = np.array([r["iterations"]])
            continue
        rk[p + "x"] = np.array(r["x"]); rk[p + "y"] = np.array(r["y"]); rk[p + "theta"] = np.array(r["theta"])
        rk[p + "parent"] = np.array(r["parent"], np.int32)
        rk[p + "sol"] = np.array([-1 if r["sol"] is None else r["sol"]])
        rk[p + "iterations"] = np.array([r["iterations"]])
        rk[p + "nearest_it"] = np.array([k for k, _ in r["nearest"]], np.int32)
        rk[p + "nearest_idx"] = np.array([j for _, j in r["nearest"]], np.int32)
        rk[p + "los"] = np.array([b for _, b in r["los"]], bool)
        rk[p + "los_it"] = np.array([k for k, _ in r["los"]], np.int32)
        # steer records: (iter, x, y, theta, steer, iccx, iccy, rad, dist), NaN icc/rad = straight
        rk[p + "steer"] = np.array([[t[0], t[1], t[2], t[3], t[4], np.nan if t[5] is None else t[5][0],
                                     np.nan if t[5] is None else t[5][1], np.nan if t[6] is None else t[6], t[7]]
                                    for t in r["steer"]], float).reshape(-1, 9)
        rk[p + "drive"] = np.array(r["drive"], float).reshape(-1, 4)
        rk[p + "arc_it"] = np.array([k for k, _ in r["arc"]], np.int32)
        rk[p + "arc_npx"] = np.array([len(px) for _, px in r["arc"]], np.int32)
        # children lists as (parent, child) edge list in append order per parent
        edges = [(pi, c) for pi, ch in enumerate(r["children"]) for c in ch]
        rk[p + "edges"] = np.array(edges, np.int32).reshape(-1, 2)
    rk["n_runs"] = np.array([len(runs)])
    np.savez_compressed(os.path.join(HERE, "rrt_kat.npz"), **rk)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
