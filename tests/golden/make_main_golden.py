"""Golden for the main.py flow (main.py:56-84) from the UNMODIFIED reference functions.

Run in the build container only (needs /root/reference):
    python tests/golden/make_main_golden.py
The loop of main.py lives under `if __name__ == "__main__"` and ends in plt.show(), so it is
restated here call for call (rrt.anglebetween, rrt.standardangle, rrt.rrt, rrt.findnearest are the
reference's own functions, rand_conf draws from numpy's global generator like in the reference);
only the matplotlib drawing calls are left out.  Output: main_kat.json.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import live_reference as L  # noqa: E402


def node_json(n):
    return None if n is None else [[float(n[0][0]), float(n[0][1])], float(n[1])]


def run_chain(rrt, waypoints):
    path = list(waypoints) + [None]
    nearest = None
    out = []
    for first, second, third in zip(path, path[1:], path[2:]):  # main.py:58-81
        angle1 = rrt.anglebetween([1, 0], np.subtract(second, first))
        if nearest is not None:
            angle1 = nearest[1]
            first = nearest[0]
        if third is None:
            angle2 = angle1
        else:
            angle2 = rrt.anglebetween([1, 0], np.subtract(third, second))
        begin = (first, rrt.standardangle(angle1))
        end = (second, rrt.standardangle(angle2))
        with L.quiet():
            solution, graph, camefrom = rrt.rrt(begin, end, debug=True)
        rec = {"begin": node_json(begin), "end": node_json(end), "solution": node_json(solution),
               "n_nodes": len(graph), "n_edges": sum(len(v) for v in graph.values()), "nearest": None, "mindist": None}
        if solution is None:
            nearest, mindist = rrt.findnearest(graph, end)
            rec["nearest"], rec["mindist"] = node_json(nearest), float(mindist)
        out.append(rec)
    return out


def main():
    search, rrt, _ = L.load()
    cases = []
    for name, seed, K, wp in (("map2", 0, 300, [(280, 0), (73, 38), (72, 39), (33, 130), (15, 190), (8, 280)]),
                              ("map1", 3, 400, [(5, 5), (43, 39), (46, 41), (90, 50)])):
        L.set_map(L.load_png(os.path.join(L.REFERENCE_DIR, name + ".png")))
        L.set_params(K=K)
        np.random.seed(seed)
        segs = run_chain(rrt, wp)
        cases.append({"map": name, "seed": seed, "K": K, "waypoints": wp, "segments": segs})
        print(name, [(s["n_nodes"], s["solution"] is not None) for s in segs])
    json.dump(cases, open(os.path.join(HERE, "main_kat.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
