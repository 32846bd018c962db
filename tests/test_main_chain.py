"""main.py flow (main.py:56-84): Theta* waypoints -> one RRT per segment -> findnearest restart.

Golden: tests/golden/main_kat.json, produced by tests/golden/make_main_golden.py from the reference's own
rrt.rrt / rrt.findnearest / rrt.anglebetween with numpy's global generator seeded like a user would.
CPU test: the oracle chained the same way reproduces it (pins the oracle on this flow).
GPU test: the drop-in theta_rrt_b200.main.chain reproduces it through the C ABI.
"""
import json
import os

import numpy as np
import pytest

from tests import util


def golden():
    return json.load(open(os.path.join(util.GOLDEN, "main_kat.json")))


def check_segments(got, case):
    assert len(got) == len(case["segments"])
    for g, ref in zip(got, case["segments"]):
        assert g["n_nodes"] == ref["n_nodes"] and g["n_edges"] == ref["n_edges"], (case["map"], g, ref)
        for key in ("begin", "end", "solution", "nearest"):
            if ref[key] is None:
                assert g[key] is None, key
                continue
            a = np.array([g[key][0][0], g[key][0][1], g[key][1]], float)
            b = np.array([ref[key][0][0], ref[key][0][1], ref[key][1]], float)
            assert np.allclose(a, b, rtol=1e-9, atol=1e-9), (key, a, b)
        if ref["mindist"] is not None:
            assert abs(g["mindist"] - ref["mindist"]) <= 1e-9 * max(1.0, abs(ref["mindist"]))


def oracle_chain(free, waypoints, K):
    """main.py:58-81 with the C oracle in place of rrt.rrt / rrt.findnearest."""
    from oracle import c_oracle as O
    from theta_rrt_b200 import samples
    from theta_rrt_b200.rrt import anglebetween, standardangle
    path = list(waypoints) + [None]
    nearest = None
    out = []
    for first, second, third in zip(path, path[1:], path[2:]):
        angle1 = anglebetween([1, 0], np.subtract(second, first))
        if nearest is not None:
            angle1, first = nearest[1], nearest[0]
        angle2 = angle1 if third is None else anglebetween([1, 0], np.subtract(third, second))
        begin = (first, standardangle(angle1))
        end = (second, standardangle(angle2))
        state = np.random.get_state()
        sxy, sth = samples.draw_stream_global(end, K - 1, free.shape)
        o = O.rrt(free, begin, end, sxy, sth, O.Params(), K=K)
        assert o["status"] in (0, 1)
        np.random.set_state(state)
        np.random.standard_normal(3 * o["iters"])  # the reference stops drawing at its `break`
        node = lambda i: ((float(o["x"][i]), float(o["y"][i])), float(o["theta"][i]))
        sel = o["it_new"] >= 0
        rec = {"begin": begin, "end": end, "solution": node(o["sol"]) if o["sol"] >= 0 else None, "n_nodes": o["n_nodes"],
               "n_edges": int(sel.sum()), "nearest": None, "mindist": None}
        if o["sol"] < 0:
            b, d = O.findnearest(o["x"], o["y"], o["theta"], o["it_near"][sel], o["it_new"][sel], end)
            nearest = node(b)
            rec["nearest"], rec["mindist"] = nearest, d
        out.append(rec)
    return out


@pytest.mark.parametrize("case", golden(), ids=lambda c: c["map"])
def test_oracle_reproduces_main_flow(maps, case):
    np.random.seed(case["seed"])
    got = oracle_chain(maps[case["map"]], [tuple(w) for w in case["waypoints"]], case["K"])
    check_segments(got, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", golden(), ids=lambda c: c["map"])
def test_dropin_main_chain_matches_reference(maps, case, capsys):
    import builtins
    from theta_rrt_b200 import main as M
    M.set_map(maps[case["map"]])
    old_k = builtins.K
    builtins.K = case["K"]
    try:
        np.random.seed(case["seed"])
        segs = M.chain([tuple(w) for w in case["waypoints"]], debug=True)
    finally:
        builtins.K = old_k
    got = [{"begin": s["begin"], "end": s["end"], "solution": s["solution"], "n_nodes": len(s["graph"]),
            "n_edges": sum(len(v) for v in s["graph"].values()), "nearest": s["nearest"], "mindist": s["mindist"]}
           for s in segs]
    check_segments(got, case)
    assert "Nodes in tree:" in capsys.readouterr().out  # rrt.py:204


@pytest.mark.gpu
def test_dropin_plan_png_in_path_out(maps, tmp_path):
    """PNG file in, Theta* waypoints computed on the GPU (they equal main.py:57), chained path out."""
    from PIL import Image
    from theta_rrt_b200 import main as M
    p = tmp_path / "map2.png"
    Image.fromarray((maps["map2"].astype(np.uint8) * 255)).convert("RGB").save(p)
    np.random.seed(0)
    segs = M.plan(str(p), start=(280, 0), goal=(8, 280))
    case = golden()[0]
    assert [s["end"][0] for s in segs] == [tuple(w) for w in M.REFERENCE_WAYPOINTS[1:]]
    assert [len(s["graph"]) for s in segs] == [r["n_nodes"] for r in case["segments"]]
    assert all(len(M.path_nodes(s)) >= 1 for s in segs)
