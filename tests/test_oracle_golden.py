"""CPU tests: the C oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) -- the pinning of the oracle that travels to the GPU box."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as O
from tests import util


def test_theta_known_answers(maps):
    for k in util.theta_kat():
        o = O.astar(maps[k["map"]], k["start"], k["goal"], thetastar=k["thetastar"])
        tag = (k["map"], k["start"], k["goal"], k["thetastar"])
        if k["path"] is False:
            assert o["path"] is False and o["status"] in (1, 2, 3), tag
            if "not valid" in k["stdout"]:
                assert o["status"] == 2, tag
            elif "inside an obstacle" in k["stdout"]:
                assert o["status"] == 3, tag
            else:
                assert o["status"] == 1, tag
            continue
        assert o["path"] == [tuple(p) for p in k["path"]], tag
        assert o["expanded"] == k["expanded"], tag
        assert o["cost"] == k["cost"], tag  # integer coordinates + sqrt + fp64 add: exact
        los_ref = np.unpackbits(np.frombuffer(bytes.fromhex(k["los_hex"]), np.uint8))[:k["n_los"]].astype(bool)
        assert o["n_los"] == k["n_los"] and np.array_equal(o["los"], los_ref), tag


def test_main_py_57_waypoints(maps):
    """The only result the reference itself pins: main.py:57."""
    o = O.astar(maps["map2"], (280, 0), (8, 280))
    assert o["path"] == [(280, 0), (73, 38), (72, 39), (33, 130), (15, 190), (8, 280)]
    assert o["expanded"] == 30384 and o["cost"] == 463.79193690045025


def test_los_golden(maps):
    z = np.load(os.path.join(util.GOLDEN, "los_kat.npz"))
    for name in ("map1", "map2"):
        assert np.array_equal(O.lineofsight_batch(maps[name], z[name + "_seg"]), z[name + "_los"]), name


def test_circle_golden():
    for c in json.load(open(os.path.join(util.GOLDEN, "circle_kat.json"))):
        got = sorted(set(O.getcircle((100, 100), c["center"], c["r"])))
        assert got == [tuple(p) for p in c["pixels"]], (c["center"], c["r"])


def test_stream_golden():
    from theta_rrt_b200 import samples
    z = np.load(os.path.join(util.GOLDEN, "stream_kat.npz"))
    for i in range(3):
        g = z[f"s{i}_goal"]
        xy, th = samples.make_stream(((g[0], g[1]), g[2]), 400, int(z[f"s{i}_seed"][0]), tuple(z[f"s{i}_shape"]))
        assert np.array_equal(xy, z[f"s{i}_xy"]) and np.array_equal(th, z[f"s{i}_th"])


def test_rrt_golden_runs(maps):
    """Discrete outputs identical to the reference up to the first iteration the margin audit marks
    as decided by rounding noise; runs without such an iteration must match completely."""
    clean = 0
    compared = 0
    for i, run in enumerate(util.rrt_runs()):
        assert not str(run["raised"])
        free = maps[str(run["map"])]
        K = int(run["K"][0])
        o = O.rrt(free, ((run["start"][0], run["start"][1]), run["start"][2]),
                  ((run["goal"][0], run["goal"][1]), run["goal"][2]), run["sxy"], run["sth"],
                  O.Params(tol_xy=float(run["tol_xy"][0])), K=K, audit_eps=util.AUDIT_EPS)
        fa = o["first_ambiguous"]
        if i < 2:
            # cfg 1 (integer start, heading 0): the ICC lands within 1e-15 of an integer -- structurally
            # ambiguous, but the oracle reproduces scipy's arithmetic bit for bit there: full equality
            fa = -1
        compared += util.compare_rrt_with_reference(o, run, fa)
        clean += fa < 0
    assert clean >= 6 and compared > 5000


def test_rrt_cfg1_numbers(maps):
    """SURVEY.md 8c sanity points: goal found at iteration 86 with 45 nodes; parents start [-1,0,1,2,3,0,4,6,1,6,6,10]."""
    run = util.rrt_runs()[0]
    o = O.rrt(maps["map1"], ((5, 5), 0.0), ((90, 50), 90.0), run["sxy"], run["sth"], O.Params(), K=300)
    assert o["status"] == 0 and o["iters"] == 86 and o["n_nodes"] == 45
    assert list(o["parent"][:12]) == [-1, 0, 1, 2, 3, 0, 4, 6, 1, 6, 6, 10]


def test_steer_drive_records(maps):
    """Single-step float parity: steer / drive outputs of the oracle against the reference's records,
    evaluated on the ORACLE's own tree state one step at a time (no drift accumulation beyond the inputs)."""
    run = util.rrt_runs()[4]
    free = maps[str(run["map"])]
    K = int(run["K"][0])
    o = O.rrt(free, ((run["start"][0], run["start"][1]), run["start"][2]), ((run["goal"][0], run["goal"][1]), run["goal"][2]),
              run["sxy"], run["sth"], O.Params(tol_xy=0.0), K=K)
    st = run["steer"]
    worst = 0.0
    for row in st:
        it = int(row[0]) - 1
        near = int(o["it_near"][it])
        s = O.steer((o["x"][near], o["y"][near]), o["theta"][near], run["sxy"][it], run["sth"][it], O.Params(tol_xy=0.0))
        ref = row[1:]
        got = np.array([s["x"], s["y"], s["theta"], s["steer"], s["icc"][0], s["icc"][1], s["rad"], s["dist"]])
        assert np.isnan(ref[5]) == s["straight"]
        m = ~np.isnan(ref)
        # angles may sit on the +-180 seam
        d = np.abs(got[m] - ref[m])
        d = np.minimum(d, np.abs(d - 360))
        worst = max(worst, float(np.max(d / np.maximum(1.0, np.abs(ref[m])))))
    assert worst < 1e-5  # inputs already carry the lineage drift (<= 1e-6 on this run)


@pytest.mark.parametrize("threads", [1, 3])
def test_batch_entry_points_match_single(maps, threads):
    from theta_rrt_b200 import samples
    free = maps["map1"]
    nq, K = 6, 201
    starts, goals = util.random_queries(free, nq, 5)
    sxy = np.empty((nq, K - 1, 2), np.int32)
    sth = np.empty((nq, K - 1))
    for q in range(nq):
        sxy[q], sth[q] = samples.make_stream(((goals[q, 0], goals[q, 1]), goals[q, 2]), K - 1, q, free.shape)
    b = O.rrt_batch(free, starts, goals, sxy, sth, K, O.Params(tol_xy=0.0), threads=threads)
    for q in range(nq):
        o = O.rrt(free, ((starts[q, 0], starts[q, 1]), starts[q, 2]), ((goals[q, 0], goals[q, 1]), goals[q, 2]), sxy[q], sth[q],
                  O.Params(tol_xy=0.0), K=K)
        n = o["n_nodes"]
        assert b["n_nodes"][q] == n and np.array_equal(b["parent"][q, :n], o["parent"])
        assert np.array_equal(b["nodes"][q, :n, 0], o["x"])
    sg = np.array([[5, 5, 90, 50], [2, 97, 97, 2], [50, 50, 80, 80]], np.int32)
    ab = O.astar_batch(free, sg, threads=threads)
    for q in range(3):
        o = O.astar(free, sg[q, :2], sg[q, 2:])
        assert ab["status"][q] == o["status"] and ab["cost"][q] == o["cost"] and ab["expanded"][q] == o["expanded"]
