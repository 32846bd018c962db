"""Oracle validation against the LIVE reference (unmodified Python under /root/reference).
Runs only in the build container (the reference does not travel to the GPU box); the same
comparisons are frozen into tests/golden/ by tests/golden/make_golden.py."""
import numpy as np
import pytest

from oracle import c_oracle as O
from tests import util

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def L():
    from oracle import live_reference
    live_reference.load()
    return live_reference


def test_theta_map1_live(L, maps):
    L.set_map(maps["map1"])
    for s, g, th in (((5, 5), (90, 50), True), ((97, 2), (2, 97), True), ((30, 70), (70, 30), False), ((50, 50), (80, 80), True)):
        r = L.run_astar(s, g, thetastar=th)
        o = O.astar(maps["map1"], s, g, thetastar=th)
        assert (o["path"] if o["status"] == 0 else False) == r["path"]
        if r["path"] is not False:
            assert o["expanded"] == r["expanded"] and o["cost"] == r["cost"]
        assert list(o["los"]) == r["los"]


def test_primitives_live(L, maps):
    search, rrt, _ = L.load()
    L.set_map(maps["map1"])
    rng = np.random.default_rng(3)
    for _ in range(300):
        a, b = rng.integers(-3, 103, 2), rng.integers(-3, 103, 2)
        assert O.bresenham(a, b) == [(int(x), int(y)) for x, y in search.bresenham(a, b)]
        assert O.lineofsight(maps["map1"], a, b) == search.lineofsight(a, b)
    for _ in range(150):
        r = float(rng.choice([0.4, 1.2, 2.33, 4.0, 9.7, 25.5, 80.1, 700.3]))
        c = rng.uniform(-r - 5, 105 + r, 2)
        ref = sorted(set((int(x), int(y)) for x, y in search.getCircle(c, r)))
        assert sorted(set(O.getcircle((100, 100), c, r))) == ref
    for _ in range(400):
        a1, a2 = rng.uniform(-200, 200, 2)
        assert abs(O.anglediff(a1, a2) - rrt.anglediff(a1, a2)) < 1e-11
        v = rng.uniform(-50, 50, 2)
        d = abs(O.anglebetween((1, 0), v) - rrt.anglebetween([1, 0], v))
        assert min(d, abs(d - 360)) < 1e-11
        assert O.l2norm(v, (3, 4)) == search.L2norm(v, (3, 4))


def test_steer_drive_getarc_live(L, maps):
    search, rrt, _ = L.load()
    L.set_map(maps["map1"])
    L.set_params()
    rng = np.random.default_rng(4)
    n_arc = 0
    for _ in range(400):
        o_xy = tuple(rng.uniform(0, 100, 2)); th = float(rng.uniform(-180, 180))
        g_xy = tuple(int(v) for v in rng.integers(0, 100, 2)); gth = float(rng.uniform(-180, 180))
        (gp, fa), u = rrt.steer(o_xy, th, g_xy, gth)
        s = O.steer(o_xy, th, g_xy, gth)
        assert s["straight"] == (u[1] is None)
        assert abs(s["x"] - gp[0]) < 1e-8 and abs(s["y"] - gp[1]) < 1e-8
        if u[1] is not None:
            assert abs(s["rad"] - u[2]) <= 1e-9 * max(1, u[2]) and abs(s["dist"] - u[3]) < 1e-8
            (pos, ang), _ = rrt.drive((o_xy, th), (u[0], u[1], u[2], u[3] / 3))
            d = O.drive(o_xy, th, (u[0], u[1], u[2], u[3] / 3))
            assert abs(d[0] - pos[0]) < 1e-8 and abs(d[1] - pos[1]) < 1e-8
        ref_px = sorted(set((int(a), int(b)) for a, b in search.getArc(o_xy, (gp[0], gp[1]), u)))
        got_px = sorted(set(O.getarc((100, 100), o_xy, (gp[0], gp[1]), u)))
        if ref_px != got_px:  # a pixel exactly on the arc's end ray may flip with 1e-13 noise: not more than one
            assert len(set(ref_px) ^ set(got_px)) <= 1
        n_arc += 1
    assert n_arc == 400


def test_rrt_live_k2001(L, maps):
    """Longer live run (the golden set stops at K=1501): discrete outputs up to the first ambiguous iteration."""
    L.set_map(maps["map1"])
    goal, start, K = ((90, 50), 90.0), ((5, 5), 0.0), 2001
    st = L.make_stream(goal, K - 1, 0)
    ref = L.run_rrt(start, goal, st, K=K, tol_xy=0)
    sxy = np.array([s[0] for s in st], np.int32); sth = np.array([s[1] for s in st])
    o = O.rrt(maps["map1"], start, goal, sxy, sth, O.Params(tol_xy=0.0), K=K)
    assert o["n_nodes"] == ref["n_nodes"] == 1214
    assert list(o["parent"]) == ref["parent"]
    assert [int(v) for v in o["it_near"] if v >= 0] == [i for _, i in ref["nearest"]]
    assert list(o["los"]) == [b for _, b in ref["los"]]


def test_findnearest_live(L, maps):
    search, rrt, _ = L.load()
    L.set_map(maps["map1"])
    goal, start, K = ((90, 50), 90.0), ((5, 5), 0.0), 300
    st = L.make_stream(goal, K - 1, 0)
    L.set_params(K=K, tol_xy=0)
    it = iter(st)
    orig = rrt.rand_conf
    rrt.rand_conf = lambda g: next(it)
    try:
        with L.quiet():
            sol, G, cf = rrt.rrt(start, goal)
        node, dist = rrt.findnearest(G, goal)
    finally:
        rrt.rand_conf = orig
        L.set_params()
    keys = list(G.keys()); index = {k: i for i, k in enumerate(keys)}
    ep = [index[p] for p in keys for c in G[p]]; ec = [index[c] for p in keys for c in G[p]]
    b, d = O.findnearest([k[0][0] for k in keys], [k[0][1] for k in keys], [k[1] for k in keys], ep, ec, goal)
    assert b == index[node] and abs(d - dist) < 1e-9


PARAM_SETS = [
    dict(bikelength=3, LEFTCONSTRAINT=-40, RIGHTCONSTRAINT=40, frontclearance=1.5, maxdrivedist=12, weightxy=0.3, tol_ang=20),
    dict(FORWARDONLY=False, bikelength=7, LEFTCONSTRAINT=-30, RIGHTCONSTRAINT=55, maxdrivedist=45, weightxy=0.9),
]


@pytest.mark.parametrize("ps", PARAM_SETS, ids=["short-bike", "reverse-allowed"])
def test_rrt_live_other_parameters(L, maps, ps):
    """The builtins.* parameters (main.py:15-32) away from their defaults: the oracle follows the live reference on
    every discrete output up to the first iteration its margin audit marks as decided by rounding noise."""
    L.set_map(maps["map1"])
    goal, start, K = ((80, 20), -45.0), ((10, 90), 30.0), 601
    st = L.make_stream(goal, K - 1, 5)
    ref = L.run_rrt(start, goal, st, K=K, tol_xy=0, **ps)
    L.set_params()
    sxy = np.array([s[0] for s in st], np.int32); sth = np.array([s[1] for s in st])
    op = O.Params(tol_xy=0.0, **{k.lower(): (int(v) if isinstance(v, bool) else v) for k, v in ps.items()})
    o = O.rrt(maps["map1"], start, goal, sxy, sth, op, K=K, audit_eps=util.AUDIT_EPS)
    if "raised" in ref and ref["raised"]:
        assert o["status"] == 4
        return
    fa = o["first_ambiguous"]
    near_ref = [i for _, i in ref["nearest"]]
    near_got = [int(v) for v in o["it_near"] if v >= 0]
    if fa < 0:
        assert o["n_nodes"] == ref["n_nodes"] and list(o["parent"]) == ref["parent"]
        assert near_got == near_ref and list(o["los"]) == [b for _, b in ref["los"]]
    else:
        n_cmp = sum(1 for it, _ in ref["nearest"] if it <= fa)
        assert near_got[:n_cmp] == near_ref[:n_cmp]
    assert o["n_nodes"] > 20  # the run really grew a tree under these parameters
