"""Live-reference harness: runs the UNMODIFIED eshira/theta-rrt Python sources.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (theta_rrt_b200/) may
import this module; it is used by tests/ (when /root/reference is present),
tests/golden/make_golden.py (fixture generation) and nothing else.

The reference lives read-only at /root/reference and exists only in the build
container, never on the GPU box.  This module
  * stubs `matplotlib` (imported at module top by main.py:4 / rrt.py:6, never
    touched on the hot path when plot=False and showtree=False),
  * imports `search` BEFORE `main`/`rrt` (import cycle: search.py:4 and
    rrt.py:3 do `from main import *`; search.getArc needs the name `rrt`,
    search.py:151),
  * sets `builtins.imarray` and the parameters of main.py:15-32,
  * injects the sample stream by replacing `rrt.rand_conf` (resolved as a
    module global at call time, rrt.py:144),
  * records the per-iteration nearest index (rrt.py:157), every
    `search.lineofsight` boolean, and converts `(G, cameFrom)` to arrays
    (node index = insertion order of G).
"""
from __future__ import annotations

import builtins
import contextlib
import io
import os
import sys
import types

import numpy as _numpy

REFERENCE_DIR = os.environ.get("THETA_RRT_REFERENCE", "/root/reference")

DEFAULT_PARAMS = dict(  # main.py:15-32
    THETASTAR=True, bikelength=5, FORWARDONLY=True, LEFTCONSTRAINT=-65,
    RIGHTCONSTRAINT=65, frontclearance=2, K=300, showtree=False,
    maxdrivedist=30, tol_xy=10, tol_ang=45, weightxy=.6, xystdv=0.4,
    anglestdv=100,
)

_mods = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "search.py"))


def _stub_matplotlib():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_trrt_stub", False):
        return
    mpl = types.ModuleType("matplotlib")
    mpl._trrt_stub = True
    pyplot = types.ModuleType("matplotlib.pyplot")
    patches = types.ModuleType("matplotlib.patches")
    mpl.pyplot = pyplot
    mpl.patches = patches
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = pyplot
    sys.modules["matplotlib.patches"] = patches


def load():
    """Import (search, rrt, main) from the reference tree exactly once."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_DIR}")
    sys.dont_write_bytecode = True
    _stub_matplotlib()
    # The product package also has modules called `search`/`rrt`/`main`, but
    # only inside the theta_rrt_b200 package namespace, so top-level names are
    # free for the reference.
    for name in ("search", "rrt", "main"):
        if name in sys.modules:
            raise RuntimeError(f"top-level module {name!r} already imported")
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import search  # noqa: F401  (must come first)
        import rrt  # noqa: F401
        import main  # noqa: F401
    finally:
        sys.path.remove(REFERENCE_DIR)
    set_params()
    _mods = (sys.modules["search"], sys.modules["rrt"], sys.modules["main"])
    return _mods


def set_params(**over):
    p = dict(DEFAULT_PARAMS)
    p.update(over)
    for k, v in p.items():
        setattr(builtins, k, v)
    return p


def set_map(imarray):
    """imarray: bool array (H, W), True = free (main.py:38-42)."""
    builtins.imarray = _numpy.asarray(imarray, dtype=bool)


def load_png(path):
    from PIL import Image
    return _numpy.array(Image.open(path).convert("1"))  # main.py:38-42


@contextlib.contextmanager
def quiet():
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf


# --------------------------------------------------------------------------
# Theta* / A*
# --------------------------------------------------------------------------
def run_astar(start, goal, thetastar=True):
    """Return dict(path=list|False, expanded=int|None, los=list[bool], cost)."""
    search, rrt, _ = load()
    builtins.THETASTAR = bool(thetastar)
    los = []
    orig = search.lineofsight

    def rec_los(a, b):
        r = orig(a, b)
        los.append(bool(r))
        return r

    search.lineofsight = rec_los
    try:
        with quiet() as out:
            path = search.astar(tuple(start), tuple(goal))
    finally:
        search.lineofsight = orig
        builtins.THETASTAR = True
    text = out.getvalue()
    expanded = None
    for line in text.splitlines():
        if line.startswith("Expanded nodes:"):
            expanded = int(line.split(":")[1])
    res = dict(path=False, expanded=expanded, los=los, cost=None, stdout=text)
    if path is not False:
        res["path"] = [(int(p[0]), int(p[1])) for p in path]
        cost = 0.0  # plain left-to-right fp64 accumulation (Python >= 3.12 sum() is compensated)
        for a, b in zip(path, path[1:]):
            cost = cost + search.L2norm(a, b)
        res["cost"] = float(cost)
    return res


# --------------------------------------------------------------------------
# RRT
# --------------------------------------------------------------------------
class _NpProxy:
    """Forwards to numpy; records every argmin result (rrt.py:157)."""

    def __init__(self, rec):
        self._rec = rec

    def __getattr__(self, name):
        return getattr(_numpy, name)

    def argmin(self, a, *args, **kw):
        i = _numpy.argmin(a, *args, **kw)
        self._rec.append(int(i))
        return i


def run_rrt(start, goal, stream, K=None, **params):
    """Run rrt.rrt with an injected sample stream.

    stream: iterable of ((x:int, y:int), theta:float) -- one per iteration.
    Returns a dict of plain arrays/lists (see keys below).  If the reference
    raises (quirk Q7: drive() on a straight-line steer), 'raised' holds the
    exception type name and the arrays describe nothing.
    """
    search, rrt, _ = load()
    stream = list(stream)
    if K is None:
        K = len(stream) + 1
    set_params(K=K, **params)
    it = iter(stream)
    nearest_rec, los_rec, steer_rec, arc_rec, drive_rec = [], [], [], [], []
    iter_of_nearest = []
    k_counter = [0]

    orig = dict(rand_conf=rrt.rand_conf, np=rrt.np, los=search.lineofsight,
                steer=rrt.steer, getArc=search.getArc, drive=rrt.drive)

    def rand_conf(goal_):
        k_counter[0] += 1
        return next(it)

    def rec_los(a, b):
        r = orig["los"](a, b)
        los_rec.append((k_counter[0], bool(r)))
        return r

    def rec_steer(*a, **kw):
        iter_of_nearest.append(k_counter[0])
        r = orig["steer"](*a, **kw)
        (gp, fa), u = r
        steer_rec.append((k_counter[0], float(gp[0]), float(gp[1]), float(fa), float(u[0]),
                          None if u[1] is None else (float(u[1][0]), float(u[1][1])),
                          None if u[2] is None else float(u[2]), float(u[3])))
        return r

    def rec_drive(*a, **kw):
        r = orig["drive"](*a, **kw)
        (pos, ang), u = r
        drive_rec.append((k_counter[0], float(pos[0]), float(pos[1]), float(ang)))
        return r

    def rec_arc(*a, **kw):
        r = orig["getArc"](*a, **kw)
        arc_rec.append((k_counter[0], [(int(p[0]), int(p[1])) for p in r]))
        return r

    rrt.rand_conf = rand_conf
    rrt.np = _NpProxy(nearest_rec)
    search.lineofsight = rec_los
    rrt.steer = rec_steer
    rrt.drive = rec_drive
    search.getArc = rec_arc
    raised = None
    sol = G = cameFrom = None
    try:
        with quiet():
            try:
                sol, G, cameFrom = rrt.rrt(start, goal, debug=False)
            except (TypeError, ValueError, IndexError) as e:  # reference quirks Q7/Q8
                raised = type(e).__name__
    finally:
        rrt.rand_conf = orig["rand_conf"]
        rrt.np = orig["np"]
        search.lineofsight = orig["los"]
        rrt.steer = orig["steer"]
        rrt.drive = orig["drive"]
        search.getArc = orig["getArc"]
        set_params()

    res = dict(raised=raised, iterations=k_counter[0],
               nearest=[(k, i) for k, i in zip(iter_of_nearest, nearest_rec)],
               los=los_rec, steer=steer_rec, drive=drive_rec, arc=arc_rec)
    if raised is not None:
        return res
    keys = list(G.keys())
    index = {n: i for i, n in enumerate(keys)}
    res["x"] = [float(n[0][0]) for n in keys]
    res["y"] = [float(n[0][1]) for n in keys]
    res["theta"] = [float(n[1]) for n in keys]
    parent, u_out = [], []
    for n in keys:
        cf = cameFrom.get(n)
        if cf is None:
            parent.append(-1)
            u_out.append(None)
        else:
            parent.append(index[cf[0]])
            u = cf[1]
            u_out.append((float(u[0]),
                          None if u[1] is None else (float(u[1][0]), float(u[1][1])),
                          None if u[2] is None else float(u[2]), float(u[3])))
    res["parent"] = parent
    res["u"] = u_out
    res["children"] = [[index[c] for c in G[n]] for n in keys]
    res["sol"] = None if sol is None else index[sol]
    res["n_nodes"] = len(keys)
    return res


def make_stream(goal, n, seed):
    """n x rand_conf(goal) after np.random.seed(seed) -- the literal reference
    generator (rrt.py:53-68), used to validate the vectorised host version."""
    search, rrt, _ = load()
    _numpy.random.seed(seed)
    goal = (goal[0], rrt.standardangle(goal[1]))
    out = []
    for _ in range(n):
        q = rrt.rand_conf(goal)
        out.append(((int(q[0][0]), int(q[0][1])), float(rrt.standardangle(q[1]))))
    return out


def find_nearest(tree_children_nodes, goal, **params):
    """rrt.findnearest on a (G) dict rebuilt by the caller."""
    search, rrt, _ = load()
    set_params(**params)
    try:
        return rrt.findnearest(tree_children_nodes, goal)
    finally:
        set_params()
