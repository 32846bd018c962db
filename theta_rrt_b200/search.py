"""Drop-in for the reference's `search` module (search.py) on the GPU path.

Same names, argument meaning and return conventions as the reference:
`astar(start, goal)` returns the list of (x, y) nodes or False (and prints the
reference's messages), `lineofsight(n1, n2)` returns a bool.  The map and the
parameters come from `builtins.imarray` / `builtins.THETASTAR` like in the
reference.  Single calls are batches of one; the batched forms
(`astar_batch`, `lineofsight_batch`) are what throughput-sensitive callers use.
"""
from __future__ import annotations

import builtins
import math

import numpy as np

from . import _context
from .planner import STATUS_NAMES


def L2norm(node0, node1):
    """search.py:13-15 (host scalar helper; the device code computes the same fp64 expression)."""
    return math.sqrt(pow((node1[0] - node0[0]), 2) + pow((node1[1] - node0[1]), 2))


def heuristic(node, goal):
    """search.py:9-11."""
    return L2norm(node, goal)


def valid(node):
    """search.py:17-24 (pure bounds arithmetic on the image shape)."""
    shape = _context.imshape()
    if (int(node[0]) < 0) or (int(node[1]) < 0):
        return False
    if (int(node[0]) >= shape[0]) or (int(node[1]) >= shape[1]):
        return False
    return True


def lineofsight_batch(segments):
    """search.lineofsight for int segments [n,4] = (x0, y0, x1, y1); returns a bool array."""
    p = _context.current_planner()
    seg = np.asarray(segments, dtype=np.int64).reshape(-1, 4).astype(np.int32)
    return p.los(seg).cpu().numpy().astype(bool)


def lineofsight(node1, node2):
    """search.py:35-41."""
    seg = [[int(node1[0]), int(node1[1]), int(node2[0]), int(node2[1])]]
    return bool(lineofsight_batch(seg)[0])


def freespace(node):
    """search.py:26-33: a zero-length line of sight tests exactly one pixel (and its bounds)."""
    if not valid(node):
        return False
    return lineofsight(node, node)


def bresenham(node1, node2):
    """search.py:43-56 (with plotLineLow / plotLineHigh, :58-94): the pixel list, canonical start first."""
    p = _context.current_planner()
    row = [int(node1[0]), int(node1[1]), int(node2[0]), int(node2[1]), 0.0, 0.0, 0.0, 0.0, 1.0]
    return [(int(x), int(y)) for x, y in p.arc_pixels([row])[0]]


def getCirclePoints(xc, yc, p, q):
    """search.py:96-105."""
    pixels = []
    for value1 in [-p, p, -q, q]:
        for value2 in [-p, p, -q, q]:
            if abs(value1) == abs(value2):
                continue
            pixel = (xc + value1, yc + value2)
            if valid(pixel):
                pixels.append(pixel)
    return pixels


def getCircle(center, r, draw=False):
    """search.py:107-142: the in-bounds pixels of the midpoint circle (plus the diagonal-gap pixels), in the reference's
    list order.  draw=True paints them into builtins.imarray like the reference does."""
    p = _context.current_planner()
    row = [0.0, 0.0, 0.0, 0.0, 0.0, float(center[0]), float(center[1]), float(r), 2.0]
    pixels = [(int(x), int(y)) for x, y in p.arc_pixels([row])[0]]
    if draw:
        for item in pixels:
            builtins.imarray[item[1], item[0]] = 0  # search.py:139-141
    return pixels


def getArc(begin, land, u):
    """search.py:144-182: pixels of the edge (begin -> land) driven with control u = (steer, icc, rad, dist)."""
    p = _context.current_planner()
    if u[1] is None:  # search.py:145-146
        return bresenham(begin, land)
    row = [float(begin[0]), float(begin[1]), float(land[0]), float(land[1]), float(u[0]), float(u[1][0]), float(u[1][1]), float(u[2]), 0.0]
    return [(int(x), int(y)) for x, y in p.arc_pixels([row])[0]]


def getneighbors(node):
    """search.py:184-194: the free 8-neighbours `node - delta`, delta in itertools.product([-1, 0, 1], repeat=2) order."""
    cand = [(int(node[0]) - dx, int(node[1]) - dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
    free = lineofsight_batch([[c[0], c[1], c[0], c[1]] for c in cand])  # a zero-length ray is search.freespace
    return [c for c, f in zip(cand, free) if f]


def pathpixels(mainpath):
    """The polyline rasterisation of search.drawpath (search.py:206-211): bresenham of consecutive waypoints, concatenated."""
    pts = [q for q in mainpath if q is not None]
    if len(pts) < 2:
        return []
    p = _context.current_planner()
    rows = [[int(a[0]), int(a[1]), int(b[0]), int(b[1]), 0.0, 0.0, 0.0, 0.0, 1.0] for a, b in zip(pts, pts[1:])]
    return [(int(x), int(y)) for seg in p.arc_pixels(rows) for x, y in seg]


def drawpath(mainpath):
    """search.py:206-219: rasterise the A* / Theta* path and scatter-plot it (only when matplotlib is installed; the
    reference imports it unconditionally)."""
    path = pathpixels(mainpath)
    if path:
        print("Found path")
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            return
        xs = [item[0] for item in path]
        ys = [item[1] for item in path]
        plt.scatter(xs, ys, s=10, c=range(len(path)), cmap="winter")


def astar_batch(queries, thetastar=None, path_cap=None):
    """search.astar for int queries [q,4] = (sx, sy, gx, gy).  Returns a ThetaResult on the host (dict of arrays).
    Paths are returned whole: the first pass stores up to path_cap nodes per query (default 4*(H+W), any-angle paths are
    short), and the queries whose path is longer are searched once more with room for the longest."""
    p = _context.current_planner()
    if thetastar is None:
        thetastar = bool(getattr(builtins, "THETASTAR", True))
    q = np.asarray(queries, dtype=np.int64).reshape(-1, 4).astype(np.int32)
    H, W = p.grid.shape
    cap = int(path_cap) if path_cap else min(H * W, 4 * (H + W))
    h = p.theta(q, thetastar=thetastar, path_cap=cap).host()
    long_ = np.nonzero(h["path_len"] > cap)[0]
    if len(long_):
        cap2 = int(h["path_len"][long_].max())
        h2 = p.theta(q[long_], thetastar=thetastar, path_cap=cap2).host()
        path = np.full((len(q), cap2, 2), -1, np.int32)
        path[:, :cap] = h["path"]
        path[long_] = h2["path"]
        h["path"] = path
    return h


def astar(start, goal):
    """search.py:221-307.  Returns [(x, y), ...] from start to goal, or False."""
    # the reference accepts anything indexable; out-of-range ints must not wrap in int32
    sx, sy, gx, gy = int(start[0]), int(start[1]), int(goal[0]), int(goal[1])
    lim = 2 ** 31 - 1
    if max(abs(sx), abs(sy), abs(gx), abs(gy)) > lim:
        print("Start or goal is not valid. Error.")
        return False
    r = astar_batch([[sx, sy, gx, gy]])
    status = int(r["status"][0])
    if status == 2:
        print("Start or goal is not valid. Error.")       # search.py:223
        return False
    if status == 3:
        print("Start or goal is inside an obstacle. Error.")  # search.py:226
        return False
    if status == 5:
        raise ValueError("attempt to get argmin of an empty sequence")  # search.py:262
    if status == 6:
        raise MemoryError("theta_rrt_b200: search workspace exhausted (" + STATUS_NAMES[status] + ")")
    if status == 1:
        return False                                         # search.py:307
    n = int(r["path_len"][0])
    print("Expanded nodes:", int(r["expanded"][0]))         # search.py:270
    return [(int(x), int(y)) for x, y in r["path"][0, :n]]


def reconstruct(node, pathmap):
    """search.py:196-204 (host helper kept for API compatibility)."""
    path = []
    while True:
        try:
            path.append(node)
            node = pathmap[node]
        except Exception:
            break
    return path[::-1]
