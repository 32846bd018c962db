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


def astar_batch(queries, thetastar=None):
    """search.astar for int queries [q,4] = (sx, sy, gx, gy).  Returns a ThetaResult on the host (dict of arrays)."""
    p = _context.current_planner()
    if thetastar is None:
        thetastar = bool(getattr(builtins, "THETASTAR", True))
    q = np.asarray(queries, dtype=np.int64).reshape(-1, 4)
    H, W = p.grid.shape
    res = p.theta(q.astype(np.int32), thetastar=thetastar, path_cap=H * W)
    return res.host()


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
