"""Per-process planner context for the drop-in modules (`search`, `rrt`, `main`).

The reference keeps its map in `builtins.imarray` and its parameters in
`builtins.*` (main.py:15-44) and reads them at call time.  The drop-in modules
do the same: each call snapshots `builtins.*` into a Params record and re-packs
the occupancy grid when `builtins.imarray` was replaced or modified.
"""
from __future__ import annotations

import builtins
import os

import numpy as np

from .grid import OccupancyGrid
from .params import Params
from .planner import Planner

_cache = {"key": None, "planner": None}


def device() -> str:
    return os.environ.get("THETA_RRT_DEVICE", "cuda:0")


def current_planner() -> Planner:
    if not hasattr(builtins, "imarray"):
        raise NameError("name 'imarray' is not defined (set builtins.imarray like main.py:42)")
    im = np.asarray(builtins.imarray)
    key = (id(builtins.imarray), im.shape, hash(im.tobytes()))
    if _cache["key"] != key:
        _cache["planner"] = Planner(OccupancyGrid(im.astype(bool), device=device()))
        _cache["key"] = key
    p = _cache["planner"]
    p.params = Params.from_builtins()
    return p


def imshape():
    return np.asarray(builtins.imarray).shape
