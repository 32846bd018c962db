"""theta_rrt_b200 -- B200-native implementation of theta-rrt's planning inner loop.

Public surface:
  Planner, OccupancyGrid, Params          batched device API (planner.py, grid.py, params.py)
  samples                                  injected sample streams (rand_conf restated)
  search, rrt, main                        drop-in modules with the reference's names
The CUDA library (libthetarrt.so, C ABI in include/thetarrt.h) must be built
(`python -m theta_rrt_b200.build`); there is no CPU fallback.
"""
from .params import Params  # noqa: F401
from . import samples  # noqa: F401


def __getattr__(name):
    # torch-dependent pieces are imported lazily so that `import theta_rrt_b200` stays cheap
    if name in ("Planner", "RrtResult", "ThetaResult"):
        from . import planner
        return getattr(planner, name)
    if name in ("OccupancyGrid", "load_png"):
        from . import grid
        return getattr(grid, name)
    if name in ("search", "rrt", "main", "build", "shard"):
        import importlib
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
