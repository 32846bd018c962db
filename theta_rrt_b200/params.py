"""Planner parameters: the `builtins.*` globals of the reference (main.py:15-32)
as a plain record.  `Params.from_builtins()` snapshots whatever the caller put
into `builtins` (the reference's own configuration convention), falling back
to the reference defaults, so code written against the reference keeps working.
"""
from __future__ import annotations

import builtins
from dataclasses import dataclass, fields

from ._lib import CParams


@dataclass
class Params:
    THETASTAR: bool = True        # main.py:15
    bikelength: float = 5         # main.py:18
    FORWARDONLY: bool = True      # main.py:19
    LEFTCONSTRAINT: float = -65   # main.py:20
    RIGHTCONSTRAINT: float = 65   # main.py:21
    frontclearance: float = 2     # main.py:22
    K: int = 300                  # main.py:25
    showtree: bool = False        # main.py:26 (drawing only; ignored)
    maxdrivedist: float = 30      # main.py:27
    tol_xy: float = 10            # main.py:28
    tol_ang: float = 45           # main.py:29
    weightxy: float = .6          # main.py:30
    xystdv: float = 0.4           # main.py:31
    anglestdv: float = 100        # main.py:32

    @classmethod
    def from_builtins(cls, **over) -> "Params":
        vals = {}
        for f in fields(cls):
            if hasattr(builtins, f.name):
                vals[f.name] = getattr(builtins, f.name)
        vals.update(over)
        return cls(**vals)

    def replace(self, **over) -> "Params":
        d = {f.name: getattr(self, f.name) for f in fields(self)}
        d.update(over)
        return Params(**d)

    def to_c(self) -> CParams:
        return CParams(thetastar=int(bool(self.THETASTAR)), forwardonly=int(bool(self.FORWARDONLY)),
                       bikelength=float(self.bikelength), leftconstraint=float(self.LEFTCONSTRAINT),
                       rightconstraint=float(self.RIGHTCONSTRAINT), frontclearance=float(self.frontclearance),
                       maxdrivedist=float(self.maxdrivedist), tol_xy=float(self.tol_xy), tol_ang=float(self.tol_ang),
                       weightxy=float(self.weightxy))
