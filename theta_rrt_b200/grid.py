"""Occupancy grids: `builtins.imarray` (main.py:38-42) as a bit-packed device tensor.

Layout (include/thetarrt.h): per map H rows of wpr=(W+31)//32 32-bit words,
pixel (x, y) = bit (x & 31) of word [y*wpr + (x >> 5)], 1 = free, padding = blocked.
Packing runs on the device (trrt_pack_grid).  `tiles` is a derived copy in overlapping 8x16-pixel strips, one set
per driving axis, for the batched line-of-sight kernel (trrt_tile_grid; layout in include/thetarrt.h, K4b).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def load_png(path) -> np.ndarray:
    """np.array(Image.open(path).convert('1')) -- exactly main.py:38-42 (incl. PIL's dither for grey images)."""
    from PIL import Image
    return np.array(Image.open(path).convert("1"))


class OccupancyGrid:
    """One or several square maps of identical size, resident on one GPU."""

    def __init__(self, free, device="cuda:0"):
        free = np.asarray(free)
        if free.ndim == 2:
            free = free[None]
        if free.ndim != 3:
            raise ValueError("free must be (H, W) or (n_maps, H, W)")
        self.n_maps, self.H, self.W = (int(s) for s in free.shape)
        if self.H != self.W:
            raise ValueError("maps must be square: the reference's bounds test (search.py:21) compares x with "
                             "shape[0] and y with shape[1]")
        self.device = torch.device(device)
        lib = _lib.load()
        self.wpr = (self.W + 31) // 32
        words = self.n_maps * lib.trrt_grid_words(self.H, self.W)
        with torch.cuda.device(self.device):
            d_free = torch.from_numpy(np.ascontiguousarray(free.astype(bool).astype(np.uint8))).to(self.device)
            self.bits = torch.empty(words, dtype=torch.int32, device=self.device)
            st = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(lib.trrt_pack_grid(d_free.data_ptr(), self.n_maps, self.H, self.W, self.bits.data_ptr(), st),
                       "trrt_pack_grid")
            torch.cuda.current_stream(self.device).synchronize()  # d_free may be freed after return
        self.nbytes = words * 4
        self._tiles = None

    @property
    def tiles(self):
        """Strip copy of the grid (include/thetarrt.h K4b, 4x the packed rows), built on first use by trrt_tile_grid
        from the packed rows."""
        if self._tiles is None:
            lib = _lib.load()
            with torch.cuda.device(self.device):
                t = torch.empty(self.n_maps * lib.trrt_tile_words(self.H, self.W), dtype=torch.int64, device=self.device)
                st = torch.cuda.current_stream(self.device).cuda_stream
                _lib.check(lib.trrt_tile_grid(self.bits.data_ptr(), self.n_maps, self.H, self.W, t.data_ptr(), st),
                           "trrt_tile_grid")
                self._tiles_ready = torch.cuda.Event()
                self._tiles_ready.record(torch.cuda.current_stream(self.device))
            self._tiles = t
        else:  # built on some stream earlier: the caller's stream must not run ahead of that build
            torch.cuda.current_stream(self.device).wait_event(self._tiles_ready)
        return self._tiles

    @property
    def shape(self):
        return (self.H, self.W)
