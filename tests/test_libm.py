"""The deterministic libm (theta_rrt_b200/csrc/trrt_libm.h, host build through the oracle .so) against glibc."""
import ctypes as C
import math

import numpy as np

from oracle import c_oracle as O

dp = C.POINTER(C.c_double)


def _sincos(x):
    lib = O.lib()
    lib.orc_tl_sincos_array.argtypes = [dp, C.c_int64, dp, dp]
    s, c = np.zeros_like(x), np.zeros_like(x)
    lib.orc_tl_sincos_array(x.ctypes.data_as(dp), len(x), s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    return s, c


def _atan2(y, x):
    lib = O.lib()
    lib.orc_tl_atan2_array.argtypes = [dp, dp, C.c_int64, dp]
    o = np.zeros_like(x)
    lib.orc_tl_atan2_array(y.ctypes.data_as(dp), x.ctypes.data_as(dp), len(x), o.ctypes.data_as(dp))
    return o


def test_sincos_close_to_glibc():
    rng = np.random.default_rng(0)
    for lo, hi in ((-math.pi, math.pi), (-8, 8), (-1e5, 1e5), (-1e-3, 1e-3)):
        x = rng.uniform(lo, hi, 200000)
        s, c = _sincos(x)
        gs, gc = np.sin(x), np.cos(x)
        assert np.max(np.abs(s - gs) / np.spacing(np.abs(gs))) <= 1.0
        assert np.max(np.abs(c - gc) / np.spacing(np.abs(gc))) <= 1.0
        assert np.mean(s == gs) > 0.985 and np.mean(c == gc) > 0.985


def test_sincos_special():
    s, c = _sincos(np.array([0.0, -0.0, 1e-300, math.pi / 2, math.pi, 0.7853981633974483]))
    assert s[0] == 0.0 and c[0] == 1.0 and math.copysign(1, s[1]) == -1.0
    assert s[2] == 1e-300 and s[3] == 1.0
    assert s[4] == math.sin(math.pi) and c[4] == -1.0


def test_atan2_close_to_glibc():
    rng = np.random.default_rng(1)
    for sy, sx in ((100, 100), (1, 1), (1e-3, 100), (100, 1e-3)):
        y, x = rng.uniform(-sy, sy, 200000), rng.uniform(-sx, sx, 200000)
        o = _atan2(y, x)
        g = np.array([math.atan2(a, b) for a, b in zip(y, x)])
        assert np.max(np.abs(o - g) / np.spacing(np.abs(g))) <= 1.0
        assert np.mean(o == g) > 0.99


def test_atan2_special_cases():
    lib = O.lib()
    lib.orc_tl_atan2.restype = C.c_double
    lib.orc_tl_atan2.argtypes = [C.c_double, C.c_double]
    for y, x in [(0.0, 1.0), (-0.0, 1.0), (0.0, -1.0), (-0.0, -1.0), (1.0, 0.0), (-1.0, 0.0), (0.0, 0.0), (-0.0, -0.0),
                 (0.0, -0.0), (5.0, 5.0), (-5.0, 5.0), (5.0, -5.0), (-5.0, -5.0), (1e-300, 1e300), (3.0, 4.0)]:
        a, b = lib.orc_tl_atan2(y, x), math.atan2(y, x)
        assert a == b and math.copysign(1, a) == math.copysign(1, b), (y, x)
