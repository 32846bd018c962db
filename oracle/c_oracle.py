"""ctypes front-end of the C parity oracle (oracle/trrt_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
theta_rrt_b200 never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)

STATUS_NAMES = {0: "OK_FOUND", 1: "OK_NOT_FOUND", 2: "ERR_ENDPOINT_INVALID", 3: "ERR_ENDPOINT_BLOCKED",
                4: "ERR_REF_RAISES_DRIVE_NONE", 5: "ERR_REF_RAISES_ARGMIN_EMPTY", 6: "ERR_CAPACITY"}


class Params(C.Structure):
    """Mirror of orc_params; defaults are main.py:15-32."""
    _fields_ = [("thetastar", C.c_int), ("bikelength", C.c_double), ("forwardonly", C.c_int),
                ("leftconstraint", C.c_double), ("rightconstraint", C.c_double), ("frontclearance", C.c_double),
                ("maxdrivedist", C.c_double), ("tol_xy", C.c_double), ("tol_ang", C.c_double), ("weightxy", C.c_double)]

    def __init__(self, **kw):
        d = dict(thetastar=1, bikelength=5, forwardonly=1, leftconstraint=-65, rightconstraint=65, frontclearance=2,
                 maxdrivedist=30, tol_xy=10, tol_ang=45, weightxy=.6)
        d.update(kw)
        super().__init__(**d)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "trrt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_l2norm.restype = C.c_double
        _lib.orc_l2norm.argtypes = [C.c_double] * 4
        _lib.orc_anglediff.restype = C.c_double
        _lib.orc_anglediff.argtypes = [C.c_double] * 2
        _lib.orc_anglebetween.restype = C.c_double
        _lib.orc_anglebetween.argtypes = [C.c_double] * 4
        _lib.orc_standardangle.restype = C.c_double
        _lib.orc_standardangle.argtypes = [C.c_double]
        _lib.orc_rotz.argtypes = [C.c_double] * 3 + [c_dp]
        _lib.orc_lineofsight.argtypes = [c_u8p, C.c_int, C.c_int] + [C.c_double] * 4
        _lib.orc_lineofsight_batch.argtypes = [c_u8p, C.c_int, C.c_int, c_i32p, C.c_int64, c_u8p, C.c_int]
        _lib.orc_bresenham.restype = C.c_int64
        _lib.orc_bresenham.argtypes = [C.c_long] * 4 + [c_i32p, C.c_int64]
        _lib.orc_getcircle.restype = C.c_int64
        _lib.orc_getcircle.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, c_i32p, C.c_int64]
        _lib.orc_getarc.restype = C.c_int64
        _lib.orc_getarc.argtypes = [C.c_int, C.c_int] + [C.c_double] * 8 + [C.c_int, c_i32p, C.c_int64]
        _lib.orc_nearest.restype = C.c_int64
        _lib.orc_nearest.argtypes = [c_dp, c_dp, C.c_int64, C.c_double, C.c_double]
        _lib.orc_nearest_batch.argtypes = [c_dp, c_dp, C.c_int64, c_i32p, C.c_int64, c_i32p, C.c_int]
        _lib.orc_steer.restype = C.c_int
        _lib.orc_steer.argtypes = [C.POINTER(Params)] + [C.c_double] * 6 + [c_dp]
        _lib.orc_drive.argtypes = [C.POINTER(Params)] + [C.c_double] * 8 + [c_dp]
        _lib.orc_rrt.restype = C.c_int
        _lib.orc_rrt.argtypes = [c_u8p, C.c_int, C.c_int, C.POINTER(Params), C.c_int, c_dp, c_dp, c_i32p, c_dp,
                                 c_dp, c_dp, c_dp, c_i32p, c_dp, c_i32p, c_i32p, c_u8p, c_i32p, c_i32p, c_i32p,
                                 c_u8p, C.c_int64, c_i64p, C.c_double, c_i32p]
        _lib.orc_rrt_batch.argtypes = [c_u8p, C.c_int, C.c_int, C.POINTER(Params), C.c_int, C.c_int64, c_dp, c_dp,
                                       c_i32p, c_dp, c_i32p, c_dp, c_i32p, c_i32p, c_i32p, c_i32p, C.c_int]
        _lib.orc_findnearest.restype = C.c_int
        _lib.orc_findnearest.argtypes = [C.POINTER(Params), c_dp, c_dp, c_dp, c_i32p, c_i32p, C.c_int64, C.c_int,
                                         c_dp, c_dp]
        _lib.orc_astar.restype = C.c_int
        _lib.orc_astar.argtypes = [c_u8p, C.c_int, C.c_int, C.c_int] + [C.c_int] * 4 + [c_i32p, C.c_int64, c_i64p,
                                                                                     c_i64p, c_dp, c_u8p, C.c_int64,
                                                                                     c_i64p, c_i64p]
        _lib.orc_astar_batch.argtypes = [c_u8p, C.c_int, C.c_int, C.c_int, C.c_int64, c_i32p, c_i32p, c_dp, c_i64p,
                                         c_i64p, c_i64p, C.c_int]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _grid(free):
    g = np.ascontiguousarray(np.asarray(free, dtype=bool).astype(np.uint8))
    return g, g.shape[0], g.shape[1]


# ------------------------------------------------------------------ scalar helpers
def l2norm(a, b):
    return lib().orc_l2norm(a[0], a[1], b[0], b[1])


def anglediff(a, b):
    return lib().orc_anglediff(a, b)


def anglebetween(v1, v2):
    return lib().orc_anglebetween(v1[0], v1[1], v2[0], v2[1])


def rotz(deg, v):
    out = np.zeros(2)
    lib().orc_rotz(deg, v[0], v[1], _p(out, c_dp))
    return out


def lineofsight(free, a, b):
    g, H, W = _grid(free)
    return bool(lib().orc_lineofsight(_p(g, c_u8p), H, W, a[0], a[1], b[0], b[1]))


def lineofsight_batch(free, seg, threads=1):
    g, H, W = _grid(free)
    seg = np.ascontiguousarray(seg, dtype=np.int32).reshape(-1, 4)
    out = np.zeros(len(seg), np.uint8)
    lib().orc_lineofsight_batch(_p(g, c_u8p), H, W, _p(seg, c_i32p), len(seg), _p(out, c_u8p), threads)
    return out.astype(bool)


def bresenham(a, b):
    cap = abs(int(a[0]) - int(b[0])) + abs(int(a[1]) - int(b[1])) + 2
    out = np.zeros((cap, 2), np.int32)
    n = lib().orc_bresenham(int(a[0]), int(a[1]), int(b[0]), int(b[1]), _p(out, c_i32p), cap)
    return [tuple(map(int, p)) for p in out[:n]]


def getcircle(shape, center, r):
    cap = 16 * (shape[0] + shape[1]) + 64
    while True:
        out = np.zeros((cap, 2), np.int32)
        n = lib().orc_getcircle(shape[0], shape[1], center[0], center[1], r, _p(out, c_i32p), cap)
        if n <= cap:
            return [tuple(map(int, p)) for p in out[:n]]
        cap = int(n)


def getarc(shape, begin, land, u):
    """u = (steer, icc|None, rad|None, dist) like the reference."""
    straight = u[1] is None
    icc = (np.nan, np.nan) if straight else u[1]
    rad = np.nan if straight else u[2]
    cap = 16 * (shape[0] + shape[1]) + 64
    while True:
        out = np.zeros((cap, 2), np.int32)
        n = lib().orc_getarc(shape[0], shape[1], begin[0], begin[1], land[0], land[1], u[0], icc[0], icc[1], rad,
                             int(straight), _p(out, c_i32p), cap)
        if n <= cap:
            return [tuple(map(int, p)) for p in out[:n]]
        cap = int(n)


def nearest(x, y, q):
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    return int(lib().orc_nearest(_p(x, c_dp), _p(y, c_dp), len(x), q[0], q[1]))


def nearest_batch(x, y, qxy, threads=1):
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    qxy = np.ascontiguousarray(qxy, np.int32).reshape(-1, 2)
    out = np.zeros(len(qxy), np.int32)
    lib().orc_nearest_batch(_p(x, c_dp), _p(y, c_dp), len(x), _p(qxy, c_i32p), len(qxy), _p(out, c_i32p), threads)
    return out


def steer(origin, theta, goal, thetagoal, params=None):
    P = params or Params()
    out = np.zeros(8)
    straight = lib().orc_steer(C.byref(P), origin[0], origin[1], theta, goal[0], goal[1], thetagoal, _p(out, c_dp))
    return dict(x=out[0], y=out[1], theta=out[2], steer=out[3], icc=(out[4], out[5]), rad=out[6], dist=out[7],
                straight=bool(straight))


def drive(origin, theta, u, params=None):
    P = params or Params()
    out = np.zeros(3)
    lib().orc_drive(C.byref(P), origin[0], origin[1], theta, u[0], u[1][0], u[1][1], u[2], u[3], _p(out, c_dp))
    return out


# ------------------------------------------------------------------ RRT
def rrt(free, start, goal, sxy, sth, params=None, K=None, log_los=True, audit_eps=0.0):
    """start/goal: ((x,y),theta).  sxy int32 [K-1,2], sth float64 [K-1]."""
    P = params or Params()
    g, H, W = _grid(free)
    sxy = np.ascontiguousarray(sxy, np.int32).reshape(-1, 2)
    sth = np.ascontiguousarray(sth, np.float64)
    if K is None:
        K = len(sth) + 1
    assert len(sth) >= K - 1
    st = np.array([start[0][0], start[0][1], start[1]], np.float64)
    gl = np.array([goal[0][0], goal[0][1], goal[1]], np.float64)
    nx, ny, nth = np.zeros(K), np.zeros(K), np.zeros(K)
    parent = np.full(K, -1, np.int32)
    u = np.full((K, 5), np.nan)
    it_near = np.full(max(K - 1, 1), -1, np.int32)
    it_new = np.full(max(K - 1, 1), -1, np.int32)
    it_code = np.full(max(K - 1, 1), 255, np.uint8)
    n_nodes, sol, iters = C.c_int32(0), C.c_int32(-1), C.c_int32(0)
    los_cap = 2 * K + 8
    los = np.zeros(los_cap, np.uint8) if log_los else None
    n_los = C.c_int64(0)
    first_amb = C.c_int32(-1)
    status = lib().orc_rrt(_p(g, c_u8p), H, W, C.byref(P), K, _p(st, c_dp), _p(gl, c_dp), _p(sxy, c_i32p),
                           _p(sth, c_dp), _p(nx, c_dp), _p(ny, c_dp), _p(nth, c_dp), _p(parent, c_i32p), _p(u, c_dp),
                           _p(it_near, c_i32p), _p(it_new, c_i32p), _p(it_code, c_u8p), C.byref(n_nodes),
                           C.byref(sol), C.byref(iters), _p(los, c_u8p), los_cap, C.byref(n_los), float(audit_eps), C.byref(first_amb))
    n = n_nodes.value
    return dict(status=status, n_nodes=n, sol=sol.value, iters=iters.value, x=nx[:n], y=ny[:n], theta=nth[:n],
                parent=parent[:n], u=u[:n], it_near=it_near[:K - 1], it_new=it_new[:K - 1], it_code=it_code[:K - 1],
                los=(los[:n_los.value].astype(bool) if log_los else None), n_los=n_los.value,
                first_ambiguous=first_amb.value)


def rrt_batch(free, starts, goals, sxy, sth, K, params=None, threads=1, want_nodes=True):
    """starts/goals float64 [q,3]; sxy int32 [q,K-1,2]; sth float64 [q,K-1]."""
    P = params or Params()
    g, H, W = _grid(free)
    starts = np.ascontiguousarray(starts, np.float64).reshape(-1, 3)
    goals = np.ascontiguousarray(goals, np.float64).reshape(-1, 3)
    nq = len(starts)
    sxy = np.ascontiguousarray(sxy, np.int32).reshape(nq, K - 1, 2)
    sth = np.ascontiguousarray(sth, np.float64).reshape(nq, K - 1)
    parent = np.full((nq, K), -1, np.int32) if want_nodes else None
    xyz = np.full((nq, K, 3), np.nan) if want_nodes else None
    n_nodes = np.zeros(nq, np.int32)
    sol = np.zeros(nq, np.int32)
    status = np.zeros(nq, np.int32)
    iters = np.zeros(nq, np.int32)
    lib().orc_rrt_batch(_p(g, c_u8p), H, W, C.byref(P), K, nq, _p(starts, c_dp), _p(goals, c_dp), _p(sxy, c_i32p),
                        _p(sth, c_dp), _p(parent, c_i32p), _p(xyz, c_dp), _p(n_nodes, c_i32p), _p(sol, c_i32p),
                        _p(status, c_i32p), _p(iters, c_i32p), threads)
    return dict(parent=parent, nodes=xyz, n_nodes=n_nodes, sol=sol, status=status, iters=iters)


def findnearest(x, y, theta, edge_parent, edge_child, goal, params=None):
    P = params or Params()
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    theta = np.ascontiguousarray(theta, np.float64)
    ep = np.ascontiguousarray(edge_parent, np.int32)
    ec = np.ascontiguousarray(edge_child, np.int32)
    gl = np.array([goal[0][0], goal[0][1], goal[1]], np.float64)
    d = C.c_double(0)
    i = lib().orc_findnearest(C.byref(P), _p(x, c_dp), _p(y, c_dp), _p(theta, c_dp), _p(ep, c_i32p), _p(ec, c_i32p),
                              len(ep), len(x), _p(gl, c_dp), C.byref(d))
    return i, d.value


# ------------------------------------------------------------------ Theta*
def astar(free, start, goal, thetastar=True, log_los=True):
    g, H, W = _grid(free)
    cap = H * W + 1
    path = np.zeros((cap, 2), np.int32)
    plen, exp, nl, npush = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
    cost = C.c_double(0)
    los_cap = H * W + 8
    los = np.zeros(los_cap, np.uint8) if log_los else None
    status = lib().orc_astar(_p(g, c_u8p), H, W, int(thetastar), int(start[0]), int(start[1]), int(goal[0]),
                             int(goal[1]), _p(path, c_i32p), cap, C.byref(plen), C.byref(exp), C.byref(cost),
                             _p(los, c_u8p), los_cap, C.byref(nl), C.byref(npush))
    return dict(status=status, path=[tuple(map(int, p)) for p in path[:plen.value]] if status == 0 else False,
                expanded=exp.value, cost=cost.value, los=(los[:nl.value].astype(bool) if log_los else None),
                n_los=nl.value, pushes=npush.value)


def astar_batch(free, sg, thetastar=True, threads=1):
    g, H, W = _grid(free)
    sg = np.ascontiguousarray(sg, np.int32).reshape(-1, 4)
    nq = len(sg)
    status = np.zeros(nq, np.int32)
    cost = np.zeros(nq)
    expanded = np.zeros(nq, np.int64)
    plen = np.zeros(nq, np.int64)
    nlos = np.zeros(nq, np.int64)
    lib().orc_astar_batch(_p(g, c_u8p), H, W, int(thetastar), nq, _p(sg, c_i32p), _p(status, c_i32p), _p(cost, c_dp),
                          _p(expanded, c_i64p), _p(plen, c_i64p), _p(nlos, c_i64p), threads)
    return dict(status=status, cost=cost, expanded=expanded, path_len=plen, n_los=nlos)
