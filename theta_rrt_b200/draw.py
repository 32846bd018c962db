"""Visual adapters: the reference's drawing calls -- rrt.draw_bicycle (rrt.py:16-40), rrt.drawpath
(:79-98), rrt.drawtree (:100-106), rrt.draw_path_segment (:224-270) -- fed from the (sol, G, cameFrom) dictionaries the
drop-in `rrt.rrt` returns, or straight from the device result tensors (`tree_display_list`).

Drawing has no effect on planning results, so nothing here runs on the GPU.  The functions first build a DISPLAY LIST of
plain primitives

    ("line",  (x0, y0), (x1, y1), color)                          a straight edge
    ("arc",   (cx, cy), radius, start_deg, sweep_deg, color)      a driven arc: matplotlib Arc(angle=start, theta1=0, theta2=sweep)
    ("bike",  (x, y), theta_deg, steer_deg, color)                a bicycle glyph

and `render()` paints it with matplotlib when that is installed (it is not in the build image; the reference imports it
unconditionally).  Without matplotlib the display list is still returned, so callers and tests can inspect what would be drawn.
"""
from __future__ import annotations

import math

import numpy as np

SEGMENT_COLORS = {"left": "magenta", "right": "dodgerblue", "straight": "red", "bike": "dodgerblue"}
TREE_COLORS = {"left": "lightgray", "right": "silver", "straight": "silver", "bike": "darkgray"}


def _wrap(a):
    while a > 180:
        a -= 360
    while a <= -180:
        a += 360
    return a


def _angle(v1, v2):
    """Angle from v2 to v1 in the image's y-down convention (the reference's anglebetween, rrt.py:73-77), degrees."""
    return _wrap(math.degrees(math.atan2(v2[0] * v1[1] - v1[0] * v2[1], v1[0] * v2[0] + v1[1] * v2[1])))


def segment_primitives(bike1, bike2, u, colors=SEGMENT_COLORS, bikes=True, bikelength=5.0, forwardonly=True):
    """Primitives of one edge bike1 -> bike2 driven with control u = (steer, icc, rad, dist) (rrt.py:224-270)."""
    out = []
    icc = u[1]
    if icc is not None:
        rad = float(u[2])
        to2 = (bike2[0][0] - icc[0], bike2[0][1] - icc[1])
        to1 = (bike1[0][0] - icc[0], bike1[0][1] - icc[1])
        a2, a1, span = _angle((1, 0), to2), _angle((1, 0), to1), _angle(to2, to1)
        # which way the arc is swept follows the sign of the steering angle, flipped when the bike would drive backwards
        th = math.radians(bike1[1])
        frame = (bikelength * math.cos(th), bikelength * math.sin(th))
        sv = (frame[0] + bike1[0][0] - icc[0], frame[1] + bike1[0][1] - icc[1])
        sv = (-sv[1], sv[0])  # rotated by 90 degrees
        steer = _wrap(-_angle(frame, sv))
        flip = bool(forwardonly) and (steer > 90 or steer < -90)
        color = colors["right"] if u[0] < 0 else colors["left"]
        if u[0] < 0 or flip:
            out.append(("arc", (float(icc[0]), float(icc[1])), rad, -a2, -span, color))
        else:
            out.append(("arc", (float(icc[0]), float(icc[1])), rad, -a1, span, color))
    else:
        out.append(("line", (float(bike1[0][0]), float(bike1[0][1])), (float(bike2[0][0]), float(bike2[0][1])), colors["straight"]))
    if bikes:
        out.append(("bike", (float(bike1[0][0]), float(bike1[0][1])), float(bike1[1]), float(u[0]), colors["bike"]))
    return out


def path_display_list(solution, camefrom, **kw):
    """rrt.drawpath (rrt.py:79-98): the solution bike in green, then every edge back to the start."""
    out = []
    if solution is None:
        return out
    out.append(("bike", (float(solution[0][0]), float(solution[0][1])), float(solution[1]), 0.0, "green"))
    a = solution
    while camefrom.get(a) is not None:
        b, u = camefrom[a]
        out += segment_primitives(b, a, u, **kw)
        a = b
    return out


def tree_display_list(graph, camefrom, **kw):
    """rrt.drawtree (rrt.py:100-106): every edge of G in grey, a grey bike on every leaf."""
    out = []
    for parent, children in graph.items():
        for child in children:
            out += segment_primitives(parent, child, camefrom[child][1], colors=TREE_COLORS, bikes=False, **kw)
        if not children:
            out.append(("bike", (float(parent[0][0]), float(parent[0][1])), float(parent[1]), 0.0, "darkgray"))
    return out


def result_display_list(host, q=0, **kw):
    """The tree of query q straight from RrtResult.host() arrays (no dictionaries): one primitive per node with a parent."""
    n = int(host["n_nodes"][q])
    x, y, th, par, u = host["node_x"][q], host["node_y"][q], host["node_theta"][q], host["parent"][q], host["u"][q]
    out = []
    for i in range(1, n):
        p = int(par[i])
        if p < 0:
            continue
        ui = u[i]
        uu = (float(ui[0]), None, None, float(ui[4])) if np.isnan(ui[3]) else (float(ui[0]), (float(ui[1]), float(ui[2])), float(ui[3]), float(ui[4]))
        out += segment_primitives(((float(x[p]), float(y[p])), float(th[p])), ((float(x[i]), float(y[i])), float(th[i])), uu,
                                  colors=TREE_COLORS, bikes=False, **kw)
    return out


def render(display_list, ax=None, bikelength=5.0):
    """Paint a display list with matplotlib.  Raises ImportError when matplotlib is not installed."""
    import matplotlib.patches as patches
    import matplotlib.pyplot as plt
    ax = ax or plt.gca()
    for prim in display_list:
        if prim[0] == "line":
            ax.plot([prim[1][0], prim[2][0]], [prim[1][1], prim[2][1]], linestyle="--", color=prim[3])
        elif prim[0] == "arc":
            _, c, r, start, sweep, color = prim
            ax.add_patch(patches.Arc(c, 2 * r, 2 * r, angle=start, theta1=0, theta2=sweep, edgecolor=color, linestyle="--", zorder=10))
        elif prim[0] == "bike":
            _, (x, y), theta, alpha, color = prim
            t = math.radians(theta)
            fx, fy = bikelength * math.cos(t), bikelength * math.sin(t)
            ax.plot([x, x + fx], [y, y + fy], color=color, linewidth=5)
            ax.quiver(x, y, fx / 2, fy / 2, facecolor="red", edgecolor="black", linewidth=0.5, headwidth=2.5, zorder=10,
                      angles="xy", scale_units="xy", scale=1)
            for da, fc in ((0.0, "yellow"), (90.0, "cyan")):  # front wheel direction and its normal
                w = math.radians(theta + alpha + da)
                ax.quiver(x + fx, y + fy, bikelength / 2 * math.cos(w), bikelength / 2 * math.sin(w), facecolor=fc, edgecolor="black",
                          linewidth=0.5, headwidth=2.5, zorder=10, angles="xy", scale_units="xy", scale=1)
    return ax
