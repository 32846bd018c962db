"""CPU model of the control logic of the fused RRT kernel's speculative window (csrc/trrt_rrt.cuh, schedule 0):
snapshot nearest -> tentative expansion -> forward fold of the inserted nodes (window minimum / equality flags) ->
parallel commit of the lanes before the first one whose nearest node is a node of its own window; the next window
starts at that lane.

The exactness of that schedule does not depend on WHAT an expansion computes, only on it being a deterministic function
of (nearest node, sample, map).  So the model runs a toy expansion on an integer lattice (many exact distance ties,
duplicate nodes, samples that coincide with nodes, rejections) and must reproduce the sequential loop of rrt.py:141-201
node for node.  The CUDA kernel itself is compared with the oracle bit for bit in test_gpu_parity.py.
"""
import numpy as np
import pytest

BLOCKED, IN_TREE, STEER, ARC, NEW, EXISTING = "blocked", "in_tree", "steer", "arc", "new", "existing"


def blocked(s):
    return (s[0] * 7 + s[1] * 13) % 11 == 0  # freespace(qrand) is false (rrt.py:148)


def expand(node, s):
    """Toy stand-in for steer / clearance / re-drive / edge test: deterministic in (node, sample)."""
    h = (node[0] * 31 + node[1] * 17 + s[0] * 5 + s[1] * 3 + s[2]) % 10
    if h == 0:
        return STEER, None
    if h == 1:
        return ARC, None
    # move up to 3 lattice steps towards the sample; heading copied from the sample: nodes equal to samples and
    # duplicates of existing nodes both happen
    dx, dy = s[0] - node[0], s[1] - node[1]
    step = lambda d: max(-3, min(3, d))  # noqa: E731
    return "accept", (node[0] + step(dx), node[1] + step(dy), s[2] if h < 6 else node[2])


def d2(s, n):
    return (s[0] - n[0]) ** 2 + (s[1] - n[1]) ** 2


def sequential(start, samples):
    nodes, parent, index, codes, nears = [start], [-1], {start: 0}, [], []
    for s in samples:
        near = -1
        if blocked(s):
            code = BLOCKED
        elif s in index:
            code = IN_TREE  # rrt.py:151
        else:
            near = min(range(len(nodes)), key=lambda i: (d2(s, nodes[i]), i))  # first minimum (np.argmin)
            kind, w = expand(nodes[near], s)
            if kind != "accept":
                code = kind
            else:
                idx = index.get(w)
                if idx is None:
                    idx = len(nodes)
                    nodes.append(w); parent.append(-1); index[w] = idx
                    code = NEW
                else:
                    code = EXISTING
                if idx != near:
                    parent[idx] = near  # rrt.py:187-188
        codes.append(code); nears.append(near)
    return nodes, parent, codes, nears


def windowed(start, samples, G, stats):
    """The kernel's schedule, lane by lane, with the same state per lane and the same folds: a window of up to G
    iterations against a snapshot; the inserting lanes are folded forward in iteration order; the window commits up to the
    first lane whose nearest node was inserted in this very window, and the next window starts at that iteration."""
    nodes, parent, index, codes, nears = [start], [-1], {start: 0}, [], []
    k0 = 0
    while k0 < len(samples):
        win = samples[k0:k0 + G]
        L = len(win)
        n0 = len(nodes)
        snap = dict(index)  # the index as the window starts
        # ---- phase A (against the snapshot)
        pre = [blocked(s) for s in win]
        q_in_tree = [(not pre[j]) and win[j] in snap for j in range(L)]
        bd, near, e, exist = [None] * L, [-1] * L, [None] * L, [-1] * L
        for j in range(L):
            if not pre[j] and not q_in_tree[j]:
                near[j] = min(range(n0), key=lambda i: (d2(win[j], nodes[i]), i))
                bd[j] = d2(win[j], nodes[near[j]])
                e[j] = expand(nodes[near[j]], win[j])
                if e[j][0] == "accept":
                    exist[j] = snap.get(e[j][1], -1)
        # ---- pass: fold the inserting lanes forward until a lane has moved
        moved = [False] * L
        done, last = [], -1
        while True:
            live = [not pre[l] and not q_in_tree[l] for l in range(L)]
            stop = [l for l in range(L) if live[l] and moved[l]]
            ins = [l for l in range(L) if l > last and live[l] and not moved[l] and e[l][0] == "accept" and exist[l] < 0]
            fs = stop[0] if stop else L
            ni = ins[0] if ins else L
            if ni >= fs:
                break
            idx_i = n0 + len(done)
            done.append(ni)
            last = ni
            w = e[ni][1]
            for l in range(ni + 1, L):  # every later lane folds the node
                if bd[l] is not None and d2(win[l], w) < bd[l]:
                    moved[l] = True  # strictly nearer than the snapshot winner: ties stay with the lower index
                if win[l] == w:
                    q_in_tree[l] = True
                if exist[l] < 0 and e[l] is not None and e[l][0] == "accept" and e[l][1] == w:
                    exist[l] = idx_i
        done = [l for l in done if l < fs]
        stats["windows"] += 1
        stats["lanes"] += fs
        assert fs >= 1  # lane 0 has no predecessor in its window
        # ---- commit of lanes [0, fs), all at once
        target = {}
        for l in range(fs):
            code, near_j = None, -1
            if pre[l]:
                code = BLOCKED
            elif q_in_tree[l]:
                code = IN_TREE
            else:
                near_j = near[l]
                kind, w = e[l]
                if kind != "accept":
                    code = kind
                else:
                    if exist[l] >= 0:
                        idx, code = exist[l], EXISTING
                    else:
                        idx, code = n0 + done.index(l), NEW
                    if idx != near_j:
                        target[idx] = near_j  # the last lane that targets a node wins (rrt.py:188)
            codes.append(code); nears.append(near_j)
        for l in done:
            w = e[l][1]
            nodes.append(w); parent.append(-1); index[w] = len(nodes) - 1
        for idx, nr in target.items():
            parent[idx] = nr
        k0 += fs
    return nodes, parent, codes, nears


def make_samples(rng, side, n, clustered):
    if clustered:  # like rand_conf (rrt.py:53-68): normal around the goal, clipped -- consecutive samples crowd together
        g = np.array([side * 3 // 4, side * 3 // 4])
        xy = np.clip(np.rint(rng.normal(g, side * 0.12, size=(n, 2))), 0, side - 1).astype(int)
    else:
        xy = rng.integers(0, side, size=(n, 2))
    return [(int(x), int(y), int(t)) for (x, y), t in zip(xy, rng.integers(0, 3, n))]


@pytest.mark.parametrize("G", [2, 4, 8, 32])
@pytest.mark.parametrize("seed,side,clustered", [(0, 14, False), (1, 30, False), (2, 60, True), (3, 90, True), (4, 40, True)])
def test_window_schedule_equals_sequential_loop(G, seed, side, clustered):
    rng = np.random.default_rng(seed)
    samples = make_samples(rng, side, 700, clustered)
    start = (side // 8, side // 8, 0)
    ref = sequential(start, samples)
    stats = {"windows": 0, "lanes": 0}
    got = windowed(start, samples, G, stats)
    assert got[0] == ref[0], "nodes"
    assert got[1] == ref[1], "parents"
    assert got[2] == ref[2], "iteration codes"
    assert got[3] == ref[3], "nearest indices"
    assert {NEW, EXISTING, BLOCKED, STEER, ARC} <= set(ref[2]), set(ref[2])  # the toy exercises the outcomes
    if clustered and G == 32:
        assert stats["lanes"] < G * (stats["windows"] - 1), stats  # some windows were cut short by a moved lane


def test_window_model_sees_samples_on_nodes():
    """`qrand in G` (rrt.py:151) against nodes inserted earlier in the same window."""
    rng = np.random.default_rng(9)
    samples = make_samples(rng, 10, 900, False)
    ref = sequential((5, 5, 0), samples)
    assert IN_TREE in ref[2]
    for G in (4, 32):
        got = windowed((5, 5, 0), samples, G, {"windows": 0, "lanes": 0})
        assert got == ref
