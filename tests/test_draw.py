"""Visual adapters: display lists built from the dictionaries / arrays the planner returns.  CPU only."""
import numpy as np

from theta_rrt_b200 import draw


def _tree():
    start = ((5.0, 5.0), 0.0)
    a = ((9.0, 7.0), 40.0)
    b = ((20.0, 7.0), 40.0)
    c = ((3.0, 9.0), -120.0)
    G = {start: [a, c], a: [b], b: [], c: []}
    came = {start: None, a: (start, (-65.0, np.array([5.0, 2.67]), 2.33, 4.6)), b: (a, (0, None, None, 1)),
            c: (start, (65.0, np.array([5.0, 7.33]), 2.33, 5.1))}
    return start, a, b, c, G, came


def test_path_display_list_walks_back_to_the_start():
    start, a, b, c, G, came = _tree()
    dl = draw.path_display_list(b, came)
    kinds = [p[0] for p in dl]
    assert kinds == ["bike", "line", "bike", "arc", "bike"]          # goal bike, straight edge a->b, arc start->a
    assert dl[0][4] == "green" and dl[1][1] == a[0] and dl[1][2] == b[0]
    arc = dl[3]
    assert arc[1] == (5.0, 2.67) and arc[2] == 2.33 and arc[5] == "dodgerblue"  # right/left colour by the sign of the steering angle
    assert draw.path_display_list(None, came) == []


def test_tree_display_list_counts_edges_and_leaves():
    start, a, b, c, G, came = _tree()
    dl = draw.tree_display_list(G, came)
    assert sum(p[0] in ("arc", "line") for p in dl) == 3 and sum(p[0] == "bike" for p in dl) == 2
    assert all(p[-1] in draw.TREE_COLORS.values() for p in dl)


def test_result_display_list_from_arrays():
    K = 4
    host = {"n_nodes": np.array([3]), "node_x": np.array([[5.0, 9.0, 20.0, 0]]), "node_y": np.array([[5.0, 7.0, 7.0, 0]]),
            "node_theta": np.array([[0.0, 40.0, 40.0, 0]]), "parent": np.array([[-1, 0, 1, -1]]),
            "u": np.array([[[np.nan] * 5, [-65.0, 5.0, 2.67, 2.33, 4.6], [0.0, np.nan, np.nan, np.nan, 1.0], [0] * 5]])}
    dl = draw.result_display_list(host, 0)
    assert [p[0] for p in dl] == ["arc", "line"] and dl[1][1] == (9.0, 7.0)


def test_dropin_adapters_return_display_lists_without_matplotlib():
    from theta_rrt_b200 import rrt as R
    start, a, b, c, G, came = _tree()
    assert [p[0] for p in R.drawpath(b, came)] == ["bike", "line", "bike", "arc", "bike"]
    assert len(R.drawtree(start, G, came)) == 5
    assert R.draw_bicycle((1, 2), 30, 5)[0][:3] == ("bike", (1.0, 2.0), 30.0)
    assert R.draw_path_segment(a, b, (0, None, None, 1), bikes=False) == [("line", a[0], b[0], "red")]
