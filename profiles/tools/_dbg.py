import sys, numpy as np
sys.path.insert(0,'/root/repo')
from tests import util
from theta_rrt_b200 import samples, OccupancyGrid, Params, Planner
z = np.load('/root/repo/tests/golden/maps.npz'); free = z['map1'].astype(bool)
nq, K = 8, 801
starts, goals = util.random_queries(free, 64, 1234)
starts, goals = starts[:nq], goals[:nq]
sxy = np.empty((nq, K - 1, 2), np.int32); sth = np.empty((nq, K - 1))
for q in range(nq):
    sxy[q], sth[q] = samples.make_stream(((goals[q, 0], goals[q, 1]), goals[q, 2]), K - 1, q, free.shape)
p = Planner(OccupancyGrid(free), Params(tol_xy=0.0))
for lanes in (32, 16, 8, 4, 2, 1):
    for sched in (0, 1):
        r = p.rrt(starts, goals, sxy, sth, K=K, counters=True, lanes=lanes, schedule=sched).host()
        print(lanes, sched, r["counters"][:, 8], r["counters"][:4, 0])
