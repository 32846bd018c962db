"""world_size-2 gloo test of the query sharding + result gather used by the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from theta_rrt_b200 import shard


def test_shard_range_partitions():
    for n in (0, 1, 7, 4096, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            r = [shard.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(s for s in shard.shard_sizes(n, w)) - min(s for s in shard.shard_sizes(n, w)) <= 1


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n, rank, world)
    # per-query record = (query id, id squared, rank)
    ids = torch.arange(lo, hi, dtype=torch.int64)
    local = torch.stack([ids, ids * ids, torch.full_like(ids, rank)], dim=1)
    full = shard.gather_records(local, n, dst=0)
    if rank == 0:
        assert full.shape == (n, 3)
        assert torch.equal(full[:, 0], torch.arange(n)) and torch.equal(full[:, 1], torch.arange(n) ** 2)
        assert full[:, 2].tolist() == [0] * (n // 2) + [1] * (n - n // 2)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [9, 64])
def test_gather_records_world2(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)


def test_bind_host_to_device_is_harmless_without_topology():
    """shard.bind_host_to_device: on a box without a GPU (or without NUMA information in sysfs) it reports that and
    leaves the CPU affinity of the process alone."""
    import os
    from theta_rrt_b200 import shard
    before = os.sched_getaffinity(0)
    info = shard.bind_host_to_device(0)
    assert isinstance(info, dict) and "numa_node" in info
    if info["numa_node"] is None:
        assert os.sched_getaffinity(0) == before
    else:
        assert os.sched_getaffinity(0) <= before
        os.sched_setaffinity(0, before)


def _mixed_worker(rank, world, port, tmp):
    """Each rank builds its cfg-5 shard (bench.make_cfg5: RRT and Theta* halves, per-query map ids, sample streams) and
    plans it with the CPU oracle on a small slice; rank 0 gathers the per-query summaries and checks that the union of
    the shards is exactly the unsharded workload, in query order."""
    import numpy as np
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = bench.make_cfg5(rank, world, nq5=64, K5=41)
    one = bench.make_cfg5(0, 1, nq5=64, K5=41)
    lo, hi = c["lo"], c["hi"]
    assert (lo, hi) == shard.shard_range(64, rank, world)
    for k in ("starts", "goals", "sxy", "sth", "sg", "mid_r", "mid_t"):
        assert np.array_equal(c[k], one[k][lo:hi]), k  # a shard is a contiguous slice of the whole, stream seeds included
    from oracle import c_oracle as O
    recs = []
    for q in range(hi - lo):
        r = O.rrt(c["maps"][c["mid_r"][q]], ((c["starts"][q, 0], c["starts"][q, 1]), c["starts"][q, 2]),
                  ((c["goals"][q, 0], c["goals"][q, 1]), c["goals"][q, 2]), c["sxy"][q], c["sth"][q], O.Params(tol_xy=0.0), K=c["K"])
        t = O.astar(c["maps"][c["mid_t"][q]], tuple(c["sg"][q, :2]), tuple(c["sg"][q, 2:]))
        recs.append([lo + q, r["n_nodes"], r["status"], t["status"], t["expanded"] if t["status"] == 0 else 0])
    full = shard.gather_records(torch.tensor(recs, dtype=torch.int64), 64, dst=0)
    if rank == 0:
        assert torch.equal(full[:, 0], torch.arange(64))
        torch.save(full, os.path.join(tmp, "full.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_cfg5_mixed_shards_world2(tmp_path):
    """BASELINE cfg 5 (mixed RRT / Theta* queries over several maps) sharded over two ranks on gloo: same records as the
    unsharded run."""
    import numpy as np
    import bench
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_mixed_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    full = torch.load(os.path.join(str(tmp_path), "full.pt")).numpy()
    from oracle import c_oracle as O
    c = bench.make_cfg5(0, 1, nq5=64, K5=41)
    for q in (0, 13, 31, 32, 63):  # both sides of the shard boundary
        r = O.rrt(c["maps"][c["mid_r"][q]], ((c["starts"][q, 0], c["starts"][q, 1]), c["starts"][q, 2]),
                  ((c["goals"][q, 0], c["goals"][q, 1]), c["goals"][q, 2]), c["sxy"][q], c["sth"][q], O.Params(tol_xy=0.0), K=c["K"])
        assert full[q, 1] == r["n_nodes"] and full[q, 2] == r["status"]
