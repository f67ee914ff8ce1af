"""N>1 host logic on CPU: contiguous sharding of a batch of matches over ranks + the result gather,
exercised with torch.distributed (gloo, world_size 2).  The per-match path has no collective."""
import os

import numpy as np
import torch.multiprocessing as mp

from lidar_slam_b200 import batch


def test_shard_range_partitions():
    for n in (0, 1, 7, 4000, 1024):
        for w in (1, 2, 3, 4, 8):
            ranges = [batch.shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1 and sizes == batch.shard_sizes(n, w)


def _worker(rank, world, port, n_items, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = batch.shard_range(n_items, rank, world)
    # each rank "matches" its block: row = (score, match id, rank)
    rows = np.stack([np.arange(lo, hi) * 0.5, np.arange(lo, hi), np.full(hi - lo, rank)], axis=1).astype(np.float64)
    table = batch.gather_rows(rows, n_items, dist)
    dist.barrier()
    if rank == 0:
        q.put(table)
    dist.destroy_process_group()


def test_gather_rows_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_items = 37
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    table = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert table.shape == (n_items, 3)
    assert np.array_equal(table[:, 1], np.arange(n_items))
    assert np.array_equal(table[:, 2], np.array([0] * 19 + [1] * 18))
    assert batch.best_hypothesis(table, 0) == n_items - 1


def test_pose6_to_matrix_convention():
    """batch.pose6_to_matrix (used by relocalize) = the generator's and the oracle's Rx Ry Rz convention."""
    import numpy as np
    from lidar_slam_b200 import batch, synth
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    for _ in range(50):
        p = rng.uniform(-3, 3, 6)
        T = batch.pose6_to_matrix(p)
        assert np.allclose(T, synth.pose6_to_matrix(p), atol=1e-15)
        assert np.allclose(T, O.pose_to_matrix(p), atol=1e-6)        # the oracle builds its matrix in float
