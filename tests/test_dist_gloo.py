"""CPU tests of the multi-GPU host logic (ambivert_b200/dist.py) with two gloo ranks: contiguous
query shards, the single gather of (best_score, best_idx), rank-count independence of the result."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ambivert_b200.dist import gather_best_hits, shard_bounds, shard_bounds_by_work


def test_shard_bounds():
    assert list(shard_bounds(10, 3)) == [0, 4, 7, 10]
    assert list(shard_bounds(2, 4)) == [0, 1, 2, 2, 2]
    assert list(shard_bounds(0, 2)) == [0, 0, 0]
    rng = np.random.default_rng(0)
    lens = rng.integers(150, 300, 1001)
    for w in (1, 2, 3, 8):
        b = shard_bounds_by_work(lens, w)
        assert b[0] == 0 and b[-1] == 1001 and (np.diff(b) >= 0).all()
        work = [lens[b[r]:b[r + 1]].sum() for r in range(w)]
        assert max(work) - min(work) <= 2 * lens.max()          # balanced to within a query or two
    assert list(shard_bounds_by_work([], 2)) == [0, 0, 0]


def fake_scores(lo, hi):
    i = np.arange(lo, hi, dtype=np.int64)
    return ((i * 7919) % 251 - 40).astype(np.int32), ((i * 104729) % 2000).astype(np.int32)


def _worker(rank, world, port, n_q, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lens = (np.arange(n_q) % 97) + 150
        bounds = shard_bounds_by_work(lens, world)
        s, i = fake_scores(int(bounds[rank]), int(bounds[rank + 1]))
        res = gather_best_hits(torch.from_numpy(s), torch.from_numpy(i), bounds, dist, rank, world)
        if rank == 0:
            ws, wi = fake_scores(0, n_q)
            ok = res is not None and (res[0].numpy() == ws).all() and (res[1].numpy() == wi).all()
            out.put(bool(ok))
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_q", [1001, 3, 1])
def test_gather_two_ranks_gloo(n_q):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() + n_q) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_q, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True


def test_gather_single_rank_is_identity():
    s, i = fake_scores(0, 10)
    rs, ri = gather_best_hits(torch.from_numpy(s), torch.from_numpy(i), shard_bounds(10, 1), None, 0, 1)
    assert (rs.numpy() == s).all() and (ri.numpy() == i).all()
