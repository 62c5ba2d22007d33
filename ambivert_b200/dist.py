"""Multi-GPU hashtable build: shard the queries over the GPUs of one box, replicate the targets, score
every shard with the CUDA library and bring the per-query best hits to rank 0 with ONE gather
(8 bytes per query) -- SURVEY.md section 8e.  One process per GPU (torchrun); torch.distributed is
the plumbing (NCCL on GPUs, gloo in the CPU tests of the host logic).

The arg-max is per query and every query lives on exactly one rank, so the result is independent of
the number of ranks; there is no other collective on the path.
"""
import numpy as np


def shard_bounds(n_items, world_size):
    """contiguous, balanced ranges: rank r owns [b[r], b[r+1])"""
    base, rem = divmod(int(n_items), int(world_size))
    sizes = [base + (1 if r < rem else 0) for r in range(world_size)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def shard_bounds_by_work(lengths, world_size):
    """contiguous ranges balanced by sum of query lengths (cells are proportional to it)"""
    lengths = np.asarray(lengths, dtype=np.int64)
    if lengths.size == 0:
        return np.zeros(world_size + 1, np.int64)
    csum = np.cumsum(lengths)
    cuts = [0]
    for r in range(1, world_size):
        cuts.append(int(np.searchsorted(csum, csum[-1] * r / world_size, side="left")))
    cuts.append(lengths.size)
    return np.maximum.accumulate(np.asarray(cuts, np.int64))


def gather_best_hits(local_score, local_idx, bounds, dist, rank, world_size, device=None):
    """ONE collective: gather each rank's (best_score, best_idx) slice to rank 0.

    local_score/local_idx: 1-D int32 torch tensors of this rank's shard (on the GPU for NCCL, on the
    CPU for gloo).  Shards are padded to the largest shard so that a single `gather` moves them; rank 0
    returns the concatenated (score, idx) tensors in global query order, other ranks return None."""
    import torch
    n_max = int(np.max(np.diff(bounds))) if world_size > 0 else 0
    packed = torch.empty((2, n_max), dtype=torch.int32, device=local_score.device)
    n_loc = local_score.numel()
    packed[0, :n_loc] = local_score
    packed[1, :n_loc] = local_idx
    if world_size == 1:
        return local_score.clone(), local_idx.clone()
    if rank == 0:
        recv = [torch.empty_like(packed) for _ in range(world_size)]
        dist.gather(packed, gather_list=recv, dst=0)
        score = torch.cat([recv[r][0, :int(bounds[r + 1] - bounds[r])] for r in range(world_size)])
        idx = torch.cat([recv[r][1, :int(bounds[r + 1] - bounds[r])] for r in range(world_size)])
        return score, idx
    dist.gather(packed, gather_list=None, dst=0)
    return None


class ShardedBestHits:
    """The hashtable-build scorer of one rank: a resident job over this rank's query shard."""

    def __init__(self, queries, targets, rank=0, world_size=1, mode=None, seed=None, **scoring):
        import torch
        from . import batch as B
        self.B, self.torch = B, torch
        self.rank, self.world_size = rank, world_size
        self.bounds = shard_bounds_by_work([len(q) for q in queries], world_size)
        lo, hi = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.lo, self.hi = lo, hi
        mode = B.GLOBAL if mode is None else mode
        seed = B.SEED_GLOBAL if seed is None else seed
        self.job = B.BestHitsJob(queries[lo:hi], targets, mode, seed, **scoring)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.d_score = torch.empty(max(hi - lo, 1), dtype=torch.int32, device=dev)[:hi - lo]
        self.d_idx = torch.empty(max(hi - lo, 1), dtype=torch.int32, device=dev)[:hi - lo]

    def step(self, dist=None):
        """launch on torch's current stream, then the single gather; rank 0 gets (score, idx) tensors"""
        torch = self.torch
        stream = torch.cuda.current_stream().cuda_stream or 1   # 0 = legacy default stream = cudaStreamLegacy (0x1)
        self.job.launch(stream, self.d_score.data_ptr(), self.d_idx.data_ptr())
        if self.world_size == 1 or dist is None:
            return self.d_score, self.d_idx
        return gather_best_hits(self.d_score, self.d_idx, self.bounds, dist, self.rank, self.world_size)

    def close(self):
        self.job.close()
