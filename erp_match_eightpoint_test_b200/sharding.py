"""Host-side sharding logic for one process per GPU.

The path shards without a data-path collective: query rows are split in contiguous blocks
per rank (train set replicated), hypothesis ids are split in contiguous ranges per rank
(Philox is keyed by the GLOBAL hypothesis id, so results do not depend on the world size).
The only exchange steps are
  * best-model selection: one 8-byte MAX all-reduce of the packed (count, ~id) word,
  * cross-check: a MIN all-reduce over per-train (d2, queryIdx).
These helpers are pure index arithmetic plus torch.distributed calls, so they run under
``gloo`` on CPU in the tests exactly as they run under NCCL on the GPUs.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of n items for this rank; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_best(count: int, hyp_id: int) -> int:
    """(count << 32) | (0xFFFFFFFF - id): MAX picks the highest count, then the lowest id."""
    return (int(count) << 32) | (0xFFFFFFFF - int(hyp_id))


def unpack_best(packed: int) -> tuple[int, int]:
    return int(packed) >> 32, 0xFFFFFFFF - (int(packed) & 0xFFFFFFFF)


def allreduce_best(packed_tensor, dist=None):
    """In-place MAX all-reduce of a 1-element int64 tensor holding the packed best.
    The packed word is < 2^63 (counts are < 2^31) so signed MAX orders it correctly."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(packed_tensor, op=dist.ReduceOp.MAX)
    return packed_tensor


def allreduce_cross_check(best_d2, best_q, dist=None):
    """Global nearest query per train row from per-rank partial results.

    best_d2: float64 tensor (nt), best_q: int32/int64 tensor (nt) of GLOBAL query ids.
    Two MIN all-reduces: the distance, then the lowest query id among ranks that attain it
    (exactly cv::BFMatcher(crossCheck) tie order)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return best_d2, best_q
    import torch

    gmin = best_d2.clone()
    dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
    cand = torch.where(best_d2 == gmin, best_q.to(torch.int64), torch.full_like(best_q, 2**31 - 1, dtype=torch.int64))
    dist.all_reduce(cand, op=dist.ReduceOp.MIN)
    return gmin, cand.to(best_q.dtype)


def concat_matches(per_rank: list[np.ndarray]) -> np.ndarray:
    """Rank-ordered concatenation keeps ascending queryIdx (feature_matcher.cpp:50-56 order)."""
    return np.concatenate(per_rank) if per_rank else np.empty(0)
