"""Gallery sharding across ranks (one process per GPU) and the exchange steps of the path.

The reference has no distributed code; this is the B200 design of section 8e:
queries replicated, gallery rows partitioned contiguously, FOUR collectives per call of engine.retrieve (each over the
whole query batch, not per query block):
  1. all_reduce(MAX) of the positives' exact scores   (owner rank holds the score, others -inf)
  2. all_reduce(MAX) of the per-shard completeness cut-off of the re-scored head (before re-scoring: every shard then
     re-scores only the rows that can still reach the gallery-wide head)
  3. all_reduce(SUM) of one flat counter buffer: the per-positive "rows ranked above" counts (additive over shards) and
     two per-query counters (re-scored rows above the best positive, candidate-buffer overflows) for the decidability check
  4. all_gather of the per-shard exact top lists (scores + indices in one buffer), merged per query.
With HOST-resident query features a fourth step precedes them: every rank uploads and fuses only its
1/world slice of a query block and `gather_query_block` assembles the fused block on every rank over
NVLink (all_gather), so the PCIe upload of a block is paid once per box instead of once per GPU.
These helpers are device-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import Tuple

import torch


def shard_range(G: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [start, end) of `rank`; the first G % world ranks get one extra row."""
    base, rem = divmod(G, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def exchange_pos_scores(pos_score: torch.Tensor, group=None) -> torch.Tensor:
    import torch.distributed as dist
    if group is not None or dist.is_initialized():
        dist.all_reduce(pos_score, op=dist.ReduceOp.MAX, group=group)
    return pos_score


def exchange_counts(pos_above: torch.Tensor, group=None) -> torch.Tensor:
    import torch.distributed as dist
    if group is not None or dist.is_initialized():
        dist.all_reduce(pos_above, op=dist.ReduceOp.SUM, group=group)
    return pos_above


def exchange_bound(cut: torch.Tensor, group=None) -> torch.Tensor:
    """Completeness cut-off of the re-scored head: MAX over the shards of each shard's kx-th best approximate score --
    at least kx gallery rows score at or above it, so rows below it cannot belong to the gallery-wide head."""
    import torch.distributed as dist
    if group is not None or dist.is_initialized():
        dist.all_reduce(cut, op=dist.ReduceOp.MAX, group=group)
    return cut


def gather_top_lists(top_score: torch.Tensor, top_idx: torch.Tensor, group=None):
    """-> ([world, Q, R] scores, [world, Q, R] global indices).  ONE all_gather: scores (as their bit patterns) and indices
    travel in the same [Q, 2, R] 32-bit buffer."""
    import torch.distributed as dist
    if not (group is not None or dist.is_initialized()):
        return top_score.contiguous()[None], top_idx.contiguous()[None]
    world = dist.get_world_size(group)
    Q, R = top_score.shape
    pack = torch.empty(Q, 2, R, dtype=torch.int32, device=top_score.device)
    pack[:, 0].copy_(top_score.contiguous().view(torch.int32))
    pack[:, 1].copy_(top_idx)
    # concatenated along dim 0 (the layout both NCCL and gloo accept), viewed as [world, Q, 2, R]
    both = torch.empty(world * Q, 2, R, dtype=torch.int32, device=top_score.device)
    dist.all_gather_into_tensor(both, pack, group=group)
    both = both.view(world, Q, 2, R)
    return both[:, :, 0].contiguous().view(torch.float32), both[:, :, 1].contiguous()


def block_slice(n: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Rows [start, end) of an n-row query block that `rank` uploads and fuses, and the common slot size m =
    ceil(n / world) of the all-gather (the last ranks may hold fewer than m, possibly zero, rows)."""
    m = -(-n // world)
    start = min(n, rank * m)
    return start, min(n, start + m), m


def gather_query_block(part: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """part: [m, ...] this rank's slot (rows beyond its slice are don't-care) -> [n, ...] the whole block."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * part.shape[0],) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(out, part.contiguous(), group=group)
    return out[:n]
