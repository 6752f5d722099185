"""Data-parallel plumbing for the DAG-loss path: one process per GPU, utterances sharded across ranks, NO
collective inside the DP (every lattice is independent -- SURVEY.md section 8(e)).  The only exchanges are the
ones the reference's trainer performs around the criterion: summing the logging scalars
(fairseq trainer.py:1469 -> distributed/utils.py:668) and, for timing, a MAX over ranks.

Works with backend "nccl" on GPUs and "gloo" on CPU (used by the world_size-2 tests).
"""
from typing import Dict, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_utt: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous utterance range [lo, hi) owned by `rank`; sizes differ by at most one (the reference shards at
    the iterator: num_shards=world_size, shard_id=rank, fairseq trainer.py:726-727)."""
    base, rem = divmod(n_utt, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world_size: int):
    """Slice every [B, ...] tensor to this rank's utterances."""
    lo, hi = shard_range(tensors[0].shape[0], rank, world_size)
    return [t[lo:hi] for t in tensors]


def sum_stats(stats: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """One flat all-reduce(SUM) of the per-rank logging scalars (loss sum, token counts, invalid sentences)."""
    rank, ws = world()
    keys = sorted(stats)
    flat = torch.stack([stats[k].detach().to(torch.float64).reshape(()) for k in keys])
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return {k: flat[i] for i, k in enumerate(keys)}


def max_over_ranks(value: float, device=None) -> float:
    rank, ws = world()
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dag_nll_sharded(loss_fn, match_all, links, output_length, target_length):
    """Criterion-style reduction over a sharded batch: every rank evaluates `loss_fn` on its utterances, masks
    infeasible lattices like NATDAGLoss._compute_dag_loss (nat_dag_loss.py:143-147) and the global mean is formed
    from all-reduced sums.  Returns (global mean nll, local loss tensor, stats)."""
    rank, ws = world()
    m, lk, ol, tl = shard_batch([match_all, links, output_length, target_length], rank, ws)
    loss = loss_fn(m, lk, ol, tl)
    invalid = loss.isinf().logical_or(loss.isnan())
    loss = loss.masked_fill(invalid, 0)
    stats = sum_stats({"nll_sum": -(loss / tl).sum(), "nsentences": torch.tensor(float(loss.shape[0])),
                       "invalid_nsentences": invalid.sum(), "ntokens": tl.sum()})
    return stats["nll_sum"] / stats["nsentences"], loss, stats
