"""Data-parallel plumbing for the DAG-loss path: one process per GPU, utterances sharded across ranks, NO
collective inside the DP (every lattice is independent -- SURVEY.md section 8(e)).  The only exchanges are the
ones the reference's trainer performs around the criterion: summing the logging scalars
(fairseq trainer.py:1469 -> distributed/utils.py:668) and, for timing, a MAX over ranks.

The data-parallel step itself has one more collective, the gradient all-reduce of the trainer
(fairseq legacy_distributed_data_parallel.py:76-165: one flat buffer, pre-divided by the world size, all_reduce):
`FlatGradAllReduce` below is that exchange, issued on a side stream so that it overlaps the tail of the backward pass.

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
    dev = next((v.device for v in stats.values() if v.is_cuda), torch.device("cpu"))   # NCCL reduces device tensors only
    flat = torch.stack([stats[k].detach().to(device=dev, dtype=torch.float64).reshape(()) for k in keys])
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
    stats = sum_stats({"nll_sum": -(loss / tl).sum(), "nsentences": loss.new_tensor(float(loss.shape[0])),
                       "invalid_nsentences": invalid.sum(), "ntokens": tl.sum()})
    return stats["nll_sum"] / stats["nsentences"], loss, stats


class FlatGradAllReduce:
    """The trainer's gradient exchange for one update (fairseq legacy_distributed_data_parallel.py:108-165,
    trainer.py:928,946): ONE flat buffer holding all gradients, divided by the world size, summed over ranks.

    `start()` is called where the backward pass has produced the gradients that go into the buffer; it makes a side
    stream wait for the producer stream, scales and all-reduces there, so whatever the producer stream runs next (the
    rest of the backward pass, the next step's forward) overlaps the exchange.  `finish()` makes the current stream
    wait for the result (the optimizer step would read it).  On CPU tensors (gloo tests) both are synchronous."""

    def __init__(self, numel: int, dtype=torch.float32, device=None):
        self.buffer = torch.zeros(numel, dtype=dtype, device=device)
        self.cuda = self.buffer.is_cuda
        self.side = torch.cuda.Stream(device=self.buffer.device) if self.cuda else None
        self._done = None

    def start(self):
        _, ws = world()
        if not self.cuda:
            if ws > 1:
                self.buffer.div_(ws)
                dist.all_reduce(self.buffer)
            return
        self.side.wait_stream(torch.cuda.current_stream(self.buffer.device))
        with torch.cuda.stream(self.side):
            if ws > 1:
                self.buffer.div_(ws)
                dist.all_reduce(self.buffer)
            self._done = torch.cuda.Event()
            self._done.record(self.side)

    def finish(self):
        if self.cuda and self._done is not None:
            torch.cuda.current_stream(self.buffer.device).wait_event(self._done)
            self._done = None
        return self.buffer
