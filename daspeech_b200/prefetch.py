"""Host -> device input prefetch for the DAG-loss path.

The lattice inputs of a step (emission plane `match_all` [B,M,L] and transition plane `links` [B,L,T], fp32) are
335 MB at the C2 shape: 6 ms over a PCIe Gen5 x16 link against 1.2 ms of kernels.  fairseq moves a sample to the
device right before the step (trainer.py:797 `_prepare_sample`), serialising copy and compute;
`DevicePrefetcher` issues the copies of step i+1 on a side stream while step i computes, so a step costs
max(copy, compute) instead of their sum.  Two device buffers per tensor (allocated by torch's caching allocator),
events for the hand-over in both directions, no host synchronisation.
"""
from typing import Iterable, Iterator, Sequence, Tuple

import torch


class DevicePrefetcher:
    """Iterate over host batches (tuples of pinned CPU tensors), yielding device tuples one step ahead.

    >>> for match, links, olen, tlen in DevicePrefetcher(host_batches, device):
    ...     loss = dag_loss(match.requires_grad_(), links.requires_grad_(), olen, tlen)
    """

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device: torch.device):
        if not torch.cuda.is_available():
            raise RuntimeError("DevicePrefetcher needs a CUDA device (there is no CPU path)")
        self._it = iter(batches)
        self._dev = device
        self._copy = torch.cuda.Stream(device=device)
        self._next = None
        self._ready = None
        self._preload()

    def _preload(self) -> None:
        try:
            host = next(self._it)
        except StopIteration:
            self._next = None
            return
        cur = torch.cuda.current_stream(self._dev)
        with torch.cuda.stream(self._copy):
            # the caching allocator may hand back a block the compute stream is still reading: order the copy after it
            self._copy.wait_stream(cur)
            self._next = tuple(t.to(self._dev, non_blocking=True) for t in host)
            self._ready = torch.cuda.Event()
            self._ready.record(self._copy)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        return self

    def __next__(self) -> Tuple[torch.Tensor, ...]:
        if self._next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self._dev)
        cur.wait_event(self._ready)
        out = self._next
        for t in out:
            t.record_stream(cur)       # the block returns to the copy stream's pool only after the compute stream is done
        self._preload()
        return out
