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


def bind_host_to_gpu(device_index: int) -> bool:
    """Best effort: restrict this process to the CPUs NVML reports as local to the GPU (same NUMA node / PCIe root),
    so that pinned buffers allocated afterwards are first-touched next to it.  On a two-socket host a pinned buffer on
    the remote node costs ~25 % of the host -> device bandwidth.  Returns True when an affinity was applied."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return True
    except Exception:
        pass
    return False


class DevicePrefetcher:
    """Iterate over host batches (tuples of pinned CPU tensors), yielding device tuples one step ahead.

    >>> for match, links, olen, tlen in DevicePrefetcher(host_batches, device):
    ...     loss = dag_loss(match.requires_grad_(), links.requires_grad_(), olen, tlen)

    Two persistent sets of device buffers are alternated (no allocator traffic in steady state).  Contract of every
    prefetcher of this kind: a batch may be used until the NEXT batch is requested; the copy that later overwrites its
    buffers waits for everything the compute stream had been given by then.
    """

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device: torch.device):
        if not torch.cuda.is_available():
            raise RuntimeError("DevicePrefetcher needs a CUDA device (there is no CPU path)")
        self._it = iter(batches)
        self._dev = device
        self._copy = torch.cuda.Stream(device=device)
        self._slots = [None, None]          # persistent device buffers
        self._released = [None, None]       # compute-stream events: the slot's previous batch is no longer needed
        self._k = 0                         # batches handed out so far
        self._next = None
        self._next_slot = 0
        self._ready = None
        self._fill = 0                      # batches copied so far
        self._preload()

    def _preload(self) -> None:
        try:
            host = next(self._it)
        except StopIteration:
            self._next = None
            return
        slot = self._fill & 1
        bufs = self._slots[slot]
        if bufs is None or len(bufs) != len(host) or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, host)):
            bufs = tuple(torch.empty(t.shape, dtype=t.dtype, device=self._dev) for t in host)
            self._slots[slot] = bufs
            # the allocator may have handed back blocks that earlier work on the compute stream still reads
            self._copy.wait_stream(torch.cuda.current_stream(self._dev))
        with torch.cuda.stream(self._copy):
            if self._released[slot] is not None:
                self._copy.wait_event(self._released[slot])
            for b, t in zip(bufs, host):
                b.copy_(t, non_blocking=True)
            self._ready = torch.cuda.Event()
            self._ready.record(self._copy)
        self._next = bufs
        self._next_slot = slot
        self._fill += 1

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        return self

    def __next__(self) -> Tuple[torch.Tensor, ...]:
        if self._next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self._dev)
        # everything enqueued so far used the batch handed out before this one: its slot may be overwritten after it
        if self._k > 0:
            ev = torch.cuda.Event()
            ev.record(cur)
            self._released[(self._next_slot + 1) & 1] = ev
        cur.wait_event(self._ready)
        out = tuple(b.detach() for b in self._next)     # fresh tensor objects (no stale .grad), same storage
        self._k += 1
        self._preload()
        return out
