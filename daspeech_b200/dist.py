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
        # highest priority: the exchange kernel's few thread blocks are placed as soon as any block of the compute
        # stream retires, instead of waiting behind the whole grid of the kernel that happens to be running
        self.side = torch.cuda.Stream(device=self.buffer.device, priority=-1) if self.cuda else None
        self._done = None
        self._premul = None

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
                # the division by the world size rides inside the collective (NCCL pre-multiplied sum): no extra
                # read+write pass over the buffer
                if self._premul is None:
                    self._premul = dist._make_nccl_premul_sum(1.0 / ws)
                dist.all_reduce(self.buffer, op=self._premul)
            self._done = torch.cuda.Event()
            self._done.record(self.side)

    def finish(self):
        if self.cuda and self._done is not None:
            torch.cuda.current_stream(self.buffer.device).wait_event(self._done)
            self._done = None
        return self.buffer


class _RawDeviceArray:
    """A cudaMalloc allocation seen through __cuda_array_interface__ (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr: int, numel: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerGradExchange:
    """The same exchange as `FlatGradAllReduce` (one flat fp32 gradient buffer -> mean over ranks, fairseq
    legacy_distributed_data_parallel.py:76-165 via trainer.py:928), moved by the COPY ENGINES over NVLink peer memory:
    reduce-scatter and all-gather as peer-to-peer cudaMemcpyAsync pushes between IPC-mapped buffers, one short kernel
    that sums the rank's own 1/world slice, two flag barriers (daspeech_b200/csrc/xchg.cu, C ABI dagb200_grad_exchange*).
    No thread block is held for the duration of the transfer, so the lattice kernels of the next step keep the SMs.

    One process per GPU of ONE node; the default process group (any backend) carries the 64-byte IPC handles once, at
    construction.  `start()` / `finish()` as in FlatGradAllReduce.  All ranks end with bit-identical buffers."""

    def __init__(self, numel: int, device=None, group=None):
        import ctypes
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = world()
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.numel = (int(numel) + 3) // 4 * 4
        self._ctypes = ctypes
        self._owned, self._opened, self._handle = [], [], ctypes.c_void_p()
        with torch.cuda.device(self.device):
            slice_ = self.lib.dagb200_grad_exchange_slice(self.numel, self.world)
            buf = self._alloc(self.numel * 4)
            flags = self._alloc(32 * 4)
            staging = self._alloc(max(1, (self.world - 1) * slice_) * 4)
            bufs, flagss, stagings = [buf], [flags], [staging]
            if self.world > 1:
                mine = torch.tensor(list(self._export(buf) + self._export(flags) + self._export(staging)), dtype=torch.uint8)
                if dist.get_backend(group) == "nccl":
                    mine = mine.to(self.device)
                every = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(every, mine, group=group)
                bufs, flagss, stagings = [], [], []
                for p, h in enumerate(every):
                    if p == self.rank:
                        bufs.append(buf)
                        flagss.append(flags)
                        stagings.append(staging)
                        continue
                    raw = bytes(h.cpu().tolist())
                    bufs.append(self._open(raw[:64]))
                    flagss.append(self._open(raw[64:128]))
                    stagings.append(self._open(raw[128:]))
            arr = ctypes.c_void_p * self.world
            _lib.check(self.lib.dagb200_grad_exchange_create(arr(*bufs), arr(*flagss), arr(*stagings), self.numel, self.rank,
                                                             self.world, ctypes.byref(self._handle)),
                       "grad_exchange_create")
        self.buffer = torch.as_tensor(_RawDeviceArray(buf, self.numel, "<f4"), device=self.device)
        self.side = torch.cuda.Stream(device=self.device, priority=-1)
        self._done = None
        if self.world > 1:
            dist.barrier(group=group)       # every rank has opened every handle before anybody's first exchange

    def _alloc(self, nbytes):
        from . import _lib
        p = self._ctypes.c_void_p()
        _lib.check(self.lib.dagb200_peer_alloc(nbytes, self._ctypes.byref(p)), "peer_alloc")
        self._owned.append(p.value)
        return p.value

    def _export(self, ptr):
        from . import _lib
        h = (self._ctypes.c_ubyte * 64)()
        _lib.check(self.lib.dagb200_peer_export(ptr, h), "peer_export")
        return tuple(h)

    def _open(self, raw):
        from . import _lib
        h = (self._ctypes.c_ubyte * 64)(*raw)
        p = self._ctypes.c_void_p()
        _lib.check(self.lib.dagb200_peer_open(h, self._ctypes.byref(p)), "peer_open")
        self._opened.append(p.value)
        return p.value

    def start(self):
        from . import _lib
        self.side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dagb200_grad_exchange(self._handle, self.side.cuda_stream), "grad_exchange")
        self._done = torch.cuda.Event()
        self._done.record(self.side)

    def finish(self):
        if self._done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done)
            self._done = None
        return self.buffer

    def phases_ms(self):
        """Durations of push, barrier, reduce, push, barrier of the most recent exchange (the first call only switches
        the recording on)."""
        from . import _lib
        ms = (self._ctypes.c_float * 5)()
        _lib.check(self.lib.dagb200_grad_exchange_phases(self._handle, ms), "grad_exchange_phases")
        return list(ms)

    def timed_out_epoch(self) -> int:
        """0 when every barrier so far completed; synchronises the device (diagnostic)."""
        from . import _lib
        e = self._ctypes.c_int(0)
        _lib.check(self.lib.dagb200_grad_exchange_status(self._handle, self._ctypes.byref(e)), "grad_exchange_status")
        return e.value

    def close(self):
        if self._handle:
            torch.cuda.synchronize(self.device)
            self.lib.dagb200_grad_exchange_destroy(self._handle)
            self._handle = self._ctypes.c_void_p()
            self.buffer = None
            for p in self._opened:
                self.lib.dagb200_peer_close(p)
            if self.world > 1:
                dist.barrier()              # nobody frees an allocation a peer still has mapped
            for p in self._owned:
                self.lib.dagb200_peer_free(p)
            self._opened, self._owned = [], []


class NvlsGradExchange:
    """The same exchange reduced INSIDE the NVSwitch: the buffer lives in symmetric memory bound to a multicast object
    (`torch.distributed._symmetric_memory` does the mapping and provides the stream-ordered barriers -- plumbing), and one
    small kernel of this library (`dagb200_grad_exchange_nvls`, csrc/xchg.cu: `multimem.ld_reduce` + `multimem.st`)
    turns every rank's 1/world slice into the mean over ranks in all copies.  A GPU sends and receives 300 MB per
    exchange whatever the world size (point-to-point reduce-scatter + all-gather: 525 MB each way at 8 GPUs), from
    8 thread blocks (measured at 8 GPUs: 0.70 ms alone with 8 or 16 blocks, 0.85 with 4, 1.57 with 2; step with the
    exchange overlapped 1.29 ms with 8 blocks, 1.64 with 16), so the 128 recurrence CTAs of the next step stay resident.
    Raises RuntimeError when the box has no multicast support; callers fall back to PeerGradExchange."""

    def __init__(self, numel: int, device=None, group=None, ctas: int = 0):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = world()
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.numel = (int(numel) + 3) // 4 * 4
        import os
        # thread blocks of the exchange kernel: enough to keep the switch busy with this rank's slice (37.5 MB at 8 ranks:
        # 8 blocks run at full speed; 150 MB at 2 ranks: 8 blocks take 1.5 ms, 16 take 0.84), as few as possible otherwise
        slice_mb = self.numel * 4 / self.world / 1e6
        auto = max(8, min(24, int(round(slice_mb / 4.7))))
        self.ctas = int(os.environ.get("DAGB200_NVLS_CTAS", ctas or auto))
        self.group = group if group is not None else dist.group.WORLD
        self.buffer = symm_mem.empty(self.numel, dtype=torch.float32, device=self.device)
        self.handle = symm_mem.rendezvous(self.buffer, self.group)
        self.mc = int(self.handle.multicast_ptr)
        if not self.mc:
            raise RuntimeError("no multicast (NVLink SHARP) mapping for symmetric memory on this box")
        self.buffer.zero_()
        self.side = torch.cuda.Stream(device=self.device, priority=-1)
        self._done = None

    def start(self):
        from . import _lib
        self.side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.side):
            self.handle.barrier(channel=0)          # every rank's gradients are in its buffer
            _lib.check(self.lib.dagb200_grad_exchange_nvls(self.mc, self.numel, self.rank, self.world, self.ctas,
                                                           self.side.cuda_stream), "grad_exchange_nvls")
            self.handle.barrier(channel=1)          # every slice has landed in every copy
            self._done = torch.cuda.Event()
            self._done.record(self.side)

    def finish(self):
        if self._done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done)
            self._done = None
        return self.buffer


def exchange_order(kind: str, world_size: int):
    """Mechanisms to try, in order.  Two ranks: the copy engines move the 150 MB slices without holding SMs (0.69 ms) and
    the in-switch form has nothing to save (it moves MORE bytes than point-to-point at N = 2); from four ranks on the
    switch reduces (300 MB per GPU and direction whatever N, against 2 * (N-1)/N * 300 MB point-to-point)."""
    if kind != "auto":
        return (kind,)
    return ("peer", "nvls", "nccl") if world_size <= 2 else ("nvls", "peer", "nccl")


def make_grad_exchange(numel: int, device, kind: str = "auto"):
    """The gradient exchange for this box: "nvls" (in-switch), "peer" (copy engines), "nccl", or "auto" = the first of
    those that can be set up.  Returns (exchange, kind)."""
    _, ws = world()
    order = exchange_order(kind, ws)
    last = None
    for k in order:
        try:
            if k == "nvls":
                if ws < 2:
                    raise RuntimeError("single rank")
                return NvlsGradExchange(numel, device), k
            if k == "peer":
                return PeerGradExchange(numel, device), k
            return FlatGradAllReduce(numel, torch.float32, device), "nccl"
        except Exception as e:   # noqa: BLE001 -- set-up failures select the next mechanism
            last = e
    raise RuntimeError("no gradient exchange could be set up: %r" % (last,))
