"""Operator layer: the reference plugin's `DASpeech/custom_ops/dag_loss.py` surface on top of libdagb200.so.

Mirrors, name for name (reference DASpeech/custom_ops/dag_loss.py):
    get_dag_kernel()                      :37-64   -> object exposing the four native callables
                                                      (dag_loss.cpp:24-29) with identical signatures
    DagLossFunc / dag_loss                :66-121, :187
    DagLossWithAlphaBetaFunc / dag_loss_with_alpha_beta   :123-185, :188
    DagBestAlignmentFunc / dag_best_alignment             :190-236
    DagLogsoftmaxGatherFunc / dag_logsoftmax_gather_inplace :238-299
    logsumexp_keepdim, torch_dag_loss, torch_dag_best_alignment, torch_dag_logsoftmax_gather_inplace
                                          :303-425  (device-agnostic torch versions; the criterions'
                                                     --torch-dag-* switches and the S2S posterior use them)

The CUDA operators call hand-written sm_100a kernels through the C ABI in include/dagb200.h (ctypes, raw
device pointers, the CURRENT torch stream).  There is no CPU or torch fallback behind them: without a GPU or
without the built library they raise.  The torch_* functions are independent re-implementations kept because
they are part of the exported surface; the CUDA operators never route through them.
"""
import os
import warnings
from typing import Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from .. import _lib

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.float64: 3}
_DEBUG = os.environ.get("DAGB200_DEBUG", "0") not in ("", "0")

# logsoftmax_gather writes its [B, L, S] result into a [B, S, L]-contiguous buffer and returns the transposed
# view: the criterions immediately call .transpose(1, 2) (nat_dag_loss.py:128) and the loss wrappers call
# .contiguous() (dag_loss.py:103) -- with this layout both are free.  Set to False for a plain contiguous result.
TRANSPOSED_GATHER_OUTPUT = True

# False (default): fp32 lattices run the blocked tensor-core recurrences (flush-to-zero contract in DESIGN.md).
# True: always run the exact log-domain kernels (bit-for-bit the reference's -inf structure, ~100x slower at C2).
EXACT_LOG_DOMAIN = os.environ.get("DAGB200_EXACT", "0") not in ("", "0")

_STATUS_TEXT = {
    1: "dag_best_alignment: target/output length should at least 2",
    2: "dag_best_alignment: graph size is too small (smaller than target length)",
    3: "dag_best_alignment: target length is too short or graph size is too large. "
       "Please increase max_transition_length or remove samples that are too short",
    4: "dag_best_alignment: no valid path",
}


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _check_status(status):
    if status is None:
        return
    bad = status.nonzero()
    if bad.numel():
        b = int(bad[0])
        raise RuntimeError("sample %d: %s" % (b, _STATUS_TEXT.get(int(status[b]), "device status %d" % int(status[b]))))


class _DagKernel:
    """Drop-in for the reference's pybind module `dag_loss_fn` (dag_loss.cpp:24-29): same four callables,
    same argument order, same return values, torch tensors in and out."""

    def __init__(self):
        if not torch.cuda.is_available():
            raise RuntimeError("You need GPU to use the custom cuda operations")
        self.lib = _lib.load()
        self._scratch = {}
        self._pending = []      # (what, pinned host copy of the per-sample status, event) of earlier calls
        self._pinned = None
        self._free = []
        self.track = True       # False: status words are produced but not copied back (CUDA-graph capture, graphs.py)

    # ---- per-sample device status (the reference's CUDA_KERNEL_ASSERTs, dag_loss.cu:68-69, dag_best_alignment.cu:67-70,118)
    # The status words are always produced (B int32).  DAGB200_DEBUG=1 checks them synchronously and raises; otherwise
    # they are copied to pinned host memory without a sync and inspected at the next operator call (or by
    # `check_pending_status()`), where a violation becomes a RuntimeWarning naming the sample.
    def _track_status(self, what, status):
        if not self.track:
            return
        if _DEBUG:
            _check_status(status)
            return
        # a small ring of persistent pinned buffers: no host allocation and NO host synchronisation on the launch path.
        # When the host runs many calls ahead of the device the ring fills up; the status of those calls is then simply
        # not tracked (the outputs still carry -inf / -1 for offending samples).
        n = status.numel()
        if self._pinned is None or self._pinned.shape[1] < n:
            self.check_pending_status(wait=True)
            self._pinned = torch.empty((16, max(n, 256)), dtype=torch.int32, pin_memory=True)
            self._free = list(range(16))
        if not self._free:
            return
        slot = self._free.pop()
        host = self._pinned[slot, :n]
        host.copy_(status, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((what, host, ev, slot))

    def check_pending_status(self, wait=False):
        keep = []
        blocked = False
        for what, host, ev, slot in self._pending:
            if wait:
                ev.synchronize()
            if blocked or not ev.query():      # oldest first: once one is still in flight the younger ones are not asked
                blocked = True
                keep.append((what, host, ev, slot))
                continue
            self._free.append(slot)
            bad = host.nonzero()
            if bad.numel():
                b = int(bad[0])
                warnings.warn("%s: sample %d: %s (%d sample(s) affected; outputs are -inf / -1 for them)" % (
                    what, b, _STATUS_TEXT.get(int(host[b]), "device status %d" % int(host[b])), bad.numel()), RuntimeWarning)
        self._pending = keep

    def release_workspaces(self):
        """Drop the cached per-(device, stream) scratch buffers (~400 MB each at the C2 shape)."""
        self._scratch.clear()

    def _workspace(self, nbytes, device):
        """Per-(device, stream) scratch for the blocked kernels, grown on demand and kept across calls: the
        contents only live for the duration of one call, and calls on a stream are ordered.  (Allocating ~400 MB
        per call makes the caching allocator fall back to synchronous cudaMalloc when it alternates with the
        268 MB gradient buffers.)"""
        if nbytes <= 0:
            return None
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        buf = self._scratch.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = None
            self._scratch.pop(key, None)
            buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._scratch[key] = buf
        return buf

    # ---- shared argument checks (dag_loss.cu:317-332 / dag_best_alignment.cu:212-227) -------------------
    @staticmethod
    def _check_lattice(match_all, links, output_length, target_length):
        _check(match_all.is_cuda, "match_all must be a CUDA tensor")
        _check(links.is_cuda, "links must be a CUDA tensor")
        _check(output_length.is_cuda, "output_length must be a CUDA tensor")
        _check(target_length.is_cuda, "target_length must be a CUDA tensor")
        _check(links.device == match_all.device and output_length.device == match_all.device and
               target_length.device == match_all.device, "match_all, links and the lengths must be on the same device")
        _check(match_all.dim() == 3, "match_all dim != 3")
        _check(links.dim() == 3, "links dim != 3")
        _check(output_length.dim() == 1, "output_length dim != 3")
        _check(target_length.dim() == 1, "target_length dim != 3")
        bsz, tarlen, prelen = match_all.shape
        _check(links.size(0) == bsz and output_length.size(0) == bsz and target_length.size(0) == bsz,
               "batch size not match")
        _check(links.size(1) == prelen, "prelen not match")
        _check(output_length.dtype == torch.long and target_length.dtype == torch.long, "length should be long")
        _check(match_all.dtype in (torch.float32, torch.float64),
               '"dag_loss" not implemented for \'%s\'' % str(match_all.dtype).replace("torch.", ""))
        _check(links.dtype == match_all.dtype, "match_all and links must have the same dtype")
        return bsz, tarlen, prelen, links.size(2)

    def dag_loss(self, match_all, links, output_length, target_length, require_gradient, config) -> Tuple[Tensor, Tensor]:
        bsz, tarlen, prelen, translen = self._check_lattice(match_all, links, output_length, target_length)
        match_all = match_all.contiguous()
        links = links.contiguous()
        output_length = output_length.contiguous()
        target_length = target_length.contiguous()
        alpha = torch.empty((bsz, tarlen, prelen), dtype=match_all.dtype, device=match_all.device)
        beta = torch.empty_like(alpha)
        self.check_pending_status()
        status = torch.empty(bsz, dtype=torch.int32, device=match_all.device)
        nbytes = 0
        if not EXACT_LOG_DOMAIN and match_all.dtype == torch.float32:
            nbytes = int(self.lib.dagb200_dag_loss_workspace_bytes(bsz, tarlen, prelen, translen))
        workspace = self._workspace(nbytes, match_all.device)
        self.lib.dagb200_set_exact(int(bool(EXACT_LOG_DOMAIN)))
        with torch.cuda.device(match_all.device):
            rc = self.lib.dagb200_dag_loss(_ptr(match_all), _ptr(links), _ptr(output_length), _ptr(target_length),
                                           _ptr(alpha), _ptr(beta), _DTYPE_CODE[match_all.dtype],
                                           bsz, tarlen, prelen, translen, int(bool(require_gradient)), int(config),
                                           _ptr(workspace), nbytes, _ptr(status), _stream())
        _lib.check(rc, "dag_loss")
        if bsz:
            self._track_status("dag_loss", status)
        return alpha, beta

    def dag_loss_backward(self, grad_output, alpha, beta, match_all, links, output_length, target_length,
                          config1, config2) -> Tuple[Tensor, Tensor]:
        bsz, tarlen, prelen = match_all.shape
        translen = links.size(2)
        grad_output = grad_output.to(match_all.dtype).contiguous()
        match_all = match_all.contiguous()
        links = links.contiguous()
        output_length = output_length.contiguous()
        target_length = target_length.contiguous()
        grad_match_all = torch.empty_like(alpha)
        grad_links = torch.empty((bsz, prelen, translen), dtype=match_all.dtype, device=match_all.device)
        nbytes = 0
        if not EXACT_LOG_DOMAIN and match_all.dtype == torch.float32:
            nbytes = int(self.lib.dagb200_dag_loss_backward_workspace_bytes(bsz, tarlen, prelen, translen))
        workspace = self._workspace(nbytes, match_all.device)
        self.lib.dagb200_set_exact(int(bool(EXACT_LOG_DOMAIN)))
        with torch.cuda.device(match_all.device):
            rc = self.lib.dagb200_dag_loss_backward_ws(_ptr(grad_output), _ptr(alpha), _ptr(beta), _ptr(match_all),
                                                       _ptr(links), _ptr(output_length), _ptr(target_length),
                                                       _ptr(grad_match_all), _ptr(grad_links),
                                                       _DTYPE_CODE[match_all.dtype], bsz, tarlen, prelen, translen,
                                                       int(config1), int(config2), _ptr(workspace), nbytes, _stream())
        _lib.check(rc, "dag_loss_backward")
        return grad_match_all, grad_links

    def dag_best_alignment(self, match_all, links, output_length, target_length, config,
                           want_alpha=True, track_status=True) -> Tuple[Tensor, Tensor]:
        bsz, tarlen, prelen, translen = self._check_lattice(match_all, links, output_length, target_length)
        match_all = match_all.contiguous()
        links = links.contiguous()
        output_length = output_length.contiguous()
        target_length = target_length.contiguous()
        dev = match_all.device
        alpha = torch.empty((bsz, tarlen, prelen), dtype=match_all.dtype, device=dev) if want_alpha else None
        path = torch.empty((bsz, prelen), dtype=torch.int32, device=dev)
        nbytes = int(self.lib.dagb200_best_alignment_workspace_bytes(bsz, tarlen, prelen, translen))
        workspace = self._workspace(max(nbytes, 1), dev)
        self.check_pending_status()
        status = torch.empty(bsz, dtype=torch.int32, device=dev)
        self.lib.dagb200_set_exact(int(bool(EXACT_LOG_DOMAIN)))
        with torch.cuda.device(dev):
            rc = self.lib.dagb200_dag_best_alignment(_ptr(match_all), _ptr(links), _ptr(output_length),
                                                     _ptr(target_length), _ptr(alpha), _ptr(path),
                                                     _DTYPE_CODE[match_all.dtype], bsz, tarlen, prelen, translen,
                                                     int(config), _ptr(workspace), nbytes, _ptr(status), _stream())
        _lib.check(rc, "dag_best_alignment")
        if bsz and track_status:
            self._track_status("dag_best_alignment", status)
        return alpha, path

    def logsoftmax_gather(self, word_ins_out, select_idx, require_gradient, want_argmax=False):
        _check(word_ins_out.is_cuda, "word_ins_out must be a CUDA tensor")
        _check(select_idx.is_cuda, "select_idx must be a CUDA tensor")
        _check(word_ins_out.dim() == 3, "word_ins_out dim != 3")
        _check(select_idx.dim() == 3, "select_idx dim != 3")
        bsz, prelen, vocabsize = word_ins_out.shape
        slen = select_idx.size(2)
        _check(select_idx.size(0) == bsz, "batch size not match")
        _check(select_idx.size(1) == prelen, "prelen size not match")
        _check(select_idx.dtype == torch.long, "select_idx should be long")
        _check(word_ins_out.is_contiguous(), "word_ins_out is not contiguous")
        _check(word_ins_out.dtype in _DTYPE_CODE,
               "logsoftmax_gather_kernel_scalar_t not implemented for '%s'" % word_ins_out.dtype)
        out_dtype = torch.float64 if word_ins_out.dtype == torch.float64 else torch.float32
        if TRANSPOSED_GATHER_OUTPUT:
            buf = torch.empty((bsz, slen, prelen), dtype=out_dtype, device=word_ins_out.device)
            result = buf.transpose(1, 2)
        else:
            result = torch.empty((bsz, prelen, slen), dtype=out_dtype, device=word_ins_out.device)
        isb, isl, iss = select_idx.stride()
        osb, osl, oss = result.stride()
        if want_argmax:
            # fused with the gather (one pass over the logits instead of two); fp64 has no fused variant
            if word_ins_out.dtype == torch.float64:
                amax = word_ins_out.argmax(-1)
            else:
                amax = torch.empty((bsz, prelen), dtype=torch.long, device=word_ins_out.device)
                with torch.cuda.device(word_ins_out.device):
                    rc = self.lib.dagb200_logsoftmax_gather_argmax(
                        _ptr(word_ins_out), _DTYPE_CODE[word_ins_out.dtype], _ptr(select_idx), isb, isl, iss,
                        _ptr(result), osb, osl, oss, _ptr(amax), bsz, prelen, vocabsize, slen,
                        int(bool(require_gradient)), _stream())
                _lib.check(rc, "logsoftmax_gather_argmax")
                return result, amax
        with torch.cuda.device(word_ins_out.device):
            rc = self.lib.dagb200_logsoftmax_gather(_ptr(word_ins_out), _DTYPE_CODE[word_ins_out.dtype],
                                                    _ptr(select_idx), isb, isl, iss, _ptr(result), osb, osl, oss,
                                                    bsz, prelen, vocabsize, slen, int(bool(require_gradient)),
                                                    _stream())
        _lib.check(rc, "logsoftmax_gather")
        return (result, amax) if want_argmax else result

    # fused replacement of the two torch ops in DagLogsoftmaxGatherFunc.backward (dag_loss.py:294-295)
    def logsoftmax_gather_backward(self, probs_inout, select_idx, grad_output) -> Tensor:
        bsz, prelen, vocabsize = probs_inout.shape
        slen = select_idx.size(2)
        want = torch.float64 if probs_inout.dtype == torch.float64 else torch.float32
        if grad_output.dtype != want:
            grad_output = grad_output.to(want)
        isb, isl, iss = select_idx.stride()
        gsb, gsl, gss = grad_output.stride()
        with torch.cuda.device(probs_inout.device):
            rc = self.lib.dagb200_logsoftmax_gather_backward(_ptr(probs_inout), _DTYPE_CODE[probs_inout.dtype],
                                                             _ptr(select_idx), isb, isl, iss,
                                                             _ptr(grad_output), gsb, gsl, gss,
                                                             bsz, prelen, vocabsize, slen, _stream())
        _lib.check(rc, "logsoftmax_gather_backward")
        return probs_inout


dag_kernel = None


def get_dag_kernel():
    """Reference: dag_loss.py:37-64 (JIT-compiles on first use).  Here the library is prebuilt in-tree."""
    global dag_kernel
    if not torch.cuda.is_available():
        raise RuntimeError("You need GPU to use the custom cuda operations")
    if dag_kernel is None:
        dag_kernel = _DagKernel()
    return dag_kernel


# =====================================================================================================
class DagLossFunc(Function):
    # kept for API compatibility: the reference's tuner mutates these (dag_loss.py:67-69, 452-454)
    config = 1
    config1 = 2
    config2 = 2

    @staticmethod
    def forward(ctx, match_all, links, output_length, target_length):
        r"""DAG log-marginal.  match_all [B, M, L] (log P(y_i | v_j)), links [B, L, T] (links[b, i, k] =
        log P(v_i -> v_{i+k+1})), output_length / target_length [B] long.  Returns [B]."""
        require_gradient = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        match_all = match_all.contiguous()
        links = links.contiguous()
        alpha, beta = get_dag_kernel().dag_loss(match_all, links, output_length, target_length,
                                                require_gradient, DagLossFunc.config)
        if require_gradient:
            res = beta[:, 0, 0].clone()
        else:
            res = alpha[range(alpha.shape[0]), target_length - 1, output_length - 1]
        ctx.save_for_backward(alpha, beta, match_all, links, output_length, target_length)
        return res

    @staticmethod
    def backward(ctx, grad_output):
        alpha, beta, match_all, links, output_length, target_length = ctx.saved_tensors
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            grad_match_all, grad_links = get_dag_kernel().dag_loss_backward(
                grad_output, alpha, beta, match_all, links, output_length, target_length,
                DagLossFunc.config1, DagLossFunc.config2)
            return grad_match_all, grad_links, None, None
        return None, None, None, None


class DagLossWithAlphaBetaFunc(Function):
    config = 1
    config1 = 2
    config2 = 2

    @staticmethod
    def forward(ctx, match_all, links, output_length, target_length):
        r"""As DagLossFunc, additionally returning (alpha, beta) [B, M, L] as non-differentiable extras
        (consumed by the S2S posterior expectation, s2s_dag_fastspeech2_loss.py:257-265)."""
        require_gradient = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        match_all = match_all.contiguous()
        links = links.contiguous()
        alpha, beta = get_dag_kernel().dag_loss(match_all, links, output_length, target_length,
                                                require_gradient, DagLossWithAlphaBetaFunc.config)
        if require_gradient:
            res = beta[:, 0, 0].clone()
        else:
            res = alpha[range(alpha.shape[0]), target_length - 1, output_length - 1]
        ctx.save_for_backward(alpha, beta, match_all, links, output_length, target_length)
        return res, (alpha, beta)

    @staticmethod
    def backward(ctx, grad_output, unused):
        alpha, beta, match_all, links, output_length, target_length = ctx.saved_tensors
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            grad_match_all, grad_links = get_dag_kernel().dag_loss_backward(
                grad_output, alpha, beta, match_all, links, output_length, target_length,
                DagLossWithAlphaBetaFunc.config1, DagLossWithAlphaBetaFunc.config2)
            return grad_match_all, grad_links, None, None
        return None, None, None, None


dag_loss = DagLossFunc.apply
dag_loss_with_alpha_beta = DagLossWithAlphaBetaFunc.apply


class DagBestAlignmentFunc(Function):
    config = 1

    @staticmethod
    def forward(ctx, match_all, links, output_length, target_length):
        r"""Viterbi alignment of the target to the DAG.  Returns path [B, L] long: path[b, j] = index of the
        target token emitted by vertex j, or -1 when the vertex is skipped."""
        match_all = match_all.contiguous()
        links = links.contiguous()
        # the max-plus lattice is discarded by this wrapper (as in the reference) -> do not materialise it
        _, path = get_dag_kernel().dag_best_alignment(match_all, links, output_length, target_length,
                                                      DagBestAlignmentFunc.config, want_alpha=False)
        path = path.to(torch.long)
        ctx.mark_non_differentiable(path)
        return path

    @staticmethod
    def backward(ctx, grad_output):
        assert False, "no backward function for best alignment"


dag_best_alignment = DagBestAlignmentFunc.apply


class DagLogsoftmaxGatherFunc(Function):

    @staticmethod
    def forward(ctx, word_ins_out, select_idx):
        r"""res = word_ins_out.log_softmax(-1, dtype=float).gather(-1, select_idx), fused.
        word_ins_out [B, L, V] is MODIFIED IN PLACE (softmax probabilities, kept for backward) when it
        requires grad -- do not use it afterwards.  select_idx [B, L, S] long (may be an expanded view)."""
        require_gradient = ctx.needs_input_grad[0]
        selected_result = get_dag_kernel().logsoftmax_gather(word_ins_out, select_idx, require_gradient)
        ctx.mark_dirty(word_ins_out)
        ctx.set_materialize_grads(False)
        if require_gradient:
            ctx.save_for_backward(word_ins_out, select_idx)
            ctx.has_backward = False
        return word_ins_out, selected_result

    @staticmethod
    def backward(ctx, grad_word_ins_out, grad_output):
        if not ctx.needs_input_grad[0]:
            return None, None
        assert grad_word_ins_out is None, "Cannot reuse word_ins_out after logsoftmax_gather"
        if grad_output is None:
            return None, None
        assert not ctx.has_backward, "Cannot backward twice in logsoftmax_gather"
        ctx.has_backward = True
        grad_input, selected_idx = ctx.saved_tensors
        get_dag_kernel().logsoftmax_gather_backward(grad_input, selected_idx, grad_output)
        return grad_input, None


dag_logsoftmax_gather_inplace = DagLogsoftmaxGatherFunc.apply


class DagLogsoftmaxGatherArgmaxFunc(Function):
    r"""`dag_logsoftmax_gather_inplace` that also returns `word_ins_out.argmax(-1)` of the RAW logits, computed in the
    same pass (the GLAT pass of the criterion takes the arg-max right before the gather, nat_dag_loss.py:209-213).
    Returns (word_ins_out, selected_result, pred_tokens); pred_tokens is long [B, L] and not differentiable."""

    @staticmethod
    def forward(ctx, word_ins_out, select_idx):
        require_gradient = ctx.needs_input_grad[0]
        selected_result, pred = get_dag_kernel().logsoftmax_gather(word_ins_out, select_idx, require_gradient, want_argmax=True)
        ctx.mark_dirty(word_ins_out)
        ctx.mark_non_differentiable(pred)
        ctx.set_materialize_grads(False)
        if require_gradient:
            ctx.save_for_backward(word_ins_out, select_idx)
            ctx.has_backward = False
        return word_ins_out, selected_result, pred

    @staticmethod
    def backward(ctx, grad_word_ins_out, grad_output, grad_pred):
        return DagLogsoftmaxGatherFunc.backward(ctx, grad_word_ins_out, grad_output)


dag_logsoftmax_gather_argmax_inplace = DagLogsoftmaxGatherArgmaxFunc.apply


# =====================================================================================================
# Device-agnostic torch versions (exported names; not used by the CUDA operators above).
def logsumexp_keepdim(x: Tensor, dim: int) -> Tensor:
    """log-sum-exp along `dim` (kept) that returns -inf, with zero gradient, for all -inf slices
    (reference dag_loss.py:303-311)."""
    peak = x.detach().amax(dim=dim, keepdim=True)
    empty = peak == float("-inf")
    shift = torch.where(empty, torch.zeros_like(peak), peak)
    total = (x - shift).exp().sum(dim=dim, keepdim=True)
    total = torch.where(empty, torch.ones_like(total), total)
    return torch.where(empty, torch.full_like(total, float("-inf")), total.log() + shift)


def _dense_forward(match_all, links, reduce_fn):
    """Shared column recursion over DENSE links [B, L, L] (links[b, i, j]: vertex i -> vertex j).
    Returns the full lattice [B, M, L]."""
    bsz, tarlen, prelen = match_all.shape
    assert links.shape[1] == links.shape[2], "links should be batch_size * prelen * prelen"
    state = match_all.new_full((bsz, prelen), float("-inf"))
    state[:, 0] = match_all[:, 0, 0]
    rows = [state]
    for step in range(1, tarlen):
        state = reduce_fn(state.unsqueeze(2) + links) + match_all[:, step]
        rows.append(state)
    return torch.stack(rows, dim=1)


def torch_dag_loss(match_all, links, output_length, target_length):
    """Torch version of dag_loss; NOTE links is the dense [B, L, L] layout here (reference dag_loss.py:325-366)."""
    lattice = _dense_forward(match_all, links, lambda x: logsumexp_keepdim(x, 1).squeeze(1))
    return lattice[torch.arange(lattice.shape[0], device=lattice.device), target_length - 1, output_length - 1]


def torch_dag_best_alignment(match_all, links, output_length, target_length):
    """Torch version of dag_best_alignment over dense links (reference dag_loss.py:388-419): explicit max-plus
    recursion with back-pointers and a batched backtrace."""
    with torch.no_grad():
        bsz, tarlen, prelen = match_all.shape
        state = match_all.new_full((bsz, prelen), float("-inf"))
        state[:, 0] = match_all[:, 0, 0]
        back = []
        for step in range(1, tarlen):
            best, arg = (state.unsqueeze(2) + links).max(dim=1)
            state = best + match_all[:, step]
            back.append(arg)
        rows = torch.arange(bsz, device=match_all.device)
        path = torch.full((bsz, prelen), -1, dtype=torch.long, device=match_all.device)
        pos = (output_length - 1).clone()
        for step in range(tarlen - 1, 0, -1):
            active = step <= (target_length - 1)
            path[rows[active], pos[active]] = step
            pos = torch.where(active, back[step - 1][rows, pos], pos)
        path[rows, pos] = 0
    return path


def torch_dag_logsoftmax_gather_inplace(word_ins_out, select_idx):
    """Unfused log_softmax + gather (reference dag_loss.py:421-425); does not modify word_ins_out."""
    logits = torch.log_softmax(word_ins_out, -1, dtype=torch.float32)
    match = logits.gather(dim=-1, index=select_idx)
    return word_ins_out, match
