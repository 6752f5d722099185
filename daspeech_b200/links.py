"""Transition log-probabilities of the DAG from the link heads.

`extract_links` mirrors `GlatLinkDecoder.extract_links` (DASpeech/models/s2t_conformer_dag.py:171-212, with
`extract_valid_links` :140-155) for the banded form the models use (`max_transition_length != -1`): same inputs
(decoder features, previous output tokens, the positional embedding and the three linear heads), same `[B, L, T]` fp32
result.  The forward runs as ONE kernel (`dagb200_extract_links`, csrc/dag_links.cu: QK^T on tcgen05, masked softmax
over the successors, gate mixture over the heads) instead of the `[B, L, L, H]` einsum and the `[B, L, T, H]`
temporaries; the backward recomputes the reference's op sequence in row chunks (bounded memory, plain torch), so the
GLAT pass and inference -- which run without gradients -- get the fused path and training stays differentiable.
"""
import ctypes
import math

import torch
import torch.nn.functional as F

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _logsumexp(x, dim):
    """log-sum-exp that returns -inf (and a zero gradient) where every term is -inf, as the reference's helper
    (s2t_conformer_dag.py:53-58) does; torch.logsumexp differentiates such rows to NaN."""
    m = x.max(dim=dim).values
    dead = m == float("-inf")
    shift = m.masked_fill(dead, 0.0)
    s = (x - shift.unsqueeze(dim)).exp().sum(dim=dim)
    return s.masked_fill(dead, 1.0).log() + shift.masked_fill(dead, float("-inf"))


def torch_extract_links(query_chunks, key_chunks, log_gates, output_length, translen):
    """The reference's op sequence (s2t_conformer_dag.py:188-202 with :140-155) on the per-head projections:
    query_chunks / key_chunks [B, L, H, F], log_gates [B, L, H] (fp32 log-softmax), output_length [B] -> links [B, L, T].
    Device-agnostic; used by the golden-vector tests and, row chunk by row chunk, by the backward of the fused op."""
    B, L, H, Fd = query_chunks.shape
    content = torch.einsum("bicf,bjcf->bijc", query_chunks.float(), key_chunks.float()) / (Fd ** 0.5)
    idx = (torch.arange(L, device=content.device).unsqueeze(1) + torch.arange(translen, device=content.device).unsqueeze(0) + 1)
    invalid = idx.unsqueeze(0) >= output_length.view(B, 1, 1)
    idx = idx.unsqueeze(0).masked_fill(invalid, 0)
    res = content.gather(2, idx.unsqueeze(-1).expand(-1, -1, -1, H))
    res = res.masked_fill(invalid.unsqueeze(-1), float("-inf"))
    nouse = invalid.all(-1)
    res = res.masked_fill(nouse.unsqueeze(-1).unsqueeze(-1), float("-inf"))
    res = F.log_softmax(res, dim=2)
    res = res.masked_fill(nouse.unsqueeze(-1).unsqueeze(-1), float("-inf"))
    return _logsumexp(res + log_gates.unsqueeze(2), -1)


def _torch_rows(query_chunks, key_chunks, log_gates, output_length, translen, i0, i1):
    """Rows i0..i1-1 of torch_extract_links without the full [B, L, L, H] product (keys i0+1 .. i1-1+T only)."""
    B, L, H, Fd = query_chunks.shape
    j0, j1 = i0 + 1, min(L, i1 + translen)
    dev = query_chunks.device
    if j1 <= j0:
        return query_chunks.new_full((B, i1 - i0, translen), float("-inf"), dtype=torch.float32)
    content = torch.einsum("bicf,bjcf->bijc", query_chunks[:, i0:i1].float(), key_chunks[:, j0:j1].float()) / (Fd ** 0.5)
    idx = torch.arange(i0, i1, device=dev).unsqueeze(1) + torch.arange(translen, device=dev).unsqueeze(0) + 1   # absolute j
    invalid = idx.unsqueeze(0) >= output_length.view(B, 1, 1)
    rel = (idx - j0).unsqueeze(0).masked_fill(invalid, 0).clamp_(0, j1 - j0 - 1)
    res = content.gather(2, rel.unsqueeze(-1).expand(-1, -1, -1, H))
    res = res.masked_fill(invalid.unsqueeze(-1), float("-inf"))
    nouse = invalid.all(-1)
    safe = res.masked_fill(nouse.unsqueeze(-1).unsqueeze(-1), 0.0)           # keeps the softmax of dead rows finite
    lp = F.log_softmax(safe, dim=2).masked_fill(nouse.unsqueeze(-1).unsqueeze(-1), float("-inf"))
    return _logsumexp(lp + log_gates[:, i0:i1].unsqueeze(2), -1)


class _ExtractLinksFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query_chunks, key_chunks, log_gates, output_length, translen):
        lib = _lib.load()
        B, L, H, Fd = query_chunks.shape
        q = query_chunks.detach().float().contiguous()
        k = key_chunks.detach().float().contiguous()
        g = log_gates.detach().float().contiguous()
        ol = output_length.contiguous()
        links = torch.empty((B, L, translen), dtype=torch.float32, device=q.device)
        nbytes = int(lib.dagb200_extract_links_workspace_bytes(B, L, H, Fd))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)              # converted operands + row normalisers
        with torch.cuda.device(q.device):
            rc = lib.dagb200_extract_links(_ptr(q), _ptr(k), _ptr(g), _ptr(ol), _ptr(links), B, L, H, Fd, translen,
                                           _ptr(ws), nbytes,
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "extract_links")
        ctx.save_for_backward(query_chunks, key_chunks, log_gates, ol)
        ctx.translen = translen
        return links

    @staticmethod
    def backward(ctx, grad_links):
        query_chunks, key_chunks, log_gates, ol = ctx.saved_tensors
        T = ctx.translen
        B, L, H, Fd = query_chunks.shape
        gq = torch.zeros_like(query_chunks, dtype=torch.float32)
        gk = torch.zeros_like(key_chunks, dtype=torch.float32)
        gg = torch.zeros_like(log_gates, dtype=torch.float32)
        rows = max(16, min(L, (1 << 26) // max(1, B * (T + 128) * H)))       # ~256 MB of fp32 scores per chunk
        for i0 in range(0, L, rows):
            i1 = min(L, i0 + rows)
            with torch.enable_grad():
                q = query_chunks.detach().float().requires_grad_()
                k = key_chunks.detach().float().requires_grad_()
                g = log_gates.detach().float().requires_grad_()
                out = _torch_rows(q, k, g, ol, T, i0, i1)
                go = grad_links[:, i0:i1].masked_fill(torch.isinf(out.detach()), 0.0)   # -inf entries carry no gradient
                dq, dk, dg = torch.autograd.grad((out.masked_fill(torch.isinf(out.detach()), 0.0) * go).sum(), [q, k, g])
            gq += dq
            gk += dk
            gg += dg
        return gq.to(query_chunks.dtype), gk.to(key_chunks.dtype), gg.to(log_gates.dtype), None, None


def extract_links_from_chunks(query_chunks, key_chunks, log_gates, output_length, translen, fused=True):
    """links [B, L, translen] from the per-head projections.  `fused=False` (or CPU tensors) runs the reference's op
    sequence; the fused path needs the feature size per head to be a multiple of 16 in [16, 128]."""
    Fd = query_chunks.shape[-1]
    if fused and query_chunks.is_cuda and Fd % 16 == 0 and 16 <= Fd <= 128 and translen > 0:
        return _ExtractLinksFunc.apply(query_chunks, key_chunks, log_gates, output_length, int(translen))
    if fused and query_chunks.is_cuda:
        raise RuntimeError("extract_links: the fused kernel needs a head size that is a multiple of 16 in [16, 128]; "
                           "pass fused=False for the reference op sequence")
    return torch_extract_links(query_chunks, key_chunks, log_gates, output_length, int(translen))


def extract_links(features, prev_output_tokens, link_positional, query_linear, key_linear, gate_linear, *, pad,
                  decoder_attention_heads, max_transition_length, links_feature="feature:position", fused=True):
    """Drop-in for `GlatLinkDecoder.extract_links(features, prev_output_tokens, link_positional, query_linear, key_linear,
    gate_linear)` (s2t_conformer_dag.py:171-212); what the reference reads from `self.args` / `self.pad` is passed by
    keyword.  Only the banded form (`max_transition_length != -1`) is provided."""
    if max_transition_length == -1:
        raise NotImplementedError("extract_links: the dense form (max_transition_length == -1) is not provided")
    parts = []
    names = links_feature.split(":")
    if "feature" in names:
        parts.append(features)
    if "position" in names or "sinposition" in names:
        parts.append(link_positional(prev_output_tokens))
    x = torch.cat(parts, dim=-1)
    B, L = features.shape[0], features.shape[1]
    H = decoder_attention_heads
    q = query_linear(x)
    k = key_linear(x)
    Fd = q.shape[-1] // H
    query_chunks = q.reshape(B, L, H, Fd)
    key_chunks = k.reshape(B, L, H, Fd)
    log_gates = F.log_softmax(gate_linear(x), dim=-1, dtype=torch.float32)
    translen = min(int(max_transition_length), L - 1)
    output_length = prev_output_tokens.ne(pad).sum(dim=-1)
    return extract_links_from_chunks(query_chunks, key_chunks, log_gates, output_length, translen, fused=fused)
