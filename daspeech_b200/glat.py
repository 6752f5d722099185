"""GLAT force-emit masking of the emission plane.

The criterion (DASpeech/criterions/nat_dag_loss.py:130-132) pins every glanced vertex to the target its alignment chose:

    glat_prev_mask = keep_word_mask.unsqueeze(1)
    match_all = match_all.masked_fill(glat_prev_mask, 0) + \
                match_all.masked_fill(~matchmask, float("-inf")).masked_fill(~glat_prev_mask, 0).detach()

`glat_force_emit(match_all, matchmask, keep_word_mask)` is the same function (values and gradient) as one streaming
kernel each way (`dagb200_glat_force_emit`, dag_glat.cu).  No CPU path.

`glat_alignment(match, links, output_length, target_length, tgt_tokens, pred_tokens)` is the Viterbi alignment of the
glancing pass together with everything the criterion derives from the path (nat_dag_loss.py:221-227: `path`,
`predict_align_mask`, `matchmask`, `oracle`, `same_num`): the Viterbi kernel followed by ONE kernel that writes the
[B, M, L] mask plane exactly once (`dagb200_glat_alignment`) instead of a [B, M+1, L] zero fill, a scatter, a slice, a
gather and a masked reduction.
"""
import torch

from . import _lib
from .custom_ops.dag_loss import _check, _ptr, _stream


class GlatForceEmitFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, match_all, matchmask, keep_word_mask):
        _check(match_all.is_cuda, "You need GPU to use the custom cuda operations")
        _check(match_all.dim() == 3 and match_all.dtype == torch.float32, "match_all should be an fp32 [bsz, tarlen, prelen] tensor")
        B, M, L = match_all.shape
        _check(matchmask.shape == (B, M, L) and matchmask.dtype == torch.bool, "matchmask should be bool [bsz, tarlen, prelen]")
        _check(keep_word_mask.shape == (B, L) and keep_word_mask.dtype == torch.bool, "keep_word_mask should be bool [bsz, prelen]")
        m = match_all.contiguous()
        mm = matchmask.contiguous()
        keep = keep_word_mask.contiguous()
        out = torch.empty_like(m)
        with torch.cuda.device(m.device):
            rc = _lib.load().dagb200_glat_force_emit(_ptr(m), _ptr(mm), _ptr(keep), _ptr(out), B, M, L, 0, _stream())
        _lib.check(rc, "glat_force_emit")
        ctx.save_for_backward(keep)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        keep, = ctx.saved_tensors
        g = grad_out.contiguous()
        B, M, L = g.shape
        gin = torch.empty_like(g)
        with torch.cuda.device(g.device):
            rc = _lib.load().dagb200_glat_force_emit(_ptr(g), None, _ptr(keep), _ptr(gin), B, M, L, 1, _stream())
        _lib.check(rc, "glat_force_emit backward")
        return gin, None, None


glat_force_emit = GlatForceEmitFunc.apply


def glat_alignment(match, links, output_length, target_length, tgt_tokens, pred_tokens=None):
    """Returns dict(path [B,L] long, predict_align_mask [B,L] bool, matchmask [B,M,L] bool, oracle [B,L] long,
    same_num [B] long or None).  Non-differentiable, like `dag_best_alignment` (custom_ops/dag_loss.py:190-236)."""
    from .custom_ops.dag_loss import get_dag_kernel, DagBestAlignmentFunc
    with torch.no_grad():
        k = get_dag_kernel()
        match = match.detach().contiguous()
        links = links.detach().contiguous()
        _, path32 = k.dag_best_alignment(match, links, output_length, target_length, DagBestAlignmentFunc.config,
                                         want_alpha=False)
        B, M, L = match.shape
        _check(tgt_tokens.shape == (B, M) and tgt_tokens.dtype == torch.long, "tgt_tokens should be long [bsz, tarlen]")
        if pred_tokens is not None:
            _check(pred_tokens.shape == (B, L) and pred_tokens.dtype == torch.long, "pred_tokens should be long [bsz, prelen]")
            pred_tokens = pred_tokens.contiguous()
        dev = match.device
        matchmask = torch.empty((B, M, L), dtype=torch.bool, device=dev)
        oracle = torch.empty((B, L), dtype=torch.long, device=dev)
        path = torch.empty((B, L), dtype=torch.long, device=dev)
        align = torch.empty((B, L), dtype=torch.bool, device=dev)
        same = torch.empty((B,), dtype=torch.long, device=dev) if pred_tokens is not None else None
        with torch.cuda.device(dev):
            rc = _lib.load().dagb200_glat_alignment(_ptr(path32), _ptr(tgt_tokens), tgt_tokens.stride(0), tgt_tokens.stride(1),
                                                    _ptr(pred_tokens), _ptr(matchmask), _ptr(oracle), _ptr(path), _ptr(align),
                                                    _ptr(same), B, M, L, _stream())
        _lib.check(rc, "glat_alignment")
    return {"path": path, "predict_align_mask": align, "matchmask": matchmask, "oracle": oracle, "same_num": same}
