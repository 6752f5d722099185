"""GLAT force-emit masking of the emission plane.

The criterion (DASpeech/criterions/nat_dag_loss.py:130-132) pins every glanced vertex to the target its alignment chose:

    glat_prev_mask = keep_word_mask.unsqueeze(1)
    match_all = match_all.masked_fill(glat_prev_mask, 0) + \
                match_all.masked_fill(~matchmask, float("-inf")).masked_fill(~glat_prev_mask, 0).detach()

`glat_force_emit(match_all, matchmask, keep_word_mask)` is the same function (values and gradient) as one streaming
kernel each way (`dagb200_glat_force_emit`, dag_glat.cu).  No CPU path.
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
