"""Alignment posterior of the S2S criterion on the lattices of `dag_loss_with_alpha_beta`.

The reference (DASpeech/criterions/s2s_dag_fastspeech2_loss.py:257-265, training strategy "expect") computes

    score = (alpha + beta - logsumexp_keepdim(alpha + beta, dim=-1)).exp()      # :259
    score.masked_fill_(torch.isnan(score), 0)                                    # :260
    score = score.to(features)                                                   # :261
    expect_features = torch.matmul(score, features)                              # :262

`dag_posterior` is lines 259-261 as ONE pass over the two lattices (`dagb200_dag_posterior`, dag_posterior.cu);
`dag_expected_features` adds line 262 (cuBLAS through torch.matmul, differentiable with respect to `features` --
alpha and beta are non-differentiable outputs in the reference too, dag_loss.py:176-179).  No CPU path.
"""
import torch

from . import _lib
from .custom_ops.dag_loss import _DTYPE_CODE, _check, _ptr, _stream


def dag_posterior(alpha: torch.Tensor, beta: torch.Tensor, dtype: torch.dtype = None) -> torch.Tensor:
    """P(a_t = j | x, y) as a `[B, M, L]` tensor of `dtype` (default: fp32); rows without a finite cell are all zero."""
    _check(alpha.is_cuda and beta.is_cuda, "You need GPU to use the custom cuda operations")
    _check(alpha.dim() == 3 and alpha.shape == beta.shape, "alpha and beta should be [bsz, tarlen, prelen] lattices of one shape")
    _check(alpha.dtype == torch.float32 and beta.dtype == torch.float32, "dag_posterior expects the fp32 lattices of dag_loss_with_alpha_beta")
    dtype = dtype or torch.float32
    _check(dtype in (torch.float32, torch.float16, torch.bfloat16), "unsupported posterior dtype %s" % dtype)
    alpha, beta = alpha.detach().contiguous(), beta.detach().contiguous()
    B, M, L = alpha.shape
    score = torch.empty((B, M, L), dtype=dtype, device=alpha.device)
    with torch.cuda.device(alpha.device):
        rc = _lib.load().dagb200_dag_posterior(_ptr(alpha), _ptr(beta), _ptr(score), _DTYPE_CODE[dtype], B, M, L, _stream())
    _lib.check(rc, "dag_posterior")
    return score


def dag_expected_features(alpha: torch.Tensor, beta: torch.Tensor, features: torch.Tensor) -> torch.Tensor:
    """z_t = sum_j P(a_t = j | x, y) v_j  (`[B, M, D]`, dtype of `features`; gradient flows into `features` only)."""
    return torch.matmul(dag_posterior(alpha, beta, features.dtype), features)
