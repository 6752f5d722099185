"""Host-side mirror of the criterion code that drives the DAG-loss hot path.

The reference's criterions (`DASpeech/criterions/nat_dag_loss.py`, `s2s_dag_fastspeech2_loss.py`) cannot be imported
without fairseq, so the three call chains that fix the operator contract are restated here as plain functions with the
reference's argument meaning and return values, in two flavours selected by `fused`:

    fused=True   the B200 path: gather with fused arg-max, Viterbi + `glat_alignment`, `glat_force_emit`, `dag_loss`,
                 `dag_posterior` -- what a maintainer gets after applying INTEGRATION.md
    fused=False  the reference's own op sequence written with the exported operators only (the four custom ops or, with
                 `use_torch_ops=True`, their `torch_*` versions: the criterions' --torch-dag-* switches)

    compute_dag_loss(...)             NATDAGLoss._compute_dag_loss            nat_dag_loss.py:114-156
    compute_dag_loss_with_alpha_beta  S2SDAGFastSpeech2Loss._compute_dag_loss_with_alpha_beta
                                                                              s2s_dag_fastspeech2_loss.py:53-91
    glat_function(...)                the closure of NATDAGLoss.forward       nat_dag_loss.py:202-264
    expected_features(...)            training strategy "expect"              s2s_dag_fastspeech2_loss.py:257-263

The tests run both flavours on the same inputs and random draws and compare losses, gradients, masks and tokens.
"""
import torch

from . import custom_ops as _ops
from .custom_ops.dag_loss import dag_logsoftmax_gather_argmax_inplace
from .glat import glat_alignment, glat_force_emit
from .posterior import dag_expected_features


def restore_valid_links(links, max_transition_length=99999):
    """Banded [B, L, T] transitions -> dense [B, L, L] (models/s2t_conformer_dag.py:157-169), for the torch_* operators."""
    bsz, prelen, translen = links.shape
    translen = min(max_transition_length, prelen - 1, translen)
    idx = torch.arange(prelen, device=links.device).unsqueeze(1) + torch.arange(translen, device=links.device).unsqueeze(0) + 1
    idx = idx.masked_fill(idx >= prelen, prelen)
    dense = links.new_full((bsz, prelen, prelen + 1), float("-inf"))
    dense.scatter_(2, idx.unsqueeze(0).expand(bsz, -1, -1), links[:, :, :translen])
    return dense[:, :, :prelen]


def _emissions(outputs, targets, use_torch_ops, want_argmax=False):
    prelen = outputs.shape[1]
    idx = targets.unsqueeze(1).expand(-1, prelen, -1)
    pred = None
    if use_torch_ops:
        if want_argmax:
            pred = outputs.argmax(-1)
        outputs, match_all = _ops.torch_dag_logsoftmax_gather_inplace(outputs, idx)
    elif want_argmax:
        outputs, match_all, pred = dag_logsoftmax_gather_argmax_inplace(outputs, idx)
    else:
        outputs, match_all = _ops.dag_logsoftmax_gather_inplace(outputs, idx)
    return outputs, match_all.transpose(1, 2), pred


def _force_emit(match_all, matchmask, keep_word_mask, fused):
    if fused:
        return glat_force_emit(match_all, matchmask, keep_word_mask)
    prev = keep_word_mask.unsqueeze(1)                                                    # nat_dag_loss.py:130-132
    return match_all.masked_fill(prev, 0) + match_all.masked_fill(~matchmask, float("-inf")).masked_fill(~prev, 0).detach()


def _reduce(loss_result, targets, target_length, output_masks, pad, name, factor):
    invalid = loss_result.isinf().logical_or(loss_result.isnan())                        # nat_dag_loss.py:143-147
    loss_result = loss_result.masked_fill(invalid, 0)
    loss = -(loss_result / target_length).mean()
    return {"name": name, "loss": loss * factor, "nll_loss": loss.detach(), "factor": factor,
            "ntokens": targets.ne(pad).sum(), "nvalidtokens": output_masks.sum(), "nsentences": targets.shape[0],
            "loss_nofactor": loss, "invalid_nsentences": invalid.sum().detach()}


def compute_dag_loss(outputs, output_masks, targets, target_masks, links, name="loss", factor=1.0, matchmask=None,
                     keep_word_mask=None, pad=1, no_force_emit=False, fused=True, use_torch_ops=False,
                     max_transition_length=99999):
    """outputs [B, L, V] logits (OVERWRITTEN by the CUDA gather when they require grad), output_masks [B, L] bool,
    targets [B, M] long, target_masks [B, M] bool, links [B, L, T]."""
    output_length = output_masks.sum(dim=-1)
    target_length = target_masks.sum(dim=-1)
    _, match_all, _ = _emissions(outputs, targets, use_torch_ops)
    if matchmask is not None and not no_force_emit:
        match_all = _force_emit(match_all, matchmask, keep_word_mask, fused and not use_torch_ops)
    if use_torch_ops:
        loss_result = _ops.torch_dag_loss(match_all, restore_valid_links(links, max_transition_length), output_length, target_length)
    else:
        assert max_transition_length != -1, "cuda dag loss does not support max_transition_length=-1. You can use a very large number such as 99999"
        loss_result = _ops.dag_loss(match_all, links, output_length, target_length)
    return _reduce(loss_result, targets, target_length, output_masks, pad, name, factor)


def compute_dag_loss_with_alpha_beta(outputs, output_masks, targets, target_masks, links, name="loss", factor=1.0,
                                     matchmask=None, keep_word_mask=None, pad=1, no_force_emit=False, fused=True):
    """As compute_dag_loss, additionally returning the (alpha, beta) lattices (CUDA operators only, as the reference)."""
    output_length = output_masks.sum(dim=-1)
    target_length = target_masks.sum(dim=-1)
    _, match_all, _ = _emissions(outputs, targets, False)
    if matchmask is not None and not no_force_emit:
        match_all = _force_emit(match_all, matchmask, keep_word_mask, fused)
    loss_result, (alpha, beta) = _ops.dag_loss_with_alpha_beta(match_all, links, output_length, target_length)
    return _reduce(loss_result, targets, target_length, output_masks, pad, name, factor), alpha, beta


def expected_features(alpha, beta, features, fused=True):
    """z_i = sum_j P(a_i = j | x, y) v_j without <bos> (s2s_dag_fastspeech2_loss.py:257-263)."""
    if fused:
        return dag_expected_features(alpha, beta, features)[:, 1:, :]
    score = (alpha + beta - _ops.logsumexp_keepdim(alpha + beta, dim=-1)).exp()
    score.masked_fill_(torch.isnan(score), 0)
    return torch.matmul(score.to(features), features)[:, 1:, :]


def glat_function(word_ins_out, tgt_tokens, prev_output_tokens, glat, links, pad=1, glance_strategy=None, fused=True,
                  use_torch_ops=False, max_transition_length=99999):
    """The glancing pass: which vertices get to see their aligned target token in the second decoder pass.
    Returns (glat_prev_output_tokens, glat_tgt_tokens, glat_info) as the reference's closure does."""
    prelen = links.shape[1]
    tarlen = tgt_tokens.shape[1]
    target_length = (~tgt_tokens.eq(pad)).sum(1)
    output_length = prev_output_tokens.ne(pad).sum(1)
    fused = fused and not use_torch_ops
    _, match, pred_tokens = _emissions(word_ins_out, tgt_tokens, use_torch_ops, want_argmax=True)
    if fused:
        al = glat_alignment(match, links, output_length, target_length, tgt_tokens, pred_tokens)
        predict_align_mask, matchmask, oracle, same_num = al["predict_align_mask"], al["matchmask"], al["oracle"], al["same_num"]
    else:
        if use_torch_ops:
            path = _ops.torch_dag_best_alignment(match, restore_valid_links(links, max_transition_length), output_length, target_length)
        else:
            path = _ops.dag_best_alignment(match, links, output_length, target_length)
        predict_align_mask = path >= 0                                                   # nat_dag_loss.py:223-227
        matchmask = torch.zeros(tgt_tokens.shape[0], tarlen + 1, prelen, device=match.device, dtype=torch.bool) \
            .scatter_(1, path.unsqueeze(1) + 1, 1)[:, 1:]
        oracle = tgt_tokens.gather(-1, path.clip(min=0))
        same_num = ((pred_tokens == oracle) & predict_align_mask).sum(1)

    if glance_strategy is None:
        keep_prob = ((target_length - same_num) / target_length * glat["context_p"]).unsqueeze(-1) * predict_align_mask.float()
    elif glance_strategy in ("number-random", "cmlm"):
        prob = torch.randn(oracle.shape, device=tgt_tokens.device, dtype=torch.float)
        prob.masked_fill_(~predict_align_mask, -100)
        if glance_strategy == "number-random":
            glance_nums = ((target_length - same_num) * glat["context_p"] + 0.5).to(torch.long)
        else:
            glance_nums = (target_length * torch.rand_like(target_length, dtype=torch.float) + 0.5).to(torch.long)
        prob_thresh = prob.sort(descending=True)[0].gather(-1, (glance_nums - 1).clip(min=0).unsqueeze(-1)).squeeze(-1)
        prob_thresh.masked_fill_(glance_nums == 0, 100)
        keep_prob = (prob >= prob_thresh.unsqueeze(-1)).to(prob.dtype)
    else:
        raise ValueError("unknown glance strategy %r" % (glance_strategy,))

    keep_word_mask = (torch.rand(prev_output_tokens.shape, device=prev_output_tokens.device) < keep_prob).bool()
    glat_prev_output_tokens = prev_output_tokens.masked_fill(keep_word_mask, 0) + oracle.masked_fill(~keep_word_mask, 0)
    glat_info = {"glat_accu": (same_num.sum() / target_length.sum()).detach(), "glat_context_p": glat["context_p"],
                 "glat_keep": keep_prob.mean().detach(), "matchmask": matchmask, "keep_word_mask": keep_word_mask,
                 "glat_prev_output_tokens": glat_prev_output_tokens}
    return glat_prev_output_tokens, tgt_tokens, glat_info
