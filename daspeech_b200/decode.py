"""Inference decoding over the DAG on the GPU.

Mirrors the decoding branch of `S2SConformerDAGFastSpeech2.forward_decoder`-style generation in
`DASpeech/models/s2s_conformer_dag_fastspeech2.py:211-304` (strategies "greedy", "lookahead", "viterbi",
"jointviterbi"), which the reference runs as Python loops over `.tolist()`-ed tensors -- one host synchronisation per
batch and a Python iteration per token.  Here:

  * greedy / lookahead: one kernel (`dagb200_decode_lookahead`): row arg-max over the banded transitions, then the walk.
  * viterbi / jointviterbi: the free-length max-plus recurrence `alpha_s[j] = max_i (alpha_{s-1}[i] + links[i][j]) (+ emission)`
    is the max-plus lattice of the training-time alignment (`dagb200_dag_best_alignment`, wave kernel, lattice output) on
    an emission plane that does not depend on a target: row s+1 of that lattice IS the reference's `scores[s]`, bit for
    bit (one add per candidate, one per cell).  `dagb200_decode_viterbi_finish` adds the end transition, applies the
    length penalty, picks the length, walks back (arg-max recomputed for the cells on the path, torch.max's tie-break)
    and removes duplicates.

`dag_decode(...)` returns (tokens [B, Nmax] long padded with `pad`, features [B, Fmax, D] (zeros beyond the length),
feature_lengths [B] long) -- what the reference hands to the FastSpeech2 decoder (`output_tokens`, `output_features`,
`output_features_length`).  The transitions are the BANDED [B, L, T] tensor the model's `extract_links` produces (the
reference first expands it with `restore_valid_links`).  No CPU path.
"""
import torch

from . import _lib
from .custom_ops.dag_loss import _check, _ptr, _stream, get_dag_kernel


def _emission_plane(vlogit, S, beta, joint):
    """[B, S+1, L] emissions of the alignment kernel such that lattice row s+1 equals the reference's scores[s]
    (s2s_conformer_dag_fastspeech2.py:249-262): row 0 carries the start vertex (plus its own emission for the joint
    variant, :251-252), every later row the vertex emissions for the joint variant; the plain variant adds them to the
    first step only (:254)."""
    B, L = vlogit.shape
    em = vlogit.new_zeros((B, S + 1, L))
    e = vlogit * beta
    if joint:
        em[:, 1:] = e.unsqueeze(1)
        em[:, 0, 0] = e[:, 0]
    else:
        em[:, 1] = e
    return em


def dag_decode(output_logits, links, output_length, features=None, strategy="lookahead", decode_beta=1.0,
               decode_viterbibeta=1.0, src_upsample_scale=0.5, pad=1, max_length=None):
    _check(output_logits.is_cuda and links.is_cuda, "You need GPU to use the custom cuda operations")
    _check(strategy in ("greedy", "lookahead", "viterbi", "jointviterbi"), "unknown decode strategy %r" % (strategy,))
    with torch.no_grad():
        B, L, T = links.shape
        links = links.float().contiguous()
        output_length = output_length.contiguous()
        vlogit, vtoken = output_logits.log_softmax(dim=-1).max(dim=-1)            # :215-216
        vlogit = vlogit.float().contiguous()
        vtoken = vtoken.contiguous()
        dev = links.device
        tokens = torch.empty((B, L), dtype=torch.long, device=dev)
        vertices = torch.empty((B, L), dtype=torch.int32, device=dev)
        lengths = torch.empty((B, 2), dtype=torch.int32, device=dev)
        lib = _lib.load()
        if strategy in ("greedy", "lookahead"):
            beta = float(decode_beta) if strategy == "lookahead" else 0.0
            with torch.cuda.device(dev):
                rc = lib.dagb200_decode_lookahead(_ptr(links), _ptr(vlogit), _ptr(vtoken), _ptr(output_length), beta, int(pad),
                                                  B, L, T, _ptr(tokens), _ptr(vertices), _ptr(lengths), _stream())
            _lib.check(rc, "decode_lookahead")
        else:
            S = int(max_length) if max_length is not None else int(L / 8 / src_upsample_scale)      # :258
            S = max(1, min(S, L - 1))
            em = _emission_plane(vlogit, S, float(decode_beta), strategy == "jointviterbi")
            olen_all = torch.full((B,), L, dtype=torch.long, device=dev)           # the recurrence runs over all vertices
            tlen = torch.full((B,), S + 1, dtype=torch.long, device=dev)
            # only the lattice is used: the end cell of this pseudo-alignment need not be reachable
            lattice, _ = get_dag_kernel().dag_best_alignment(em, links, olen_all, tlen, 1, want_alpha=True, track_status=False)
            scratch = torch.empty((B, S), dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                rc = lib.dagb200_decode_viterbi_finish(_ptr(lattice), _ptr(links), _ptr(vtoken), _ptr(output_length),
                                                       float(decode_viterbibeta), int(pad), B, S, L, T, _ptr(tokens),
                                                       _ptr(vertices), _ptr(lengths), _ptr(scratch), _stream())
            _lib.check(rc, "decode_viterbi_finish")
        nmax = lengths.max(dim=0).values.tolist()                                  # ONE host synchronisation per batch
        tokens = tokens[:, :max(nmax[0], 1)]
        flen = lengths[:, 1].long()
        if features is None:
            return tokens, vertices[:, :max(nmax[1], 1)], flen
        vt = vertices[:, :max(nmax[1], 1)].long()
        feats = features.gather(1, vt.clamp(min=0).unsqueeze(-1).expand(-1, -1, features.shape[-1]))
        feats = feats.masked_fill((vt < 0).unsqueeze(-1), 0)
        return tokens, feats, flen
