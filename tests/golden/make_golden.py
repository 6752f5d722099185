"""Generate tests/golden/*.npz by running the REFERENCE's own torch implementations.

Run in the build container only (it imports /root/reference/DASpeech/custom_ops/dag_loss.py by file
path; that file needs nothing but torch).  The GPU box has no /root/reference, so the vectors are
committed and this script is kept as their provenance:

    python tests/golden/make_golden.py

Reference functions exercised (DASpeech/custom_ops/dag_loss.py):
    torch_dag_loss (:325-366) + autograd        -> loss, grad_match, grad_links
    __torch_max_loss (:369-386)                 -> Viterbi score
    torch_dag_best_alignment (:388-419)         -> path
    torch_dag_logsoftmax_gather_inplace (:421)  -> match + autograd grad of the logits
    restore_valid_links (models/s2t_conformer_dag.py:157-169 semantics, via oracle.dense_links which is
                         checked against the reference's own test helper dag_loss.py:439-448 below)
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402  (input generator + layout helper only)

REF = "/root/reference/DASpeech/custom_ops/dag_loss.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_dag_loss", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_restore_valid_links(links):
    """The reference's scatter-based band->dense conversion (same recipe as dag_loss.py:439-448)."""
    B, L, T = links.shape
    idx = torch.arange(L).unsqueeze(1) + torch.arange(T).unsqueeze(0) + 1
    inval = idx >= L
    idx = idx.masked_fill(inval, L)
    res = torch.full((B, L, L + 1), float("-inf"), dtype=links.dtype)
    res.scatter_(2, idx.unsqueeze(0).expand(B, -1, -1), links)
    return res[:, :, :L]


def dp_case(ref, name, B, L, M, T, seed, ragged, glat_frac=0.0, dtype=torch.float64, kill_sample=None):
    match, links, olen, tlen = oracle.make_lattice(B, L, M, T, seed=seed, ragged=ragged,
                                                   dtype=np.float32, glat_frac=glat_frac)
    if kill_sample is not None:
        # make one sample infeasible: vertex 0 has no outgoing edge at all
        links[kill_sample, 0, :] = -np.inf
    rng = np.random.default_rng(seed + 1000)
    go = (rng.random(B) + 0.5).astype(np.float32)

    m = torch.tensor(match, dtype=dtype, requires_grad=True)
    lk = torch.tensor(links, dtype=dtype, requires_grad=True)
    ol, tl = torch.tensor(olen), torch.tensor(tlen)
    dense = ref_restore_valid_links(lk)
    assert np.array_equal(dense.detach().numpy(), oracle.dense_links(lk.detach().numpy()))
    loss = ref.torch_dag_loss(m, dense, ol, tl)
    out = dict(match=match, links=links, olen=olen, tlen=tlen, grad_output=go,
               loss=loss.detach().numpy())
    feasible = torch.isfinite(loss)
    if bool(feasible.all()):
        gm, gl = torch.autograd.grad((loss * torch.tensor(go, dtype=dtype)).sum(), [m, lk])
        out["grad_match"] = gm.numpy()
        out["grad_links"] = gl.numpy()
        with torch.no_grad():
            score = getattr(ref, "__torch_max_loss")(m.detach(), dense.detach(), ol, tl)
        path = ref.torch_dag_best_alignment(m.detach().clone(), dense.detach(), ol, tl)
        out["viterbi_score"] = score.numpy()
        out["viterbi_path"] = path.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", out["loss"])


def gather_case(ref, name, B, L, V, S, seed, half):
    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((B, L, V)) * 2).astype(np.float16 if half else np.float32)
    tgt = rng.integers(0, V, size=(B, S)).astype(np.int64)
    tgt[:, 1] = tgt[:, 0]  # duplicated target id: scatter_add must accumulate
    x = torch.tensor(logits.astype(np.float32), requires_grad=True)
    idx = torch.tensor(tgt).unsqueeze(1).expand(-1, L, -1)
    _, sel = ref.torch_dag_logsoftmax_gather_inplace(x, idx)
    w = rng.standard_normal((B, L, S)).astype(np.float32)
    g = torch.autograd.grad((sel * torch.tensor(w)).sum(), [x])[0]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), logits=logits, targets=tgt, selected=sel.detach().numpy(),
                        grad_selected=w, grad_logits=g.numpy())
    print(name, float(sel.mean()))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    ref = load_reference()
    # BASELINE config C1: B=2, L=64, M=32, V=512
    dp_case(ref, "c1_full", 2, 64, 32, 63, seed=1, ragged=False)
    dp_case(ref, "c1_ragged", 2, 64, 32, 63, seed=2, ragged=True)
    dp_case(ref, "c1_band8", 3, 64, 32, 8, seed=3, ragged=True)
    dp_case(ref, "c1_glat", 2, 64, 32, 63, seed=4, ragged=True, glat_frac=0.3)
    dp_case(ref, "c1_infeasible", 2, 64, 32, 63, seed=5, ragged=True, kill_sample=1)
    dp_case(ref, "tiny_min", 2, 5, 2, 4, seed=6, ragged=False)
    dp_case(ref, "mid_ragged", 3, 160, 48, 32, seed=7, ragged=True)
    dp_case(ref, "c1_full_fp32", 2, 64, 32, 63, seed=1, ragged=False, dtype=torch.float32)
    gather_case(ref, "gather_c1_fp32", 2, 64, 512, 32, seed=11, half=False)
    gather_case(ref, "gather_c1_fp16", 2, 64, 512, 32, seed=12, half=True)
    gather_case(ref, "gather_odd", 3, 7, 1237, 5, seed=13, half=False)


if __name__ == "__main__":
    main()
