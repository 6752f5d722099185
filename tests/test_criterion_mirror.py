"""daspeech_b200.criterions (the host-side mirror of the criterion call chains) against
vectors produced by the UNMODIFIED reference criterion (tests/golden/make_golden_criterion.py: NATDAGLoss.forward with
its glat_function closure and _compute_dag_loss, run on CPU with the criterion's own --torch-dag-* switches).

CPU: the torch flavour of the mirror reproduces loss, gradients, masks and glanced tokens (same seed, same RNG calls).
GPU: the fused B200 flavour reproduces the deterministic outputs and, fed with the golden's masks, loss and gradients;
fused and unfused flavours agree on the GPU under the same seed."""
import os

import numpy as np
import pytest
import torch

from daspeech_b200 import criterions as C

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAD = 1
CASES = ["criterion_plain", "criterion_glat", "criterion_glat_number_random"]
STRATEGY = {"criterion_plain": None, "criterion_glat": None, "criterion_glat_number_random": "number-random"}


def load(name, dev):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    t = {k: torch.tensor(g[k]).to(dev) for k in ("logits", "links", "tgt", "prev")}
    return g, t


def run_chain(t, glat_p, strategy, seed, dev, fused, use_torch_ops, masks=None):
    """What NATDAGLoss.forward + the model's forward do around the operators (nat_dag_loss.py:186-283,
    models/s2t_conformer_dag.py:236-266), with the network replaced by fixed logits / links."""
    logits = t["logits"].clone().requires_grad_()
    links = t["links"].clone().requires_grad_()
    tgt, prev = t["tgt"], t["prev"]
    info = None
    if masks is not None:
        info = masks
    elif glat_p > 0:
        torch.manual_seed(seed)
        with torch.no_grad():
            _, _, info = C.glat_function(logits.detach().clone(), tgt, prev, {"context_p": glat_p}, links.detach(), pad=PAD,
                                         glance_strategy=strategy, fused=fused, use_torch_ops=use_torch_ops)
    out = C.compute_dag_loss(logits * 1, prev.ne(PAD), tgt, tgt.ne(PAD), links * 1, name="dag-loss", factor=1,
                             matchmask=None if info is None else info["matchmask"],
                             keep_word_mask=None if info is None else info["keep_word_mask"], pad=PAD, fused=fused,
                             use_torch_ops=use_torch_ops)
    out["loss"].backward()
    return out, info, logits.grad, links.grad


def relerr(x, ref):
    x, ref = np.asarray(x, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.mark.parametrize("name", CASES)
def test_torch_flavour_reproduces_the_reference_criterion_on_cpu(name):
    g, t = load(name, "cpu")
    out, info, glog, glk = run_chain(t, float(g["glat_p"]), STRATEGY[name], int(g["seed"]), "cpu", fused=False, use_torch_ops=True)
    assert abs(float(out["loss"]) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert relerr(glog.numpy(), g["grad_logits"]) <= 1e-5 and relerr(glk.numpy(), g["grad_links"]) <= 1e-5
    assert int(out["ntokens"]) == int(g["ntokens"]) and int(out["nvalidtokens"]) == int(g["nvalidtokens"])
    assert int(out["invalid_nsentences"]) == int(g["invalid_nsentences"])
    if info is not None:
        assert np.array_equal(info["matchmask"].numpy(), g["matchmask"])
        assert np.array_equal(info["keep_word_mask"].numpy(), g["keep_word_mask"])
        assert np.array_equal(info["glat_prev_output_tokens"].numpy(), g["glat_prev_output_tokens"])
        assert abs(float(info["glat_accu"]) - float(g["glat_accu"])) <= 1e-6
        assert abs(float(info["glat_keep"]) - float(g["glat_keep"])) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_fused_flavour_matches_the_reference_criterion_on_gpu(name):
    g, t = load(name, "cuda")
    glat_p = float(g["glat_p"])
    masks = None
    if glat_p > 0:
        # deterministic part of the glancing pass: alignment, mask plane, oracle tokens, match count
        with torch.no_grad():
            torch.manual_seed(1)
            _, _, info = C.glat_function(t["logits"].clone(), t["tgt"], t["prev"], {"context_p": glat_p}, t["links"], pad=PAD,
                                         glance_strategy=STRATEGY[name], fused=True)
        assert np.array_equal(info["matchmask"].cpu().numpy(), g["matchmask"])
        assert abs(float(info["glat_accu"]) - float(g["glat_accu"])) <= 1e-6
        keep = info["keep_word_mask"]
        assert not (keep & ~info["matchmask"].any(1)).any()              # only aligned vertices are glanced
        glanced = info["glat_prev_output_tokens"][keep]
        oracle = t["tgt"].gather(-1, torch.tensor(g["matchmask"]).to("cuda").float().argmax(1))[keep]
        assert torch.equal(glanced, oracle)
        masks = {"matchmask": torch.tensor(g["matchmask"]).cuda(), "keep_word_mask": torch.tensor(g["keep_word_mask"]).cuda()}
    # loss and gradients with the golden's own masks
    out, _, glog, glk = run_chain(t, glat_p, STRATEGY[name], int(g["seed"]), "cuda", fused=True, use_torch_ops=False, masks=masks)
    assert abs(float(out["loss"]) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    assert relerr(glog.cpu().numpy(), g["grad_logits"]) <= 1e-4 and relerr(glk.cpu().numpy(), g["grad_links"]) <= 1e-4
    assert int(out["ntokens"]) == int(g["ntokens"]) and int(out["invalid_nsentences"]) == int(g["invalid_nsentences"])
    # fused and unfused operator sequences, same seed on the same device: identical masks, matching numbers
    a = run_chain(t, glat_p, STRATEGY[name], 11, "cuda", fused=True, use_torch_ops=False)
    b = run_chain(t, glat_p, STRATEGY[name], 11, "cuda", fused=False, use_torch_ops=False)
    if glat_p > 0:
        assert torch.equal(a[1]["matchmask"], b[1]["matchmask"]) and torch.equal(a[1]["keep_word_mask"], b[1]["keep_word_mask"])
        assert torch.equal(a[1]["glat_prev_output_tokens"], b[1]["glat_prev_output_tokens"])
    assert abs(float(a[0]["loss"]) - float(b[0]["loss"])) <= 1e-6 * abs(float(b[0]["loss"]))
    assert relerr(a[2].cpu().numpy(), b[2].cpu().numpy()) <= 1e-5 and relerr(a[3].cpu().numpy(), b[3].cpu().numpy()) <= 1e-5


@pytest.mark.gpu
def test_s2s_chain_expected_features_fused_vs_reference_ops():
    """S2S criterion, training strategy "expect" (s2s_dag_fastspeech2_loss.py:53-91, 257-263): lattices from
    dag_loss_with_alpha_beta, posterior expectation of the decoder features; fused kernel vs the criterion's torch ops."""
    g, t = load("criterion_plain", "cuda")
    feats = torch.randn(t["logits"].shape[0], t["logits"].shape[1], 24, device="cuda", requires_grad=True)
    res = []
    for fused in (True, False):
        logits = t["logits"].clone().requires_grad_()
        out, alpha, beta = C.compute_dag_loss_with_alpha_beta(logits * 1, t["prev"].ne(PAD), t["tgt"], t["tgt"].ne(PAD),
                                                              t["links"].clone().requires_grad_(), pad=PAD, fused=fused)
        z = C.expected_features(alpha, beta, feats, fused=fused)
        res.append((float(out["loss"]), z))
    assert abs(res[0][0] - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    assert res[0][1].shape == (t["tgt"].shape[0], t["tgt"].shape[1] - 1, 24)
    assert torch.allclose(res[0][1], res[1][1], rtol=1e-4, atol=1e-5)
