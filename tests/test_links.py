"""extract_links against golden vectors produced by the UNMODIFIED reference functions
(tests/golden/make_golden_links.py runs `extract_links` / `extract_valid_links` of s2t_conformer_dag.py:140-212):
the torch mirror on CPU, the fused tcgen05 kernel + chunked backward on the GPU."""
import glob
import os

import numpy as np
import pytest
import torch

from daspeech_b200 import links as dl

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "links")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "links_*.npz")))


def _modules(g, dev):
    D = g["features"].shape[-1]
    H = int(g["H"])
    ql, kl, gl = torch.nn.Linear(2 * D, D), torch.nn.Linear(2 * D, D), torch.nn.Linear(2 * D, H)
    with torch.no_grad():
        ql.weight.copy_(torch.tensor(g["qw"])); ql.bias.copy_(torch.tensor(g["qb"]))
        kl.weight.copy_(torch.tensor(g["kw"])); kl.bias.copy_(torch.tensor(g["kb"]))
        gl.weight.copy_(torch.tensor(g["gw"])); gl.bias.copy_(torch.tensor(g["gb"]))
    pos = torch.tensor(g["pos"], device=dev)
    return ql.to(dev), kl.to(dev), gl.to(dev), (lambda t: pos)


def _run(g, dev, fused):
    ql, kl, gl, link_positional = _modules(g, dev)
    features = torch.tensor(g["features"], device=dev, requires_grad=True)
    tokens = torch.tensor(g["tokens"], device=dev)
    links = dl.extract_links(features, tokens, link_positional, ql, kl, gl, pad=int(g["pad"]),
                             decoder_attention_heads=int(g["H"]), max_transition_length=int(g["T"]), fused=fused)
    fin = torch.isfinite(links)
    (links.masked_fill(~fin, 0.0) * torch.tensor(g["w"], device=dev)).sum().backward()
    return links.detach().cpu().numpy(), features.grad.cpu().numpy(), ql.weight.grad.cpu().numpy(), \
        kl.weight.grad.cpu().numpy(), gl.weight.grad.cpu().numpy()


def _check(out, g, tol):
    links, gf, gqw, gkw, ggw = out
    ref = g["links"]
    assert links.shape == ref.shape
    assert np.array_equal(np.isfinite(links), np.isfinite(ref))
    fin = np.isfinite(ref)
    assert np.abs(links[fin] - ref[fin]).max() <= tol
    for mine, name in ((gf, "grad_features"), (gqw, "grad_qw"), (gkw, "grad_kw"), (ggw, "grad_gw")):
        r = g[name]
        assert np.abs(mine - r).max() <= 10 * tol * max(1.0, np.abs(r).max()), name


@pytest.mark.parametrize("name", CASES)
def test_torch_mirror_equals_the_reference_function(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    _check(_run(g, "cpu", fused=False), g, 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_fused_kernel_equals_the_reference_function(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    _check(_run(g, "cuda", fused=True), g, 2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 1024, 8, 64, 1023, (1024, 700)), (3, 300, 4, 32, 64, (300, 17, 129)), (1, 70, 1, 128, 9, (70,))])
def test_fused_forward_against_the_op_sequence_on_the_gpu(shape):
    """Larger lattices (C2-sized rows) against the reference's op sequence run by torch on the same device."""
    B, L, H, Fd, T, lens = shape
    gen = torch.Generator(device="cuda").manual_seed(L + H)
    q = torch.randn(B, L, H, Fd, device="cuda", generator=gen)
    k = torch.randn(B, L, H, Fd, device="cuda", generator=gen) * 1.5
    lg = torch.log_softmax(torch.randn(B, L, H, device="cuda", generator=gen), dim=-1)
    ol = torch.tensor(lens, device="cuda")
    mine = dl.extract_links_from_chunks(q, k, lg, ol, T, fused=True)
    ref = torch.cat([dl._torch_rows(q.double(), k.double(), lg.double(), ol, T, i0, min(L, i0 + 128)) for i0 in range(0, L, 128)], dim=1)
    # transitions more than 87 nats below their row's total may flush to -inf (kernel header); everything else matches
    big = ref > -80
    assert torch.isfinite(mine[big]).all()
    assert not (torch.isfinite(mine) & ~torch.isfinite(ref)).any()
    assert (mine[big].double() - ref[big]).abs().max() <= 2e-4
