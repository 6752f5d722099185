"""Generate tests/golden/links/links_*.npz by running the UNMODIFIED reference functions `logsumexp`, `extract_valid_links` and
`extract_links` of /root/reference/DASpeech/models/s2t_conformer_dag.py on CPU.

The model file imports fairseq (absent offline), so the three function definitions are cut out of the file's syntax
tree and executed as they stand -- the two methods as members of a bare class that carries the `args` / `pad`
attributes they read.  Nothing of the reference is copied into the repository; run in the build container only:

    python tests/golden/make_golden_links.py
"""
import ast
import os
import types

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/DASpeech/models/s2t_conformer_dag.py"


def load_reference():
    tree = ast.parse(open(SRC).read())
    ns = {"torch": torch, "F": F, "Tensor": Tensor}
    fns = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name in ("logsumexp", "extract_valid_links", "extract_links")]
    assert sorted(f.name for f in fns) == ["extract_links", "extract_valid_links", "logsumexp"]
    mod = ast.Module(body=[f for f in fns if f.name == "logsumexp"], type_ignores=[])
    exec(compile(mod, SRC, "exec"), ns)
    cls = ast.ClassDef(name="RefDecoder", bases=[], keywords=[], decorator_list=[],
                       body=[f for f in fns if f.name != "logsumexp"])
    mod = ast.Module(body=[cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, SRC, "exec"), ns)
    return ns["RefDecoder"]


def case(name, B, L, H, Fd, T, lengths, seed):
    torch.manual_seed(seed)
    D = H * Fd
    Ref = load_reference()
    dec = Ref()
    dec.pad = 1
    dec.args = types.SimpleNamespace(max_transition_length=T, decoder_attention_heads=H, decoder_embed_dim=D,
                                     links_feature="feature:position")
    features = torch.randn(B, L, D)
    tokens = torch.full((B, L), 5, dtype=torch.long)
    for b, n in enumerate(lengths):
        tokens[b, n:] = 1
    pos = torch.nn.Embedding(L + 2, D)
    link_positional = lambda t: pos(torch.arange(L).unsqueeze(0).expand(t.shape[0], -1))   # noqa: E731
    ql, kl, gl = torch.nn.Linear(2 * D, D), torch.nn.Linear(2 * D, D), torch.nn.Linear(2 * D, H)
    with torch.no_grad():
        ql.weight.mul_(2.0)
        kl.weight.mul_(2.0)
    features.requires_grad_()
    links = dec.extract_links(features, tokens, link_positional, ql, kl, gl)
    w = torch.randn_like(links)
    fin = torch.isfinite(links)
    (links.masked_fill(~fin, 0.0) * w).sum().backward()
    x = torch.cat([features, link_positional(tokens)], dim=-1)
    np.savez_compressed(os.path.join(HERE, "links", name + ".npz"),
                        features=features.detach().numpy(), tokens=tokens.numpy(), pos=link_positional(tokens).detach().numpy(),
                        qw=ql.weight.detach().numpy(), qb=ql.bias.detach().numpy(), kw=kl.weight.detach().numpy(),
                        kb=kl.bias.detach().numpy(), gw=gl.weight.detach().numpy(), gb=gl.bias.detach().numpy(),
                        links=links.detach().numpy(), w=w.numpy(), grad_features=features.grad.numpy(),
                        grad_qw=ql.weight.grad.numpy(), grad_kw=kl.weight.grad.numpy(), grad_gw=gl.weight.grad.numpy(),
                        H=H, T=T, pad=1)
    print(name, tuple(links.shape), "finite", int(fin.sum()))


if __name__ == "__main__":
    case("links_small", 3, 40, 2, 16, 99999, [40, 33, 2], 1)
    case("links_band", 2, 150, 4, 32, 37, [150, 97], 2)
    case("links_wide", 2, 136, 8, 16, 99999, [136, 71], 3)
