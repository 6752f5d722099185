"""Accuracy of the dp2 fallback (DAGB200_DP=2) on a multi-pass lattice (M > 256) against the exact log-domain kernels."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
dev = "cuda:0"
for (B, L, M, T) in ((1, 640, 530, 639), (2, 352, 300, 351), (2, 300, 40, 299)):
    match, links, olen, tlen = oracle.make_lattice(B, L, M, T, seed=3, ragged=False)
    t = [torch.tensor(x, device=dev) for x in (match, links, olen, tlen)]
    def run(exact):
        ops.EXACT_LOG_DOMAIN = exact
        m = t[0].clone().requires_grad_(); lk = t[1].clone().requires_grad_()
        loss, (a, b) = ops.dag_loss_with_alpha_beta(m, lk, t[2], t[3])
        loss.sum().backward()
        ops.EXACT_LOG_DOMAIN = False
        return loss.detach(), a, b, m.grad, lk.grad
    f = run(False); e = run(True)
    fin = torch.isfinite(e[1]) & torch.isfinite(f[1])
    print((B, L, M, T), "loss rel", float(((f[0] - e[0]) / e[0]).abs().max()),
          "alpha max abs", float((f[1][fin] - e[1][fin]).abs().max()), "finite-set mismatch", int((torch.isfinite(e[1]) != torch.isfinite(f[1])).sum()),
          "gm rel", float((f[3] - e[3]).abs().max() / e[3].abs().max()), "gl rel", float((f[4] - e[4]).abs().max() / e[4].abs().max()))
