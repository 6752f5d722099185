import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle, build_ref
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
from daspeech_b200.csrc import build_shim
k = ops.get_dag_kernel()
B, L, M, T = 3, 200, 30, 199
match, links, olen, tlen = oracle.make_lattice(B, L, M, T, seed=31, ragged=True)
cu = lambda a: torch.as_tensor(a).cuda()
m, lk, ol, tl = cu(match), cu(links), cu(olen), cu(tlen)
if len(sys.argv) > 1:
    ref = build_ref.load_ref()
    print("ref module", ref.__file__)
    r0 = ref.dag_loss(m, lk, ol, tl, True, 1)
mod = build_shim.load()
print("shim module", mod.__file__)
for rep in range(2):
    a0, b0 = k.dag_loss(m, lk, ol, tl, True, 1)
    a1, b1 = mod.dag_loss(m, lk, ol, tl, True, 1)
    torch.cuda.synchronize()
    for n, x, y in (("alpha", a0, a1), ("beta", b0, b1)):
        d = (x.view(torch.int32) != y.view(torch.int32))
        print(rep, n, "mismatch", int(d.sum()), "nan", int(torch.isnan(x).sum()), int(torch.isnan(y).sum()))
        if d.any():
            for i in d.nonzero()[:5]:
                print("  ", i.tolist(), float(x[tuple(i)]), float(y[tuple(i)]))
    if len(sys.argv) > 1:
        print("shim == ref?", torch.equal(a1, r0[0]), torch.equal(b1, r0[1]))
