import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
shape = (5, 257, 64, 64, True, 0.25)
B, L, M, T, ragged, glat = shape
match, links, olen, tlen = oracle.make_lattice(B, L, M, T, seed=sum(shape[:4]), ragged=ragged, glat_frac=glat)
k = ops.get_dag_kernel()
cu = lambda a: torch.as_tensor(a).cuda()
a, b = k.dag_loss(cu(match), cu(links), cu(olen), cu(tlen), True, 1)
a = a.cpu().numpy(); b = b.cpu().numpy()
l64, oa, ob = oracle.dag_loss(match, links, olen, tlen, True, np.float64)
for name, mine, orc, other in (("alpha", a, oa, ob), ("beta", b, ob, oa)):
    mism = np.isfinite(mine) != np.isfinite(orc)
    print(name, "mismatch cells", mism.sum(), "mine finite & oracle -inf:", (np.isfinite(mine) & ~np.isfinite(orc)).sum())
    idx = np.argwhere(mism)[:12]
    for (bb, t, j) in idx:
        post = orc[bb, t, j] + other[bb, t, j] - match[bb, t, j] - l64[bb]
        tp = t - 1 if name == "alpha" else t + 1
        rowmax = np.max(orc[bb, tp][np.isfinite(orc[bb, tp])]) if np.isfinite(orc[bb, tp]).any() else None
        print("  b", bb, "t", t, "j", j, "O", olen[bb], "Tn", tlen[bb], "oracle", orc[bb, t, j], "mine", mine[bb, t, j], "logpost", post, "prev-row max", rowmax)
    fin = np.isfinite(mine) & np.isfinite(orc)
    print(name, "max abs diff on common finite", np.abs(mine[fin] - orc[fin]).max())
