"""Run a few dag_loss forward/backward steps at C2 (for ncu captures)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 3
match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, 1023, 4096, 1234)
for _ in range(n):
    a, b = k.dag_loss(match, links, olen, tlen, True, 1)
    gm, gl = k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)
    if "--viterbi" in sys.argv:
        k.dag_best_alignment(match, links, olen, tlen, 1, want_alpha=False)
torch.cuda.synchronize()
print("ok", float(b[:, 0, 0].mean()))
