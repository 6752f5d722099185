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
if "--lsg" in sys.argv:
    B, L, V, M = 64, 1024, 4096, 256
    for dt in (torch.float16, torch.float32):
        logits = (torch.randn(B, L, V, device=dev) * 2).to(dt)
        idx = torch.randint(4, V, (B, M), device=dev).unsqueeze(1).expand(-1, L, -1)
        gsel = torch.randn(B, M, L, device=dev).transpose(1, 2)
        for _ in range(n):
            k.logsoftmax_gather(logits, idx, True)
            k.logsoftmax_gather_backward(logits, idx, gsel)
torch.cuda.synchronize()
print("ok", float(b[:, 0, 0].mean()))
