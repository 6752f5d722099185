"""Back-to-back dag_loss launches at C2 (no sync in between), then a check of the loss: catches rare races."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, 1023, 4096, 1234)
a, b = k.dag_loss(match, links, olen, tlen, True, 1)
torch.cuda.synchronize()
ref = b[:, 0, 0].clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    a, b = k.dag_loss(match, links, olen, tlen, True, 1)
e1.record()
torch.cuda.synchronize()
print("ok fwd ms %.4f" % (e0.elapsed_time(e1) / n), "max |dZ|", float((b[:, 0, 0] - ref).abs().max()), float(ref.mean()))
