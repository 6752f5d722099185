import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1023
match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, T, 4096, 1234)
for _ in range(3):
    k.dag_best_alignment(match, links, olen, tlen, 1, want_alpha=False)
torch.cuda.synchronize()
