import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, 1023, 4096, 1234)
for _ in range(3): a, b = k.dag_loss(match, links, olen, tlen, True, 1)
torch.cuda.synchronize()
