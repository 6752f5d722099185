"""No-gradient logsoftmax_gather (the GLAT pass) at C2, fp16: timing / profiling target."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
B, L, V, M = 64, 1024, 4096, 256
logits = (torch.randn(B, L, V, device=dev) * 2).half()
idx = torch.randint(4, V, (B, M), device=dev).unsqueeze(1).expand(-1, L, -1)
for _ in range(3):
    k.logsoftmax_gather(logits, idx, False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    k.logsoftmax_gather(logits, idx, False)
b.record(); torch.cuda.synchronize()
print("no-grad fp16 gather ms", a.elapsed_time(b) / 10)
