"""Time dag_best_alignment at C2 (full lengths and ragged), with and without the lattice output."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for T in (1023, 32):
    match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, T, 4096, 1234)
    print("T", T, "no lattice %.4f ms" % timeit(lambda: k.dag_best_alignment(match, links, olen, tlen, 1, want_alpha=False)),
          "with lattice %.4f ms" % timeit(lambda: k.dag_best_alignment(match, links, olen, tlen, 1, want_alpha=True)))
