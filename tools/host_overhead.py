import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, 1023, 4096, 1)
for _ in range(3):
    a, b = k.dag_loss(match, links, olen, tlen, True, 1); gm, gl = k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    a, b = k.dag_loss(match, links, olen, tlen, True, 1)
t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print("fwd: host enqueue %.3f ms/call, total %.3f ms/call" % ((t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3))
t0 = time.perf_counter()
for _ in range(N):
    gm, gl = k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)
t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print("bwd: host enqueue %.3f ms/call, total %.3f ms/call" % ((t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(N):
    a, b = k.dag_loss(match, links, olen, tlen, True, 1)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
