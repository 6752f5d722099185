import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, 1023, 4096, 1234)
def run(n):
    for _ in range(n):
        a, b = k.dag_loss(match, links, olen, tlen, True, 1)
        gm, gl = k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)
def timeit(n=20):
    run(5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); run(n); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n
print("with status tracking: gpu ms/step %.3f host enqueue ms/step %.3f" % timeit())
orig = k._track_status
k._track_status = lambda *a, **kw: None
print("without:              gpu ms/step %.3f host enqueue ms/step %.3f" % timeit())
k._track_status = orig
print("with again:           gpu ms/step %.3f host enqueue ms/step %.3f" % timeit())
