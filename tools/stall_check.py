import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, 1023, 4096, 1234)
def step():
    a, b = k.dag_loss(match, links, olen, tlen, True, 1)
    gm, gl = k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)
    return a, b, gm, gl
for rep in range(6):
    for _ in range(5): out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    e0.record()
    for s in range(20):
        t0 = time.perf_counter()
        a, b = k.dag_loss(match, links, olen, tlen, True, 1)
        t1 = time.perf_counter()
        gm, gl = k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)
        ts.append((t1 - t0, time.perf_counter() - t1))
    e1.record(); torch.cuda.synchronize()
    worst = max(range(20), key=lambda i: sum(ts[i]))
    print("rep", rep, "ms/step %.3f" % (e0.elapsed_time(e1) / 20), "worst host iter", worst, "fwd %.2f ms bwd %.2f ms" % (ts[worst][0] * 1e3, ts[worst][1] * 1e3),
          "mallocs", torch.cuda.memory_stats().get("num_device_alloc"))
