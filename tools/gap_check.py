import ctypes, importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel(); lib = k.lib
dev = torch.device("cuda", 0)
match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, 1023, 4096, 1234)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (t1 - t0) / n * 1e3
a, b = k.dag_loss(match, links, olen, tlen, True, 1)
print("fwd only  (ms gpu, ms host enqueue):", timed(lambda: k.dag_loss(match, links, olen, tlen, True, 1)))
print("bwd only  :", timed(lambda: k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)))
def both():
    a2, b2 = k.dag_loss(match, links, olen, tlen, True, 1)
    k.dag_loss_backward(go, a2, b2, match, links, olen, tlen, 2, 2)
print("fwd+bwd   :", timed(both))
st = torch.cuda.memory_stats()
print("cudaMalloc calls", st.get("num_device_alloc"), "retries", st.get("num_alloc_retries"), "reserved MB", st.get("reserved_bytes.all.current") / 2**20)
lib.dagb200_set_profile(1)
buf = (ctypes.c_float * 5)()
both(); lib.dagb200_get_profile(ctypes.cast(buf, ctypes.c_void_p), 5); print("profile", [round(x, 3) for x in buf])
