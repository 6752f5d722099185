import ctypes, importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel(); lib = k.lib
dev = torch.device("cuda", 0)
match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, 1023, 4096, 1)
def prof():
    buf = (ctypes.c_float * 5)(); lib.dagb200_get_profile(ctypes.cast(buf, ctypes.c_void_p), 5); return [round(x, 3) for x in buf]
for _ in range(3):
    a, b = k.dag_loss(match, links, olen, tlen, True, 1); gm, gl = k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)
torch.cuda.synchronize()
lib.dagb200_set_profile(1)
for mode in ("fwd-only", "alternate", "alternate-keep-outputs"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    keep = []
    e0.record()
    for s in range(20):
        if mode == "fwd-only":
            a, b = k.dag_loss(match, links, olen, tlen, True, 1)
        else:
            a, b = k.dag_loss(match, links, olen, tlen, True, 1)
            gm, gl = k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)
            if mode.endswith("keep-outputs") and s < 4: keep.append((a, b, gm, gl))
    e1.record(); torch.cuda.synchronize()
    print(mode, "ms/step %.3f" % (e0.elapsed_time(e1) / 20), "last-step kernels [prep, dp, gm, gl, vit] =", prof(),
          "reserved MB", torch.cuda.memory_reserved() >> 20, "num_alloc_retries", torch.cuda.memory_stats()["num_alloc_retries"],
          "segments", torch.cuda.memory_stats()["segment.all.allocated"])
