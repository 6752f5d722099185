"""Forward (alpha+beta) kernel timings through the library's own profile marks + a parity spot check.  argv: B L M T"""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
B, L, M, T = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (64, 1024, 256, 1023))]
match, links, olen, tlen, go = bench.make_inputs(torch, dev, B, L, M, T, 4096, 1234)
lib = k.lib
for _ in range(5): a, b = k.dag_loss(match, links, olen, tlen, True, 1)
torch.cuda.synchronize()
lib.dagb200_set_profile(1)
buf = (ctypes.c_float * 5)()
acc = [0.0] * 5
n = 10
for _ in range(n):
    a, b = k.dag_loss(match, links, olen, tlen, True, 1)
    lib.dagb200_get_profile(ctypes.cast(buf, ctypes.c_void_p), 5)
    for i in range(5): acc[i] += max(buf[i], 0) / n
lib.dagb200_set_profile(0)
print("prep %.4f ms  alpha_beta %.4f ms" % (acc[0], acc[1]))
# parity spot check against the exact log-domain kernels
lib.dagb200_set_exact(1)
a2, b2 = k.dag_loss(match, links, olen, tlen, True, 1)
lib.dagb200_set_exact(0)
fin = torch.isfinite(a2)
print("alpha: finite sets equal", bool((torch.isfinite(a) == fin).all()), "max |diff| %.3e" % float((a - a2)[fin].abs().max()),
      "| beta: finite sets equal", bool((torch.isfinite(b) == torch.isfinite(b2)).all()),
      "max |diff| %.3e" % float((b - b2)[torch.isfinite(b2)].abs().max()))
