import ctypes, importlib, os, sys, subprocess, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel(); lib = k.lib
dev = torch.device("cuda", 0)
match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, 1023, 4096, 1)
for _ in range(3):
    a, b = k.dag_loss(match, links, olen, tlen, True, 1)
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "50"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(400):
    a, b = k.dag_loss(match, links, olen, tlen, True, 1)
e1.record(); torch.cuda.synchronize()
print("fwd ms/call %.3f" % (e0.elapsed_time(e1) / 400))
time.sleep(0.1); p.terminate()
out = p.stdout.read().strip().splitlines()
print("nvidia-smi samples:", out[:3], "...", out[-4:])
