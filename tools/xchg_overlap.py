"""Which kernels of the step overlap the gradient exchange?  torchrun --nproc-per-node N tools/xchg_overlap.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from daspeech_b200.dist import PeerGradExchange, FlatGradAllReduce
import importlib
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
k = ops.get_dag_kernel()
B, L, M, T, V = 64, 1024, 256, 1023, 4096
match, links, olen, tlen, go = bench.make_inputs(torch, dev, B, L, M, T, V, 1234 + rank)
ex = {"peer": PeerGradExchange(75_000_000, dev), "nccl": FlatGradAllReduce(75_000_000, torch.float32, dev)}
alpha, beta = k.dag_loss(match, links, olen, tlen, True, 1)
a = torch.randn(4096, 4096, device=dev, dtype=torch.bfloat16)
x = torch.randn(64 << 20, device=dev)

def fwd(): k.dag_loss(match, links, olen, tlen, True, 1)
def bwd(): k.dag_loss_backward(go, alpha, beta, match, links, olen, tlen, 2, 2)
def mm():
    for _ in range(8): torch.mm(a, a)
def ew():
    for _ in range(4): x.mul_(1.0001)
def nothing(): pass

def timed(fn, e, n=10):
    def loop(n):
        for _ in range(n):
            fn()
            if e is not None:
                e.finish(); e.start()
        if e is not None: e.finish()
    loop(3)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(); loop(n); a1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return a0.elapsed_time(a1) / n, host

for name, fn in (("nothing", nothing), ("fwd", fwd), ("bwd", bwd), ("mm", mm), ("ew", ew)):
    row = []
    for ename, e in (("none", None), ("peer", ex["peer"]), ("nccl", ex["nccl"])):
        ms, host = timed(fn, e)
        row.append("%s %.3f (host %.3f)" % (ename, ms, host))
    if rank == 0: print("%-8s" % name, " | ".join(row), flush=True)
dist.barrier()
