"""Timeline of the overlapped step loop (CUDA events on both streams).  torchrun --nproc-per-node N tools/xchg_timeline.py"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, importlib
from daspeech_b200.dist import PeerGradExchange, FlatGradAllReduce
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
hp = len(sys.argv) > 3 and sys.argv[3] == "1"
both = len(sys.argv) > 4 and sys.argv[4] == "1"
if hp:
    dist.init_process_group("nccl", device_id=dev, pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
else:
    dist.init_process_group("nccl", device_id=dev)
k = ops.get_dag_kernel()
B, L, M, T, V = 64, 1024, 256, 1023, 4096
match, links, olen, tlen, go = bench.make_inputs(torch, dev, B, L, M, T, V, 1234 + rank)
kind = sys.argv[1] if len(sys.argv) > 1 else "peer"
if both: other = FlatGradAllReduce(75_000_000, torch.float32, dev)
e = PeerGradExchange(75_000_000, dev) if kind == "peer" else FlatGradAllReduce(75_000_000, torch.float32, dev)
E = lambda: torch.cuda.Event(enable_timing=True)
def loop(n, log=None):
    for i in range(n):
        t = [E() for _ in range(5)]
        t[0].record()
        alpha, beta = k.dag_loss(match, links, olen, tlen, True, 1)
        t[1].record()
        k.dag_loss_backward(go, alpha, beta, match, links, olen, tlen, 2, 2)
        t[2].record()
        e.finish()
        e.side.wait_stream(torch.cuda.current_stream())
        t[3].record(e.side)
        e.start()
        t[4].record(e.side)
        if log is not None: log.append(t)
    e.finish()
loop(5)
dist.barrier(); torch.cuda.synchronize()
log = []
base = E(); base.record()
loop(nsteps, log)
torch.cuda.synchronize()
if rank == 0:
    print(kind, nsteps, hp, both, "period %.3f" % ((base.elapsed_time(log[-1][0]) - base.elapsed_time(log[1][0])) / (len(log) - 2)))
    for i, t in enumerate(log[-3:]):
        print(kind, "step %d: fwd %.3f-%.3f bwd -%.3f | exchange %.3f-%.3f" % ((i,) + tuple(base.elapsed_time(x) for x in t)), flush=True)
dist.barrier()
dist.destroy_process_group()
