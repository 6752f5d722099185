"""Does torch symmetric memory (VMM + multicast over NVSwitch) work on this box?  torchrun --nproc-per-node N tools/symm_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
numel = 75_000_000
try:
    t = symm_mem.empty(numel, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    if rank == 0:
        print("rendezvous ok; multicast_ptr", hex(hdl.multicast_ptr), "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs],
              "signal_pad", [hex(p) for p in hdl.signal_pad_ptrs], "pad size", hdl.signal_pad_size, flush=True)
except Exception as e:
    print("rank", rank, "rendezvous failed:", repr(e)[:600], flush=True)
    sys.exit(0)
gname = dist.group.WORLD.group_name
def timed(fn, n=10):
    fn(); fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
    try:
        op = getattr(torch.ops.symm_mem, name)
        t.fill_(rank + 1.0)
        op(t, "sum", gname)
        torch.cuda.synchronize()
        ms = timed(lambda: op(t, "sum", gname))
        if rank == 0: print(name, "%.3f ms" % ms, "busbw %.0f GB/s" % (numel * 4 * 2 * (world - 1) / world / ms / 1e6), flush=True)
    except Exception as e:
        if rank == 0: print(name, "failed:", repr(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
