"""Time logsoftmax_gather forward / backward at C2 (B=64, L=1024, V=4096, S=256), fp16 and fp32."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel()
dev = torch.device("cuda", 0)
B, L, V, M = 64, 1024, 4096, 256
peak = 6545.6
def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, dt, esz in (("fp16", torch.float16, 2), ("fp32", torch.float32, 4)):
    torch.manual_seed(0)
    logits = (torch.randn(B, L, V, device=dev) * 2).to(dt)
    idx = torch.randint(4, V, (B, M), device=dev).unsqueeze(1).expand(-1, L, -1)
    gsel = torch.randn(B, M, L, device=dev).transpose(1, 2)
    ref = torch.log_softmax(logits.float(), -1).gather(-1, idx)
    got = k.logsoftmax_gather(logits.clone(), idx, True)
    err = float((got - ref).abs().max())
    ms_f = timeit(lambda: k.logsoftmax_gather(logits, idx, True))
    by_f = 2 * esz * B * L * V + 4 * B * L * M + 8 * B * M
    ms_b = timeit(lambda: k.logsoftmax_gather_backward(logits, idx, gsel))
    by_b = 2 * esz * B * L * V + 4 * B * L * M
    print("%s fwd %.4f ms (%.1f%% of %.0f GB/s)  bwd %.4f ms (%.1f%%)  max|err| %.2e" %
          (name, ms_f, 100 * by_f / ms_f / 1e6 / peak, peak, ms_b, 100 * by_b / ms_b / 1e6 / peak, err))
