"""Host -> device bandwidth from pinned memory: one copy vs the same bytes split over several streams."""
import torch, time
dev = torch.device("cuda", 0)
n = 268 * 1024 * 1024 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
for ns in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    ch = n // ns
    def fn():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * ch:(i + 1) * ch].copy_(h[i * ch:(i + 1) * ch], non_blocking=True)
    t = timeit(fn)
    print("streams", ns, "%.2f ms  %.1f GB/s" % (t * 1e3, n * 4 / t / 1e9))
