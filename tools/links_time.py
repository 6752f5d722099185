"""extract_links at the C2 shape: fused kernel vs the reference's op sequence run by torch on the same device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daspeech_b200 import links as dl
B, L, H, Fd = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (64, 1024, 8, 64)))
T = L - 1
gen = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(B, L, H, Fd, device="cuda", generator=gen); k = torch.randn(B, L, H, Fd, device="cuda", generator=gen)
lg = torch.log_softmax(torch.randn(B, L, H, device="cuda", generator=gen), -1)
ol = torch.full((B,), L, device="cuda")
def timed(fn, n):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
with torch.no_grad():
    ms = timed(lambda: dl.extract_links_from_chunks(q, k, lg, ol, T, fused=True), 10)
    print("fused extract_links B=%d L=%d H=%d F=%d: %.3f ms" % (B, L, H, Fd, ms))
    Bs = min(B, 8)   # the op sequence needs ~2.1 GB per 8 utterances for the [B,L,L,H] product alone
    torch.cuda.reset_peak_memory_stats()
    ms_ref = timed(lambda: dl.torch_extract_links(q[:Bs], k[:Bs], lg[:Bs], ol[:Bs], T), 3)
    print("reference op sequence on %d utterances: %.3f ms (x%d = %.1f ms), peak memory %.1f GB" % (
        Bs, ms_ref, B // Bs, ms_ref * B / Bs, torch.cuda.max_memory_allocated() / 1e9))
