"""GPU, informational: time the UNMODIFIED reference CUDA kernels (oracle/_ref/dag_loss_fn.so) next to ours on
the same C2 tensors and leave the numbers in gpurun_out/reference_timing.json (the "GPU reference bar" of
BASELINE.md section 5).  Asserts only that both produce the same loss."""
import importlib
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, n=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


@pytest.mark.parametrize("T", [1023, 32])
def test_time_reference_kernels_at_c2(T):
    from oracle import build_ref
    ref = build_ref.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref/dag_loss_fn.so not present")
    ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
    k = ops.get_dag_kernel()
    B, L, M, V = 64, 1024, 256, 4096
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    match = torch.log(torch.rand(B, M, L, device=dev, generator=g) * 0.98 + 0.01)
    raw = torch.randn(B, L, T, device=dev, generator=g)
    i = torch.arange(L, device=dev).view(1, L, 1)
    kk = torch.arange(T, device=dev).view(1, 1, T)
    valid = (i + kk + 1) < L
    links = torch.log_softmax(raw.masked_fill(~valid, float("-inf")), -1).masked_fill(~valid, float("-inf")).contiguous()
    del raw
    ol = torch.full((B,), L, dtype=torch.long, device=dev)
    tl = torch.full((B,), M, dtype=torch.long, device=dev)
    go = torch.ones(B, device=dev)
    res = {"shape": dict(B=B, L=L, M=M, T=T, V=V)}
    a0, b0 = ref.dag_loss(match, links, ol, tl, True, 1)
    a1, b1 = k.dag_loss(match, links, ol, tl, True, 1)
    torch.cuda.synchronize()
    assert torch.allclose(b0[:, 0, 0], b1[:, 0, 0], rtol=1e-4)
    res["ref_fwd_ms"] = _time(lambda: ref.dag_loss(match, links, ol, tl, True, 1))
    res["new_fwd_ms"] = _time(lambda: k.dag_loss(match, links, ol, tl, True, 1))
    res["ref_bwd_ms"] = _time(lambda: ref.dag_loss_backward(go, a0, b0, match, links, ol, tl, 2, 2))
    res["new_bwd_ms"] = _time(lambda: k.dag_loss_backward(go, a1, b1, match, links, ol, tl, 2, 2))
    if (M - 1) * T + 1 >= L:
        res["ref_viterbi_ms"] = _time(lambda: ref.dag_best_alignment(match, links, ol, tl, 1))
        res["new_viterbi_ms"] = _time(lambda: k.dag_best_alignment(match, links, ol, tl, 1, want_alpha=False))
    for name, dt in (("fp16", torch.float16), ("fp32", torch.float32)):
        x = (torch.randn(B, L, V, device=dev) * 2).to(dt)
        idx = torch.randint(4, V, (B, M), device=dev).unsqueeze(1).expand(-1, L, -1)
        res["ref_gather_%s_ms" % name] = _time(lambda: ref.logsoftmax_gather(x, idx, True))
        res["new_gather_%s_ms" % name] = _time(lambda: k.logsoftmax_gather(x, idx, True))
        del x
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "reference_timing.json")
    allres = json.load(open(path)) if os.path.exists(path) else {}
    allres["T%d" % T] = res
    json.dump(allres, open(path, "w"), indent=1)
    print(json.dumps(res))
