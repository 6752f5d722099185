"""CPU: host-side operator layer -- exported names, the device-agnostic torch_* functions against the golden
vectors of the reference, and 'no GPU => loud failure' for the CUDA operators."""
import glob
import os

import numpy as np
import pytest
import torch

import daspeech_b200
from daspeech_b200 import custom_ops
import importlib
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")  # the name `dag_loss` is shadowed by the function
from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DP_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith(("gather", "criterion")))


def test_exported_names_match_reference_surface():
    names = ["dag_loss", "dag_loss_with_alpha_beta", "dag_best_alignment", "dag_logsoftmax_gather_inplace",
             "torch_dag_loss", "torch_dag_best_alignment", "torch_dag_logsoftmax_gather_inplace", "logsumexp_keepdim"]
    for n in names:
        assert callable(getattr(custom_ops, n)) and callable(getattr(daspeech_b200, n))
    # tuner-visible class attributes (reference dag_loss.py:67-69,191)
    assert (ops.DagLossFunc.config, ops.DagLossFunc.config1, ops.DagLossFunc.config2) == (1, 2, 2)
    assert ops.DagBestAlignmentFunc.config == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_cuda_ops_refuse_to_run_without_gpu():
    m = torch.zeros(1, 2, 4)
    lk = torch.zeros(1, 4, 3)
    ln = torch.tensor([4])
    with pytest.raises(RuntimeError, match="You need GPU"):
        ops.dag_loss(m, lk, ln, torch.tensor([2]))
    with pytest.raises(RuntimeError, match="You need GPU"):
        ops.dag_best_alignment(m, lk, ln, torch.tensor([2]))
    with pytest.raises(RuntimeError, match="You need GPU"):
        ops.dag_logsoftmax_gather_inplace(torch.zeros(1, 4, 8), torch.zeros(1, 4, 2, dtype=torch.long))


@pytest.mark.parametrize("name", DP_CASES)
def test_torch_versions_against_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    m = torch.tensor(g["match"], dtype=torch.float64, requires_grad=True)
    lk = torch.tensor(g["links"], dtype=torch.float64, requires_grad=True)
    # differentiable band -> dense scatter so grad_links can be compared
    B, L, T = g["links"].shape
    idx = (torch.arange(L).unsqueeze(1) + torch.arange(T).unsqueeze(0) + 1)
    dense = torch.full((B, L, L + 1), float("-inf"), dtype=torch.float64).scatter(
        2, idx.clamp(max=L).unsqueeze(0).expand(B, -1, -1), lk)[:, :, :L]
    ol, tl = torch.tensor(g["olen"]), torch.tensor(g["tlen"])
    loss = ops.torch_dag_loss(m, dense, ol, tl)
    ref = torch.tensor(g["loss"], dtype=torch.float64)
    fin = torch.isfinite(ref)
    tol = 1e-5 if name.endswith("fp32") else 1e-10
    assert torch.equal(torch.isfinite(loss), fin)
    assert torch.allclose(loss[fin], ref[fin], rtol=tol, atol=tol)
    if "grad_match" in g:
        gm, gl = torch.autograd.grad((loss * torch.tensor(g["grad_output"], dtype=torch.float64)).sum(), [m, lk])
        gt = 2e-5 if name.endswith("fp32") else 1e-9
        assert np.allclose(gm.numpy(), g["grad_match"], rtol=gt, atol=gt)
        assert np.allclose(gl.numpy(), g["grad_links"], rtol=gt, atol=gt)
        path = ops.torch_dag_best_alignment(m.detach(), dense.detach(), ol, tl)
        assert np.array_equal(path.numpy(), g["viterbi_path"])


def test_logsumexp_keepdim_handles_empty_slices():
    x = torch.tensor([[0.0, -1.0], [float("-inf"), float("-inf")]], requires_grad=True)
    y = ops.logsumexp_keepdim(x, 1)
    assert y.shape == (2, 1) and y[1, 0] == float("-inf")
    assert torch.allclose(y[0, 0], torch.logsumexp(x[0].detach(), 0))
    y[0].sum().backward()
    assert torch.isfinite(x.grad).all()


def test_torch_gather_matches_golden():
    g = dict(np.load(os.path.join(GOLDEN, "gather_c1_fp32.npz")))
    x = torch.tensor(g["logits"])
    B, L, V = x.shape
    idx = torch.tensor(g["targets"]).unsqueeze(1).expand(-1, L, -1)
    same, sel = ops.torch_dag_logsoftmax_gather_inplace(x, idx)
    assert same is x and np.allclose(sel.numpy(), g["selected"], rtol=1e-5, atol=1e-6)


def test_prefetcher_has_no_cpu_path_and_numa_binding_is_best_effort():
    """daspeech_b200.prefetch: the prefetcher is device plumbing only (no CPU fallback); the NUMA helper never raises."""
    import pytest
    from daspeech_b200 import prefetch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="needs a CUDA device"):
            prefetch.DevicePrefetcher(iter([(torch.zeros(2),)]), torch.device("cpu"))
    assert prefetch.bind_host_to_gpu(0) in (True, False)


def test_gradient_exchange_order_by_world_size():
    from daspeech_b200.dist import exchange_order
    assert exchange_order("auto", 2) == ("peer", "nvls", "nccl")
    assert exchange_order("auto", 4)[0] == "nvls" and exchange_order("auto", 8)[-1] == "nccl"
    assert exchange_order("nccl", 8) == ("nccl",)


def test_link_rows_helper_equals_the_full_op_sequence_with_gradients():
    """daspeech_b200.links._torch_rows (the chunked recomputation behind the fused op's backward) against the full
    mirror of the reference op sequence: same values, same -inf pattern, same gradients, no NaN from dead rows."""
    import torch
    from daspeech_b200 import links as dl
    torch.manual_seed(0)
    B, L, H, Fd, T = 2, 90, 3, 16, 40
    q = torch.randn(B, L, H, Fd, requires_grad=True)
    k = torch.randn(B, L, H, Fd, requires_grad=True)
    g = torch.log_softmax(torch.randn(B, L, H), -1).requires_grad_()
    ol = torch.tensor([90, 37])
    full = dl.torch_extract_links(q, k, g, ol, T)
    rows = torch.cat([dl._torch_rows(q, k, g, ol, T, i0, min(L, i0 + 32)) for i0 in range(0, L, 32)], 1)
    fin = torch.isfinite(full)
    assert torch.equal(fin, torch.isfinite(rows))
    assert float((full - rows)[fin].abs().max()) <= 1e-6
    w = torch.randn_like(full)
    g1 = torch.autograd.grad((full.masked_fill(~fin, 0) * w).sum(), [q, k, g])
    g2 = torch.autograd.grad((rows.masked_fill(~fin, 0) * w).sum(), [q, k, g])
    for a, b in zip(g1, g2):
        assert torch.isfinite(b).all()
        assert float((a - b).abs().max()) <= 1e-5
