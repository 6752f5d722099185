"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the operator layer, i.e. the
C ABI of libdagb200.so.  Checkers: the CPU oracle (oracle/), the committed golden vectors of the reference's
torch path (tests/golden/), and -- when oracle/_ref/ holds the prebuilt UNMODIFIED reference CUDA extension --
the reference kernels themselves on the same tensors.

Tolerances (BASELINE.json north_star): arg-max alignment indices bit-exact; fp32 loss within 1e-4 relative;
fp32 gradients within 1e-4 of the tensor's scale (|g - g_ref|_inf <= 1e-4 * |g_ref|_inf).  At the full C2 shape
fp32 log-domain arithmetic itself sits ~1-2e-4 from the fp64 truth (DESIGN.md "Numerics"), so there the bound is
max(1e-4, 1.5 x the error of the fp32 restatement of the reference arithmetic) against the fp64 oracle.
"""
import glob
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu

ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DP_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith("gather"))
GATHER_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "gather*.npz")))
DEV = "cuda"


def cu(a, dtype=None):
    t = torch.as_tensor(a)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def relerr(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    s = np.abs(ref).max()
    return float(np.abs(x - ref).max() / (s if s > 0 else 1.0))


def check_lattice_side_outputs(a, b, oa, ob, match, z):
    """alpha/beta returned by dag_loss_with_alpha_beta against the fp64 oracle.  The blocked fp32 path may flush a
    cell to -inf when ALL of its predecessors are > 87 nats below the best predecessor of its 32-vertex block
    (DESIGN.md numerics contract) -- never a cell that carries posterior mass, and never the other way round."""
    for mine, orc in ((a, oa), (b, ob)):
        assert not (np.isfinite(mine) & ~np.isfinite(orc)).any()
    with np.errstate(invalid="ignore"):
        post = oa + ob - match.astype(np.float64) - np.asarray(z, dtype=np.float64)[:, None, None]
    relevant = np.isfinite(post) & (post > np.log(1e-20))
    for mine, orc in ((a, oa), (b, ob)):
        assert np.isfinite(mine[relevant]).all()
        assert np.allclose(mine[relevant], orc[relevant], rtol=2e-4, atol=2e-3)


def run_loss(match, links, olen, tlen, go, dtype=torch.float32, with_ab=False):
    m = cu(match, dtype).requires_grad_()
    lk = cu(links, dtype).requires_grad_()
    ol, tl = cu(olen), cu(tlen)
    if with_ab:
        loss, (alpha, beta) = ops.dag_loss_with_alpha_beta(m, lk, ol, tl)
    else:
        loss = ops.dag_loss(m, lk, ol, tl)
        alpha = beta = None
    fin = torch.isfinite(loss)
    gm, gl = torch.autograd.grad((torch.where(fin, loss, torch.zeros_like(loss)) * cu(go, dtype)).sum(), [m, lk])
    return loss.detach().cpu().numpy(), gm.cpu().numpy(), gl.cpu().numpy(), alpha, beta


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", DP_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_dag_loss_against_reference_golden(name, dtype):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    loss, gm, gl, alpha, beta = run_loss(g["match"], g["links"], g["olen"], g["tlen"], g["grad_output"], dtype, True)
    ref = g["loss"]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(loss), fin)
    tol = 1e-4 if dtype == torch.float32 else 1e-9
    if name.endswith("fp32"):
        tol = max(tol, 1e-5)
    assert np.allclose(loss[fin], ref[fin], rtol=tol, atol=0)
    if "grad_match" in g:
        assert relerr(gm, g["grad_match"]) <= tol
        assert relerr(gl, g["grad_links"]) <= tol
    else:  # infeasible sample: zero gradients, never NaN (dag_loss.cu:395,463)
        assert np.isfinite(gm).all() and np.isfinite(gl).all()
        assert not gm[~fin].any() and not gl[~fin].any()
    # forward-only path returns alpha[Tn-1, O-1] (dag_loss.py:110)
    with torch.no_grad():
        l2 = ops.dag_loss(cu(g["match"], dtype), cu(g["links"], dtype), cu(g["olen"]), cu(g["tlen"])).cpu().numpy()
    assert np.allclose(l2[fin], ref[fin], rtol=tol, atol=0)
    # alpha/beta side outputs against the oracle (same -inf structure)
    npdt = np.float32 if dtype == torch.float32 else np.float64
    _, oa, ob = oracle.dag_loss(g["match"], g["links"], g["olen"], g["tlen"], True, npdt)
    a, b = alpha.cpu().numpy(), beta.cpu().numpy()
    assert np.array_equal(np.isfinite(a), np.isfinite(oa)) and np.array_equal(np.isfinite(b), np.isfinite(ob))
    assert np.allclose(a[np.isfinite(oa)], oa[np.isfinite(oa)], rtol=1e-5, atol=1e-4)
    assert np.allclose(b[np.isfinite(ob)], ob[np.isfinite(ob)], rtol=1e-5, atol=1e-4)


SHAPES = [
    # B, L, M, T, ragged, glat
    (3, 2, 2, 1, False, 0.0),        # smallest legal lattice
    (4, 33, 7, 32, True, 0.0),       # L not a multiple of the warp
    (2, 300, 40, 299, True, 0.0),    # 512-thread CTA
    (2, 300, 40, 17, True, 0.0),     # banded
    (2, 1100, 24, 1099, True, 0.0),  # L > block size (strided sweep)
    (5, 257, 64, 64, True, 0.25),    # GLAT force-emit 0 / -inf emissions
    (67, 96, 20, 95, True, 0.0),     # many samples
    (2, 352, 300, 351, True, 0.0),   # more than 256 target rows: two passes of the column-major recurrences
    (1, 640, 530, 639, False, 0.0),  # three passes, 17 row chunks (odd split between the Viterbi cluster CTAs)
    (2, 320, 290, 40, True, 0.1),    # two passes, banded transitions, forced emissions
    (1, 2080, 24, 2079, False, 0.0), # 65 vertex blocks: beyond the shared memory of the tcgen05 / mma.sync recurrences (dp2 fallback)
]


@pytest.mark.parametrize("shape", SHAPES)
def test_dag_loss_against_oracle(shape):
    B, L, M, T, ragged, glat = shape
    match, links, olen, tlen = oracle.make_lattice(B, L, M, T, seed=sum(shape[:4]), ragged=ragged, glat_frac=glat)
    go = np.random.default_rng(1).random(B).astype(np.float32) + 0.5
    loss, gm, gl, alpha, beta = run_loss(match, links, olen, tlen, go, torch.float32, True)
    ol64, oa, ob = oracle.dag_loss(match, links, olen, tlen, True, np.float64)
    ogm, ogl = oracle.dag_loss_backward(go, oa, ob, match, links, olen, tlen, np.float64)
    fin = np.isfinite(ol64)
    assert np.array_equal(np.isfinite(loss), fin)
    assert np.allclose(loss[fin], ol64[fin], rtol=1e-4, atol=0)
    tol_m = tol_l = 1e-4
    if M > 128:
        # lattice values ~1e3 carry an fp32 ulp of 6e-5..1.2e-4: the reference's own fp32 arithmetic (restated in the
        # oracle's fp32 build) is then only accurate to ~1e-4, so the bar is relative to it (DESIGN.md "Numerics")
        _, a32, b32 = oracle.dag_loss(match, links, olen, tlen, True, np.float32)
        gm32, gl32 = oracle.dag_loss_backward(go, a32, b32, match, links, olen, tlen, np.float32)
        tol_m = max(1e-4, 1.5 * relerr(np.where(fin[:, None, None], gm32, 0), np.where(fin[:, None, None], ogm, 0)))
        tol_l = max(1e-4, 1.5 * relerr(np.where(fin[:, None, None], gl32, 0), np.where(fin[:, None, None], ogl, 0)))
    em, el = relerr(gm, np.where(fin[:, None, None], ogm, 0)), relerr(gl, np.where(fin[:, None, None], ogl, 0))
    print("grad errors", em, el, "tolerances", tol_m, tol_l)
    assert em <= tol_m
    assert el <= tol_l
    check_lattice_side_outputs(alpha.cpu().numpy(), beta.cpu().numpy(), oa, ob, match, ol64)


def _flat_lattice(B, L, M, T, slack, seed):
    """Flat transition scores and a tight live band (O - Tn = slack): every cell of the band matters and
    alpha/beta vary by >100 nats inside a 128-vertex window -- the stress case for block-local frames."""
    rng = np.random.default_rng(seed)
    tlen = np.full(B, M, dtype=np.int64)
    olen = np.minimum(tlen + slack, L).astype(np.int64)
    match = np.log(rng.random((B, M, L)) * 0.5 + 0.5).astype(np.float32)
    i = np.arange(L)[:, None]; k = np.arange(T)[None, :]
    valid = (i + k + 1)[None] < olen[:, None, None]
    cnt = np.maximum(valid.sum(-1, keepdims=True), 1)
    links = np.where(valid, -np.log(cnt), -np.inf).astype(np.float32)
    return match, links, olen, tlen


@pytest.mark.parametrize("cfg", [(2, 320, 256, 319, 60), (2, 200, 150, 199, 3), (2, 512, 96, 511, 400), (1, 1024, 256, 32, 700)])
def test_dag_loss_flat_tight_band(cfg):
    B, L, M, T, slack = cfg
    match, links, olen, tlen = _flat_lattice(B, L, M, T, slack, seed=L + M)
    go = np.ones(B, np.float32)
    loss, gm, gl, alpha, beta = run_loss(match, links, olen, tlen, go, torch.float32, True)
    l64, a64, b64 = oracle.dag_loss(match, links, olen, tlen, True, np.float64)
    gm64, gl64 = oracle.dag_loss_backward(go, a64, b64, match, links, olen, tlen, np.float64)
    l32, a32, b32 = oracle.dag_loss(match, links, olen, tlen, True, np.float32)
    gm32, gl32 = oracle.dag_loss_backward(go, a32, b32, match, links, olen, tlen, np.float32)
    fin = np.isfinite(l64)
    assert np.array_equal(np.isfinite(loss), fin)
    assert np.allclose(loss[fin], l64[fin], rtol=1e-4, atol=0)
    for mine, truth, ref32 in ((gm, gm64, gm32), (gl, gl64, gl32)):
        assert relerr(mine, truth) <= max(1e-4, 1.5 * relerr(ref32, truth)), (relerr(mine, truth), relerr(ref32, truth))
    # every cell that carries posterior mass must be finite in alpha and beta
    post = a64 + b64 - match.astype(np.float64) - l64[:, None, None]
    relevant = np.isfinite(post) & (post > np.log(1e-12))
    assert np.isfinite(alpha.cpu().numpy()[relevant]).all() and np.isfinite(beta.cpu().numpy()[relevant]).all()


def test_exact_log_domain_switch():
    """EXACT_LOG_DOMAIN routes fp32 lattices to the log-domain kernels (no workspace): same results."""
    match, links, olen, tlen = oracle.make_lattice(3, 200, 30, 199, seed=21, ragged=True)
    go = np.ones(3, np.float32)
    fast = run_loss(match, links, olen, tlen, go, torch.float32, True)
    ops.EXACT_LOG_DOMAIN = True
    try:
        exact = run_loss(match, links, olen, tlen, go, torch.float32, True)
    finally:
        ops.EXACT_LOG_DOMAIN = False
    assert np.allclose(fast[0], exact[0], rtol=1e-5)
    assert relerr(fast[1], exact[1]) <= 1e-4 and relerr(fast[2], exact[2]) <= 1e-4
    assert torch.equal(torch.isfinite(fast[3]), torch.isfinite(exact[3]))
    assert torch.equal(torch.isfinite(fast[4]), torch.isfinite(exact[4]))


@pytest.mark.parametrize("config", [1, 2, 3, 4])
@pytest.mark.parametrize("shape", SHAPES[1:])
def test_viterbi_indices_bit_exact(shape, config):
    B, L, M, T, ragged, glat = shape
    match, links, olen, tlen = oracle.make_lattice(B, L, M, T, seed=7 + sum(shape[:4]), ragged=ragged, glat_frac=glat)
    if glat > 0:
        # quantise so that exact ties are frequent and the tie-break order matters
        links = np.where(np.isfinite(links), np.round(links * 2) / 2, links).astype(np.float32)
        match = np.where(np.isfinite(match), np.round(match), match).astype(np.float32)
    oalpha, opath, _ = oracle.dag_best_alignment(match, links, olen, tlen, config, np.float32)
    old = ops.DagBestAlignmentFunc.config
    try:
        ops.DagBestAlignmentFunc.config = config
        path = ops.dag_best_alignment(cu(match), cu(links), cu(olen), cu(tlen))
        alpha, p32 = ops.get_dag_kernel().dag_best_alignment(cu(match), cu(links), cu(olen), cu(tlen), config)
    finally:
        ops.DagBestAlignmentFunc.config = old
    assert path.dtype == torch.long and not path.requires_grad
    feasible = np.isfinite(oalpha[np.arange(B), tlen - 1, olen - 1])
    got = path.cpu().numpy()
    assert np.array_equal(got[feasible], opath[feasible].astype(np.int64))
    assert np.array_equal(p32.cpu().numpy()[feasible], opath[feasible])
    a = alpha.cpu().numpy()
    assert np.array_equal(a, oalpha)  # max-plus values are exact fp32 adds: bit-identical


def test_viterbi_crafted_ties():
    B, M, L, T = 1, 3, 12, 11
    match = np.zeros((B, M, L), np.float32)
    links = np.zeros((B, L, T), np.float32)
    links[0, 8, 0] = -1.0
    links[0, 4, 4] = -1.0
    olen, tlen = np.array([L]), np.array([M])
    for config in (1, 2, 3, 4):
        _, opath, _ = oracle.dag_best_alignment(match, links, olen, tlen, config, np.float32)
        _, p = ops.get_dag_kernel().dag_best_alignment(cu(match), cu(links), cu(olen), cu(tlen), config)
        assert np.array_equal(p.cpu().numpy(), opath)


@pytest.mark.parametrize("name", [n for n in DP_CASES if n != "c1_infeasible"])
def test_viterbi_against_reference_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    path = ops.dag_best_alignment(cu(g["match"]), cu(g["links"]), cu(g["olen"]), cu(g["tlen"]))
    assert np.array_equal(path.cpu().numpy(), g["viterbi_path"])


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GATHER_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16, torch.float64])
def test_logsoftmax_gather_against_golden(name, dtype):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    if dtype in (torch.float16, torch.bfloat16) and "fp16" not in name:
        logits_np = g["logits"].astype(np.float32)
    else:
        logits_np = g["logits"]
    x0 = cu(logits_np.astype(np.float32)).to(dtype)
    B, L, V = x0.shape
    idx = cu(g["targets"]).unsqueeze(1).expand(-1, L, -1)
    # reference values for THIS dtype's (rounded) logits
    xr = x0.detach().double().cpu().requires_grad_()
    sel_ref = torch.log_softmax(xr, -1).gather(-1, idx.cpu())
    w = torch.tensor(g["grad_selected"]).double()
    gref = torch.autograd.grad((sel_ref * w).sum(), [xr])[0]

    leaf = x0.clone().requires_grad_()
    work = leaf * 1  # non-leaf, as in the model (the op overwrites it in place)
    out, sel = ops.dag_logsoftmax_gather_inplace(work, idx)
    assert sel.shape == (B, L, g["targets"].shape[1]) and sel.dtype == (torch.float64 if dtype == torch.float64 else torch.float32)
    assert out.data_ptr() == work.data_ptr()
    tol = {torch.float32: 2e-5, torch.float64: 1e-10, torch.float16: 2e-5, torch.bfloat16: 2e-5}[dtype]
    assert np.allclose(sel.detach().cpu().numpy(), sel_ref.detach().numpy(), rtol=tol, atol=10 * tol)
    if dtype == torch.float32 and "fp16" not in name:
        assert np.allclose(sel.detach().cpu().numpy(), g["selected"], rtol=1e-4, atol=1e-4)
    # logits were overwritten with probabilities
    probs = torch.softmax(xr.detach(), -1)
    ptol = {torch.float32: 1e-5, torch.float64: 1e-10, torch.float16: 1e-3, torch.bfloat16: 8e-3}[dtype]
    assert np.allclose(out.detach().double().cpu().numpy(), probs.numpy(), rtol=ptol, atol=ptol * 1e-1)
    grad = torch.autograd.grad((sel * cu(w).to(sel.dtype)).sum(), [leaf])[0]
    gtol = {torch.float32: 1e-4, torch.float64: 1e-9, torch.float16: 2e-2, torch.bfloat16: 6e-2}[dtype]
    assert relerr(grad.double().cpu().numpy(), gref.numpy()) <= gtol
    # no-grad call must leave the logits untouched (logsoftmax_gather.cu:296)
    x1 = x0.clone()
    with torch.no_grad():
        _, sel2 = ops.dag_logsoftmax_gather_inplace(x1, idx)
    assert torch.equal(x1, x0)
    assert torch.allclose(sel2, sel.detach())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shared_idx", [True, False])
@pytest.mark.parametrize("VS", [(1024, 40), (4104, 300)])
def test_logsoftmax_gather_duplicates_and_ragged_groups(dtype, shared_idx, VS):
    """Targets with many duplicates (the scatter must accumulate, dag_loss.py:295), a vertex count that is not a
    multiple of the kernels' 8-row groups, expanded (stride-0) and genuinely strided index tensors."""
    torch.manual_seed(5)
    B, L = 3, 21
    V, S = VS
    x0 = (torch.randn(B, L, V, device=DEV) * 2).to(dtype)
    tg = torch.randint(0, 16, (B, S), device=DEV) * 8 + torch.randint(0, 3, (B, S), device=DEV)   # few distinct ids
    if shared_idx:
        idx = tg.unsqueeze(1).expand(-1, L, -1)
    else:
        idx = (tg.unsqueeze(1) + torch.arange(L, device=DEV).view(1, L, 1)) % V
    w = torch.randn(B, L, S, device=DEV)
    xr = x0.detach().double().requires_grad_()
    sel_ref = torch.log_softmax(xr, -1).gather(-1, idx)
    gref = torch.autograd.grad((sel_ref * w.double()).sum(), [xr])[0]
    leaf = x0.clone().requires_grad_()
    _, sel = ops.dag_logsoftmax_gather_inplace(leaf * 1, idx)
    assert torch.allclose(sel.double(), sel_ref.detach(), rtol=2e-5, atol=2e-4)
    grad = torch.autograd.grad((sel * w).sum(), [leaf])[0]
    gtol = {torch.float32: 1e-4, torch.float16: 2e-2, torch.bfloat16: 6e-2}[dtype]
    assert relerr(grad.double().cpu().numpy(), gref.cpu().numpy()) <= gtol


def test_device_prefetcher_overlapped_inputs_give_identical_results():
    """daspeech_b200.prefetch.DevicePrefetcher: inputs copied on a side stream one step ahead; every step must see
    exactly its own batch (different data per step) and produce the loss of the plain, serial path."""
    from daspeech_b200.prefetch import DevicePrefetcher
    host = []
    want = []
    for s in range(4):
        match, links, olen, tlen = oracle.make_lattice(3, 40, 12, 39, seed=100 + s, ragged=True)
        tens = (torch.tensor(match).pin_memory(), torch.tensor(links).pin_memory(), torch.tensor(olen).pin_memory(),
                torch.tensor(tlen).pin_memory())
        host.append(tens)
        dev_t = [t.to(DEV) for t in tens]
        dev_t[0].requires_grad_()     # same code path as below: with a gradient the loss is read from beta[0,0]
        want.append(ops.dag_loss(*dev_t).detach().cpu())
    got = []
    for m, lk, ol, tl in DevicePrefetcher(iter(host), torch.device(DEV)):
        m.requires_grad_()
        loss = ops.dag_loss(m, lk, ol, tl)
        loss.sum().backward()
        got.append(loss.detach().cpu())
    assert len(got) == 4
    for a, b in zip(got, want):
        assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 40, 12, 39), (2, 1024, 40, 1023), (2, 301, 17, 64)])
def test_dag_posterior_matches_the_criterion_formula(shape, dtype):
    """daspeech_b200.posterior: the S2S criterion's posterior (s2s_dag_fastspeech2_loss.py:259-262) from the lattices of
    dag_loss_with_alpha_beta, against the reference's own sequence of torch ops (with logsumexp_keepdim as exported)."""
    from daspeech_b200.posterior import dag_posterior, dag_expected_features
    B, L, M, T = shape
    match, links, olen, tlen = oracle.make_lattice(B, L, M, T, seed=11 + L, ragged=True)
    m, lk, ol, tl = cu(match), cu(links), cu(olen), cu(tlen)
    m.requires_grad_()
    _, (alpha, beta) = ops.dag_loss_with_alpha_beta(m, lk, ol, tl)
    alpha[0, -1] = float("-inf")                      # a row without any finite cell: the reference's NaN -> 0
    feats = torch.randn(B, L, 48, device=DEV).to(dtype).requires_grad_()
    ref = (alpha + beta - ops.logsumexp_keepdim(alpha + beta, -1)).exp()
    ref.masked_fill_(torch.isnan(ref), 0)
    got = dag_posterior(alpha, beta, dtype)
    assert got.dtype == dtype and got.shape == ref.shape
    tol = {torch.float32: 2e-5, torch.float16: 1e-3, torch.bfloat16: 8e-3}[dtype]
    assert torch.allclose(got.float(), ref.to(dtype).float(), rtol=tol, atol=tol * 1e-2)
    assert torch.all(got[0, -1] == 0)
    rows = got.float().sum(-1)
    live = torch.isfinite(alpha + beta).any(-1)
    assert torch.allclose(rows[live], torch.ones_like(rows[live]), atol=5 * tol)
    z = dag_expected_features(alpha, beta, feats)
    zr = torch.matmul(ref.to(feats), feats.detach())
    assert torch.allclose(z.float(), zr.float(), rtol=10 * tol, atol=10 * tol)
    z.float().sum().backward()
    assert feats.grad is not None and torch.isfinite(feats.grad).all() and m.grad is None


@pytest.mark.parametrize("shape", [(3, 12, 40), (2, 64, 1024), (2, 7, 33)])
def test_glat_force_emit_matches_the_criterion_expression(shape):
    """daspeech_b200.glat: the force-emit masking of nat_dag_loss.py:130-132, values and gradient, against the
    criterion's own torch expression."""
    from daspeech_b200.glat import glat_force_emit
    B, M, L = shape
    torch.manual_seed(3)
    base = torch.randn(B, M, L, device=DEV) - 5
    base[0, 1, 2] = float("-inf")
    matchmask = torch.zeros(B, M, L, dtype=torch.bool, device=DEV)
    matchmask.scatter_(1, torch.randint(0, M, (B, 1, L), device=DEV), True)      # one aligned target per vertex
    keep = torch.rand(B, L, device=DEV) < 0.4
    w = torch.randn(B, M, L, device=DEV)

    m_ref = base.clone().requires_grad_()
    prev = keep.unsqueeze(1)
    ref = m_ref.masked_fill(prev, 0) + m_ref.masked_fill(~matchmask, float("-inf")).masked_fill(~prev, 0).detach()
    m_new = base.clone().requires_grad_()
    got = glat_force_emit(m_new, matchmask, keep)
    assert torch.equal(got, ref)
    fin = torch.isfinite(ref)
    (ref[fin] * w[fin]).sum().backward()
    (got[fin] * w[fin]).sum().backward()
    assert torch.equal(m_new.grad, m_ref.grad)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16, torch.float64])
@pytest.mark.parametrize("VS", [(512, 32), (4104, 40), (1001, 7)])
def test_logsoftmax_gather_with_fused_argmax(dtype, VS):
    """dag_logsoftmax_gather_argmax_inplace: the gather plus `word_ins_out.argmax(-1)` of the raw logits in one pass
    (nat_dag_loss.py:209-213); TMA-staged shapes and the generic arg-max kernel; ties resolve to the first index."""
    torch.manual_seed(9)
    V, S = VS
    B, L = 2, 19
    x0 = (torch.randn(B, L, V, device=DEV) * 2).to(dtype)
    x0[0, 0, 5] = x0[0, 0].max() + 1            # an exact tie: both indices hold the row maximum
    x0[0, 0, 77] = x0[0, 0, 5]
    idx = torch.randint(0, V, (B, S), device=DEV).unsqueeze(1).expand(-1, L, -1)
    want_pred = x0.argmax(-1)
    want_pred[0, 0] = 5
    ref = torch.log_softmax(x0.double(), -1).gather(-1, idx)
    leaf = x0.clone().requires_grad_()
    out, sel, pred = ops.dag_logsoftmax_gather_argmax_inplace(leaf * 1, idx)
    assert pred.dtype == torch.long and not pred.requires_grad
    assert torch.equal(pred, want_pred)
    tol = 1e-10 if dtype == torch.float64 else 2e-5
    assert torch.allclose(sel.double(), ref, rtol=tol, atol=10 * tol)
    w = torch.randn_like(sel)
    g1 = torch.autograd.grad((sel * w).sum(), [leaf])[0]
    leaf2 = x0.clone().requires_grad_()
    _, sel2 = ops.dag_logsoftmax_gather_inplace(leaf2 * 1, idx)
    g2 = torch.autograd.grad((sel2 * w).sum(), [leaf2])[0]
    # duplicate targets are scattered with shared-memory atomics: equal up to the order of the additions
    assert relerr(g1.double().cpu().numpy(), g2.double().cpu().numpy()) <= (1e-5 if dtype in (torch.float32, torch.float64) else 2e-2)
    with torch.no_grad():
        x1 = x0.clone()
        _, _, pred2 = ops.dag_logsoftmax_gather_argmax_inplace(x1, idx)
    assert torch.equal(x1, x0) and torch.equal(pred2, want_pred)


def test_logsoftmax_gather_layouts_and_errors():
    x = torch.randn(2, 5, 64, device=DEV)
    idx_full = torch.randint(0, 64, (2, 5, 3), device=DEV)   # genuinely strided (non-expanded) indices
    _, a = ops.dag_logsoftmax_gather_inplace(x.clone(), idx_full)
    ref = torch.log_softmax(x, -1).gather(-1, idx_full)
    assert torch.allclose(a, ref, atol=1e-5)
    ops.TRANSPOSED_GATHER_OUTPUT = False
    try:
        _, b = ops.dag_logsoftmax_gather_inplace(x.clone(), idx_full)
        assert b.is_contiguous() and torch.allclose(b, ref, atol=1e-5)
    finally:
        ops.TRANSPOSED_GATHER_OUTPUT = True
    assert a.transpose(1, 2).is_contiguous()
    with pytest.raises(RuntimeError, match="select_idx should be long"):
        ops.dag_logsoftmax_gather_inplace(x.clone(), idx_full.int())
    with pytest.raises(RuntimeError, match="not contiguous"):
        ops.dag_logsoftmax_gather_inplace(x.transpose(0, 1), idx_full.transpose(0, 1))
    with pytest.raises(RuntimeError, match="length should be long"):
        ops.dag_loss(torch.zeros(1, 2, 4, device=DEV), torch.zeros(1, 4, 3, device=DEV),
                     torch.tensor([4], device=DEV, dtype=torch.int32), torch.tensor([2], device=DEV))
    with pytest.raises(RuntimeError, match="prelen not match"):
        ops.dag_loss(torch.zeros(1, 2, 4, device=DEV), torch.zeros(1, 5, 3, device=DEV),
                     torch.tensor([4], device=DEV), torch.tensor([2], device=DEV))
    with pytest.raises(RuntimeError, match="not implemented for 'float16'"):
        ops.dag_loss(torch.zeros(1, 2, 4, device=DEV).half(), torch.zeros(1, 4, 3, device=DEV).half(),
                     torch.tensor([4], device=DEV), torch.tensor([2], device=DEV))


def test_criterion_style_chain_end_to_end():
    """logits -> gather -> transpose -> (GLAT-style masking) -> dag_loss -> -(loss/len).mean() -> backward,
    the call sequence of NATDAGLoss._compute_dag_loss (nat_dag_loss.py:114-156), against the torch versions."""
    torch.manual_seed(0)
    B, L, M, V = 3, 40, 9, 128
    logits = (torch.randn(B, L, V, device=DEV) * 2).half()
    tgt = torch.randint(4, V, (B, M), device=DEV)
    _, lk_np, olen, tlen = oracle.make_lattice(B, L, M, L - 1, seed=5, ragged=True)
    links = cu(lk_np)
    ol, tl = cu(olen), cu(tlen)

    def chain(use_cuda):
        leaf = logits.clone().requires_grad_()
        lk = links.clone().requires_grad_()
        outputs = leaf * 1
        idx = tgt.unsqueeze(1).expand(-1, L, -1)
        if use_cuda:
            outputs, match = ops.dag_logsoftmax_gather_inplace(outputs, idx)
        else:
            outputs, match = ops.torch_dag_logsoftmax_gather_inplace(outputs, idx)
        match = match.transpose(1, 2)
        if use_cuda:
            loss = ops.dag_loss(match, lk, ol, tl)
        else:
            dense = torch.full((B, L, L + 1), float("-inf"), device=DEV)
            ii = (torch.arange(L, device=DEV).unsqueeze(1) + torch.arange(L - 1, device=DEV).unsqueeze(0) + 1).clamp(max=L)
            dense = dense.scatter(2, ii.unsqueeze(0).expand(B, -1, -1), lk)[:, :, :L]
            loss = ops.torch_dag_loss(match, dense, ol, tl)
        total = -(loss / tl).mean()
        g1, g2 = torch.autograd.grad(total, [leaf, lk])
        return total.item(), g1.float(), g2

    l_c, g1_c, g2_c = chain(True)
    l_t, g1_t, g2_t = chain(False)
    assert abs(l_c - l_t) <= 1e-4 * abs(l_t)
    assert relerr(g2_c.cpu().numpy(), g2_t.cpu().numpy()) <= 1e-4
    assert relerr(g1_c.cpu().numpy(), g1_t.cpu().numpy()) <= 2e-2  # fp16 logits gradient


def test_no_grad_beta_is_zero_like_reference():
    match, links, olen, tlen = oracle.make_lattice(2, 20, 6, 19, seed=3, ragged=True)
    with torch.no_grad():
        loss, (alpha, beta) = ops.dag_loss_with_alpha_beta(cu(match), cu(links), cu(olen), cu(tlen))
    assert not beta.any()  # reference returns its at::zeros beta untouched (dag_loss.cu:340,355)
    oa = oracle.dag_alpha(match, links, olen, tlen, np.float32)
    assert np.allclose(loss.cpu().numpy(), oa[np.arange(2), tlen - 1, olen - 1], rtol=1e-5)


def test_non_default_stream_and_empty_batch():
    match, links, olen, tlen = oracle.make_lattice(2, 50, 8, 49, seed=9, ragged=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        m, lk = cu(match), cu(links)
        loss = ops.dag_loss(m, lk, cu(olen), cu(tlen))
    s.synchronize()
    ol, _, _ = oracle.dag_loss(match, links, olen, tlen, False, np.float32)
    assert np.allclose(loss.cpu().numpy(), ol, rtol=1e-5)
    e = ops.dag_loss(torch.zeros(0, 4, 8, device=DEV), torch.zeros(0, 8, 7, device=DEV),
                     torch.zeros(0, dtype=torch.long, device=DEV), torch.zeros(0, dtype=torch.long, device=DEV))
    assert e.shape == (0,)


# --------------------------------------------------------------------------------------------------
def _ref_ext():
    from oracle import build_ref
    try:
        return build_ref.load_ref()
    except Exception:  # pragma: no cover
        return None


REF_SHAPES = [(2, 64, 32, 63, 1), (4096 // 50, 400, 50, 32, 2), (4, 1024, 256, 1023, 3), (3, 777, 100, 64, 4)]


@pytest.mark.parametrize("shape", REF_SHAPES)
def test_against_compiled_reference_cuda_extension(shape):
    """Differential test against the UNMODIFIED reference kernels (oracle/_ref/dag_loss_fn.so)."""
    ref = _ref_ext()
    if ref is None:
        pytest.skip("oracle/_ref/dag_loss_fn.so not present")
    B, L, M, T, seed = shape
    match, links, olen, tlen = oracle.make_lattice(B, L, M, T, seed=seed, ragged=True)
    m, lk, ol, tl = cu(match), cu(links), cu(olen), cu(tlen)
    go = torch.rand(B, device=DEV) + 0.5
    k = ops.get_dag_kernel()
    a1, b1 = k.dag_loss(m, lk, ol, tl, True, 1)
    gm1, gl1 = k.dag_loss_backward(go, a1, b1, m, lk, ol, tl, 2, 2)
    a0, b0 = ref.dag_loss(m, lk, ol, tl, True, 1)
    gm0, gl0 = ref.dag_loss_backward(go, a0, b0, m, lk, ol, tl, 2, 2)
    torch.cuda.synchronize()
    z1, z0 = b1[:, 0, 0], b0[:, 0, 0]
    assert torch.allclose(z1, z0, rtol=1e-4, atol=0)
    # alpha/beta: same -inf structure as the reference kernels up to the flush-to-zero contract
    check_lattice_side_outputs(a1.cpu().numpy(), b1.cpu().numpy(), a0.double().cpu().numpy(), b0.double().cpu().numpy(),
                               match, z0.double().cpu().numpy())
    # both are fp32 log-domain: judge both against the fp64 truth when the lattice is big
    scale = 1.0
    if M * L >= 100000:
        _, oa, ob = oracle.dag_loss(match, links, olen, tlen, True, np.float64)
        ogm, ogl = oracle.dag_loss_backward(go.cpu().numpy(), oa, ob, match, links, olen, tlen, np.float64)
        e_ref = max(relerr(gm0.cpu().numpy(), ogm), relerr(gl0.cpu().numpy(), ogl))
        e_new = max(relerr(gm1.cpu().numpy(), ogm), relerr(gl1.cpu().numpy(), ogl))
        assert e_new <= max(1e-4, 1.5 * e_ref), (e_new, e_ref)
        scale = max(1.0, 2.5 * e_ref / 1e-4)
    assert relerr(gm1.cpu().numpy(), gm0.cpu().numpy()) <= 1e-4 * scale
    assert relerr(gl1.cpu().numpy(), gl0.cpu().numpy()) <= 1e-4 * scale
    # Viterbi: indices bit-exact against the reference CUDA kernel
    av0, p0 = ref.dag_best_alignment(m, lk, ol, tl, 1)
    av1, p1 = k.dag_best_alignment(m, lk, ol, tl, 1)
    torch.cuda.synchronize()
    assert torch.equal(p1, p0)
    assert torch.equal(av1, av0)
    # gather (fp16 logits, fast-math reference): reference tolerance rtol 1e-3 / atol 1e-4 (dag_loss.py:567)
    V = 1000
    x = (torch.randn(B, L, V, device=DEV) * 2).half()
    idx = torch.randint(0, V, (B, M), device=DEV).unsqueeze(1).expand(-1, L, -1)
    xa, xb = x.clone(), x.clone()
    s0 = ref.logsoftmax_gather(xa, idx, True)
    s1 = k.logsoftmax_gather(xb, idx, True)
    torch.cuda.synchronize()
    assert torch.allclose(s1, s0, rtol=1e-3, atol=1e-4)
    assert torch.allclose(xb.float(), xa.float(), rtol=2e-3, atol=1e-6)


# --------------------------------------------------------------------------------------------------
def test_full_size_c2_properties():
    """BASELINE config C2 (B=64, L=1024, M=256, V=4096, T=1023): size-independent identities."""
    B, L, M, V = 64, 1024, 256, 4096
    T = L - 1
    g = torch.Generator(device=DEV).manual_seed(1234)
    tgt = torch.randint(4, V, (B, M), device=DEV, generator=g)
    logits = (torch.randn(B, L, V, device=DEV, generator=g) * 2).half().requires_grad_()
    tl = torch.randint(M // 2, M + 1, (B,), device=DEV, generator=g)
    ol = torch.maximum(torch.randint(L // 2, L + 1, (B,), device=DEV, generator=g), tl)
    raw = torch.randn(B, L, T, device=DEV, generator=g)
    i = torch.arange(L, device=DEV).view(1, L, 1)
    k = torch.arange(T, device=DEV).view(1, 1, T)
    valid = (i + k + 1) < ol.view(B, 1, 1)
    links = torch.log_softmax(raw.masked_fill(~valid, float("-inf")), -1)
    links = links.masked_fill(~valid, float("-inf")).requires_grad_()
    del raw
    work = logits * 1
    _, match = ops.dag_logsoftmax_gather_inplace(work, tgt.unsqueeze(1).expand(-1, L, -1))
    match = match.transpose(1, 2)
    assert match.is_contiguous()
    match.retain_grad()
    loss, (alpha, beta) = ops.dag_loss_with_alpha_beta(match, links, ol, tl)
    assert torch.isfinite(loss).all()
    za = alpha[torch.arange(B, device=DEV), tl - 1, ol - 1]
    assert torch.allclose(za, loss, rtol=1e-5)                      # forward and backward chains agree on Z
    go = torch.rand(B, device=DEV, generator=g) + 0.5
    (loss * go).sum().backward()
    gm, gl = match.grad, links.grad
    assert torch.isfinite(gm).all() and torch.isfinite(gl).all()
    rowsum = gm.sum(-1)                                              # posterior over vertices sums to go for t < Tn
    tmask = torch.arange(M, device=DEV).view(1, M) < tl.view(B, 1)
    # fp32 log-domain storage: |alpha|,|beta| ~ 1e3 carry ~1e-4 ulp, accumulated over M steps -> ~1e-3..1e-2
    assert torch.allclose(rowsum[tmask], go.view(B, 1).expand(B, M)[tmask], rtol=1e-2)
    assert not rowsum[~tmask].any()
    assert torch.allclose(gl.sum((1, 2)), go * (tl - 1), rtol=1e-2)  # expected number of transitions
    assert not gl.masked_select(~valid).any()
    # gradient of the logits sums to ~0 over the vocabulary (softmax Jacobian)
    gsum = logits.grad.float().sum(-1)
    assert gsum.abs().max() <= 5e-2
    # Viterbi path: one vertex per target token, monotone, ends at O-1, score <= log-marginal
    path = ops.dag_best_alignment(match.detach(), links.detach(), ol, tl)
    onpath = path >= 0
    assert torch.equal(onpath.sum(1), tl)
    assert (path[:, 0] == 0).all() and (path[torch.arange(B, device=DEV), ol - 1] == tl - 1).all()
    vals = path.masked_fill(~onpath, -1).cummax(1).values
    assert ((path == vals) | ~onpath).all()
    # two samples against the fp64 oracle at full size
    sub = [0, B - 1]
    mm = match.detach()[sub].cpu().numpy(); ll = links.detach()[sub].cpu().numpy()
    oo = ol[sub].cpu().numpy(); tt = tl[sub].cpu().numpy(); gg = go[sub].cpu().numpy()
    l64, a64, b64 = oracle.dag_loss(mm, ll, oo, tt, True, np.float64)
    gm64, gl64 = oracle.dag_loss_backward(gg, a64, b64, mm, ll, oo, tt, np.float64)
    l32, a32, b32 = oracle.dag_loss(mm, ll, oo, tt, True, np.float32)
    gm32, gl32 = oracle.dag_loss_backward(gg, a32, b32, mm, ll, oo, tt, np.float32)
    assert np.allclose(loss[sub].detach().cpu().numpy(), l64, rtol=1e-4, atol=0)
    for mine, truth, ref32 in ((gm[sub].cpu().numpy(), gm64, gm32), (gl[sub].cpu().numpy(), gl64, gl32)):
        assert relerr(mine, truth) <= max(1e-4, 1.5 * relerr(ref32, truth)), (relerr(mine, truth), relerr(ref32, truth))
    _, opath, _ = oracle.dag_best_alignment(mm, ll, oo, tt, 1, np.float32)
    assert np.array_equal(path[sub].cpu().numpy(), opath.astype(np.int64))
