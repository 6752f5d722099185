"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the operator layer, i.e. the
C ABI of libdagb200.so.  Checkers: the CPU oracle (oracle/), the committed golden vectors of the reference's
torch path (tests/golden/), and -- when oracle/_ref/ holds the prebuilt UNMODIFIED reference CUDA extension --
the reference kernels themselves on the same tensors.

Tolerances (BASELINE.json north_star): arg-max alignment indices bit-exact; fp32 loss within 1e-4 relative;
fp32 gradients ELEMENT-WISE: elem_err(g, g_ref) = max_i |g_i - ref_i| / max(|ref_i|, 1e-6 * max|ref|), i.e. a relative
bound for every element within six decades of the largest one and the matching absolute bound below that floor
(tiny posteriors count).  The bound is tol = max(1e-4, 1.5 x the same statistic of the oracle's fp32 build), both
against the fp64 oracle: the lattices are fp32 tensors at the boundary, so |alpha|, |beta| ~ 1e3 carry an ulp of
6e-5..1.2e-4 that no implementation can undercut (the reference tolerates rtol 1e-3 on the loss and torch.allclose
defaults on the gradients, custom_ops/dag_loss.py:478,492-493).  DESIGN.md section 6 tabulates the bounds per shape.
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
                  if not os.path.basename(p).startswith(("gather", "criterion")))
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


GRAD_FLOOR = 1e-6


def elem_err(x, ref, floor=GRAD_FLOOR):
    """max_i |x_i - ref_i| / max(|ref_i|, floor * max|ref|): element-wise relative error down to `floor` of the largest
    element, absolute (at the floor's scale) below it."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    s = np.abs(ref).max()
    if not s > 0:
        return float(np.abs(x).max())
    return float((np.abs(x - ref) / np.maximum(np.abs(ref), floor * s)).max())


def grad_tol(ref32, truth):
    """Element-wise bound for a gradient tensor: 1e-4, or 1.5 x what the reference's own fp32 arithmetic (oracle fp32
    build) achieves on the same lattice against the fp64 truth, whichever is larger."""
    return max(1e-4, 1.5 * elem_err(ref32, truth))


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
        tgm, tgl = g["grad_match"], g["grad_links"]
        if name.endswith("fp32"):
            # this golden was produced by the reference's torch path in fp32: element-wise the truth is the fp64 oracle
            _, a64, b64 = oracle.dag_loss(g["match"], g["links"], g["olen"], g["tlen"], True, np.float64)
            tgm, tgl = oracle.dag_loss_backward(g["grad_output"], a64, b64, g["match"], g["links"], g["olen"], g["tlen"], np.float64)
            assert relerr(tgm, g["grad_match"]) <= 1e-4 and relerr(tgl, g["grad_links"]) <= 1e-4
        etol_m = etol_l = 1e-9
        if dtype == torch.float32:
            _, a32, b32 = oracle.dag_loss(g["match"], g["links"], g["olen"], g["tlen"], True, np.float32)
            gm32, gl32 = oracle.dag_loss_backward(g["grad_output"], a32, b32, g["match"], g["links"], g["olen"], g["tlen"], np.float32)
            etol_m, etol_l = grad_tol(gm32, tgm), grad_tol(gl32, tgl)
        em, el = elem_err(gm, tgm), elem_err(gl, tgl)
        print("elem grad errors", em, el, "bounds", etol_m, etol_l)
        assert em <= etol_m and el <= etol_l, (em, el, etol_m, etol_l)
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
    (1, 2080, 24, 2079, False, 0.0), # 65 vertex blocks: beyond the shared memory of the tcgen05 recurrences (exact log-domain kernels)
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
    _, a32, b32 = oracle.dag_loss(match, links, olen, tlen, True, np.float32)
    gm32, gl32 = oracle.dag_loss_backward(go, a32, b32, match, links, olen, tlen, np.float32)
    f3 = fin[:, None, None]
    ogm, ogl = np.where(f3, ogm, 0), np.where(f3, ogl, 0)
    tol_m, tol_l = grad_tol(np.where(f3, gm32, 0), ogm), grad_tol(np.where(f3, gl32, 0), ogl)
    em, el = elem_err(gm, ogm), elem_err(gl, ogl)
    print("elem grad errors", em, el, "bounds", tol_m, tol_l)
    assert em <= tol_m, (em, tol_m)
    assert el <= tol_l, (el, tol_l)
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
        assert elem_err(mine, truth) <= grad_tol(ref32, truth), (elem_err(mine, truth), elem_err(ref32, truth))
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
    assert elem_err(fast[1], exact[1]) <= 2e-4 and elem_err(fast[2], exact[2]) <= 2e-4   # two fp32 paths, 1e-4 each
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
    assert elem_err(g2_c.cpu().numpy(), g2_t.cpu().numpy()) <= 2e-4    # two fp32 paths against each other
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


def test_blocked_recurrences_repeatable_back_to_back_on_two_streams():
    """The tcgen05 recurrences hand data between warps through mailboxes, mbarriers and the async proxy; a missing fence
    shows up as a rare run-to-run difference.  200 back-to-back launches at the C2 shape, alternating between two
    streams that run concurrently (each with its own workspace), must reproduce loss, alpha and beta BIT FOR BIT."""
    B, L, M = 32, 1024, 256
    T = L - 1
    g = torch.Generator(device=DEV).manual_seed(77)
    match = torch.log(torch.rand(B, M, L, device=DEV, generator=g) * 0.98 + 0.01)
    tl = torch.randint(M // 2, M + 1, (B,), device=DEV, generator=g)
    ol = torch.maximum(torch.randint(L // 2, L + 1, (B,), device=DEV, generator=g), tl)
    raw = torch.randn(B, L, T, device=DEV, generator=g)
    i = torch.arange(L, device=DEV).view(1, L, 1)
    kk = torch.arange(T, device=DEV).view(1, 1, T)
    valid = (i + kk + 1) < ol.view(B, 1, 1)
    links = torch.log_softmax(raw.masked_fill(~valid, float("-inf")), -1).masked_fill(~valid, float("-inf")).contiguous()
    del raw, valid
    k = ops.get_dag_kernel()
    a0, b0 = k.dag_loss(match, links, ol, tl, True, 1)
    torch.cuda.synchronize()
    assert torch.isfinite(b0[:, 0, 0]).all()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    bad = [torch.zeros((), dtype=torch.int64, device=DEV) for _ in streams]
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    for it in range(200):
        si = it & 1
        with torch.cuda.stream(streams[si]):
            a, b = k.dag_loss(match, links, ol, tl, True, 1)
            bad[si] += (a.view(torch.int32) != a0.view(torch.int32)).sum() + (b.view(torch.int32) != b0.view(torch.int32)).sum()
            del a, b
    for s in streams:
        s.synchronize()
    assert int(bad[0]) == 0 and int(bad[1]) == 0, (int(bad[0]), int(bad[1]))


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
    # both are fp32 log-domain at the boundary: judge both, element-wise, against the fp64 truth; the new kernels must
    # be within max(1e-4, 1.5 x the reference CUDA kernels' own error), and within the sum of the two of each other
    _, oa, ob = oracle.dag_loss(match, links, olen, tlen, True, np.float64)
    ogm, ogl = oracle.dag_loss_backward(go.cpu().numpy(), oa, ob, match, links, olen, tlen, np.float64)
    for mine, theirs, truth in ((gm1, gm0, ogm), (gl1, gl0, ogl)):
        mine, theirs = mine.cpu().numpy(), theirs.cpu().numpy()
        e_new, e_ref = elem_err(mine, truth), elem_err(theirs, truth)
        print("elem grad errors vs fp64: new", e_new, "reference CUDA", e_ref)
        assert e_new <= max(1e-4, 1.5 * e_ref), (e_new, e_ref)
        assert elem_err(mine, theirs) <= max(1e-4, e_new + e_ref)
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


def test_pybind_shim_is_a_drop_in_for_the_reference_native_module():
    """daspeech_b200/csrc/dag_loss_fn_b200.so called exactly as the reference's Python layer calls its native module
    (custom_ops/dag_loss.py:105,118,227,272 and the two torch ops of :294-295): same results, bit for bit, as the
    ctypes operator layer on the same tensors."""
    from daspeech_b200.csrc import build_shim
    mod = build_shim.load()
    k = ops.get_dag_kernel()
    B, L, M, T, V = 3, 200, 30, 199, 1000
    match, links, olen, tlen = oracle.make_lattice(B, L, M, T, seed=31, ragged=True)
    m, lk, ol, tl = cu(match), cu(links), cu(olen), cu(tlen)
    go = torch.rand(B, device=DEV) + 0.5
    a0, b0 = k.dag_loss(m, lk, ol, tl, True, 1)
    a1, b1 = mod.dag_loss(m, lk, ol, tl, True, 1)
    assert torch.equal(a0, a1) and torch.equal(b0, b1)
    a2, b2 = mod.dag_loss(m, lk, ol, tl, False, 1)
    assert torch.equal(a2, a0) and not b2.any()
    gm0, gl0 = k.dag_loss_backward(go, a0, b0, m, lk, ol, tl, 2, 2)
    gm1, gl1 = mod.dag_loss_backward(go, a1, b1, m, lk, ol, tl, 2, 2)
    assert torch.equal(gm0, gm1) and torch.equal(gl0, gl1)
    av0, p0 = k.dag_best_alignment(m, lk, ol, tl, 1)
    av1, p1 = mod.dag_best_alignment(m, lk, ol, tl, 1)
    assert p1.dtype == torch.int32 and torch.equal(p0, p1) and torch.equal(av0, av1)
    # gather + the reference wrapper's own backward ops on the buffer the module overwrote
    x = (torch.randn(B, L, V, device=DEV) * 2).half()
    idx = torch.randint(0, V, (B, M), device=DEV).unsqueeze(1).expand(-1, L, -1)
    xa, xb = x.clone(), x.clone()
    s0 = k.logsoftmax_gather(xa, idx, True)
    s1 = mod.logsoftmax_gather(xb, idx, True)
    assert s1.shape == (B, L, M) and torch.equal(s0, s1) and torch.equal(xa, xb)
    g = torch.randn(B, L, M, device=DEV)
    gi = xb.mul_(g.sum(-1, keepdim=True).neg().to(xb.dtype))        # dag_loss.py:294
    gi.scatter_add_(-1, idx, g.to(xb.dtype))                         # dag_loss.py:295
    k.logsoftmax_gather_backward(xa, idx, g)
    assert relerr(xa.float().cpu().numpy(), gi.float().cpu().numpy()) <= 2e-2
    with pytest.raises(RuntimeError, match="length should be long"):
        mod.dag_loss(m, lk, ol.int(), tl, True, 1)
    with pytest.raises(RuntimeError, match="select_idx should be long"):
        mod.logsoftmax_gather(x.clone(), idx.int(), True)


def test_logsoftmax_gather_backward_has_no_vocabulary_limit():
    """Vocabularies beyond the shared-memory staging limit (> ~51 k fp32 elements) take the global-memory backward: the
    forward must not succeed on a shape whose backward then fails (the reference's mul_ + scatter_add_ has no limit)."""
    torch.manual_seed(2)
    B, L, V, S = 1, 5, 66000, 9
    for dtype, gtol in ((torch.float32, 1e-4), (torch.float16, 2e-2), (torch.float64, 1e-9)):
        x0 = (torch.randn(B, L, V, device=DEV) * 2).to(dtype)
        tg = torch.randint(0, V, (B, S), device=DEV)
        tg[0, 1] = tg[0, 0]                                           # a duplicate target: the scatter accumulates
        idx = tg.unsqueeze(1).expand(-1, L, -1)
        w = torch.randn(B, L, S, device=DEV)
        xr = x0.double().requires_grad_()
        sel_ref = torch.log_softmax(xr, -1).gather(-1, idx)
        gref = torch.autograd.grad((sel_ref * w.double()).sum(), [xr])[0]
        leaf = x0.clone().requires_grad_()
        _, sel = ops.dag_logsoftmax_gather_inplace(leaf * 1, idx)
        assert torch.allclose(sel.double(), sel_ref.detach(), rtol=2e-5, atol=2e-4)
        grad = torch.autograd.grad((sel * w.to(sel.dtype)).sum(), [leaf])[0]
        assert relerr(grad.double().cpu().numpy(), gref.cpu().numpy()) <= gtol


def test_workspace_contents_never_leak_between_calls():
    """The cached scratch is uninitialised and reused: poison it with NaN bit patterns, then run a ragged lattice whose
    tiles beyond each utterance's length are skipped by the precompute -- results must not change."""
    match, links, olen, tlen = oracle.make_lattice(3, 300, 40, 299, seed=77, ragged=True)
    go = np.ones(3, np.float32)
    k = ops.get_dag_kernel()
    first = run_loss(match, links, olen, tlen, go, torch.float32, True)
    for buf in list(k._scratch.values()):
        buf.fill_(0xFF)
    again = run_loss(match, links, olen, tlen, go, torch.float32, True)
    assert np.array_equal(first[0], again[0]) and np.array_equal(first[1], again[1]) and np.array_equal(first[2], again[2])
    assert torch.equal(first[3], again[3]) and torch.equal(first[4], again[4])
    p0 = ops.dag_best_alignment(cu(match), cu(links), cu(olen), cu(tlen))
    for buf in list(k._scratch.values()):
        buf.fill_(0xFF)
    assert torch.equal(p0, ops.dag_best_alignment(cu(match), cu(links), cu(olen), cu(tlen)))
    k.release_workspaces()
    assert not k._scratch


def test_sharded_nll_on_cuda_single_process():
    """daspeech_b200.dist on CUDA tensors (the gloo test covers the two-rank logic on CPU): every logging scalar lives on
    the loss device, so the flat stats tensor can be all-reduced over NCCL."""
    from daspeech_b200 import dist as ddist
    match, links, olen, tlen = oracle.make_lattice(4, 40, 9, 39, seed=3, ragged=True)
    mean, local, stats = ddist.dag_nll_sharded(ops.dag_loss, cu(match), cu(links), cu(olen), cu(tlen))
    assert all(v.is_cuda for v in stats.values())
    l64, _, _ = oracle.dag_loss(match, links, olen, tlen, False, np.float64)
    assert abs(float(mean) - float(-(l64 / tlen).mean())) <= 1e-4 * abs(float(mean))
    assert int(stats["nsentences"]) == 4 and int(stats["ntokens"]) == int(tlen.sum())
    ex = ddist.FlatGradAllReduce(1 << 16, torch.float32, torch.device(DEV))
    ex.buffer.fill_(2.0)
    ex.start()
    assert float(ex.finish().sum()) == 2.0 * (1 << 16)     # world size 1: unchanged


def test_device_status_is_reported_without_debug_mode():
    """Per-sample precondition violations (the reference's CUDA_KERNEL_ASSERTs) are always recorded; outside
    DAGB200_DEBUG they surface as a RuntimeWarning at the next status check instead of a sticky device assert."""
    import warnings
    match, links, olen, tlen = oracle.make_lattice(2, 20, 6, 19, seed=3, ragged=False)
    tlen[1] = 1                                                      # target length < 2 (dag_loss.cu:68)
    k = ops.get_dag_kernel()
    k.check_pending_status(wait=True)
    a, b = k.dag_loss(cu(match), cu(links), cu(olen), cu(tlen), True, 1)
    assert torch.isinf(b[1]).all() and torch.isfinite(b[0, 0, 0])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        k.check_pending_status(wait=True)
    assert any("at least 2" in str(x.message) for x in w)


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
        print("C2 elem grad error", elem_err(mine, truth), "fp32 reference arithmetic", elem_err(ref32, truth))
        assert elem_err(mine, truth) <= grad_tol(ref32, truth), (elem_err(mine, truth), elem_err(ref32, truth))
    _, opath, _ = oracle.dag_best_alignment(mm, ll, oo, tt, 1, np.float32)
    assert np.array_equal(path[sub].cpu().numpy(), opath.astype(np.int64))


def test_cuda_graph_replay_of_the_step_equals_the_eager_calls():
    """daspeech_b200.graphs.GraphedDagLossStep: the six kernels of a step captured once and replayed -- same bytes as
    the eager calls, also after the inputs were rewritten in place."""
    from daspeech_b200.graphs import GraphedDagLossStep
    B, L, M, T = 6, 320, 70, 319
    k = ops.get_dag_kernel()
    go = torch.rand(B, device=DEV) + 0.5
    lat = [oracle.make_lattice(B, L, M, T, seed=s, ragged=True) for s in (5, 6)]
    m, lk, ol, tl = (cu(x) for x in lat[0])
    step = GraphedDagLossStep(m, lk, ol, tl, go)
    for match, links, olen, tlen in lat:
        m.copy_(cu(match)); lk.copy_(cu(links)); ol.copy_(cu(olen)); tl.copy_(cu(tlen))
        a1, b1, gm1, gl1 = (t.clone() for t in step.replay())
        a0, b0 = k.dag_loss(m, lk, ol, tl, True, 1)
        gm0, gl0 = k.dag_loss_backward(go, a0, b0, m, lk, ol, tl, 2, 2)
        for x, y in ((a1, a0), (b1, b0), (gm1, gm0), (gl1, gl0)):
            assert torch.equal(x, y)
