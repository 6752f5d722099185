"""CPU: pin the C oracle (oracle/dag_oracle.c) against the golden vectors produced by the
reference's own torch implementations (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DP_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith(("gather", "criterion")))
GATHER_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "gather*.npz")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def test_golden_present():
    assert len(DP_CASES) >= 7 and len(GATHER_CASES) >= 3


@pytest.mark.parametrize("name", DP_CASES)
def test_loss_and_grads_f64(name):
    g = load(name)
    loss, alpha, beta = oracle.dag_loss(g["match"], g["links"], g["olen"], g["tlen"], True, np.float64)
    ref = g["loss"].astype(np.float64)
    tol = 1e-5 if name.endswith("fp32") else 1e-10
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(loss), fin)
    assert np.allclose(loss[fin], ref[fin], rtol=tol, atol=tol)
    # forward-only result (alpha[Tn-1, O-1]) must agree with Z = beta[0,0]
    loss_a, _, _ = oracle.dag_loss(g["match"], g["links"], g["olen"], g["tlen"], False, np.float64)
    assert np.allclose(loss_a[fin], ref[fin], rtol=tol, atol=tol)
    if "grad_match" in g:
        gm, gl = oracle.dag_loss_backward(g["grad_output"], alpha, beta, g["match"], g["links"],
                                          g["olen"], g["tlen"], np.float64)
        gtol = 2e-5 if name.endswith("fp32") else 1e-10
        assert np.allclose(gm, g["grad_match"], rtol=gtol, atol=gtol)
        assert np.allclose(gl, g["grad_links"], rtol=gtol, atol=gtol)
        # sanity identities
        for b in range(len(ref)):
            Tn = int(g["tlen"][b])
            assert np.allclose(gm[b, :Tn].sum(-1), float(g["grad_output"][b]), rtol=1e-9)
            assert np.allclose(gl[b].sum(), float(g["grad_output"][b]) * (Tn - 1), rtol=1e-9)


@pytest.mark.parametrize("name", DP_CASES)
def test_loss_and_grads_f32(name):
    """fp32 build of the oracle stays within the 1e-4 relative contract of the fp64 truth."""
    g = load(name)
    loss, alpha, beta = oracle.dag_loss(g["match"], g["links"], g["olen"], g["tlen"], True, np.float32)
    ref = g["loss"]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(loss), fin)
    assert np.allclose(loss[fin], ref[fin], rtol=1e-5)
    if "grad_match" in g:
        gm, gl = oracle.dag_loss_backward(g["grad_output"], alpha, beta, g["match"], g["links"],
                                          g["olen"], g["tlen"], np.float32)
        assert np.abs(gm - g["grad_match"]).max() <= 1e-4 * np.abs(g["grad_match"]).max()
        assert np.abs(gl - g["grad_links"]).max() <= 1e-4 * np.abs(g["grad_links"]).max()


def test_infeasible_sample_is_minus_inf_with_zero_grads():
    g = load("c1_infeasible")
    loss, alpha, beta = oracle.dag_loss(g["match"], g["links"], g["olen"], g["tlen"], True, np.float32)
    assert loss[1] == -np.inf and np.isfinite(loss[0])
    gm, gl = oracle.dag_loss_backward(np.ones(2, np.float32), alpha, beta, g["match"], g["links"],
                                      g["olen"], g["tlen"], np.float32)
    assert not gm[1].any() and not gl[1].any()
    assert np.isfinite(gm).all() and np.isfinite(gl).all()


def path_score(path, match, links, olen, tlen):
    """Re-score an alignment the way the reference self-test does (dag_loss.py:497-512)."""
    B, M, L = match.shape
    out = np.zeros(B)
    for b in range(B):
        verts = [j for j in range(L) if path[b, j] >= 0]
        assert [path[b, j] for j in verts] == list(range(int(tlen[b])))
        assert verts[0] == 0 and verts[-1] == int(olen[b]) - 1
        s = float(match[b, 0, 0])
        for t in range(1, len(verts)):
            i, j = verts[t - 1], verts[t]
            s += float(links[b, i, j - i - 1]) + float(match[b, t, j])
        out[b] = s
    return out


@pytest.mark.parametrize("name", [n for n in DP_CASES if n != "c1_infeasible"])
def test_viterbi(name):
    g = load(name)
    for dt in (np.float32, np.float64):
        alpha, path, _ = oracle.dag_best_alignment(g["match"], g["links"], g["olen"], g["tlen"], 1, dt)
        B = len(g["olen"])
        score = alpha[np.arange(B), g["tlen"] - 1, g["olen"] - 1]
        assert np.allclose(score, g["viterbi_score"], rtol=1e-5)
        # continuous random inputs: no ties, so the reference torch path is THE path
        assert np.array_equal(path.astype(np.int64), g["viterbi_path"])
        assert np.allclose(path_score(path, g["match"], g["links"], g["olen"], g["tlen"]), score, rtol=1e-5)


def test_viterbi_tie_break_order():
    """Crafted exact ties: winner follows lane priority bit-reverse(0..W-1), then smaller delta
    (dag_best_alignment.cu:100-111)."""
    B, M, L, T = 1, 3, 12, 11
    match = np.zeros((B, M, L), np.float32)
    links = np.zeros((B, L, T), np.float32)   # every candidate ties at 0
    olen = np.array([L]); tlen = np.array([M])
    _, _, trace = oracle.dag_best_alignment(match, links, olen, tlen, 1, np.float32)
    # row t=2, cell j: candidates delta=1..j-1 reach alpha[1][j-delta] (finite for j-delta>=1).
    # width 4: lane 0 holds delta 1,5,9 -> first strict max is delta=1 whenever j-1>=1
    assert trace[0, 2, 5] == 4
    # make delta=1 and delta=2 (lanes 0,1) worse so lanes 2 (delta=3) and 3 (delta=4) tie: lane 2 wins
    links2 = links.copy()
    j = 9
    links2[0, j - 1, 0] = -1.0
    links2[0, j - 2, 1] = -1.0
    links2[0, j - 5, 4] = -1.0   # delta=5 (lane 0)
    _, _, trace2 = oracle.dag_best_alignment(match, links2, olen, tlen, 1, np.float32)
    # t=1 values: alpha[1][i] = links[0][i-1] = 0 for all i, so at t=2 candidates are the link values
    assert trace2[0, 2, j] == j - 3
    # lane 1 (delta=2) vs lane 2 (delta=3) tie, lane 0 out: bit-reversed priority => lane 2 (delta=3) wins
    links3 = links.copy()
    links3[0, j - 1, 0] = -1.0
    links3[0, j - 5, 4] = -1.0
    _, _, trace3 = oracle.dag_best_alignment(match, links3, olen, tlen, 1, np.float32)
    assert trace3[0, 2, j] == j - 3
    # config 4 (width 32): single lane per delta<=32, priority is bit-reversal of (delta-1) over 5 bits
    _, _, trace4 = oracle.dag_best_alignment(match, links3, olen, tlen, 4, np.float32)
    # candidates delta in {2,3,4,6,7,8} tie (delta=1,5 penalised); lanes 1,2,3,5,6,7 -> bitrev5: 16,8,24,20,12,28 -> lane 2 (delta=3)
    assert trace4[0, 2, j] == j - 3


@pytest.mark.parametrize("name", GATHER_CASES)
def test_gather(name):
    g = load(name)
    logits = g["logits"].astype(np.float32)
    B, L, V = logits.shape
    idx = np.broadcast_to(g["targets"][:, None, :], (B, L, g["targets"].shape[1]))
    sel, probs = oracle.logsoftmax_gather(logits, idx, True, np.float32)
    assert np.allclose(sel, g["selected"], rtol=1e-5, atol=1e-5)
    assert np.allclose(probs.sum(-1), 1.0, rtol=1e-5)
    gin = oracle.logsoftmax_gather_backward(probs, idx, g["grad_selected"], np.float32)
    assert np.allclose(gin, g["grad_logits"], rtol=1e-4, atol=1e-5)
