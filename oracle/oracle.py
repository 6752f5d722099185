"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; the product never imports this).

numpy front-end over oracle/libdag_oracle.so (dag_oracle.c), the plain-C restatement of the
reference CUDA operators of DASpeech/custom_ops (file:line citations in dag_oracle_impl.h).

Allowed importers: tests/, __graft_entry__.smoke(), and the cpu_baseline / `--impl reference`
legs of bench.py.  Pinned against tests/golden/ (vectors produced by the reference's own torch
functions, tests/golden/make_golden.py) and against the compiled reference CUDA extension
(oracle/_ref/, on the GPU box).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdag_oracle.so")
_lib = None

_c_i64p = ctypes.POINTER(ctypes.c_int64)
_c_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
            for f in ("dag_oracle.c", "dag_oracle_impl.h")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libdag_oracle.so"],
                              env={**os.environ, "CC": "gcc"})
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "_f32", ctypes.c_float
    if dtype == np.float64:
        return "_f64", ctypes.c_double
    raise TypeError(dtype)


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _prep(match, links, olen, tlen, dtype):
    match = np.ascontiguousarray(match, dtype=dtype)
    links = np.ascontiguousarray(links, dtype=dtype)
    olen = np.ascontiguousarray(olen, dtype=np.int64)
    tlen = np.ascontiguousarray(tlen, dtype=np.int64)
    B, M, L = match.shape
    assert links.shape[0] == B and links.shape[1] == L
    return match, links, olen, tlen, B, M, L, links.shape[2]


def dag_alpha(match, links, olen, tlen, dtype=np.float32):
    """alpha [B,M,L] -- dag_loss.cu:71-131."""
    sfx, ct = _suffix(dtype)
    match, links, olen, tlen, B, M, L, T = _prep(match, links, olen, tlen, dtype)
    alpha = np.empty((B, M, L), dtype=dtype)
    getattr(lib(), "oracle_alpha" + sfx)(_p(match, ct), _p(links, ct), _p(olen, ctypes.c_int64),
                                         _p(tlen, ctypes.c_int64), _p(alpha, ct), B, M, L, T)
    return alpha


def dag_beta(match, links, olen, tlen, dtype=np.float32):
    """beta [B,M,L] -- dag_loss.cu:206-265."""
    sfx, ct = _suffix(dtype)
    match, links, olen, tlen, B, M, L, T = _prep(match, links, olen, tlen, dtype)
    beta = np.empty((B, M, L), dtype=dtype)
    getattr(lib(), "oracle_beta" + sfx)(_p(match, ct), _p(links, ct), _p(olen, ctypes.c_int64),
                                        _p(tlen, ctypes.c_int64), _p(beta, ct), B, M, L, T)
    return beta


def dag_loss(match, links, olen, tlen, require_gradient=True, dtype=np.float32):
    """(loss [B], alpha, beta) with the wrapper's choice of Z -- dag_loss.py:102-112."""
    alpha = dag_alpha(match, links, olen, tlen, dtype)
    olen64 = np.asarray(olen, dtype=np.int64)
    tlen64 = np.asarray(tlen, dtype=np.int64)
    B = alpha.shape[0]
    if require_gradient:
        beta = dag_beta(match, links, olen, tlen, dtype)
        loss = beta[:, 0, 0].copy()
    else:
        beta = np.full_like(alpha, -np.inf)
        loss = alpha[np.arange(B), tlen64 - 1, olen64 - 1].copy()
    return loss, alpha, beta


def dag_loss_backward(grad_output, alpha, beta, match, links, olen, tlen, dtype=np.float32):
    """(grad_match [B,M,L], grad_links [B,L,T]) -- dag_loss.cu:395-399, 461-484."""
    sfx, ct = _suffix(dtype)
    match, links, olen, tlen, B, M, L, T = _prep(match, links, olen, tlen, dtype)
    alpha = np.ascontiguousarray(alpha, dtype=dtype)
    beta = np.ascontiguousarray(beta, dtype=dtype)
    go = np.ascontiguousarray(grad_output, dtype=dtype)
    gm = np.empty((B, M, L), dtype=dtype)
    gl = np.empty((B, L, T), dtype=dtype)
    getattr(lib(), "oracle_grad_match" + sfx)(_p(go, ct), _p(alpha, ct), _p(beta, ct), _p(match, ct),
                                              _p(gm, ct), B, M, L)
    getattr(lib(), "oracle_grad_links" + sfx)(_p(go, ct), _p(alpha, ct), _p(beta, ct), _p(links, ct),
                                              _p(olen, ctypes.c_int64), _p(tlen, ctypes.c_int64),
                                              _p(gl, ct), B, M, L, T)
    return gm, gl


def dag_best_alignment(match, links, olen, tlen, config=1, dtype=np.float32):
    """(alpha_max [B,M,L], path [B,L] int32, trace [B,M,L] int32) -- dag_best_alignment.cu:72-122,178-184.
    config 1..4 selects the reference's TRANS_BLOCK_SIZE 4/8/16/32 (tie-break order)."""
    sfx, ct = _suffix(dtype)
    match, links, olen, tlen, B, M, L, T = _prep(match, links, olen, tlen, dtype)
    width = {1: 4, 2: 8, 3: 16, 4: 32}[int(config)]
    alpha = np.empty((B, M, L), dtype=dtype)
    trace = np.empty((B, M, L), dtype=np.int32)
    path = np.empty((B, L), dtype=np.int32)
    getattr(lib(), "oracle_viterbi" + sfx)(_p(match, ct), _p(links, ct), _p(olen, ctypes.c_int64),
                                           _p(tlen, ctypes.c_int64), _p(alpha, ct),
                                           _p(trace, ctypes.c_int32), _p(path, ctypes.c_int32),
                                           B, M, L, T, width)
    return alpha, path, trace


def _idx_strides(idx):
    assert idx.dtype == np.int64 and idx.ndim == 3
    return tuple(int(s // 8) for s in idx.strides)


def logsoftmax_gather(logits, idx, require_gradient=False, dtype=np.float32):
    """(selected [B,L,S], probs [B,L,V] or None) -- logsoftmax_gather.cu:268-308.
    `idx` may be a stride-0 broadcast view (np.broadcast_to), as the criterion passes."""
    sfx, ct = _suffix(dtype)
    logits = np.ascontiguousarray(logits, dtype=dtype)
    B, L, V = logits.shape
    idx = np.asarray(idx)
    S = idx.shape[2]
    sb, sl, ss = _idx_strides(idx)
    out = np.empty((B, L, S), dtype=dtype)
    probs = np.empty((B, L, V), dtype=dtype) if require_gradient else None
    base = ctypes.cast(idx.ctypes.data, _c_i64p)
    getattr(lib(), "oracle_logsoftmax_gather" + sfx)(
        _p(logits, ct), base, ctypes.c_int64(sb), ctypes.c_int64(sl), ctypes.c_int64(ss),
        _p(out, ct), _p(probs, ct) if probs is not None else None, B, L, V, S)
    return out, probs


def logsoftmax_gather_backward(probs, idx, grad_out, dtype=np.float32):
    """grad_logits [B,L,V] -- dag_loss.py:293-295."""
    sfx, ct = _suffix(dtype)
    probs = np.ascontiguousarray(probs, dtype=dtype)
    grad_out = np.ascontiguousarray(grad_out, dtype=dtype)
    B, L, V = probs.shape
    idx = np.asarray(idx)
    S = idx.shape[2]
    sb, sl, ss = _idx_strides(idx)
    gin = np.empty((B, L, V), dtype=dtype)
    base = ctypes.cast(idx.ctypes.data, _c_i64p)
    getattr(lib(), "oracle_logsoftmax_gather_backward" + sfx)(
        _p(probs, ct), base, ctypes.c_int64(sb), ctypes.c_int64(sl), ctypes.c_int64(ss),
        _p(grad_out, ct), _p(gin, ct), B, L, V, S)
    return gin


# --------------------------------------------------------------------------------------
# Layout helper used by tests: banded links [B,L,T] -> dense [B,L,L] as the reference's torch
# path expects (restore_valid_links, models/s2t_conformer_dag.py:157-169 / dag_loss.py:439-448).
def dense_links(links):
    links = np.asarray(links)
    B, L, T = links.shape
    dense = np.full((B, L, L), -np.inf, dtype=links.dtype)
    for i in range(L):
        n = min(T, L - 1 - i)
        if n > 0:
            dense[:, i, i + 1:i + 1 + n] = links[:, i, :n]
    return dense


# Synthetic lattice generator shared by tests and bench.
def make_lattice(B, L, M, T=None, seed=0, ragged=True, dtype=np.float32, glat_frac=0.0):
    rng = np.random.default_rng(seed)
    T = (L - 1) if T is None else T
    if ragged:
        tlen = rng.integers(max(2, M // 2), M + 1, size=B)
        lo = np.maximum(tlen, L // 2)
        olen = rng.integers(lo, L + 1)
        # Viterbi feasibility precondition (dag_best_alignment.cu:69)
        olen = np.minimum(olen, (tlen - 1) * T + 1)
        olen = np.maximum(olen, tlen)
    else:
        tlen = np.full(B, M)
        olen = np.full(B, L)
    match = np.log(rng.random((B, M, L)) * 0.98 + 0.01).astype(dtype)
    raw = rng.standard_normal((B, L, T))
    i = np.arange(L)[:, None]
    k = np.arange(T)[None, :]
    valid = (i + k + 1)[None] < olen[:, None, None]
    raw = np.where(valid, raw, -np.inf)
    mx = np.max(raw, axis=-1, keepdims=True)
    mx = np.where(np.isfinite(mx), mx, 0.0)
    ex = np.exp(raw - mx)
    s = ex.sum(-1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        links = np.where(s > 0, raw - mx - np.log(np.where(s > 0, s, 1.0)), -np.inf)
    links = links.astype(dtype)
    if glat_frac > 0:
        # GLAT force-emit shape (nat_dag_loss.py:130-132): forced vertices emit exactly one token
        for b in range(B):
            nf = int(glat_frac * min(olen[b], tlen[b]))
            verts = np.sort(rng.choice(np.arange(1, olen[b] - 1), size=min(nf, max(olen[b] - 2, 0)), replace=False))
            toks = np.sort(rng.choice(np.arange(1, tlen[b] - 1), size=min(len(verts), max(tlen[b] - 2, 0)), replace=False))
            for v, tk in zip(verts[:len(toks)], toks):
                match[b, :, v] = -np.inf
                match[b, tk, v] = 0.0
    return match, links, olen.astype(np.int64), tlen.astype(np.int64)
