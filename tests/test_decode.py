"""GPU inference decoding (daspeech_b200.decode) against a transcription of the reference's decoding branch
(DASpeech/models/s2s_conformer_dag_fastspeech2.py:211-304: dense links, torch.max, Python walks over .tolist()-ed
back-pointers) on the same tensors: token sequences, gathered features and lengths must be identical."""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu
PAD = 1


def dense_links(links):
    bsz, prelen, translen = links.shape
    idx = torch.arange(prelen, device=links.device).unsqueeze(1) + torch.arange(translen, device=links.device).unsqueeze(0) + 1
    idx = idx.masked_fill(idx >= prelen, prelen)
    res = links.new_full((bsz, prelen, prelen + 1), float("-inf"))
    res.scatter_(2, idx.unsqueeze(0).expand(bsz, -1, -1), links)
    return res[:, :, :prelen]


def reference_decode(output_logits, links, output_length, features, strategy, decode_beta, viterbibeta, upsample, pad):
    """Line-by-line transcription of s2s_conformer_dag_fastspeech2.py:213-304 (returns python lists)."""
    links = dense_links(links).clone()
    logits_n = output_logits.log_softmax(dim=-1)
    unreduced_logits, unreduced_tokens = logits_n.max(dim=-1)
    unreduced_tokens = unreduced_tokens.tolist()
    toks, feats = [], []
    if strategy in ("lookahead", "greedy"):
        lengths = output_length.tolist()
        if strategy == "lookahead":
            links_idx = (links + unreduced_logits.unsqueeze(1) * decode_beta).max(dim=-1)[1].cpu().tolist()
        else:
            links_idx = links.max(dim=-1)[1].cpu().tolist()
        for i, length in enumerate(lengths):
            last = unreduced_tokens[i][0]
            j = 0
            res, rf = [last], []
            steps = 0
            while j != length - 1 and steps < links.shape[1]:
                j = links_idx[i][j]
                now = unreduced_tokens[i][j]
                if now != pad and now != last:
                    res.append(now)
                    rf.append(j)
                last = now
                steps += 1
                if j == 0:
                    break
            toks.append(res)
            feats.append(rf)
    else:
        scores, indexs = [], []
        alpha_t = links[:, 0].clone()
        if strategy == "jointviterbi":
            alpha_t += unreduced_logits[:, 0].unsqueeze(1) * decode_beta
        batch_size, graph_length, _ = links.size()
        alpha_t += unreduced_logits * decode_beta
        scores.append(alpha_t)
        max_length = int(graph_length / 8 / upsample)
        for _ in range(max_length - 1):
            alpha_t, index = torch.max(alpha_t.unsqueeze(-1) + links, dim=1)
            if strategy == "jointviterbi":
                alpha_t += unreduced_logits * decode_beta
            scores.append(alpha_t)
            indexs.append(index)
        indexs = torch.stack(indexs, dim=0) if indexs else torch.zeros(0, batch_size, graph_length, dtype=torch.long)
        scores = torch.stack(scores, dim=0)
        link_last = torch.gather(links, -1, (output_length - 1).view(batch_size, 1, 1).repeat(1, graph_length, 1)).view(1, batch_size, graph_length)
        scores = scores + link_last
        scores, max_idx = torch.max(scores, dim=-1)
        lengths = torch.arange(max_length).unsqueeze(-1).repeat(1, batch_size) + 1
        length_penalty = (lengths ** viterbibeta).to(scores.device)
        scores = scores / length_penalty
        _, pred_length = torch.max(scores, dim=0)
        pred_length = pred_length + 1
        initial_idx = torch.gather(max_idx, 0, (pred_length - 1).view(1, batch_size)).view(batch_size).tolist()
        indexs = indexs.tolist()
        pred_length = pred_length.tolist()
        for i, length in enumerate(pred_length):
            j = initial_idx[i]
            last = unreduced_tokens[i][j]
            res, rf = [last], [j]
            for k in range(length - 1):
                j = indexs[length - k - 2][i][j]
                now = unreduced_tokens[i][j]
                if now != pad and now != last:
                    res.insert(0, now)
                    rf.insert(0, j)
                last = now
            toks.append(res)
            feats.append(rf)
    return toks, feats


def make_case(B, L, V, T, seed, peaked):
    rng = np.random.default_rng(seed)
    _, links, olen, _ = oracle.make_lattice(B, L, 4, T, seed=seed, ragged=True)
    logits = rng.standard_normal((B, L, V)).astype(np.float32) * (4.0 if peaked else 1.0)
    # repeated tokens on neighbouring vertices and some pad predictions exercise the de-duplication
    for b in range(B):
        for j in range(1, L, 3):
            logits[b, j] = logits[b, j - 1]
        logits[b, 5::7, PAD] += 20.0
    feats = rng.standard_normal((B, L, 8)).astype(np.float32)
    return torch.tensor(logits).cuda(), torch.tensor(links).cuda(), torch.tensor(olen).cuda(), torch.tensor(feats).cuda()


@pytest.mark.parametrize("strategy", ["greedy", "lookahead", "viterbi", "jointviterbi"])
@pytest.mark.parametrize("shape", [(3, 64, 12, 63, False), (2, 200, 30, 199, True), (2, 96, 9, 16, False), (1, 320, 50, 319, True)])
def test_decode_matches_the_reference_walks(strategy, shape):
    from daspeech_b200.decode import dag_decode
    B, L, V, T, peaked = shape
    logits, links, olen, feats = make_case(B, L, V, T, seed=L + V, peaked=peaked)
    ref_t, ref_f = reference_decode(logits, links, olen, feats, strategy, 1.0, 1.0, 0.5, PAD)
    tokens, fout, flen = dag_decode(logits, links, olen, feats, strategy=strategy, decode_beta=1.0, decode_viterbibeta=1.0,
                                    src_upsample_scale=0.5, pad=PAD)
    tokens = tokens.cpu().tolist()
    for b in range(B):
        n = len(ref_t[b])
        assert tokens[b][:n] == ref_t[b], (b, tokens[b][:n + 2], ref_t[b])
        assert all(x == PAD for x in tokens[b][n:])
        assert int(flen[b]) == len(ref_f[b])
        if ref_f[b]:
            want = feats[b, torch.tensor(ref_f[b], device="cuda")]
            assert torch.equal(fout[b, :len(ref_f[b])], want)
        assert not fout[b, len(ref_f[b]):].any()
