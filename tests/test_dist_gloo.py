"""CPU, world_size 2 over gloo: utterance sharding + scalar all-reduce give the single-process result.
(The DP itself has no multi-GPU coupling; the lattice function used here is the device-agnostic torch version.)"""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, B, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    torch.set_num_threads(1)
    ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
    from daspeech_b200 import dist as ddist
    match, links, olen, tlen = oracle.make_lattice(B, 24, 7, 23, seed=11, ragged=True)
    links[B - 1, 0, :] = -np.inf  # last utterance infeasible -> masked, counted
    dense = torch.tensor(oracle.dense_links(links))
    mean, local, stats = ddist.dag_nll_sharded(ops.torch_dag_loss, torch.tensor(match), dense,
                                               torch.tensor(olen), torch.tensor(tlen))
    lo, hi = ddist.shard_range(B, rank, ws)
    assert local.shape[0] == hi - lo
    t = ddist.max_over_ranks(float(rank + 1))
    # the trainer's flat gradient exchange: pre-divided by the world size, summed -> the mean over ranks
    ar = ddist.FlatGradAllReduce(1000, torch.float32)
    ar.buffer.copy_(torch.arange(1000, dtype=torch.float32) * (rank + 1))
    ar.start()
    gmean = ar.finish().clone()
    if rank == 0:
        torch.save({"mean": mean, "stats": stats, "tmax": t, "gmean": gmean}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from daspeech_b200 import dist as ddist
    for n in (0, 1, 5, 64, 67):
        for ws in (1, 2, 3, 8):
            spans = [ddist.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_two_rank_gloo_matches_single_process(tmp_path):
    B, ws = 5, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(ws, _free_port(), B, out), nprocs=ws, join=True)
    got = torch.load(out)
    ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
    match, links, olen, tlen = oracle.make_lattice(B, 24, 7, 23, seed=11, ragged=True)
    links[B - 1, 0, :] = -np.inf
    loss = ops.torch_dag_loss(torch.tensor(match), torch.tensor(oracle.dense_links(links)),
                              torch.tensor(olen), torch.tensor(tlen))
    invalid = loss.isinf() | loss.isnan()
    assert int(invalid.sum()) == 1
    ref = -(loss.masked_fill(invalid, 0) / torch.tensor(tlen)).mean()
    assert abs(float(got["mean"]) - float(ref)) < 1e-6
    assert int(got["stats"]["invalid_nsentences"]) == 1 and int(got["stats"]["nsentences"]) == B
    assert int(got["stats"]["ntokens"]) == int(tlen.sum())
    assert got["tmax"] == 2.0
    assert torch.allclose(got["gmean"], torch.arange(1000, dtype=torch.float32) * 1.5)
