"""The copy-engine gradient exchange (daspeech_b200/csrc/xchg.cu, dist.PeerGradExchange) against the arithmetic it
stands for (fairseq legacy_distributed_data_parallel.py:76-165: buffer / world, summed over ranks): two processes, one
per GPU when the box has two, else both on cuda:0 (IPC mapping and the flag barrier work the same on one device)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fill(numel, rank, it):
    g = torch.Generator().manual_seed(1000 * it + rank)
    return torch.randn(numel, generator=g, dtype=torch.float32) * (1.0 + rank)


def _worker(rank, ws, port, ndev, numel, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    from daspeech_b200.dist import PeerGradExchange
    ex = PeerGradExchange(numel, device=dev)
    worst, identical = 0.0, True
    for it in range(6):
        ex.buffer[:numel].copy_(_fill(numel, rank, it))
        ex.start()
        got = ex.finish()[:numel].cpu()            # .cpu() synchronises the current stream, which waited for the exchange
        want = sum(_fill(numel, r, it).double() for r in range(ws)) / ws
        worst = max(worst, float((got.double() - want).abs().max() / want.abs().max()))
        every = [torch.empty_like(got) for _ in range(ws)]
        dist.all_gather(every, got)
        identical = identical and all(torch.equal(every[0], e) for e in every)
    timed_out = ex.timed_out_epoch()
    dist.barrier()
    ex.close()
    if rank == 0:
        torch.save({"worst": worst, "identical": identical, "timed_out": timed_out}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("numel", [1_000_003, 4096, 5])
def test_peer_exchange_is_the_mean_over_ranks(tmp_path, numel):
    ndev = torch.cuda.device_count()
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), ndev, numel, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["timed_out"] == 0
    assert r["identical"], "ranks must end with bit-identical buffers"
    assert r["worst"] < 1e-6


def _nvls_worker(rank, ws, port, numel, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
    from daspeech_b200.dist import NvlsGradExchange
    res = {"skipped": None}
    try:
        ex = NvlsGradExchange(numel, device=dev)
    except Exception as e:   # no multicast support on this box
        res["skipped"] = repr(e)[:200]
        ex = None
    if ex is not None:
        worst, identical = 0.0, True
        for it in range(4):
            ex.buffer[:numel].copy_(_fill(numel, rank, it))
            ex.start()
            got = ex.finish()[:numel].clone()
            want = sum(_fill(numel, r, it).double() for r in range(ws)) / ws
            worst = max(worst, float((got.double().cpu() - want).abs().max() / want.abs().max()))
            every = [torch.empty_like(got) for _ in range(ws)]
            dist.all_gather(every, got)
            identical = identical and all(torch.equal(every[0], e) for e in every)
        res.update(worst=worst, identical=identical)
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("numel", [1_000_004, 4096])
def test_nvls_exchange_is_the_mean_over_ranks(tmp_path, numel):
    if torch.cuda.device_count() < 2:
        pytest.skip("the in-switch exchange needs two GPUs of one NVSwitch domain")
    out = str(tmp_path / "r.pt")
    mp.spawn(_nvls_worker, args=(2, _free_port(), numel, out), nprocs=2, join=True)
    r = torch.load(out)
    if r["skipped"]:
        pytest.skip(r["skipped"])
    assert r["identical"], "ranks must end with bit-identical buffers"
    assert r["worst"] < 1e-6
