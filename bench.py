#!/usr/bin/env python
"""bench.py -- DAG-loss hot path on B200: DP cells/s, HBM-roofline fraction, CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "C2"): dag_loss forward+backward on a synthetic lattice, B=64 utterances per
GPU, L=1024 graph vertices, M=256 target tokens, T=L-1 transitions, fp32, ragged lengths off (full lattice).
One "step" = one dag_loss forward (alpha+beta) + one dag_loss backward (grad_match+grad_links) over the batch,
inputs resident in HBM.  A DP cell is one (b, t, j) of the padded B x M x L lattice; value = cells/s over all ranks.
The step's working set (1.27 GB algorithmic) is ~10x the 126 MB L2, so no explicit L2 flush is needed.

Multi-GPU (torchrun, one rank per GPU): utterances are sharded (B per GPU, weak scaling), no collective inside the
DP; every step additionally carries the trainer's gradient exchange -- the mean over ranks of ONE 300 MB fp32 buffer
(fairseq legacy_distributed_data_parallel.py:76-165), issued on a side stream right after the backward kernels so that
it overlaps the next step's kernels (at most one exchange in flight; the step that follows waits for it before
starting its own).  The exchange runs on the copy engines over NVLink peer memory (dagb200_grad_exchange; NCCL with
DAGB200_EXCHANGE=nccl).  `collective` reports its stand-alone duration, the step time with the exchange fully exposed
and the same step with NCCL's all-reduce instead; `parts.no_collective` keeps the replica-only number.

Prints ONE JSON line (rank 0).  Extra objects: roofline (dominant kernel, HBM bound, measured peak from
MEASURED_PEAKS.json, plus the tensor-core roofline of the same kernel), cpu_baseline (CPU oracle = C port of the
reference arithmetic, all host threads, bounded sample), e2e (same metric through the public operator API with
pinned HOST buffers, H2D/D2H inside the timed region), clocks (NVML samples during the timed region), parts (other
kernels of the path, the banded T=32 and tuner-family shapes, and the UNMODIFIED reference CUDA kernels of
oracle/_ref timed on the same tensors in the same run; informational).
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dag_loss_fwd_bwd_dp_cells_per_sec"
UNIT = "cells/s"
C2 = dict(B=64, L=1024, M=256, V=4096)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=C2["B"], help="utterances per GPU")
    ap.add_argument("--prelen", type=int, default=C2["L"])
    ap.add_argument("--tarlen", type=int, default=C2["M"])
    ap.add_argument("--translen", type=int, default=0, help="0 = L-1")
    ap.add_argument("--vocab", type=int, default=C2["V"])
    ap.add_argument("--cpu-sample", type=int, default=8, help="utterances in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parts", action="store_true", help="skip the informational per-kernel parts")
    ap.add_argument("--grad-mb", type=float, default=300.0, help="size of the gradient all-reduce buffer (MB, fp32); N>1 only")
    return ap.parse_args()


def algorithmic_bytes(B, M, L, T):
    N, E = B * M * L, B * L * T
    return {"fwd": 4 * (3 * N + E), "bwd": 4 * (4 * N + 2 * E), "fwd_bwd": 4 * (7 * N + 3 * E),
            "viterbi": 4 * (N + E) + 4 * B * L}


def edge_relaxations(B, M, L, T):
    """R = sum_b sum_{t>=1} sum_{j=t}^{L-1} min(j, T) at full lengths."""
    r = 0
    for t in range(1, M):
        # sum_{j=t}^{L-1} min(j, T)
        lo = t
        if lo <= T:
            hi = min(T, L - 1)
            r += (lo + hi) * (hi - lo + 1) // 2
            r += T * max(0, L - 1 - hi)
        else:
            r += T * (L - lo)
    return r * B


def measured_peaks():
    """(HBM GB/s, bf16 TFLOP/s sustained, source).  The step is a long back-to-back kernel sequence: sustained figure."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), \
                "measured (MEASURED_PEAKS.json: hbm_gbs, bf16_tflops_sustained)"
        except Exception:
            pass
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """NVML samples of SM clock + throttle reasons while the timed region runs."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in {**self.BAD, **self.NOTE}.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(float(os.environ.get("BENCH_CLOCK_PERIOD", "0.02")))

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        if self.nv is not None and not self.samples:   # the region was shorter than one NVML query
            self._stop.clear()
            self._stop.set()
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            except Exception:
                pass

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------------
def make_inputs(torch, dev, B, L, M, T, V, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    match = torch.log(torch.rand(B, M, L, device=dev, generator=g) * 0.98 + 0.01)
    raw = torch.randn(B, L, T, device=dev, generator=g)
    i = torch.arange(L, device=dev).view(1, L, 1)
    k = torch.arange(T, device=dev).view(1, 1, T)
    valid = (i + k + 1) < L
    links = torch.log_softmax(raw.masked_fill(~valid, float("-inf")), -1).masked_fill(~valid, float("-inf"))
    del raw
    olen = torch.full((B,), L, dtype=torch.long, device=dev)
    tlen = torch.full((B,), M, dtype=torch.long, device=dev)
    go = torch.full((B,), 1.0, device=dev)
    return match.contiguous(), links.contiguous(), olen, tlen, go


def cpu_baseline(args, T, with_torch_path=True):
    """C port of the reference arithmetic (oracle/dag_oracle.c, OpenMP over all host threads) on the first
    `cpu_sample` utterances of the same workload; optionally also the torch restatement of torch_dag_loss."""
    import numpy as np
    from oracle import oracle
    oracle.set_num_threads(os.cpu_count() or 1)     # torchrun exports OMP_NUM_THREADS=1: size the pool ourselves
    Bs = max(1, min(args.cpu_sample, args.batch))
    L, M = args.prelen, args.tarlen
    match, links, olen, tlen = oracle.make_lattice(Bs, L, M, T, seed=1234, ragged=False)
    go = np.ones(Bs, np.float32)
    oracle.dag_loss(match[:1], links[:1], olen[:1], tlen[:1], True, np.float32)  # warm (page in, omp pool)
    t0 = time.perf_counter()
    _, a, b = oracle.dag_loss(match, links, olen, tlen, True, np.float32)
    oracle.dag_loss_backward(go, a, b, match, links, olen, tlen, np.float32)
    dt = time.perf_counter() - t0
    out = {"value": Bs * M * L / dt, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
           "sample": "first %d utterances of the workload (B=%d,L=%d,M=%d,T=%d), fp32 fwd+bwd, %.2f s; "
                     "C/OpenMP restatement of the reference arithmetic (oracle/dag_oracle.c)" % (Bs, Bs, L, M, T, dt),
           "seconds": dt, "host_cpus": os.cpu_count()}
    if with_torch_path:
        try:
            import torch
            ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
            m = torch.tensor(match[:1], requires_grad=True)
            lk = torch.tensor(links[:1])
            dense = torch.tensor(oracle.dense_links(links[:1])).requires_grad_()
            t0 = time.perf_counter()
            loss = ops.torch_dag_loss(m, dense, torch.tensor(olen[:1]), torch.tensor(tlen[:1]))
            torch.autograd.grad(loss.sum(), [m, dense])
            dt2 = time.perf_counter() - t0
            out["torch_path"] = {"value": M * L / dt2, "unit": UNIT, "threads": torch.get_num_threads(),
                                 "kind": "port (this repo's torch restatement of torch_dag_loss, not the reference file)",
                                 "sample": "1 utterance, dense-links torch restatement of torch_dag_loss + autograd "
                                           "(the reference's CPU algorithm, dag_loss.py:325-366), %.2f s" % dt2}
            del lk
        except Exception as e:  # pragma: no cover
            out["torch_path"] = {"error": str(e)[:200]}
    return out


def reference_cuda_times(torch, k, match, links, olen, tlen, go, B, L, M, V, timeit, n=10):
    """Times the reference's own CUDA kernels (the compiled, unmodified extension) as a baseline leg, like cpu_baseline."""
    import glob
    import importlib.util
    so = sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "dag_loss_fn*.so")))
    if not so:
        return {"unavailable": "oracle/_ref/dag_loss_fn.so not present (built from /root/reference by oracle/build_ref.py)"}
    try:
        spec = importlib.util.spec_from_file_location("dag_loss_fn", so[0])
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        dev = match.device
        out = {"iterations": n, "so": os.path.relpath(so[0], ROOT)}
        a0, b0 = ref.dag_loss(match, links, olen, tlen, True, 1)
        a1, b1 = k.dag_loss(match, links, olen, tlen, True, 1)
        out["loss_max_rel_diff"] = float(((b0[:, 0, 0] - b1[:, 0, 0]).abs() / b0[:, 0, 0].abs()).max())
        out["ref_fwd_ms"] = timeit(lambda: ref.dag_loss(match, links, olen, tlen, True, 1), n)
        out["new_fwd_ms"] = timeit(lambda: k.dag_loss(match, links, olen, tlen, True, 1), n)
        out["ref_bwd_ms"] = timeit(lambda: ref.dag_loss_backward(go, a0, b0, match, links, olen, tlen, 2, 2), n)
        out["new_bwd_ms"] = timeit(lambda: k.dag_loss_backward(go, a1, b1, match, links, olen, tlen, 2, 2), n)
        out["ref_viterbi_ms"] = timeit(lambda: ref.dag_best_alignment(match, links, olen, tlen, 1), n)
        out["new_viterbi_ms"] = timeit(lambda: k.dag_best_alignment(match, links, olen, tlen, 1, want_alpha=False), n)
        del a0, b0, a1, b1
        for name, dt in (("fp16", torch.float16), ("fp32", torch.float32)):
            x = (torch.randn(B, L, V, device=dev) * 2).to(dt)
            idx = torch.randint(4, V, (B, M), device=dev).unsqueeze(1).expand(-1, L, -1)
            out["ref_gather_%s_ms" % name] = timeit(lambda: ref.logsoftmax_gather(x, idx, True), n)
            out["new_gather_%s_ms" % name] = timeit(lambda: k.logsoftmax_gather(x, idx, True), n)
            del x, idx
        return out
    except Exception as e:  # pragma: no cover
        return {"error": str(e)[:300]}


def run_reference(args):
    """`--impl reference`: the CPU arm.  The reference's CPU implementation of this path is Python/torch that lives
    in /root/reference and cannot travel to the GPU box; what runs here is its C/OpenMP port (the oracle)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import oracle
    oracle.set_num_threads(os.cpu_count() or 1)     # torchrun exports OMP_NUM_THREADS=1: size the pool ourselves
    L, M = args.prelen, args.tarlen
    T = args.translen or (L - 1)
    Bs = max(1, min(args.cpu_sample, args.batch))
    match, links, olen, tlen = oracle.make_lattice(Bs, L, M, T, seed=1234, ragged=False)
    go = np.ones(Bs, np.float32)

    def step():
        _, a, b = oracle.dag_loss(match, links, olen, tlen, True, np.float32)
        oracle.dag_loss_backward(go, a, b, match, links, olen, tlen, np.float32)

    for _ in range(min(args.warmup, 1)):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = Bs * M * L / dt
    sample = ("each step = first %d utterances of the workload, fp32 fwd+bwd, %d timed steps (capped at 5), "
              "C/OpenMP port of the reference arithmetic" % (Bs, steps))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, T),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, T):
    return {"workload": "C2 dag_loss fwd+bwd: B=%d utterances/GPU, L=%d vertices, M=%d targets, T=%d transitions, "
                        "full lengths, fp32" % (args.batch, args.prelen, args.tarlen, T),
            "global_batch": args.batch * args.gpus, "prelen": args.prelen, "tarlen": args.tarlen, "translen": T,
            "vocab": args.vocab,
            "parallelism": ("dp1 (single GPU)" if args.gpus == 1 else
                            "dp%d: utterance-sharded, no collective inside the DP; one NCCL all-reduce of a %.0f MB fp32 "
                            "gradient buffer per step (side stream, overlaps the next step's kernels)" % (args.gpus, args.grad_mb)),
            "l2": "inputs larger than L2 (1.27 GB touched per step vs 126 MB L2); no flush"}


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: plain `python bench.py --gpus N` re-executes itself under torchrun (one rank per GPU)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   "--nproc-per-node=%d" % args.gpus, "--master-addr", "127.0.0.1",
                                   "--master-port", os.environ.get("MASTER_PORT", "29541"),
                                   os.path.abspath(__file__)] + sys.argv[1:])
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the communicator's kernels run on a high-priority stream with a bounded number of thread blocks: the
        # exchange is NVLink-bound, not SM-bound, and every block it holds is one the lattice kernels do not get
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=os.environ.get("DAGB200_NCCL_PRIO", "1") == "1")
        ctas = int(os.environ.get("DAGB200_NCCL_CTAS", "0"))
        if ctas > 0:
            opts.config.max_ctas = ctas
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    B, L, M, V = args.batch, args.prelen, args.tarlen, args.vocab
    T = args.translen or (L - 1)
    K, W = args.steps, max(args.warmup, 3)
    k = ops.get_dag_kernel()
    match, links, olen, tlen, go = make_inputs(torch, dev, B, L, M, T, V, 1234 + rank)
    bytes_ = algorithmic_bytes(B, M, L, T)
    peak_gbs, peak_tf, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the trainer's gradient exchange (N > 1 only): one flat fp32 buffer, pre-divided, all-reduced on a side stream
    from daspeech_b200.dist import FlatGradAllReduce, make_grad_exchange
    # the step's gradient exchange (DAGB200_EXCHANGE): "nvls" = reduced inside the NVSwitch by this library's multimem
    # kernel, "peer" = copy engines over NVLink peer memory, "nccl"; default "auto" = the first that can be set up on
    # this box.  NCCL's all-reduce is timed after the main region for comparison.
    grad_numel = int(args.grad_mb * 1e6 / 4)
    exchange_kind = os.environ.get("DAGB200_EXCHANGE", "auto")
    exchange = exchange_nccl = None
    if world > 1:
        exchange_nccl = FlatGradAllReduce(grad_numel, torch.float32, dev)
        if exchange_kind == "nccl":
            exchange = exchange_nccl
        else:
            exchange, exchange_kind = make_grad_exchange(grad_numel, dev, exchange_kind)
            if exchange_kind == "nccl":
                exchange = exchange_nccl

    def step():
        alpha, beta = k.dag_loss(match, links, olen, tlen, True, 1)
        gm, gl = k.dag_loss_backward(go, alpha, beta, match, links, olen, tlen, 2, 2)
        return alpha, beta, gm, gl

    trace = [] if os.environ.get("DAGB200_BENCH_TRACE") else None
    # The six kernels of a step are captured once into a CUDA graph and replayed (one driver call per step): on a shared
    # host the eager launch path (~0.1 ms per step when idle) has been seen at 0.6 ms and more, which makes the GPU wait
    # for launches.  DAGB200_BENCH_GRAPH=0 times the eager operator calls instead; the e2e leg always uses them.
    graphed = None
    if os.environ.get("DAGB200_BENCH_GRAPH", "1") == "1" and trace is None:
        from daspeech_b200.graphs import GraphedDagLossStep
        try:
            graphed = GraphedDagLossStep(match, links, olen, tlen, go)
        except Exception as e:   # noqa: BLE001 -- a box that cannot capture times the eager calls, and says so
            print("CUDA graph capture failed (%r): timing the eager operator calls" % (e,), file=sys.stderr)
            torch.cuda.synchronize()
            graphed = None

    def run_steps(n, overlap=True, with_exchange=True, exchange=exchange):
        """n steps; with the exchange of step i overlapping the kernels of step i+1 (overlap) or fully exposed."""
        out = None
        for _ in range(n):
            if trace is not None:
                tr = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
                trace.append(tr)
                tr[0].record()
            if graphed is not None:
                alpha, beta, gm, gl = graphed.replay()
            else:
                alpha, beta = k.dag_loss(match, links, olen, tlen, True, 1)
                if trace is not None:
                    tr[1].record()
                gm, gl = k.dag_loss_backward(go, alpha, beta, match, links, olen, tlen, 2, 2)
                if trace is not None:
                    tr[2].record()
            if exchange is not None and with_exchange:
                exchange.finish()          # the previous step's exchange (a no-op the first time)
                if trace is not None:
                    exchange.side.wait_stream(torch.cuda.current_stream())
                    tr[3].record(exchange.side)
                exchange.start()           # this step's gradients: side stream, after the backward kernels
                if trace is not None:
                    tr[4].record(exchange.side)
                if not overlap:
                    exchange.finish()
            out = (alpha, beta, gm, gl)
        if exchange is not None and with_exchange:
            exchange.finish()
        return out

    # warm-up with exactly the allocation pattern of the timed loop (same names kept alive), so that the caching
    # allocator is in steady state and no cudaMalloc (a device-wide sync) lands inside the timed region
    # settle phase (untimed, bounded): a fresh box pages the image, creates the context, loads the kernels lazily and
    # sizes the allocator pools during the first launches; run until two consecutive batches of steps take the same time
    # (within 5 %) with the host enqueueing faster than the GPU executes, or three seconds have passed, then do the W
    # warm-up steps proper.  (Seen on fresh boxes: the first CUDA process after a CPU-heavy one enqueues 10x slower for
    # a while -- 1.7 ms per step host-bound against 1.05 ms of kernels.)
    def settle(limit_s):
        t_settle, last = time.perf_counter(), None
        while time.perf_counter() - t_settle < limit_s:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run_steps(5)
            t_enq = time.perf_counter() - t0
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if last is not None and abs(dt - last) <= 0.05 * last and t_enq <= 0.6 * dt:
                break
            last = dt
    settle(3.0)
    alpha, beta, gm, gl = run_steps(W)
    barrier()
    # ---- timed region: exactly K steps -------------------------------------------------------------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    if trace is not None:
        del trace[:]
    ev[0].record()
    t_host0 = time.perf_counter()
    alpha, beta, gm, gl = run_steps(K)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / K
    ev[1].record()
    if trace is not None:
        torch.cuda.synchronize()
        rel = lambda t: ev[0].elapsed_time(t)
        ex_part = (lambda t: " | %.2f %.2f" % (rel(t[3]), rel(t[4]) - rel(t[3]))) if exchange is not None else (lambda t: "")
        print("rank %d (fwd start, fwd ms, bwd ms | exchange start, ms): %s | end %.3f" % (
            rank, "  ".join("%.2f %.2f %.2f%s" % (rel(t[0]), rel(t[1]) - rel(t[0]), rel(t[2]) - rel(t[1]), ex_part(t))
                            for t in trace),
            rel(ev[1])), file=sys.stderr, flush=True)
        trace = None
    # The K steps are now queued on the stream (the host enqueues a step in < 0.1 ms, the GPU needs ~1 ms for it).
    # Clocks / throttle reasons are sampled while the GPU works through them: NVML queries take the driver lock, so
    # sampling WHILE launching would stall the launches and show up as idle gaps between kernels.
    with ClockSampler(local) as clk:
        barrier()

    def max_ms(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_per_step = max_ms(ev[0].elapsed_time(ev[1])) / K
    # The K steps are enqueued ahead of the GPU (0.1 ms of host time per step against ~1 ms of kernels).  If the host
    # needed longer per step than 70 % of the measured step, the GPU was waiting for launches, not working: measure the K
    # steps once more after another settle phase and say so in the line (both numbers are reported).
    remeasured = None
    if max_ms(1.0 if host_enqueue_ms > 0.7 * ms_per_step else 0.0) > 0.5:
        first = {"ms_per_step": ms_per_step, "host_enqueue_ms_per_step": host_enqueue_ms}
        settle(3.0)
        run_steps(W)
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        t_host0 = time.perf_counter()
        alpha, beta, gm, gl = run_steps(K)
        host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / K
        ev[1].record()
        barrier()
        ms_per_step = max_ms(ev[0].elapsed_time(ev[1])) / K
        remeasured = {"first_attempt": first, "host_enqueue_ms_per_step": host_enqueue_ms,
                      "reason": "first attempt was launch-bound (host enqueue slower than the kernels); K steps timed again"}
    cells = B * M * L * world
    value = cells / (ms_per_step * 1e-3)
    loss_check = float(beta[:, 0, 0].float().mean().item())

    def timed(fn, n):
        barrier()
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(n)
        b2.record()
        barrier()
        return max_ms(a.elapsed_time(b2)) / n

    collective = None
    if exchange is not None:
        kc = max(3, min(K, 10))
        exposed_ms = timed(lambda n: run_steps(n, overlap=False), kc)
        replica_ms = timed(lambda n: run_steps(n, with_exchange=False), kc)

        def only_exchange(n, ex=exchange):
            for _ in range(n):
                ex.start()
                ex.finish()
        only_exchange(2)
        ar_ms = timed(only_exchange, kc)
        phases = None
        if hasattr(exchange, "phases_ms"):
            exchange.phases_ms()
            only_exchange(2)
            phases = dict(zip(("push_scatter", "barrier_b", "reduce", "push_gather", "barrier_c"),
                              exchange.phases_ms()))
        nbytes = exchange.buffer.numel() * 4
        busbw = lambda ms: nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9
        kinds = {"nvls": "mean over ranks of one flat fp32 gradient buffer reduced inside the NVSwitch: multimem.ld_reduce + "
                         "multimem.st on a multicast mapping, one kernel of 8-24 thread blocks per rank "
                         "(dagb200_grad_exchange_nvls, daspeech_b200/csrc/xchg.cu); mapping and barriers: torch symmetric memory",
                 "peer": "mean over ranks of one flat fp32 gradient buffer: reduce-scatter + all-gather as peer-to-peer "
                         "copy-engine transfers over NVLink, one short reduce kernel, flag barriers "
                         "(dagb200_grad_exchange, daspeech_b200/csrc/xchg.cu)",
                 "nccl": "nccl all_reduce (pre-multiplied sum) of one flat fp32 gradient buffer"}
        collective = {"kind": kinds.get(exchange_kind, exchange_kind), "bytes": nbytes, "collective_ms": ar_ms,
                      "busbw_gbs": busbw(ar_ms),
                      "step_ms_overlapped": ms_per_step, "step_ms_exposed": exposed_ms, "step_ms_no_collective": replica_ms,
                      "reference": "fairseq legacy_distributed_data_parallel.py:76-165 via trainer.py:928"}
        if phases:
            collective["phases_ms_standalone"] = phases
        if exchange is not exchange_nccl:
            run_steps(3, exchange=exchange_nccl)
            nccl_ms = timed(lambda n: run_steps(n, exchange=exchange_nccl), kc)
            only_exchange(2, exchange_nccl)
            nccl_ar = timed(lambda n: only_exchange(n, exchange_nccl), kc)
            collective["nccl_comparison"] = {"step_ms_overlapped": nccl_ms, "collective_ms": nccl_ar,
                                             "busbw_gbs": busbw(nccl_ar)}
            if hasattr(exchange, "timed_out_epoch"):
                collective["barrier_timeouts"] = exchange.timed_out_epoch()

    # ---- per-kernel durations: CUDA events recorded by the library on the launch stream around each of its
    # kernels (dagb200_set_profile), averaged over a few extra steps right after the timed region -----------
    import ctypes
    lib = k.lib
    prof = [0.0] * 5
    kp = max(3, min(K, 10))
    lib.dagb200_set_profile(1)
    buf = (ctypes.c_float * 5)()
    for _ in range(kp):
        step()
        lib.dagb200_get_profile(ctypes.cast(buf, ctypes.c_void_p), 5)
        for i in range(5):
            prof[i] += max(buf[i], 0.0) / kp
    lib.dagb200_set_profile(0)
    kern = {"dag_rowmax_kernel+dag_tiles_kernel": prof[0], "dag_alpha_beta_tcgen05_kernel": prof[1],
            "grad_planes4_kernel": prof[2], "grad_fmax_kernel+grad_links_tcgen05_kernel": prof[3]}
    # algorithmic bytes of the launch each kernel belongs to: the forward pair
    # (precompute + recurrences) moves 4(3N+E), the backward pair 4(4N+2E)
    N_, E_ = B * M * L, B * L * T
    GL = "grad_fmax_kernel+grad_links_tcgen05_kernel"
    kbytes = {"dag_alpha_beta_tcgen05_kernel": bytes_["fwd"], GL: 4 * (2 * N_ + 2 * E_),
              "dag_rowmax_kernel+dag_tiles_kernel": 4 * E_, "grad_planes4_kernel": 4 * 4 * N_}
    dom_name = max(("dag_alpha_beta_tcgen05_kernel", GL), key=lambda n: kern[n])
    fwd_ms = prof[0] + prof[1]
    bwd_ms = prof[2] + prof[3]
    dom_ms = kern[dom_name]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom_name)
    except Exception:
        pass
    ach = kbytes[dom_name] / (dom_ms * 1e-3) / 1e9
    # tensor-core work of the blocked kernels: every edge relaxation is one MAC, executed as three
    # bf16 MMAs (hi*hi, lo*hi, hi*lo); alpha and beta each relax R edges, grad_links contracts the same R products
    R = edge_relaxations(B, M, L, T)
    kflop = {"dag_alpha_beta_tcgen05_kernel": 2 * R * 3 * 2, GL: R * 3 * 2}
    compute = {"flop": kflop[dom_name], "achieved": kflop[dom_name] / (dom_ms * 1e-3) / 1e12, "peak": peak_tf,
               "unit": "TFLOP/s", "frac": kflop[dom_name] / (dom_ms * 1e-3) / 1e12 / peak_tf,
               "edge_relaxations_per_pass": R,
               "note": "algorithmic bf16x3 tensor flop of the kernel over the measured sustained bf16 peak; the kernel is "
                       "bound by neither roofline but by its serial fp64 chain and the single-thread MMA issue path"}
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                "traffic": traffic, "kernel": dom_name, "kernel_ms": dom_ms, "compute": compute,
                "algorithmic_bytes_per_launch": kbytes[dom_name], "peak_source": peak_src,
                "kernels_ms": kern,
                "step": {"algorithmic_bytes": bytes_["fwd_bwd"], "ms": ms_per_step,
                         "achieved": bytes_["fwd_bwd"] / (ms_per_step * 1e-3) / 1e9,
                         "frac": bytes_["fwd_bwd"] / (ms_per_step * 1e-3) / 1e9 / peak_gbs,
                         "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
                         "compute_frac": 3 * R * 3 * 2 / (ms_per_step * 1e-3) / 1e12 / peak_tf,
                         "note": "HBM roofline of the whole fwd+bwd step; the recurrences are tensor/issue bound "
                                 "at T=L-1 (DESIGN.md section 5)"}}

    # ---- e2e: public autograd API, pinned host inputs, H2D + D2H inside the timed region -------------
    from daspeech_b200.prefetch import bind_host_to_gpu
    numa_bound = bind_host_to_gpu(torch.cuda.current_device() if world == 1 else int(os.environ.get("LOCAL_RANK", 0)))
    h_match = match.cpu().pin_memory()
    h_links = links.cpu().pin_memory()
    h_olen, h_tlen = olen.cpu().pin_memory(), tlen.cpu().pin_memory()
    h_loss = torch.empty(B, dtype=torch.float32).pin_memory()
    h2d = h_match.numel() * 4 + h_links.numel() * 4 + 16 * B
    d2h = 4 * B

    from daspeech_b200.prefetch import DevicePrefetcher

    def host_batches(n):
        for _ in range(n):
            yield (h_match, h_links, h_olen, h_tlen)       # every step copies its inputs from pinned host memory

    def e2e_run(n):
        # the copies of step i+1 travel on a side stream while step i computes (daspeech_b200/prefetch.py)
        for m, lk, ol, tl in DevicePrefetcher(host_batches(n), dev):
            m.requires_grad_()
            lk.requires_grad_()
            loss = ops.dag_loss(m, lk, ol, tl)
            total = -(loss / tl).mean()
            total.backward()
            h_loss.copy_(loss.detach(), non_blocking=True)
            del m, lk, loss, total

    del alpha, beta, gm, gl
    e2e_k = max(3, min(K, 10))
    e2e_run(5)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(e2e_k)
    e1.record()
    barrier()
    e2e_ms = max_ms(e0.elapsed_time(e1)) / e2e_k
    e2e = {"value": cells / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_ms, "steps": e2e_k, "host_bound_to_gpu_numa_node": numa_bound,
           "h2d_gbs_per_rank": h2d / (e2e_ms * 1e-3) / 1e9,
           "limiter": "host->device copy of the step's inputs (%.0f MB per rank and step over PCIe; the kernels of a step "
                      "take %.2f ms and hide behind it)" % (h2d / 1e6, ms_per_step),
           "api": "daspeech_b200.dag_loss(match_all, links, output_length, target_length) + .backward(), every step's "
                  "inputs copied from pinned host memory (DevicePrefetcher: the copy of step i+1 overlaps step i), "
                  "per-utterance loss read back"}

    # ---- informational parts: the other kernels of the path (rank 0 only) ---------------------------
    parts = {}
    if rank == 0 and not args.no_parts:
        def timeit(fn, n=5):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b2.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b2) / n

        vit_ms = timeit(lambda: k.dag_best_alignment(match, links, olen, tlen, 1, want_alpha=False))
        parts["dag_best_alignment"] = {"ms": vit_ms, "algorithmic_bytes": bytes_["viterbi"],
                                       "gbs": bytes_["viterbi"] / vit_ms / 1e6}
        for name, dt, esz in (("fp16", torch.float16, 2), ("fp32", torch.float32, 4)):
            logits = (torch.randn(B, L, V, device=dev) * 2).to(dt)
            idx = torch.randint(4, V, (B, M), device=dev).unsqueeze(1).expand(-1, L, -1)
            gsel = torch.randn(B, M, L, device=dev).transpose(1, 2)
            ms_am = timeit(lambda: logits.argmax(-1))                                   # what the GLAT pass does first
            ms_fa = timeit(lambda: k.logsoftmax_gather(logits, idx, False, want_argmax=True))
            ms_f = timeit(lambda: k.logsoftmax_gather(logits, idx, True))
            by_f = 2 * esz * B * L * V + 4 * B * L * M + 8 * B * M
            ms_b = timeit(lambda: k.logsoftmax_gather_backward(logits, idx, gsel))
            by_b = 2 * esz * B * L * V + 4 * B * L * M
            parts["logsoftmax_gather_" + name] = {"fwd_ms": ms_f, "fwd_gbs": by_f / ms_f / 1e6,
                                                  "fwd_frac": by_f / ms_f / 1e6 / peak_gbs,
                                                  "bwd_ms": ms_b, "bwd_gbs": by_b / ms_b / 1e6,
                                                  "bwd_frac": by_b / ms_b / 1e6 / peak_gbs,
                                                  "glat_pass_gather_with_fused_argmax_ms": ms_fa,
                                                  "torch_argmax_alone_ms": ms_am}
            del logits, idx, gsel

        # next row of the path: the S2S criterion's alignment posterior, fused vs the criterion's torch ops
        from daspeech_b200.posterior import dag_posterior
        a2, b2_ = k.dag_loss(match, links, olen, tlen, True, 1)
        def torch_posterior():
            sc = (a2 + b2_ - ops.logsumexp_keepdim(a2 + b2_, -1)).exp()
            sc.masked_fill_(torch.isnan(sc), 0)
            return sc
        ms_t = timeit(torch_posterior)
        for name, dt, esz in (("fp32", torch.float32, 4), ("fp16", torch.float16, 2)):
            ms_p = timeit(lambda: dag_posterior(a2, b2_, dt))
            by_p = (8 + esz) * B * M * L
            parts["dag_posterior_" + name] = {"ms": ms_p, "algorithmic_bytes": by_p, "gbs": by_p / ms_p / 1e6,
                                              "frac": by_p / ms_p / 1e6 / peak_gbs, "torch_ops_ms": ms_t}
        del a2, b2_
        # next row: GLAT force-emit masking, fused vs the criterion's torch expression
        from daspeech_b200.glat import glat_force_emit
        mm = torch.zeros(B, M, L, dtype=torch.bool, device=dev)
        mm.scatter_(1, torch.randint(0, M, (B, 1, L), device=dev), True)
        keepm = torch.rand(B, L, device=dev) < 0.3
        prevm = keepm.unsqueeze(1)
        ms_g = timeit(lambda: glat_force_emit(match, mm, keepm))
        ms_gt = timeit(lambda: match.masked_fill(prevm, 0) + match.masked_fill(~mm, float("-inf")).masked_fill(~prevm, 0))
        by_g = 9 * B * M * L
        parts["glat_force_emit"] = {"ms": ms_g, "algorithmic_bytes": by_g, "gbs": by_g / ms_g / 1e6,
                                    "frac": by_g / ms_g / 1e6 / peak_gbs, "torch_ops_ms": ms_gt}
        del mm, keepm, prevm
        # next row: the transition log-probabilities from the link heads (H = 8 heads of 64 features,
        # the models' decoder width), fused tcgen05 forward vs the model's op sequence run by torch on 8 utterances
        try:
            from daspeech_b200 import links as dlinks
            Hh, Fh = 8, 64
            gq = torch.Generator(device=dev).manual_seed(7)
            qh = torch.randn(B, L, Hh, Fh, device=dev, generator=gq)
            kh = torch.randn(B, L, Hh, Fh, device=dev, generator=gq)
            lgh = torch.log_softmax(torch.randn(B, L, Hh, device=dev, generator=gq), -1)
            with torch.no_grad():
                ms_l = timeit(lambda: dlinks.extract_links_from_chunks(qh, kh, lgh, olen, T, fused=True))
                nb = min(B, 8)
                ms_lt = timeit(lambda: dlinks.torch_extract_links(qh[:nb], kh[:nb], lgh[:nb], olen[:nb], T), 3)
            parts["extract_links"] = {"shape": {"B": B, "L": L, "H": Hh, "F": Fh, "T": T}, "ms": ms_l,
                                      "torch_op_sequence_ms": ms_lt * B / nb,
                                      "note": "op sequence timed on %d utterances ([B,L,L,H] product: 2.1 GB per 8) and scaled" % nb}
            del qh, kh, lgh
        except Exception as e:   # noqa: BLE001 -- informational part
            parts["extract_links"] = {"error": repr(e)[:200]}

        # the banded secondary configuration (T = 32, the reference tuner's setting) and the tuner's shape family
        def fwd_bwd(B_, L_, M_, T_, n=10):
            m_, lk_, ol_, tl_, go_ = make_inputs(torch, dev, B_, L_, M_, T_, V, 99)

            def f():
                a_, b_ = k.dag_loss(m_, lk_, ol_, tl_, True, 1)
                k.dag_loss_backward(go_, a_, b_, m_, lk_, ol_, tl_, 2, 2)
            ms = timeit(f, n)
            by = algorithmic_bytes(B_, M_, L_, T_)["fwd_bwd"]
            vit = timeit(lambda: k.dag_best_alignment(m_, lk_, ol_, tl_, 1, want_alpha=False), n) if (M_ - 1) * T_ + 1 >= L_ else None
            return {"shape": {"B": B_, "L": L_, "M": M_, "T": T_}, "fwd_bwd_ms": ms, "cells_per_s": B_ * M_ * L_ / (ms * 1e-3),
                    "algorithmic_bytes": by, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak_gbs, "dag_best_alignment_ms": vit}
        parts["c2_T32"] = fwd_bwd(B, L, M, 32)
        parts["tuner_shape"] = fwd_bwd(81, 400, 50, 32)

        # the GPU reference bar: the UNMODIFIED reference CUDA kernels, compiled from
        # /root/reference by oracle/build_ref.py into oracle/_ref/dag_loss_fn.so, on the same tensors in the same run
        parts["reference_cuda"] = reference_cuda_times(torch, k, match, links, olen, tlen, go, B, L, M, V, timeit)

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(args, T),
                "utt_per_sec": B * world / (ms_per_step * 1e-3), "clocks": clk.summary(), "e2e": e2e,
                "gpu_launches": (6 + ({"peer": 3, "nvls": 1}.get(exchange_kind, 0) if world > 1 else 0)) * K, "roofline": roofline, "loss_check": loss_check, "parts": parts}
        line["host_enqueue_ms_per_step"] = host_enqueue_ms
        line["launch_path"] = ("cuda graph replay of the step's six kernels (daspeech_b200.graphs.GraphedDagLossStep)"
                               if graphed is not None else "eager operator calls")
        if remeasured is not None:
            line["remeasured"] = remeasured
        if collective is not None:
            line["collective"] = collective
            parts["no_collective"] = {"ms_per_step": collective["step_ms_no_collective"],
                                      "value": cells / (collective["step_ms_no_collective"] * 1e-3), "unit": UNIT}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args, T)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
