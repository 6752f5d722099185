#!/bin/bash
# A/B of the gradient-exchange settings at N GPUs: tools/scale_ab.sh N  (run on the GPU box, from the repo root)
N=${1:-2}
run() {
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 295$((RANDOM % 90 + 10)) bench.py --gpus $N --steps 20 --warmup 5 --no-parts --no-cpu 2> gpurun_out/ab_$tag.err \
      | grep '^{' > gpurun_out/ab_$tag.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_$tag.json")); c = d["collective"]
    print("$tag", "N=%d" % d["n_gpus"], "step %.3f" % d["ms_per_step"], "exposed %.3f" % c["step_ms_exposed"],
          "none %.3f" % c["step_ms_no_collective"], "ar %.3f" % c["collective_ms"], "busbw %.0f" % c["busbw_gbs"], "nccl", c.get("nccl_comparison"), "timeouts", c.get("barrier_timeouts"))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/ab_$tag.err").read()[-1500:])
PY
}

run peer DAGB200_EXCHANGE=peer



