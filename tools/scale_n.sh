#!/bin/bash
# tools/scale_n.sh N tag [ENV=VAL ...]: one short bench run at N GPUs, prints the collective summary
N=$1; tag=$2; shift 2
env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 295$((RANDOM % 90 + 10)) bench.py --gpus $N --steps 20 --warmup 5 --no-parts --no-cpu 2> gpurun_out/sc_$tag.err \
    | grep '^{' > gpurun_out/sc_$tag.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sc_$tag.json")); c = d["collective"]
    print("$tag", "N=%d" % d["n_gpus"], "step %.3f" % d["ms_per_step"], "exposed %.3f" % c["step_ms_exposed"],
          "none %.3f" % c["step_ms_no_collective"], "alone %.3f" % c["collective_ms"],
          {k: round(v, 3) for k, v in (c.get("phases_ms_standalone") or {}).items()},
          "nccl %.3f" % c["nccl_comparison"]["step_ms_overlapped"] if "nccl_comparison" in c else "", c["kind"][:40], "timeouts", c.get("barrier_timeouts"))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/sc_$tag.err").read()[-1500:])
PY
