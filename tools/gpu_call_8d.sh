#!/bin/bash
# 8-GPU call: weak-scaling bench with the phase timeline and the cfg3 strong-scaling leg
set -u
OUT=gpurun_out; TAG=${1:-r2g8d}; N=${2:-8}; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 30 --warmup 5 --dist-timeline > $OUT/${TAG}_n$N.json 2> $OUT/${TAG}_n$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_n$N.json").read().strip().splitlines()[-1])
    print("N=$N: value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "host issue ms", round(d["host_issue_ms_per_step"],3))
    print("dist_loss_check", d.get("dist_loss_check")); print("cfg3", d.get("cfg3_strong")); print("timeline", d.get("dist_timeline")); print("clocks", d.get("clocks"))
except Exception as e: print("parse failed", e); print(open("$OUT/${TAG}_n$N.err").read()[-1500:])
PY
