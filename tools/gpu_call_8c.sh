#!/bin/bash
set -u
OUT=gpurun_out; TAG=${1:-r2g8c}; N=${2:-8}; mkdir -p $OUT
for AHEAD in 2 0 1; do
FACL_MAX_AHEAD=$AHEAD timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 30 --warmup 5 --no-cfg3 > $OUT/${TAG}_ahead$AHEAD.json 2> $OUT/${TAG}_ahead$AHEAD.err; echo "bench ahead=$AHEAD rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_ahead$AHEAD.json").read().strip().splitlines()[-1])
    print("ahead=$AHEAD: value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],3), "host issue ms", round(d["host_issue_ms_per_step"],3))
except Exception as e: print("parse failed", e); print(open("$OUT/${TAG}_ahead$AHEAD.err").read()[-1500:])
PY
done
