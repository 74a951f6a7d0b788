#!/bin/bash
set -u
OUT=gpurun_out; TAG=${1:-r2d2}; N=${2:-2}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > $OUT/${TAG}_dist_test.log 2>&1; echo "dist test rc=$?"; tail -2 $OUT/${TAG}_dist_test.log
for AHEAD in 2 0; do
FACL_MAX_AHEAD=$AHEAD timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 30 --warmup 5 --no-cfg3 --dist-timeline > $OUT/${TAG}_ahead$AHEAD.json 2> $OUT/${TAG}_ahead$AHEAD.err; echo "bench ahead=$AHEAD rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_ahead$AHEAD.json").read().strip().splitlines()[-1])
    print("ahead=$AHEAD: value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],3), "host issue ms", round(d["host_issue_ms_per_step"],3))
    for k,v in d["dist_timeline"].items(): print("   ", k, v)
except Exception as e: print("parse failed", e); print(open("$OUT/${TAG}_ahead$AHEAD.err").read()[-1500:])
PY
done
