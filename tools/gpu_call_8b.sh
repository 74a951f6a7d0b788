#!/bin/bash
set -u
OUT=gpurun_out; TAG=${1:-r2g8b}; N=${2:-8}; mkdir -p $OUT
run() {  # name, env...
  NAME=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --no-cfg3 --dist-timeline > $OUT/${TAG}_$NAME.json 2> $OUT/${TAG}_$NAME.err; echo "bench $NAME rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_$NAME.json").read().strip().splitlines()[-1])
    print("$NAME: value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]))
    for k,v in d["dist_timeline"].items(): print("   ", k, v)
except Exception as e: print("parse failed", e); print(open("$OUT/${TAG}_$NAME.err").read()[-1500:])
PY
}
run default FACL_X=0
run maxctas4 NCCL_MAX_CTAS=4
run maxctas16 NCCL_MAX_CTAS=16
