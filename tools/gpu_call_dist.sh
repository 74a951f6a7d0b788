#!/bin/bash
# multi-GPU call: NCCL parity test + bench at N ranks (usage: tools/gpu_call_dist.sh <tag> <N>)
set -u
OUT=gpurun_out; TAG=${1:-r2d}; N=${2:-2}; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -rf -s > $OUT/${TAG}_dist_test.log 2>&1; echo "dist test rc=$?"; grep -E "passed|failed|error|skipped" $OUT/${TAG}_dist_test.log | tail -3
grep -E "2-rank NCCL" $OUT/${TAG}_dist_test.log | head -30
grep -E "Error|assert" $OUT/${TAG}_dist_test.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"; tail -5 $OUT/${TAG}_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "kernel ms", d["kernel_ms_per_step"])
    print("dist_loss_check", d["dist_loss_check"]); print("cfg3", d["cfg3_strong"])
    print({k["name"]: round(k["ms_per_step"],3) for k in d["kernels"] if k["name"].startswith("loss")})
except Exception as e: print("bench parse failed", e)
PY
