#!/bin/bash
set -u
OUT=gpurun_out; TAG=${1:-r2c2}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -rf -s > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" $OUT/${TAG}_tests.log | tail -3
grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests.log | head -20
bash tools/role_profile.sh $TAG
timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err; echo "bench bf16 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_bf16.json").read().strip().splitlines()[-1])
    print("bf16 (compliant) value", d["value"], "ms", d["ms_per_step"])
    for k in d["kernels"][:10]: print(" ", k["name"], round(k["ms_per_step"],3))
except Exception as e: print("bench parse failed", e)
PY
