#!/bin/bash
set -u
OUT=gpurun_out; TAG=${1:-r2c3}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -rf > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" $OUT/${TAG}_tests.log | tail -3
grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests.log | head -20
bash tools/role_profile.sh $TAG 2>&1 | grep -v "pass D mma warp" | head -40
for PREC in fp32 bf16 bf16_fast; do
timeout 600 python bench.py --precision $PREC --steps 20 --warmup 5 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_bench_$PREC.json 2> $OUT/${TAG}_bench_$PREC.err; echo "bench $PREC rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_$PREC.json").read().strip().splitlines()[-1])
    print("$PREC value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["launches_per_step"], "loss", d["config"]["loss_at_end"])
    for k in d["kernels"][:12]: print(" ", k["name"], round(k["ms_per_step"],3))
    print("  loss tags:", {k["name"]: round(k["ms_per_step"],3) for k in d["kernels"] if k["name"].startswith("loss")})
except Exception as e: print("bench parse failed", e); print(open("$OUT/${TAG}_bench_$PREC.err").read()[-2000:])
PY
done
