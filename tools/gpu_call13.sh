#!/bin/bash
set -u
OUT=gpurun_out; TAG=${1:-r2c13}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -rf -s > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" $OUT/${TAG}_tests.log | tail -3
grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests.log | head -20
grep -B2 -A30 "test_fused_l1_backward_bf16_8x20x2048" $OUT/${TAG}_tests.log | grep "fused, matched" | head -24
for PREC in bf16; do
timeout 600 python bench.py --precision $PREC --steps 30 --warmup 5 --no-cpu-baseline --no-api-path > $OUT/${TAG}_bench_$PREC.json 2> $OUT/${TAG}_bench_$PREC.err; echo "bench $PREC rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_$PREC.json").read().strip().splitlines()[-1])
    print("$PREC value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "cfg3", d["cfg3_strong"]["value"], d["cfg3_strong"]["ms_per_step"])
    for k in d["kernels"][:8]: print(" ", k["name"], round(k["ms_per_step"],3))
except Exception as e: print("bench parse failed", e); print(open("$OUT/${TAG}_bench_$PREC.err").read()[-2000:])
PY
done
