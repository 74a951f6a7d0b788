#!/bin/bash
# quick check of a net3DV_1 kernel change: encoder parity tests, role accounting, two bench lines
set -u
OUT=gpurun_out; TAG=${1:-r2l1}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -rf -x > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" $OUT/${TAG}_tests.log | tail -3
grep -E "^FAILED|^ERROR|^E  " $OUT/${TAG}_tests.log | head -20
bash tools/role_profile.sh $TAG 2>&1 | grep -E "==|pass D|fwd pass" | awk '!seen[$3$4$5$6]++' | head -16
for PREC in fp32 bf16_fast; do
timeout 600 python bench.py --precision $PREC --steps 30 --warmup 5 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_bench_$PREC.json 2> $OUT/${TAG}_bench_$PREC.err; echo "bench $PREC rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_$PREC.json").read().strip().splitlines()[-1])
    print("$PREC value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d.get("loss_at_end"))
    for k in d["kernels"][:6]: print(" ", k["name"], round(k["ms_per_step"],3))
except Exception as e: print("bench parse failed", e); print(open("$OUT/${TAG}_bench_$PREC.err").read()[-2000:])
PY
done
