#!/bin/bash
set -u
OUT=gpurun_out; TAG=${1:-r2c5}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -rf > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" $OUT/${TAG}_tests.log | tail -3
grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests.log | head -20
for PDL in 1 0; do
FACL_PDL=$PDL timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-cfg3 > $OUT/${TAG}_bench_pdl$PDL.json 2> $OUT/${TAG}_bench_pdl$PDL.err; echo "bench PDL=$PDL rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_pdl$PDL.json").read().strip().splitlines()[-1])
    print("PDL=$PDL value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["launches_per_step"], "kernel sum", d["kernel_ms_per_step"])
    print("  api_path", d["api_path"])
except Exception as e: print("bench parse failed", e); print(open("$OUT/${TAG}_bench_pdl$PDL.err").read()[-2000:])
PY
done
FACL_PDL=1 timeout 600 python bench.py --precision bf16_fast --steps 30 --warmup 5 --no-cpu-baseline --no-cfg3 --no-api-path 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16_fast PDL', d['value'], d['ms_per_step'])"
