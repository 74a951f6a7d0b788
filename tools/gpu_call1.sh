#!/bin/bash
# round 2, call 1: parity tests of the new build, bf16 per-layer diagnostics, bench (fp32 default + bf16), launch list, ncu of bf16 mode
set -u
OUT=gpurun_out
TAG=${1:-r2c1}
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2; nproc; free -g | sed -n 2p
timeout 1200 python -m pytest tests -m gpu -q -rf -s > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" $OUT/${TAG}_tests.log | tail -3
grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests.log | head -20
timeout 600 python tools/bf16_layers.py > $OUT/${TAG}_bf16_layers.log 2>&1; echo "bf16 diag rc=$?"; grep "split kept" $OUT/${TAG}_bf16_layers.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["launches_per_step"])
    print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "cpu", d["cpu_baseline"])
    print("api_path", d["api_path"]); print("cfg3", d["cfg3_strong"])
    for k in d["kernels"][:14]: print(" ", k["name"], round(k["ms_per_step"],3))
except Exception as e: print("bench parse failed", e)
PY
timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err; echo "bench bf16 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_bf16.json").read().strip().splitlines()[-1])
    print("bf16 value", d["value"], "ms", d["ms_per_step"])
    for k in d["kernels"][:10]: print(" ", k["name"], round(k["ms_per_step"],3))
except Exception as e: print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err; echo "reference arm rc=$?"; cut -c1-400 $OUT/${TAG}_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:l1_bwd_c_kernel|l1_bwd_d_kernel|l1_fwd_kernel" -c 4 -f -o $OUT/${TAG}_bf16_hot \
    python bench.py --precision bf16 --steps 1 --warmup 1 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_ncu_bf16.log 2>&1; echo "ncu bf16 rc=$?"
ls -la $OUT | tail -12
