#!/bin/bash
# One GPU-box call: parity tests, plain bench, ncu launch list, ncu --set full of the hot kernels.
# usage: tools/gpu_profile.sh <tag> [full-kernel-regex]   (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-r1}
KREGEX=${2:-"l1_bwd_c_kernel|l1_bwd_d_kernel|group_kernel|l1_fwd_kernel"}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"])
for k in d["kernels"]: print(k["name"], round(k["ms_per_step"],3))
print(d.get("cpu_baseline"))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -c 5 -f -o $OUT/${TAG}_hot \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_hot.log 2>&1; echo "ncu hot rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:gemm_img_kernel" --launch-skip 9 --launch-count 12 -f -o $OUT/${TAG}_gemm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la $OUT | tail -12
