#!/bin/bash
# profile call: plain bench (sanity), ncu launch list, ncu --set full of the hot kernels (fp32 default mode), and of pass C/D in bf16_fast
set -u
OUT=gpurun_out; TAG=${1:-r2ncu}; mkdir -p $OUT
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "roof", d["roofline"]["frac"])
print("api", d["api_path"]["value"], d["api_path"]["prefetched"]["value"], "cfg3", d["cfg3_strong"]["value"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:l1_bwd_c_kernel|l1_bwd_d_kernel|l1_fwd_kernel|group_kernel" -c 5 -f -o $OUT/${TAG}_hot \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_ncu_hot.log 2>&1; echo "ncu hot rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:gemm_img_kernel|act_image_kernel" --launch-skip 12 --launch-count 14 -f -o $OUT/${TAG}_gemm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la $OUT | tail -8
