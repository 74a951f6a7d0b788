#!/bin/bash
# ncu --set full of the fused net3DV_1 kernels in the bf16 (mixed) mode of configs[2]
set -u
OUT=gpurun_out; TAG=${1:-r2ncubf}; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:l1_bwd_c_kernel|l1_bwd_d_kernel|l1_fwd_kernel" -c 4 -f -o $OUT/${TAG}_hot \
    python bench.py --precision bf16 --steps 1 --warmup 1 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_ncu_hot.log 2>&1; echo "ncu rc=$?"
ls -la $OUT/${TAG}_hot.ncu-rep
