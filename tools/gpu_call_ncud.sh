#!/bin/bash
# ncu --set full (with source) of one fused kernel: usage tools/gpu_call_ncud.sh <tag> <kernel regex> [precision]
set -u
OUT=gpurun_out; TAG=${1:-r2ncud}; K=${2:-l1_bwd_d_kernel}; PREC=${3:-fp32}; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" -c 1 -f -o $OUT/${TAG} \
    python bench.py --precision $PREC --steps 1 --warmup 1 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $OUT/${TAG}.ncu-rep
