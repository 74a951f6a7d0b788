#!/bin/bash
set -u
OUT=gpurun_out; TAG=${1:-r2roles}; mkdir -p $OUT
FACL_LIB_PATH=$PWD/facl_b200/libfacl_b200_prof.so timeout 300 python bench.py --precision fp32 --steps 2 --warmup 1 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_roles_fp32.log 2>&1
grep -E "^fwd pass B MMA|pass_b=1" $OUT/${TAG}_roles_fp32.log | tail -12
