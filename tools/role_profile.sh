#!/bin/bash
# Per-role wait accounting of the fused net3DV_1 kernels (clock64, diagnostic library built by `make -C facl_b200/csrc prof`).
# usage: tools/role_profile.sh <tag>      -> gpurun_out/<tag>_roles_{fp32,bf16_fast}.log
set -u
TAG=${1:-roles}
OUT=gpurun_out
mkdir -p $OUT
for PREC in fp32 bf16_fast; do
  FACL_LIB_PATH=$PWD/facl_b200/libfacl_b200_prof.so timeout 300 python bench.py --precision $PREC --steps 2 --warmup 1 \
      --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_roles_$PREC.log 2>&1
  echo "== $PREC"; grep -E "^pass [CD]|^fwd pass" $OUT/${TAG}_roles_$PREC.log | sort | uniq -c | sort -rn | head -24
done
