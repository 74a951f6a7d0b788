#!/bin/bash
# role accounting of every fused net3DV_1 kernel with the `make prof` library (fp32 and bf16_fast): gpurun_out/<tag>_roles_*.log
set -u
OUT=gpurun_out; TAG=${1:-r2roles}; mkdir -p $OUT
for PREC in fp32 bf16_fast; do
  FACL_LIB_PATH=$PWD/facl_b200/libfacl_b200_prof.so timeout 200 python bench.py --precision $PREC --steps 2 --warmup 1 --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_roles_$PREC.log 2>&1
  echo "== $PREC"; grep -E "^pass [CD]|^fwd pass" $OUT/${TAG}_roles_$PREC.log | sort | uniq -c | sort -rn | awk '{$1="";print}' | awk '!seen[$1$2$3$4$5]++' | cut -c1-330
done
