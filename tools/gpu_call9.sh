#!/bin/bash
set -u
OUT=gpurun_out; TAG=${1:-r2c9}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -rf > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" $OUT/${TAG}_tests.log | tail -3
grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests.log | head -20; grep -E "^E  " $OUT/${TAG}_tests.log | head -30
