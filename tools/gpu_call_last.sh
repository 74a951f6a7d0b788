#!/bin/bash
# last check of a round: whole GPU suite, smoke, default bench line
set -u
OUT=gpurun_out; TAG=${1:-r2last}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -rf > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/${TAG}_tests.log
grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests.log | head
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 300 python bench.py --no-cfg3 --no-api-path > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k["name"]: round(k["ms_per_step"],3) for k in d["kernels"][:5]})
PY
