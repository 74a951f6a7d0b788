#!/bin/bash
# final 1-GPU validation: whole GPU suite, smoke, default bench line, bf16 bench line
set -u
OUT=gpurun_out; TAG=${1:-r2fin}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rf > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --precision bf16 --no-cpu-baseline > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err; echo "bench bf16 rc=$?"
timeout 600 python bench.py --precision bf16_fast --no-cpu-baseline --no-cfg3 --no-api-path > $OUT/${TAG}_bench_bf16_fast.json 2> $OUT/${TAG}_bench_bf16_fast.err; echo "bench bf16_fast rc=$?"
python - <<PY
import json
for n in ("bench","bench_bf16","bench_bf16_fast"):
    try:
        d=json.loads(open("$OUT/${TAG}_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches", d.get("gpu_launches"), "roof", round(d["roofline"]["frac"],4))
        print("  api", d.get("api_path") and (round(d["api_path"]["value"]), d["api_path"].get("prefetched",{}).get("value")), "cfg3", d.get("cfg3_strong") and round(d["cfg3_strong"]["value"]), "cpu", d.get("cpu_baseline") and d["cpu_baseline"]["value"])
        print("  ", {k["name"]: round(k["ms_per_step"],3) for k in d["kernels"][:8]})
    except Exception as e: print(n, "parse failed", e)
PY
