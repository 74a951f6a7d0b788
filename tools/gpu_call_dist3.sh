#!/bin/bash
# 2-GPU call: the whole GPU suite (incl. the 2-rank NCCL test), smoke, bench at N=1 and N=2 (+ phase timeline)
set -u
OUT=gpurun_out; TAG=${1:-r2d3}; N=${2:-2}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -rf > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/${TAG}_tests.log
grep -E "^FAILED|Error" $OUT/${TAG}_tests.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench N=1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --dist-timeline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"; tail -5 $OUT/${TAG}_bench_n$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.loads(open("$OUT/${TAG}_bench_n%d.json" % n).read().strip().splitlines()[-1])
        print("N=%d value" % n, d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"))
        print("  dist_loss_check", d.get("dist_loss_check")); print("  cfg3", d.get("cfg3_strong")); print("  timeline", d.get("dist_timeline"))
        print("  cpu", d.get("cpu_baseline")); print("  roofline", d.get("roofline")); print("  clocks", d.get("clocks"))
    except Exception as e: print("bench parse failed", n, e)
PY
