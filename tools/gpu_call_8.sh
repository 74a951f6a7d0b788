#!/bin/bash
# 8-GPU call: weak-scaling bench line (with dist_loss_check + cfg3_strong) and BASELINE configs[4] extraction over 100k sequences
set -u
OUT=gpurun_out; TAG=${1:-r2g8}; N=${2:-8}; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "kernel ms", d["kernel_ms_per_step"])
    print("dist_loss_check", d["dist_loss_check"]); print("cfg3", d["cfg3_strong"]["value"], d["cfg3_strong"]["ms_per_step"])
    print({k["name"]: round(k["ms_per_step"],3) for k in d["kernels"] if k["name"].startswith("loss")})
except Exception as e: print("bench parse failed", e); print(open("$OUT/${TAG}_bench_n$N.err").read()[-1500:])
PY
for G in 10 20; do
SAVE=""; if [ $G -eq 10 ]; then SAVE="--save-dir /tmp/facl_feat"; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/extract_bench.py --sequences 100000 --views $G $SAVE > $OUT/${TAG}_extract_g$G.json 2> $OUT/${TAG}_extract_g$G.err; echo "extract G=$G rc=$?"
tail -1 $OUT/${TAG}_extract_g$G.json | cut -c1-1200; tail -3 $OUT/${TAG}_extract_g$G.err
done
ls /tmp/facl_feat/rank0 2>/dev/null | wc -l; du -sh /tmp/facl_feat 2>/dev/null
