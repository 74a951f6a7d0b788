"""GPU: per-kernel timing of a training-mode forward at cfg2 size (fused vs per-layer L1)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facl_b200 import _lib, synth, utils_my
from facl_b200.train import TrainStep, default_opt
import bench

B, G, N = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 20, 2048
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
opt = default_opt(batchSize=B, SAMPLE_NUM=N)
tr = TrainStep(opt, num_crop=G, precision=prec, radius2=0.16)
pts = torch.from_numpy(synth.make_sequences(B, G, N, seed=1)).cuda()
data1 = pts.permute(1, 0, 2, 3).reshape(-1, N, 4).contiguous()
L = _lib.lib()
for fused in (True, False):
    tr.netR.fused_l1 = fused
    for it in range(4):
        if it == 2:
            torch.cuda.synchronize(); L.facl_timing_enable(1)
        with torch.no_grad():
            xt, yt = tr.group(data1)
            x, _, _, xg = tr.netR(xt, yt, 1)
    torch.cuda.synchronize(); L.facl_timing_enable(0)
    n = 43
    tms, tc = (C.c_float * n)(), (C.c_int * n)()
    L.facl_timing_collect(tms, tc, n)
    rows = sorted(((tms[i] / 2, bench.tag_name(i)) for i in range(n) if tc[i]), reverse=True)
    print(f"fused_l1={fused} prec={prec}: total kernel ms/forward = {sum(r[0] for r in rows):.3f}")
    for ms, name in rows[:10]:
        print(f"   {name:22s} {ms:8.3f} ms")
