import sys, time, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facl_b200 import synth
from facl_b200.train import TrainStep, default_opt
B, G, N = 64, 20, 2048
tr = TrainStep(default_opt(batchSize=B, SAMPLE_NUM=N), num_crop=G, precision="fp32", radius2=0.16)
batch = torch.from_numpy(synth.make_sequences(B, G, N, seed=3)).cuda()
for _ in range(3): tr.step(batch)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): loss = tr.step(batch)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print({"op": "reference-shaped API step (group_points_3DV -> netR -> losses -> backward -> Adam via torch autograd)", "ms": ms, "sequences_per_s": B / ms * 1e3, "loss": float(loss)})
