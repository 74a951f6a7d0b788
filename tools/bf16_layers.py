#!/usr/bin/env python
"""Which layers cost the bf16-mode error?  Runs the training-mode forward of the CUDA encoder in bf16 mode with every
subset of {net3DV_1 layers 1-2 (fused), net3DV_3 layers 3, 4, 5} kept on the bf16x3 split products and prints the error of
x / x_global against the fp64 oracle, at the fixture size and at 8 x 20 x 2048.  (GPU box; diagnostic, not a test.)"""
import itertools
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle                                                     # noqa: E402
from facl_b200 import cn3d_model_conbag as MODELL, synth, utils_my  # noqa: E402
from facl_b200.train import default_opt                           # noqa: E402


def rel2(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm())


def run(name, pts, sd0, S, K, r2):
    B, G, N, _ = pts.shape
    sd64 = {k: (v.clone().double() if v.dtype.is_floating_point else v.clone()) for k, v in sd0.items()}
    clouds = pts.permute(1, 0, 2, 3).reshape(-1, N, 4).float()
    t0 = time.time()
    with torch.no_grad():
        xt, yt, _ = oracle.group_points(clouds, S, K, r2)
        ox, _, _, oxg = oracle.encoder_forward(oracle.EncoderParams(sd64, training=True), xt.double(), yt.double(), gost=G)
    print(f"[{name}] oracle fp64 forward {time.time() - t0:.1f} s", flush=True)
    dev_clouds = clouds.cuda()
    groups = {"L12": (1, 2), "L3": (3,), "L4": (4,), "L5": (5,)}
    rows = []
    for r in range(len(groups) + 1):
        for combo in itertools.combinations(groups, r):
            layers = tuple(l for g in combo for l in groups[g])
            opt = default_opt(batchSize=B, SAMPLE_NUM=N, sample_num_level1=S, knn_K=K)
            net = MODELL.PointNet_Plus_fine(opt, gost=G, sample_num_level1=S, knn_K=K)
            net.load_state_dict({k: v.clone() for k, v in sd0.items()})
            net = net.cuda()
            net.precision = "bf16_fast"
            net.bf16_split_layers = layers
            net.train()
            with torch.no_grad():
                gx, gy = utils_my._group(dev_clouds, S, K, r2)
                x, _, _, xg = net(gx, gy, 1)
            rows.append(("+".join(combo) or "none", rel2(x, ox), rel2(xg, oxg)))
            del net
    opt = default_opt(batchSize=B, SAMPLE_NUM=N, sample_num_level1=S, knn_K=K)
    net = MODELL.PointNet_Plus_fine(opt, gost=G, sample_num_level1=S, knn_K=K)
    net.load_state_dict({k: v.clone() for k, v in sd0.items()})
    net = net.cuda()
    net.train()
    with torch.no_grad():
        gx, gy = utils_my._group(dev_clouds, S, K, r2)
        x, _, _, xg = net(gx, gy, 1)
    rows.append(("fp32 mode", rel2(x, ox), rel2(xg, oxg)))
    for n, e, eg in rows:
        print(f"[{name}] split kept on {n:16s} x err {e:.3e}   x_global err {eg:.3e}", flush=True)


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "train_step.npz"))
    sd = oracle.init_state_dict(seed=int(z["seed_sd"]))
    for k in list(sd):
        if "sd0/" + k in z.files:
            sd[k] = torch.from_numpy(z["sd0/" + k]).clone()
    B, G, N, S, K = (int(v) for v in z["cfg"])
    run("fixture 4x3x128", torch.from_numpy(z["points"]), sd, S, K, float(z["r2"]))
    Bb = int(os.environ.get("BF16_DIAG_B", "8"))
    pts = torch.from_numpy(synth.make_sequences(Bb, 20, 2048, seed=31))
    run(f"synthetic {Bb}x20x2048", pts, oracle.init_state_dict(seed=12), 64, 64, 0.16)


if __name__ == "__main__":
    main()
