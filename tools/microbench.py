#!/usr/bin/env python
"""BASELINE.json configs[3] / configs[4]: FPS + grouping sweep (timing and HBM roofline; parity of the same sweep against
the oracle is in tests/test_gpu_kernels.py) and forward-only feature extraction throughput.  One JSON object per line.

    python tools/microbench.py [--quick]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from facl_b200 import ops, synth
from facl_b200.train import TrainStep, default_opt, extract_features

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def clouds(V, N, seed):
    pts = synth.make_sequences(V, 1, N, seed=seed)              # (V,1,N,4)
    return torch.from_numpy(pts.reshape(V, N, 4)).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated subset of: fps,group,augment,group2,probe,extract")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))

    def want(name):
        return not only or name in only
    V = 256 if args.quick else 1280
    dev = torch.device("cuda")
    # ---------------- FPS (K1) ----------------
    for N in ([2048] if args.quick else [1024, 2048, 4096, 8192, 16384]) if want("fps") else []:
        pts = clouds(V, N, seed=N)
        start = torch.zeros(V, dtype=torch.int32, device=dev)
        for m in ([64] if args.quick else [64, 128, 512]):
            ms = timed(lambda: ops.fps(pts, m, start))
            nbytes = V * (12 * N + 4 * m)
            print(json.dumps(dict(op="fps", V=V, N=N, m=m, ms=ms, us_per_cloud=ms * 1e3 / V, ns_per_cloud_pick=ms * 1e6 / V / m,
                                  algorithmic_GBs=nbytes / ms / 1e6, hbm_frac=nbytes / ms / 1e6 / HBM)),
                  flush=True)
    # ---------------- grouping (K2) ----------------
    S = 64
    combos = [(2048, 64, 0.16)] if args.quick else \
        [(N, 64, 0.16) for N in (1024, 2048, 4096, 8192, 16384)] + [(2048, K, 0.16) for K in (16, 32, 128)] + \
        [(2048, 64, r2) for r2 in (0.0025, 0.01, 0.06)]
    for N, K, r2 in combos if want("group") else []:
        pts = clouds(V, N, seed=N + K)
        ms = timed(lambda: ops.group_points_raw(pts, S, K, r2, want_idx=False))
        rows, idx = ops.group_points_raw(pts[:2].contiguous(), S, K, r2)
        redirected = float((idx.cpu() == torch.arange(S)[None, :, None]).float().mean())
        nbytes = V * (16 * N + 16 * S * K + 12 * S)
        print(json.dumps(dict(op="group", M=V, N=N, S=S, K=K, r2=r2, ms=ms, us_per_cloud=ms * 1e3 / V, algorithmic_GBs=nbytes / ms / 1e6,
                              hbm_frac=nbytes / ms / 1e6 / HBM, frac_slots_at_centre=redirected)), flush=True)
    # ---------------- view augmentation ----------------
    from facl_b200 import cn3D_data_set as ds
    rng = np.random.default_rng(0)
    for B, N, P in ([(64, 512, 2048)] if args.quick else [(64, 512, 2048), (64, 2048, 2048), (256, 2048, 4096)]) if want("augment") else []:
        srcs = []
        for s in range(4):
            a = rng.uniform(-0.5, 0.5, (B * P, 8)).astype(np.float32)
            a[rng.uniform(size=B * P) < 0.5, 4] = 0
            a[rng.uniform(size=B * P) < 0.5, 7] = 0
            srcs.append(ds.Ragged(torch.from_numpy(a).cuda(), (torch.arange(B + 1, dtype=torch.int32) * P).cuda(), P))
        aug = ds.ViewAugmenter(num_point=N)
        ms = timed(lambda: aug.get_data_train(srcs))
        G = len(ds.GET_DATA_TRAIN)
        nbytes = B * G * N * (16 + 16) + 2 * B * P * 4      # gathered rows (xyz + channel) in, views out, temporal-channel scans
        print(json.dumps(dict(op="augment_views", B=B, G=G, N=N, P=P, rng="philox", ms=ms, sequences_per_s=B / ms * 1e3,
                              algorithmic_GBs=nbytes / ms / 1e6, hbm_frac=nbytes / ms / 1e6 / HBM)), flush=True)
    # ---------------- level-2 grouping ----------------
    for M, C, S1, S2, K in ([(256, 131, 512, 128, 64)] if args.quick else [(1280, 131, 512, 128, 64), (1280, 259, 64, 64, 32)]) if want("group2") else []:
        feats = torch.randn(M, C, S1, device=dev)
        feats[:, :3] = torch.rand(M, 3, S1, device=dev) - 0.5
        ms = timed(lambda: ops.group_level2(feats, S2, K, 0.02), iters=5)
        nbytes = M * (4 * C * S1 + 4 * C * S2 * K)
        print(json.dumps(dict(op="group_level2", M=M, C=C, S1=S1, S2=S2, K=K, ms=ms, us_per_cloud=ms * 1e3 / M,
                              algorithmic_GBs=nbytes / ms / 1e6, hbm_frac=nbytes / ms / 1e6 / HBM)), flush=True)
    # ---------------- linear probe step ----------------
    if want("probe"):
        from facl_b200 import linercls
        for rows in ([64] if args.quick else [64, 1024]):
            tr = linercls.ProbeTrainer()
            x = torch.randn(rows, 22 * 512, device=dev)
            y = torch.randint(0, 120, (rows,), device=dev)
            ms = timed(lambda: tr.step(x, y), iters=20)
            flops = 2 * 2 * rows * 22 * 512 * 120
            print(json.dumps(dict(op="probe_step", rows=rows, features=22 * 512, classes=120, ms=ms, samples_per_s=rows / ms * 1e3,
                                  algorithmic_TFLOPs=flops / ms / 1e9, loss=tr.pop_meters()[0] / 23)), flush=True)
    # ---------------- forward-only feature extraction (configs[4]) ----------------
    for G in ([10] if args.quick else [10, 20]) if want("extract") else []:
        B, N = 64, 2048
        opt = default_opt(batchSize=B, SAMPLE_NUM=N)
        tr = TrainStep(opt, num_crop=G, precision="fp32", radius2=0.16)
        batch = torch.from_numpy(synth.make_sequences(B, G, N, seed=7)).cuda()
        ms = timed(lambda: extract_features(tr.netR, opt, batch, radius2=0.16), iters=5, warm=2)
        feat = extract_features(tr.netR, opt, batch, radius2=0.16)
        print(json.dumps(dict(op="extract", B=B, G=G, N=N, ms=ms, sequences_per_s=B / ms * 1e3, out_shape=list(feat.shape),
                              finite=bool(torch.isfinite(feat).all()))), flush=True)


if __name__ == "__main__":
    main()
