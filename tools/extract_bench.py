#!/usr/bin/env python
"""BASELINE configs[4]: forward-only feature extraction over 100 000 synthetic NTU-120-shaped sequences on N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/extract_bench.py --sequences 100000 --views 10 [--batch 256] [--save-dir /tmp/facl_feat]

Replaces the loop of reference training_code/extract_motion_feature.py:162-184,217-221 (identical in
extract_apperance_feature.py): eval-mode netR on each batch -> cat(x, x_global) -> .cpu().numpy() -> one
`(G+1)*512` float32 .npy row per video.  Eval-mode BatchNorm uses running statistics, so every cloud is independent:
the job is REPLICAS ONLY -- the sequences are split evenly over the ranks and no collective touches the data
path; the only communication is the barrier / max-reduction of the timing.

Three numbers per run (max over ranks, whole job):
  device : batches already resident in HBM, features left on the device                          (kernel throughput)
  e2e    : every batch copied from PINNED host memory, its (B, (G+1)*512) feature rows copied back to pinned host memory
           (what `.cpu()` at extract_motion_feature.py:183 costs), double-buffered on side streams
  files  : e2e + one np.save per video by a writer thread (extract_motion_feature.py:217-221), when --save-dir is given
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import queue
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=100000)
    ap.add_argument("--views", type=int, default=10, help="G: 10 = num_crop of the reference extractor, 20 = BASELINE's frame count")
    ap.add_argument("--points", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--save-dir", default=None)
    args = ap.parse_args()

    from facl_b200 import cn3d_model_conbag as MODELL, synth, utils_my
    from facl_b200.train import default_opt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    B, G, N = args.batch, args.views, args.points
    mine = args.sequences // world + (1 if rank < args.sequences % world else 0)      # replicas only: an even split
    nbatches = (mine + B - 1) // B

    opt = default_opt(batchSize=B, SAMPLE_NUM=N)
    torch.manual_seed(1)
    net = MODELL.PointNet_Plus_fine(opt, gost=G, sample_num_level1=64, knn_K=64).to(dev)
    net.precision = args.precision
    net.eval()                                                                         # extract_motion_feature.py:156

    host_in = [torch.from_numpy(synth.make_sequences(B, G, N, seed=500 + rank * 10 + i)).pin_memory() for i in range(2)]
    dev_in = [h.to(dev) for h in host_in]
    F = (G + 1) * 512

    def forward(batch):
        """extract_motion_feature.py:171-184 on the device: -> (B, (G+1)*512) rows in save_single_feature's layout (:217-221)"""
        with torch.no_grad():
            data1 = batch.permute(1, 0, 2, 3).reshape(-1, N, 4)
            xt, yt = utils_my.group_points_3DV_2048(data1, 64, 64, SAMPLE_NUM=N)
            x, _, _, xg = net(xt, yt)
            feat = torch.cat((x, xg), dim=0)                                           # :182
            return feat.reshape(G + 1, B, 512).permute(1, 0, 2).reshape(B, F).contiguous()

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for i in range(3):
        forward(dev_in[i % 2])
    sync_all()

    # ---- device-resident ------------------------------------------------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(nbatches):
        forward(dev_in[i % 2])
    e1.record()
    sync_all()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))

    # ---- end to end: H2D of every batch, D2H of every feature block, optional file writer -------------------------------------
    def run_e2e(save_dir):
        copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream()
        stage = [torch.empty((B, G, N, 4), dtype=torch.float32, device=dev) for _ in range(2)]
        host_out = [torch.empty((B, F), dtype=torch.float32).pin_memory() for _ in range(4)]
        in_ready = [torch.cuda.Event() for _ in range(2)]
        in_free = [None, None]
        out_done = [None] * 4
        q = queue.Queue()
        written = [0]

        def writer():
            while True:
                item = q.get()
                if item is None:
                    return
                slot, ev, base, n = item
                ev.synchronize()
                rows = host_out[slot].numpy()
                for j in range(n):
                    np.save(os.path.join(save_dir, f"S{rank:03d}V{base + j:07d}.npy"), rows[j])    # one file per video (:221)
                written[0] += n
                q.task_done()

        th = None
        if save_dir:
            os.makedirs(save_dir, exist_ok=True)
            th = threading.Thread(target=writer, daemon=True)
            th.start()

        def start_h2d(i):
            s = i % 2
            with torch.cuda.stream(copy_in):
                if in_free[s] is not None:
                    copy_in.wait_event(in_free[s])
                stage[s].copy_(host_in[i % 2], non_blocking=True)
                in_ready[s].record(copy_in)

        t0 = time.perf_counter()
        start_h2d(0)
        for i in range(nbatches):
            s = i % 2
            if i + 1 < nbatches:
                start_h2d(i + 1)
            main.wait_event(in_ready[s])
            rows = forward(stage[s])
            ev = torch.cuda.Event()
            ev.record(main)
            in_free[s] = ev
            o = i % 4
            if save_dir and out_done[o] is not None:
                q.join()                                             # the writer must be done with this pinned slot (4 deep)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(ev)
                host_out[o].copy_(rows, non_blocking=True)
                rows.record_stream(copy_out)
                done = torch.cuda.Event()
                done.record(copy_out)
            out_done[o] = done
            if save_dir:
                n = min(B, mine - i * B)
                q.put((o, done, i * B, n))
        torch.cuda.synchronize()
        if th is not None:
            q.join()
            q.put(None)
            th.join()
        ms = (time.perf_counter() - t0) * 1e3
        sync_all()
        return max_over_ranks(ms), written[0]

    e2e_ms, _ = run_e2e(None)
    files_ms, nfiles = (None, 0)
    if args.save_dir:
        files_ms, nfiles = run_e2e(os.path.join(args.save_dir, f"rank{rank}"))

    if rank == 0:
        total = args.sequences
        line = dict(metric="feature extraction sequences/sec (eval-mode forward, BASELINE configs[4])", unit="sequences/s",
                    n_gpus=world, sequences=total, per_gpu=mine, batch=B,
                    config=dict(workload=f"extract_motion_feature forward over {total} synthetic sequences x {G} views x {N} pts, "
                                         f"{args.precision}, {world} GPU(s), replicas only", views=G, points=N,
                                feature_row_floats=F),
                    device=dict(value=total / (dev_ms * 1e-3), seconds=dev_ms * 1e-3),
                    e2e=dict(value=total / (e2e_ms * 1e-3), seconds=e2e_ms * 1e-3, h2d_bytes_per_sequence=G * N * 16,
                             d2h_bytes_per_sequence=F * 4),
                    files=None if files_ms is None else dict(value=total / (files_ms * 1e-3), seconds=files_ms * 1e-3,
                                                             files_written_rank0=nfiles, bytes_per_file=F * 4 + 128),
                    data="synthetic", scaling="replicas")
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
