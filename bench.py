#!/usr/bin/env python
"""Benchmark of the FACL hot path on B200: training sequences/s (grouping + encoder fwd + global/circle loss +
backward + Adam) on synthetic NTU-shaped point-cloud sequences.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference ...                     # the reference's own modules on the host CPU (oracle/_ref)

Prints ONE JSON line (rank 0).  N=1 workload = BASELINE.json configs[1]: batch 64 x 20 views x 2048 points, fp32.
For N>1 (torchrun) every rank keeps that per-GPU batch (weak scaling); embeddings are all-gathered for the
contrastive losses and parameter gradients all-reduced.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

CFG2 = dict(B=64, G=20, N=2048, S=64, K=64, r2=0.16, precision="fp32")      # BASELINE.json configs[1]
CFG3 = dict(B=256, G=20, N=2048, precision="bf16")                            # BASELINE.json configs[2] (global batch)
CPU_SAMPLE = dict(B=8, G=20, N=2048)                                          # bounded CPU sample of the same workload


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle sampling during the timed region."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.rows = gpu_index, threading.Event(), []

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=float(self.rows[0][1]) if self.rows else None,
                    power_w_max=max(float(r[2]) for r in self.rows), samples=len(self.rows), reasons=reasons)


def reference_leg(B, G, N, steps, warmup, threads):
    """Times the reference's own CPU implementation of the step on the host cores.

    kind "reference": the UNMODIFIED reference modules from oracle/_ref/ (tools/vendor_reference.sh): utils_my.group_points_3DV_2048
    -> cn3d_model_conbag.PointNet_Plus_fine -> utils_my.global_contrast / circle_contrast -> autograd backward -> torch.optim.Adam,
    the call sequence of cn3d_train_motion_GL.py:224-335.  kind "port": the oracle restatement, only when oracle/_ref is absent.
    Must run in a process that does not use the GPU (the reference's hard-coded .cuda() calls are patched to identity)."""
    from facl_b200 import synth
    from oracle import ref_step
    torch.set_num_threads(threads)
    pts = torch.from_numpy(synth.make_sequences(B, G, N, seed=1))
    times, phases, loss = [], {}, float("nan")
    if ref_step.available():
        kind = "reference"
        tr = ref_step.ReferenceTrainer(B, G, N, S=CFG2["S"], K=CFG2["K"], seed=1, threads=threads)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            loss, ph = tr.step(pts)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
                for k, v in ph.items():
                    phases[k] = phases.get(k, 0.0) + v / steps
    else:
        import oracle
        kind = "port"
        sd = oracle.init_state_dict(seed=1)
        order = synth.view_order(G, 1)
        state = {}
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            res = oracle.train_step(sd, pts, order, S=CFG2["S"], K=CFG2["K"], r2=CFG2["r2"], adam_state=state)
            state = res["adam_state"]
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            loss = float(res["loss"])
    return sum(times) / len(times), loss, kind, phases


def reference_batch(requested):
    """Batch of the CPU arm: the configuration's own (64 sequences) when the host has the memory for the reference's
    stored activations (~0.47 GB per sequence at 20 views x 2048 points, measured), else the largest power of two that fits."""
    try:
        import psutil
        avail = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        avail = 64.0
    B = requested
    while B > 4 and 0.5 * B + 4 > 0.8 * avail:
        B //= 2
    return B


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    G, N = args.G or CFG2["G"], args.N or CFG2["N"]
    want = args.ref_batch or args.B or CFG2["B"]
    B = reference_batch(want)
    # one step of the full batch is ~15-40 s of CPU work: the K / W the driver passes are clamped so the run ends in minutes
    steps, warmup = max(1, min(args.steps, 1 if B >= 32 else 2)), max(0, min(args.warmup, 1))
    sec, loss, kind, phases = reference_leg(B, G, N, steps, warmup, threads)
    val = B / sec
    what = "unmodified reference modules (oracle/_ref: utils_my.group_points_3DV_2048, cn3d_model_conbag.PointNet_Plus_fine, " \
           "utils_my.global_contrast / circle_contrast, torch.optim.Adam), torch CPU" if kind == "reference" else \
           "oracle port (oracle/_ref absent: run tools/vendor_reference.sh where /root/reference exists)"
    sample = f"{what}; {B} sequences x {G} views x {N} pts per step, {warmup} warm-up + {steps} timed step(s)"
    full = args.B or CFG2["B"]
    note = "the configuration's own global batch" if B == full else f"bounded sample: batch {B} of the configuration's {full} sequences per step"
    line = dict(metric="train sequences/sec (encoder fwd+bwd+InfoNCE)", value=val, unit="sequences/s", n_gpus=args.gpus,
                steps=steps, warmup=warmup, ms_per_step=sec * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=f"motion-stream contrastive training, batch {B} x {G} views x {N} pts, S=K=64, r2={CFG2['r2']}, "
                                     f"fp32 (BASELINE configs[1])", global_batch=B, parallelism="cpu", note=note, loss_at_end=loss),
                cpu_baseline=dict(value=val, unit="sequences/s", cores=threads, kind=kind, sample=sample,
                                  phase_seconds={k: round(v, 3) for k, v in phases.items()}),
                e2e=dict(value=val, unit="sequences/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def cpu_baseline_subprocess(B, steps=2, warmup=1, timeout=600):
    """The cpu_baseline leg of the GPU arm: the reference arm above on a bounded sample, in a CHILD process (the reference's
    .cuda() patch must not leak into the process that drives the GPU)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--ref-batch", str(B), "--steps", str(steps),
           "--warmup", str(warmup)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        line = json.loads(out.stdout.strip().splitlines()[-1])
        return line["cpu_baseline"]
    except Exception as e:                                       # the baseline is a reported number, never a reason to lose the bench line
        return dict(value=None, unit="sequences/s", cores=os.cpu_count() or 1, kind="unavailable", sample=f"failed: {e!r}")


TAG_NAMES = {27: "group_kernel", 28: "fps", 29: "pack_weight", 30: "bn_finalize", 31: "pool_misc", 32: "pool_scatter",
             33: "loss_gemm", 34: "loss_misc", 35: "adam", 36: "transpose", 37: "memset", 38: "l1_misc", 39: "l1_pass_a",
             40: "l1_pass_b", 41: "l1_pass_c", 42: "l1_pass_d", 43: "act_image", 44: "augment_views", 45: "group_level2"}
CIN = [4, 64, 64, 259, 256, 512, 1024, 1024, 512]
COUT = [64, 64, 256, 256, 512, 1024, 1024, 512, 64]


def tag_name(t):
    if t < 27:
        return f"gemm_tc L{t // 3} {'fwd wgrad dgrad'.split()[t % 3]}"
    return TAG_NAMES.get(t, str(t))


def run_ours(args):
    import ctypes as C
    from facl_b200 import _lib, synth
    from facl_b200.train import FusedTrainStep, TrainStep, default_opt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = dict(CFG2)
    for k in ("B", "G", "N"):
        if getattr(args, k) is not None:
            cfg[k] = getattr(args, k)
    if args.precision:
        cfg["precision"] = args.precision
    B, G, N = cfg["B"], cfg["G"], cfg["N"]
    opt = default_opt(batchSize=B, SAMPLE_NUM=N)
    tr = TrainStep(opt, num_crop=G, precision=cfg["precision"], radius2=cfg["r2"], device=f"cuda:{local_rank}", seed=1)
    from facl_b200.dist import DistributedFusedTrainStep
    if world > 1:
        fused = DistributedFusedTrainStep(tr, B, G, N, r2=cfg["r2"])
    else:
        fused = FusedTrainStep(tr, B, G, N, r2=cfg["r2"])
    nb = 2                                                           # distinct synthetic batches, cycled
    host = [torch.from_numpy(synth.make_sequences(B, G, N, seed=100 + rank * 10 + i)).pin_memory() for i in range(nb)]
    dev = [h.cuda(non_blocking=True) for h in host]
    L = _lib.lib()

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput (per-kernel event timing OFF: it adds two event records per launch) ----------
    for i in range(args.warmup):
        fused.step(dev[i % nb])
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.facl_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    th0 = time.perf_counter()
    for i in range(args.steps):
        fused.step(dev[i % nb])
    host_issue_ms = (time.perf_counter() - th0) * 1e3 / args.steps      # host time to ISSUE a step (no synchronisation inside)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = L.facl_launch_count() - launches0
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    loss_dev = float(fused.loss2[2])
    # ---- the same steps again with every launch site bracketed by CUDA events: per-kernel breakdown and roofline ----
    ksteps = min(args.steps, 10)
    L.facl_timing_enable(1 if rank == 0 else 0)
    for i in range(ksteps):
        fused.step(dev[i % nb])
    sync_all()
    L.facl_timing_enable(0)
    ntags = 46
    tms, tcnt = (C.c_float * ntags)(), (C.c_int * ntags)()
    L.facl_timing_collect(tms, tcnt, ntags)

    # ---- end to end: pinned host batch in, loss back to the host, every step ------------------------------
    e2e_steps = max(2, args.steps)
    fused.step(host[0], want_host_loss=True, next_batch=host[1 % nb])    # the H2D of batch i+1 overlaps step i (prefetching loader)
    sync_all()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        fused.step(host[(i + 1) % nb], want_host_loss=True, next_batch=host[(i + 2) % nb])
        torch.cuda.current_stream().synchronize()                    # loss.item() of the reference loop (:335)
        _ = float(fused.loss_host[0])
    e1.record()
    sync_all()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    # ---- N > 1: is the sharded loss the loss of the global batch?  One more step with a known view order; the embeddings of
    # every rank are gathered and rank 0 recomputes the losses through the world-size-1 facl_contrast_losses ----------------
    dist_check = None
    dist_timeline = None
    if dist is not None and args.dist_timeline:
        # where a sharded step spends its time: CUDA events at every phase / collective boundary, 10 steps, max over ranks
        res = {}
        for mode, ov in (("bucketed_overlapped", True), ("single_allreduce", False)):
            fused.overlap = ov
            for i in range(3):
                fused.step(dev[i % nb])
            sync_all()
            fused.enable_timeline(True)
            for i in range(10):
                fused.step(dev[i % nb])
            tl = fused.timeline_ms()
            fused.enable_timeline(False)
            t = torch.tensor(list(tl.values()), device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[mode] = {k: round(float(v), 4) for k, v in zip(tl.keys(), t)}
            res[mode]["sum"] = round(float(t.sum()), 4)
        fused.overlap = True
        dist_timeline = res
    if dist is not None:
        from facl_b200 import losses as facl_losses
        from facl_b200.dist import reference_order_from_keys
        order = synth.view_order(G, 9)
        fused.step(dev[0], order=order)
        sync_all()
        xs = [torch.empty_like(fused.x) for _ in range(world)]
        xgs = [torch.empty_like(fused.xg) for _ in range(world)]
        dist.all_gather(xs, fused.x)
        dist.all_gather(xgs, fused.xg)
        if rank == 0:
            x_ref = reference_order_from_keys(torch.cat(xs, 0), G, B * world, B)
            lg, lc = facl_losses.contrast_losses(x_ref, torch.cat(xgs, 0), G, B * world, order=order)
            single, sharded = float(lg) + float(lc), float(fused.loss2[2])
            dist_check = dict(rel_err=abs(sharded - single) / abs(single), loss_sharded=sharded, loss_single_process=single,
                              what="all-reduced loss of the sharded step vs facl_contrast_losses (world size 1) on the gathered embeddings")
            del x_ref
        del xs, xgs

    # ---- the reference-shaped API path (N = 1): utils_my.group_points_3DV_2048 -> PointNet_Plus_fine -> contrast losses ->
    # backward -> Adam through torch autograd, pinned host batch in, loss.item() out -- what a script gets after the two-line
    # import swap of INTEGRATION.md, without the fused C-ABI step ------------------------------------------------------------
    api_path = None
    if dist is None and not args.no_api_path:
        asteps = max(2, min(args.steps, 10))
        order = synth.view_order(G, 9)
        for i in range(2):
            tr.step(host[i % nb], order=order)
        sync_all()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ta = time.perf_counter()
        a0.record()
        for i in range(asteps):
            _ = float(tr.step(host[i % nb], order=order))             # .item(): the reference loop's loss.item() (:335)
        a1.record()
        sync_all()
        ams = max(a0.elapsed_time(a1), (time.perf_counter() - ta) * 1e3)
        # the same loop fed through DevicePrefetcher (the H2D copy of batch i+1 under step i), the one-line change INTEGRATION.md shows
        from facl_b200.train import DevicePrefetcher
        sync_all()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tb = None
        for i, batch in enumerate(DevicePrefetcher([host[i % nb] for i in range(asteps + 2)], device=f"cuda:{local_rank}")):
            if i == 2:                                                # two untimed steps: the prefetcher allocates its two device slots
                sync_all()
                tb = time.perf_counter()
                b0.record()
            _ = float(tr.step(batch, order=order))
        b1.record()
        sync_all()
        bms = max(b0.elapsed_time(b1), (time.perf_counter() - tb) * 1e3)
        api_path = dict(value=B * asteps / (ams * 1e-3), unit="sequences/s", ms_per_step=ams / asteps, steps=asteps,
                        what="TrainStep.step: reference-shaped module calls through torch autograd, host batch copied synchronously as "
                             "the reference does (.cuda() at :228), loss.item() out",
                        prefetched=dict(value=B * asteps / (bms * 1e-3), ms_per_step=bms / asteps,
                                        what="same loop over facl_b200.train.DevicePrefetcher(loader): next batch's H2D under the current step"))

    # ---- BASELINE configs[2]: appearance stream, bf16, GLOBAL batch 256 sharded over the N ranks (strong scaling) ---------------
    cfg3 = None
    if not args.no_cfg3 and (B, G, N) == (CFG2["B"], CFG2["G"], CFG2["N"]) and CFG3["B"] % world == 0:
        del fused, tr, dev
        torch.cuda.empty_cache()
        Bl3 = CFG3["B"] // world
        opt3 = default_opt(batchSize=Bl3, SAMPLE_NUM=N)
        tr3 = TrainStep(opt3, num_crop=G, precision="bf16", radius2=cfg["r2"], device=f"cuda:{local_rank}", seed=1)
        if world > 1:
            f3 = DistributedFusedTrainStep(tr3, Bl3, G, N, r2=cfg["r2"])
        else:
            f3 = FusedTrainStep(tr3, Bl3, G, N, r2=cfg["r2"])
        dev3 = [torch.from_numpy(synth.make_sequences(Bl3, G, N, seed=300 + rank * 10 + i)).cuda() for i in range(2)]
        w3, s3 = max(3, min(args.warmup, 5)), max(3, min(args.steps, 10))
        for i in range(w3):
            f3.step(dev3[i % 2])
        sync_all()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(s3):
            f3.step(dev3[i % 2])
        c1.record()
        sync_all()
        cms = c0.elapsed_time(c1)
        if dist is not None:
            t = torch.tensor([cms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cms = float(t)
        cfg3 = dict(value=CFG3["B"] * s3 / (cms * 1e-3), unit="sequences/s", ms_per_step=cms / s3, steps=s3, warmup=w3,
                    global_batch=CFG3["B"], per_gpu_batch=Bl3, n_gpus=world, scaling="strong",
                    dtype="bf16 mixed: single bf16 products in net3DV_3 layers 2-3, bf16x3 split products in net3DV_1 and the first "
                          "net3DV_3 layer (the 2e-2 tolerance needs them), fp32 accumulate / BN statistics / loss / master weights",
                    workload=f"appearance-stream contrastive training, global batch {CFG3['B']} x {G} views x {N} pts, bf16, "
                             f"B sharded over {world} GPU(s), all-gathered negatives (BASELINE configs[2])",
                    loss_at_end=float(f3.loss2[2]))
        del f3, tr3, dev3
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    M = G * B
    R3, R1 = M * 64, M * 64 * 64
    rows = [R1, R1, R1, R3, R3, R3, M + B, M + B, M]
    # algorithmic FLOPs per launch of the fused net3DV_1 passes: A = layer-2 forward (statistics pass),
    # B = layer-2 + layer-3 forward, C = layer-3 data + weight gradient, D = layer-2 data + weight gradient + layer-1 weight gradient
    l1_flops = {39: 2.0 * 64 * 64 * R1, 40: 2.0 * (64 * 64 + 64 * 256) * R1, 41: 4.0 * 64 * 256 * R1,
                42: (4.0 * 64 * 64 + 2.0 * 4 * 64) * R1}
    # algorithmic HBM bytes per launch of the memory-bound kernels
    hbm_bytes = {27: M * (16 * N + 16 * 64 * 64 + 12 * 64)}
    per_tag = []
    for t in range(ntags):
        if tcnt[t] == 0:
            continue
        flops = None
        if t < 27:
            l = t // 3
            flops = 2.0 * CIN[l] * COUT[l] * rows[l]                  # algorithmic FLOPs per step of this GEMM (head: both row sets)
            if l == 3 and t % 3 == 2:
                flops = 2.0 * 256 * 256 * rows[l]
        elif t in l1_flops:
            flops = l1_flops[t]
        per_tag.append(dict(tag=t, name=tag_name(t), ms_per_step=tms[t] / ksteps, launches_per_step=tcnt[t] / ksteps,
                            flops=flops))
    per_tag.sort(key=lambda d: -d["ms_per_step"])
    kernel_ms = sum(d["ms_per_step"] for d in per_tag)
    top = per_tag[0]
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and (B, G, N) == (CFG2["B"], CFG2["G"], CFG2["N"]) and cfg["precision"] == "fp32":
        tj = json.load(open(tpath))
        traffic = tj.get(top["name"])                                   # dram bytes per launch from the committed ncu --set full capture
        traffic_source = tj.get("_source")
    # bf16 (the 2e-2 mode) keeps the split products in net3DV_1, where every roofline kernel lives; only bf16_fast issues single products there
    split = 1.0 if cfg["precision"] == "bf16_fast" else 3.0
    if top["flops"]:
        achieved = top["flops"] / (top["ms_per_step"] * 1e-3) / 1e12   # per-step FLOPs of the tag / per-step time of the tag
        roof = dict(kernel=top["name"], bound="tensor", achieved=achieved, peak=peaks["tensor_sustained"], unit="TFLOP/s",
                    frac=achieved / peaks["tensor_sustained"], traffic=traffic, traffic_source=traffic_source, peak_source=peaks["source"] + ", sustained bf16",
                    share_of_step=top["ms_per_step"] / (ms / args.steps),
                    note=f"algorithmic FLOPs; the {cfg['precision']} mode issues {split:.0f} bf16 products per FLOP, so its ceiling is "
                         f"{peaks['tensor_sustained'] / split:.0f} TFLOP/s (frac_of_mode_ceiling)",
                    frac_of_mode_ceiling=achieved * split / peaks["tensor_sustained"],
                    limiter="shared-memory bandwidth, not the tensor pipe: on these 64-wide tiles a 128 x N x 16 tcgen05.mma fetches 8192/N + 64 B of "
                            "operands per cycle of math; pass C moves 312 KB of operands + ~130 KB of role traffic per 64-row tile at 128 B/cycle/SM "
                            "= 3 450 of its ~4 600 cycles (profiles/r2_role_profile.md)")
    else:
        nbytes = hbm_bytes.get(top["tag"], 0)
        achieved = nbytes / (top["ms_per_step"] * 1e-3) / 1e9
        roof = dict(kernel=top["name"], bound="hbm", achieved=achieved, peak=peaks["hbm"], unit="GB/s",
                    frac=achieved / peaks["hbm"], traffic=traffic, traffic_source=traffic_source, peak_source=peaks["source"],
                    share_of_step=top["ms_per_step"] / (ms / args.steps))
    # whole-step tensor roofline: 15.93 GFLOP per sequence against the sustained bf16 peak
    step_flops = 796e6 * M
    step_tf = step_flops / (ms / args.steps * 1e-3) / 1e12

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_subprocess(CPU_SAMPLE["B"])            # ~10 s of CPU work: 1 warm-up + 2 timed steps of 8 sequences
    h2d = B * G * N * 4 * 4 + G * 4
    line = dict(metric="train sequences/sec (encoder fwd+bwd+InfoNCE)", value=world * B * args.steps / (ms * 1e-3),
                unit="sequences/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32 (bf16x3 split tensor-core products, fp32 accumulate)" if cfg["precision"] == "fp32" else
                      "bf16 (fp32 accumulate, fp32 BN statistics / loss / master weights)",
                data="synthetic",
                config=dict(workload=f"motion-stream contrastive training, batch {B} x {G} views x {N} pts per GPU, "
                                     f"S=K=64, r2={cfg['r2']}, {cfg['precision']} (BASELINE configs[1])",
                            global_batch=world * B, parallelism=f"dp{world}", l2="working set (activations, >10 GB) >> 126 MB L2",
                            loss_at_end=loss_dev),
                e2e=dict(value=world * B * e2e_steps / (e2e_ms * 1e-3), unit="sequences/s", h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=4, ms_per_step=e2e_ms / e2e_steps),
                gpu_launches=int(launches), launches_per_step=launches / args.steps, host_issue_ms_per_step=host_issue_ms,
                roofline=roof, step_tensor_tflops=step_tf, step_tensor_frac=step_tf / peaks["tensor_sustained"],
                kernel_ms_per_step=kernel_ms, kernels=per_tag, cpu_baseline=cpu, clocks=sampler.summary(),
                api_path=api_path, cfg3_strong=cfg3, dist_loss_check=dist_check, dist_timeline=dist_timeline)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--B", type=int, default=None)
    ap.add_argument("--G", type=int, default=None)
    ap.add_argument("--N", type=int, default=None)
    ap.add_argument("--precision", default=None, choices=[None, "fp32", "bf16", "bf16_fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-api-path", action="store_true")
    ap.add_argument("--no-cfg3", action="store_true")
    ap.add_argument("--dist-timeline", action="store_true", help="N > 1: per-phase / per-collective milliseconds of the sharded step")
    ap.add_argument("--ref-batch", type=int, default=None, help="batch of the CPU reference arm (default: the configuration's 64)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
