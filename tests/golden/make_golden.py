#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (tangent-T/FACL) on seeded inputs.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed from the reference, untouched:
  * training_code/utils_my.py         group_points_3DV, group_points_3DV_2048, group_points_3DV_nums,
                                      global_contrast, circle_contrast, Info_NCE
                                      (`.cuda()` is patched to identity -- this container has no GPU)
  * training_code/cn3d_model_conbag.py PointNet_Plus_fine (forward + autograd backward), PointNet_Plus
  * training_code/cn3D_data_set.py:675-694 farthest_point_sampling_fast -- the module itself cannot be
                                      imported (imageio missing), so the method's source lines are
                                      exec'd from the file text at run time; nothing is copied here.
  * training_code/cn3D_data_set.py:285-350, 654-663, 708-713, 734-748, 765-776 -- get_data_train,
                                      get_temporal_augment_data, reverse_transform, rotate_trans,
                                      jitter_point_cloud, lifted the same way and driven in the order of
                                      NTU_RGBD_new.__getitem__ (:105-121) with np.random.seed(...)
  * torch.optim.Adam with the reference's hyper-parameters (cn3d_train_motion_GL.py:180).

While writing each fixture the script also checks that `oracle/` reproduces it (the pin), and
aborts if it does not.
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/training_code"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

torch.Tensor.cuda = lambda self, *a, **k: self          # reference losses hard-code .cuda()
torch.nn.Module.cuda = lambda self, *a, **k: self

import utils_my as ref_utils                              # noqa: E402
import cn3d_model_conbag as ref_model                     # noqa: E402
import oracle                                             # noqa: E402
from oracle.encoder import L1_LAYERS, L3_LAYERS           # noqa: E402
from facl_b200 import synth                               # noqa: E402

torch.set_num_threads(8)


def lift_reference_fps():
    lines = open(os.path.join(REF, "cn3D_data_set.py")).read().split("\n")[674:694]
    ns = {"np": np}
    exec(textwrap.dedent("\n".join(lines)), ns)
    fn = ns["farthest_point_sampling_fast"]

    def run(pc, m, seed):
        np.random.seed(seed)
        start = np.random.randint(0, pc.shape[0])
        np.random.seed(seed)
        return fn(None, pc, m).ravel().astype(np.int32), start
    return run


def make_opt(B, N, S=64, K=64):
    return types.SimpleNamespace(temperal_num=3, knn_K=K, ball_radius=0.16, ball_radius2=0.25,
                                 sample_num_level1=S, sample_num_level2=64, INPUT_FEATURE_NUM=4,
                                 Num_Class=512, batchSize=B, pooling="concatenation", SAMPLE_NUM=N)


def sorted_rows(a):
    """Sort the K rows of every (m,s) group lexicographically -> order-free comparison."""
    M, S, K, D = a.shape
    flat = a.reshape(M * S, K, D)
    out = np.empty_like(flat)
    for i in range(flat.shape[0]):
        r = flat[i]
        out[i] = r[np.lexsort(r.T[::-1])]
    return out.reshape(M, S, K, D)


def gen_fps():
    ref_fps = lift_reference_fps()
    cases = []
    rng = np.random.default_rng(7)
    for ci, (N, m, dt, dup) in enumerate([(256, 16, np.float32, False), (1024, 64, np.float32, False),
                                          (2048, 64, np.float32, False), (2048, 64, np.float64, False),
                                          (512, 64, np.float32, True), (4096, 128, np.float32, False),
                                          (2048, 512, np.float32, False), (100, 100, np.float32, False)]):
        pts = synth.make_sequences(1, 1, N, seed=100 + ci, skeleton=(ci % 2 == 1), resample=dup,
                                   dtype=dt)[0, 0, :, 0:3]
        idx, start = ref_fps(pts, m, seed=ci)
        mine = oracle.farthest_point_sampling(pts, m, start)
        assert np.array_equal(idx, mine), f"oracle FPS != reference (case {ci})"
        if dt == np.float64:   # fp32 and fp64 picks agree on such clouds
            assert np.array_equal(oracle.farthest_point_sampling(pts.astype(np.float32), m, start), idx)
        cases.append(dict(pc=pts, m=m, start=start, idx=idx))
    np.savez_compressed(os.path.join(HERE, "fps.npz"),
                        n_cases=len(cases),
                        **{f"{k}_{i}": np.asarray(c[k]) for i, c in enumerate(cases) for k in c})
    print("fps.npz:", len(cases), "cases")


def gen_group():
    out = {}
    cases = [  # name, fn, M, N, S, K, r2, resample
        ("3dv", "group_points_3DV", 3, 512, 64, 64, 0.06, False),
        ("2048", "group_points_3DV_2048", 2, 2048, 64, 64, 0.16, False),
        ("small_r", "group_points_3DV_nums", 3, 256, 32, 16, 0.0025, False),
        ("tiny_r", "group_points_3DV_nums", 2, 300, 64, 64, 0.01, False),
        ("dups", "group_points_3DV", 3, 512, 64, 64, 0.06, True),
    ]
    for ci, (name, fn, M, N, S, K, r2, resample) in enumerate(cases):
        pts = torch.from_numpy(synth.make_sequences(M, 1, N, seed=200 + ci, skeleton=(ci == 1),
                                                    resample=resample)[:, 0])
        opt = make_opt(1, N, S, K)
        before = pts.clone()
        if fn == "group_points_3DV":
            xt, yt = ref_utils.group_points_3DV(pts, opt)
            assert opt.knn_K == 64 and opt.ball_radius == 0.06            # the side effect (:259-261)
        elif fn == "group_points_3DV_2048":
            xt, yt = ref_utils.group_points_3DV_2048(pts, K, S, SAMPLE_NUM=N)
        else:
            # group_points_3DV_nums hard-codes ball_radius=0.06 (:299); to exercise small radii the
            # same reference code is run with the constant overridden through the opt it reads.
            src = open(os.path.join(REF, "utils_my.py")).read().split("\n")[292:328]
            src = [l for l in src if "opt.ball_radius = 0.06" not in l]
            ns = {"torch": torch}
            exec("\n".join(src), ns)
            opt.ball_radius = r2
            xt, yt = ns["group_points_3DV_nums"](pts, opt, S, K)
        assert torch.equal(before, pts), "reference mutated its input?"
        assert xt.shape == (M, 4, S, K) and yt.shape == (M, 3, S, 1)
        ref_rows = xt.permute(0, 2, 3, 1).contiguous().numpy()              # (M,S,K,4)
        oxt, oyt, oidx = oracle.group_points(pts, S, K, r2)
        o_rows = oxt.permute(0, 2, 3, 1).contiguous().numpy()
        assert np.array_equal(sorted_rows(ref_rows), sorted_rows(o_rows)), f"oracle grouping != reference ({name})"
        assert torch.equal(oyt, yt)
        redirected = float((oracle.knn_ball_indices(pts.numpy(), S, K, r2)[1] > np.float32(r2)).mean())
        out[f"{name}_points"] = pts.numpy()
        out[f"{name}_cfg"] = np.array([S, K], dtype=np.int64)
        out[f"{name}_r2"] = np.array(r2, dtype=np.float64)
        out[f"{name}_rows_sorted"] = sorted_rows(ref_rows)
        out[f"{name}_idx_sorted"] = np.sort(oidx.numpy(), axis=2).astype(np.int16)
        out[f"{name}_tie_free"] = np.array(not resample)
        print(f"group case {name}: redirected fraction {redirected:.3f}")
    out["names"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(HERE, "group.npz"), **out)


def sample_positions(shape, n=64, seed=0):
    numel = int(np.prod(shape))
    rng = np.random.default_rng(seed + numel)
    return np.sort(rng.choice(numel, size=min(n, numel), replace=False))


def gen_encoder_and_step():
    """One full training step of the reference (fp32 AND fp64) on a small seeded batch.

    The fp64 run is the real pin: oracle(fp64) must equal reference(fp64) to ~1e-9, which proves the
    restated math is the reference's math.  In fp32 the step is discontinuous (max-pool / ReLU
    routing flips under rounding), so fp32 agreement is only required up to a small multiple of the
    reference's own fp32-vs-fp64 deviation; that deviation is stored per tensor as `noise/*`."""
    B, G, N, S, K = 4, 3, 128, 64, 64
    r2 = 0.06
    pts = synth.make_sequences(B, G, N, seed=300, skeleton=True)
    order = synth.view_order(G, seed=1)
    sd0 = oracle.init_state_dict(seed=11)
    # non-trivial BN affine so that sign / shift handling is exercised
    g = torch.Generator().manual_seed(5)
    for k in sd0:
        if k.endswith(".weight") and sd0[k].dim() == 1:
            sd0[k] = 0.5 + torch.rand(sd0[k].shape, generator=g)
            sd0[k][::7] *= -1.0
        if k.endswith(".bias") and k.replace(".bias", ".running_mean") in sd0:
            sd0[k] = 0.2 * (torch.rand(sd0[k].shape, generator=g) - 0.5)

    opt = make_opt(B, N, S, K)
    crit = torch.nn.CrossEntropyLoss()
    clouds = torch.from_numpy(pts).permute(1, 0, 2, 3).reshape(-1, N, 4).type(torch.FloatTensor)
    xt, yt = ref_utils.group_points_3DV(clouds, opt)

    def run_reference(dtype):
        net = ref_model.PointNet_Plus_fine(opt, gost=G, sample_num_level1=S, knn_K=K)
        assert list(net.state_dict().keys()) == oracle.STATE_KEYS
        net.load_state_dict({k: v.clone() for k, v in sd0.items()})
        net = net.to(dtype)
        net.train()
        optim = torch.optim.Adam(net.parameters(), lr=0.0003, betas=(0.5, 0.999), eps=1e-06)
        x, code, x_nor, x_global = net(xt.to(dtype), yt.to(dtype), 1)
        if dtype == torch.float32:
            loss_g = ref_utils.global_contrast(G, x_global, x, opt, crit)
            np.random.seed(1)
            loss_c = ref_utils.circle_contrast(G, x, B, crit)
            np.random.seed(1)
            chk = np.arange(G)
            np.random.shuffle(chk)
            assert np.array_equal(chk, order)
        else:   # the reference loss helpers force FloatTensor; in fp64 use the (pinned) closed form
            loss_g = oracle.global_contrast(G, x_global, x, B)
            loss_c = oracle.circle_contrast(G, x, B, order)
        loss = loss_c + loss_g
        optim.zero_grad()
        loss.backward()
        grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p))
                 for k, p in net.named_parameters()}
        optim.step()
        sd1 = {k: v.clone() for k, v in net.state_dict().items()}
        return dict(net=net, x=x.detach(), x_global=x_global.detach(), code=code.detach(), x_nor=x_nor.detach(),
                    loss=(float(loss), float(loss_g), float(loss_c)), grads=grads, sd1=sd1)

    r32 = run_reference(torch.float32)
    r64 = run_reference(torch.float64)

    def run_oracle(dtype):
        osd = {k: (v.clone().to(dtype) if v.dtype.is_floating_point else v.clone()) for k, v in sd0.items()}
        res = oracle.train_step(osd, torch.from_numpy(pts), order, S=S, K=K, r2=r2, dtype=dtype)
        res["sd1"] = osd
        return res

    o32 = run_oracle(torch.float32)
    o64 = run_oracle(torch.float64)

    def rel2(a, b):
        a = a.double().reshape(-1)
        b = b.double().reshape(-1)
        return float((a - b).norm() / b.norm().clamp_min(1e-300))

    # exact-math pin (fp64 vs fp64)
    assert rel2(o64["x"], r64["x"]) < 1e-9 and rel2(o64["x_global"], r64["x_global"]) < 1e-9
    assert abs(o64["loss"] - r64["loss"][0]) < 1e-9 * abs(r64["loss"][0])
    gscale = max(float(v.abs().max()) for v in r64["grads"].values())
    noise = {}
    for k, gref in r64["grads"].items():
        og = o64["grads"][k].reshape(gref.shape)
        if float(gref.norm()) < 1e-9 * gscale * gref.numel() ** 0.5:
            assert float(og.abs().max()) < 1e-9 * gscale, k      # exactly-zero gradients (bias before BN)
        else:
            assert rel2(og, gref) < 1e-8, (k, rel2(og, gref))
        noise[k] = rel2(r32["grads"][k], gref)
    for k, v in r64["sd1"].items():
        if v.dtype.is_floating_point:
            assert rel2(o64["sd1"][k].reshape(v.shape), v) < 1e-8, k
        else:
            assert int(o64["sd1"][k]) == int(v), k
    # fp32: within the reference's own rounding sensitivity
    print("fp32: x rel2", rel2(o32["x"], r32["x"]), "noise floor", rel2(r32["x"], r64["x"]))
    assert rel2(o32["x"], r32["x"]) < 4 * rel2(r32["x"], r64["x"]) + 1e-6
    assert abs(o32["loss"] - r32["loss"][0]) < 4 * abs(r32["loss"][0] - r64["loss"][0]) + 1e-5 * abs(r64["loss"][0])
    worst = 0.0
    for k, gref in r32["grads"].items():
        if noise[k] > 1.0:      # zero-gradient tensors: rounding noise on both sides
            assert float(o32["grads"][k].abs().max()) < 1e-4 * gscale, k
            continue
        e = rel2(o32["grads"][k].reshape(gref.shape), gref)
        worst = max(worst, e / max(noise[k], 1e-6))
        assert e < 6 * noise[k] + 1e-5, (k, e, noise[k])
    print(f"oracle pinned: loss fp32 {r32['loss'][0]:.6f} fp64 {r64['loss'][0]:.6f}; worst grad err / noise = {worst:.2f}")

    # ---- eval-mode forward (feature extraction, extract_motion_feature.py:156-184) --------
    net = r32["net"]
    net.eval()
    with torch.no_grad():
        ex, _, _, exg = net(xt, yt)
        feat = torch.cat((ex, exg), dim=0).numpy()
    ep = oracle.EncoderParams({k: v.clone().reshape(sd0[k].shape) for k, v in r32["sd1"].items()}, training=False)
    with torch.no_grad():
        ox, _, _, oxg = oracle.encoder_forward(ep, xt, yt, gost=G)
    assert rel2(torch.cat((ox, oxg), 0), torch.from_numpy(feat)) < 2e-5

    # ---- PointNet_Plus (the class the scripts construct) returns x only, same numbers ------
    net2 = ref_model.PointNet_Plus(opt, gost=G)
    net2.load_state_dict({k: v.clone() for k, v in sd0.items()})
    net2.train()
    assert torch.equal(net2(xt, yt, 1), r32["x"])

    out = dict(points=pts, order=order, cfg=np.array([B, G, N, S, K]), r2=np.array(r2), seed_sd=np.array(11),
               x=r32["x"].numpy(), x_global=r32["x_global"].numpy(), code=r32["code"].numpy(),
               x_nor=r32["x_nor"].numpy(), loss=np.array(r32["loss"]), loss64=np.array(r64["loss"]),
               x64=r64["x"].numpy(), x_global64=r64["x_global"].numpy(), eval_feat=feat)
    # BN affine that the generator perturbed must be stored (small vectors); big matrices come from the seed
    for k, v in sd0.items():
        if v.dim() <= 1:
            out["sd0/" + k] = v.numpy()
    for k, v in r32["grads"].items():
        pos = sample_positions(v.shape, 256, seed=3)
        out["grad_pos/" + k] = pos
        out["grad_val/" + k] = v.reshape(-1).numpy()[pos]
        out["grad64_val/" + k] = r64["grads"][k].reshape(-1).numpy()[pos]
        out["grad64_norm/" + k] = np.array(float(r64["grads"][k].norm()))
        out["noise/" + k] = np.array(noise[k])
    for k, v in r32["sd1"].items():
        if v.dim() <= 1:
            out["sd1/" + k] = v.numpy()
        else:
            pos = sample_positions(v.shape, 256, seed=4)
            out["sd1_pos/" + k] = pos
            out["sd1_val/" + k] = v.reshape(-1).numpy()[pos]
    # Info_NCE logits on the first two views
    logits, labels = ref_utils.Info_NCE(r32["x"][: 2 * B], opt)
    ol, _ = oracle.info_nce_logits(r32["x"][: 2 * B], B)
    assert rel2(ol, logits) < 1e-6
    out["info_nce_logits"] = logits.numpy()
    np.savez_compressed(os.path.join(HERE, "train_step.npz"), **out)
    print("train_step.npz written")


def gen_fine_geometry():
    """PointNet_Plus_fine with its DEFAULT geometry (sample_num_level1=32, knn_K=128, cn3d_model_conbag.py:142): one training
    step of the reference in fp32 and fp64 on a small batch grouped by utils_my.group_points_3DV_nums (:293-328).  Pins the
    oracle -- and through it the CUDA path -- at a max-pool group that is not 64 wide."""
    B, G, N, S, K = 2, 3, 256, 32, 128
    r2 = 0.06
    pts = synth.make_sequences(B, G, N, seed=310, skeleton=True)
    order = synth.view_order(G, seed=2)
    sd0 = oracle.init_state_dict(seed=13)
    g = torch.Generator().manual_seed(6)
    for k in sd0:
        if k.endswith(".weight") and sd0[k].dim() == 1:
            sd0[k] = 0.5 + torch.rand(sd0[k].shape, generator=g)
            sd0[k][::5] *= -1.0
        if k.endswith(".bias") and k.replace(".bias", ".running_mean") in sd0:
            sd0[k] = 0.2 * (torch.rand(sd0[k].shape, generator=g) - 0.5)
    opt = make_opt(B, N, S, K)
    clouds = torch.from_numpy(pts).permute(1, 0, 2, 3).reshape(-1, N, 4).type(torch.FloatTensor)
    xt, yt = ref_utils.group_points_3DV_nums(clouds, opt, S, K)
    assert xt.shape == (G * B, 4, S, K) and opt.ball_radius == 0.06

    def rel2(a, b):
        a = a.double().reshape(-1)
        b = b.double().reshape(-1)
        return float((a - b).norm() / b.norm().clamp_min(1e-300))

    def run_reference(dtype):
        net = ref_model.PointNet_Plus_fine(opt, gost=G)                      # default sample_num_level1=32, knn_K=128
        assert net.sample_num_level1 == S and net.knn_K == K
        net.load_state_dict({k: v.clone() for k, v in sd0.items()})
        net = net.to(dtype)
        net.train()
        x, code, x_nor, x_global = net(xt.to(dtype), yt.to(dtype), 1)
        loss = oracle.circle_contrast(G, x, B, order) + oracle.global_contrast(G, x_global, x, B)   # pinned by gen_losses
        loss.backward()
        grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in net.named_parameters()}
        return dict(x=x.detach(), x_global=x_global.detach(), loss=float(loss), grads=grads)

    r32, r64 = run_reference(torch.float32), run_reference(torch.float64)
    osd = {k: (v.clone().double() if v.dtype.is_floating_point else v.clone()) for k, v in sd0.items()}
    o64 = oracle.train_step(osd, torch.from_numpy(pts), order, S=S, K=K, r2=r2, dtype=torch.float64, apply_update=False)
    assert rel2(o64["x"], r64["x"]) < 1e-9 and rel2(o64["x_global"], r64["x_global"]) < 1e-9
    assert abs(o64["loss"] - r64["loss"]) < 1e-9 * abs(r64["loss"])
    gscale = max(float(v.abs().max()) for v in r64["grads"].values())
    out = dict(points=pts, order=order, cfg=np.array([B, G, N, S, K]), r2=np.array(r2), seed_sd=np.array(13),
               x64=r64["x"].numpy(), x_global64=r64["x_global"].numpy(), loss64=np.array(r64["loss"]), loss=np.array(r32["loss"]))
    for k, v in sd0.items():
        if v.dim() <= 1:
            out["sd0/" + k] = v.numpy()
    for k, gref in r64["grads"].items():
        og = o64["grads"][k].reshape(gref.shape)
        if float(gref.norm()) < 1e-9 * gscale * gref.numel() ** 0.5:
            assert float(og.abs().max()) < 1e-9 * gscale, k
        else:
            assert rel2(og, gref) < 1e-8, (k, rel2(og, gref))
        pos = sample_positions(gref.shape, 256, seed=5)
        out["grad_pos/" + k] = pos
        out["grad64_val/" + k] = gref.reshape(-1).numpy()[pos]
        out["grad64_norm/" + k] = np.array(float(gref.norm()))
        out["noise/" + k] = np.array(rel2(r32["grads"][k], gref))
    np.savez_compressed(os.path.join(HERE, "fine_geometry.npz"), **out)
    print("fine_geometry.npz written; fp32 noise x", rel2(r32["x"], r64["x"]), "loss", r32["loss"], r64["loss"])


def gen_heads():
    """The disabled loss heads, from the reference functions themselves: distributed_sinkhorn (cn3d_model_conbag.py:391-406, also
    with an overflowed entry for shoot_infs), the inline SwAV block of cn3d_train_motion_GL.py:240-261 (re-issued here line by line
    around the reference's distributed_sinkhorn, queue = None), utils_my.KMeans / grouping / CLD_Loss (:152-198)."""
    from oracle import heads as oh
    out = {}
    g = torch.Generator().manual_seed(77)
    # ---- Sinkhorn
    for i, (K, B, scale, inf) in enumerate([(64, 8, 1.0, False), (64, 64, 3.0, False), (16, 5, 1.0, True)]):
        Q = torch.exp(torch.randn(K, B, generator=g) * scale)
        if inf:
            Q[3, 2] = float("inf")
        ref = ref_model.distributed_sinkhorn(Q.clone(), 3)
        mine = oh.distributed_sinkhorn(Q.clone(), 3)
        assert torch.allclose(ref, mine, rtol=1e-5, atol=1e-7), i
        out[f"sk_q_{i}"], out[f"sk_out_{i}"] = Q.numpy(), ref.numpy()
    out["sk_n"] = np.array(3)
    # ---- SwAV block (num_crop views of B samples, 64 prototypes)
    G, B = 5, 8
    x = torch.randn(G * B, 512, generator=g)
    W = torch.randn(64, 512, generator=g) * 0.05
    xr = x.clone().requires_grad_(True)
    Wr = W.clone().requires_grad_(True)
    x_nor = torch.nn.functional.normalize(xr, dim=1, p=2)          # cn3d_model_conbag.py:231
    code = x_nor @ Wr.t()                                          # :232 (mapping, bias-free)
    softmax = torch.nn.Softmax(dim=1)
    loss_swa = 0
    for crop_id in range(G - 1):                                   # cn3d_train_motion_GL.py:240-261, queue is None
        with torch.no_grad():
            po = code[B * crop_id:B * (crop_id + 1), :]
            po = po / 0.03
            po = torch.exp(po).t()
            q = ref_model.distributed_sinkhorn(po, 3)[-B:]
        subloss = 0
        for v in np.delete(np.arange(G - 1), crop_id):
            p = softmax(code[B * v: B * (v + 1)] / 0.1)
            subloss = subloss - torch.mean(torch.sum(q * torch.log(p), dim=1))
        loss_swa = loss_swa + subloss
    loss_swa = loss_swa / (G - 1)
    loss_swa.backward()
    x2, W2 = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
    code2 = torch.nn.functional.normalize(x2, dim=1, p=2) @ W2.t()
    mine = oh.swav_loss(code2, G, B)
    mine.backward()
    assert abs(float(mine) - float(loss_swa)) < 1e-5 * abs(float(loss_swa)) and torch.allclose(x2.grad, xr.grad, rtol=1e-4, atol=1e-7)
    out.update(swav_x=x.numpy(), swav_w=W.numpy(), swav_cfg=np.array([G, B]), swav_loss=np.array(float(loss_swa)),
               swav_dx=xr.grad.numpy(), swav_dw=Wr.grad.numpy(), swav_code=code.detach().numpy())
    # ---- k-means / CLD on clustered unit vectors (assignments stable under rounding)
    G, B = 6, 32
    centers = torch.nn.functional.normalize(torch.randn(12, 512, generator=g), dim=1)
    pick = torch.randint(0, 12, (G * B,), generator=g)
    feats = torch.nn.functional.normalize(centers[pick] + 0.05 * torch.randn(G * B, 512, generator=g), dim=1)
    fr = feats.clone().requires_grad_(True)
    cl, c = ref_utils.KMeans(fr[: 3 * B], 60, 5)
    ol, oc = oh.kmeans(feats[: 3 * B], 60, 5)
    assert torch.equal(cl, ol) and torch.allclose(c, oc, rtol=1e-5, atol=1e-6)
    opt = make_opt(B, 128)
    loss_cld = ref_utils.CLD_Loss(0, G, fr, opt)
    loss_cld.backward()
    f2 = feats.clone().requires_grad_(True)
    mine = oh.cld_loss(G, f2, B)
    mine.backward()
    assert abs(float(mine) - float(loss_cld)) < 1e-5 * abs(float(loss_cld)) and torch.allclose(f2.grad, fr.grad, rtol=1e-3, atol=1e-6)
    out.update(cld_x=feats.numpy(), cld_cfg=np.array([G, B]), cld_loss=np.array(float(loss_cld)), cld_dx=fr.grad.numpy(),
               km_labels=cl.numpy(), km_centroids=c.detach().numpy())
    np.savez_compressed(os.path.join(HERE, "heads.npz"), **out)
    print("heads.npz written: swav", float(loss_swa), "cld", float(loss_cld))


def gen_losses():
    """Loss-only fixtures at a few (G,B,C) with larger magnitudes (values reach hundreds)."""
    out = {}
    crit = torch.nn.CrossEntropyLoss()
    cases = [(10, 8, 512, 1.0), (20, 8, 512, 0.7), (4, 5, 64, 2.0), (2, 3, 16, 1.0)]
    for ci, (G, B, C, scale) in enumerate(cases):
        g = torch.Generator().manual_seed(400 + ci)
        x = (torch.randn(G * B, C, generator=g) * scale).requires_grad_(True)
        xg = (torch.randn(B, C, generator=g) * scale).requires_grad_(True)
        opt = make_opt(B, 128)
        lg = ref_utils.global_contrast(G, xg, x, opt, crit)
        np.random.seed(ci)
        lc = ref_utils.circle_contrast(G, x, B, crit)
        np.random.seed(ci)
        order = np.arange(G)
        np.random.shuffle(order)
        (lg + lc).backward()
        x2 = x.detach().clone().requires_grad_(True)
        xg2 = xg.detach().clone().requires_grad_(True)
        olg = oracle.global_contrast(G, xg2, x2, B)
        olc = oracle.circle_contrast(G, x2, B, order)
        (olg + olc).backward()
        assert abs(float(olg) - float(lg)) <= 2e-5 * abs(float(lg)) + 1e-4, (float(olg), float(lg))
        assert abs(float(olc) - float(lc)) <= 2e-5 * abs(float(lc)) + 1e-4, (float(olc), float(lc))
        assert float((x2.grad - x.grad).abs().max()) <= 1e-4 * float(x.grad.abs().max()) + 1e-6
        assert float((xg2.grad - xg.grad).abs().max()) <= 1e-4 * float(xg.grad.abs().max()) + 1e-6
        out[f"x_{ci}"] = x.detach().numpy()
        out[f"xg_{ci}"] = xg.detach().numpy()
        out[f"cfg_{ci}"] = np.array([G, B, C])
        out[f"order_{ci}"] = order
        out[f"loss_{ci}"] = np.array([float(lg), float(lc)])
        out[f"dx_{ci}"] = x.grad.numpy()
        out[f"dxg_{ci}"] = xg.grad.numpy()
        print(f"loss case {ci}: global {float(lg):.4f} circle {float(lc):.4f}")
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)


def sorted_cols(a):
    """(M,C,S,K): sort the K columns of every (m,s) group lexicographically over the C channels."""
    M, C, S, K = a.shape
    rows = np.ascontiguousarray(a.transpose(0, 2, 3, 1)).reshape(M * S, K, C)
    out = np.empty_like(rows)
    for i in range(rows.shape[0]):
        r = rows[i]
        out[i] = r[np.lexsort(r.T[::-1])]
    return out.reshape(M, S, K, C)


def gen_group2():
    """Level-2 grouping: utils_my.group_points_2 (K=64) and group_points_2_3DV (K=32, radius 0.11), run unmodified."""
    out = {}
    cases = [("l2", "group_points_2", 2, 9, 512, 128, 64, 0.02), ("l2_3dv", "group_points_2_3DV", 2, 11, 256, 64, 32, 0.11),
             ("l2_small", "group_points_2", 2, 7, 96, 32, 64, 0.005)]
    for ci, (name, fn, M, C, S1, S2, K, r2) in enumerate(cases):
        g = torch.Generator().manual_seed(700 + ci)
        xyz = torch.from_numpy(synth.make_sequences(M, 1, S1, seed=700 + ci)[:, 0, :, 0:3]).permute(0, 2, 1)
        feats = torch.cat([xyz, torch.randn(M, C - 3, S1, generator=g)], 1).contiguous()
        before = feats.clone()
        xt, yt = getattr(ref_utils, fn)(feats, S1, S2, K, torch.tensor(r2))
        assert torch.equal(before, feats)
        assert xt.shape == (M, C, S2, K) and yt.shape == (M, 3, S2, 1)
        oxt, oyt, oidx = oracle.group_points_level2(feats, S2, K, r2)
        assert np.array_equal(sorted_cols(xt.numpy()), sorted_cols(oxt.numpy())), f"oracle level-2 grouping != reference ({name})"
        assert torch.equal(oyt, yt)
        red = float((oidx == torch.arange(S2)[None, :, None]).float().mean())
        out[f"{name}_feats"] = feats.numpy()
        out[f"{name}_cfg"] = np.array([S2, K], dtype=np.int64)
        out[f"{name}_r2"] = np.array(r2, dtype=np.float64)
        out[f"{name}_sorted"] = sorted_cols(xt.numpy())
        print(f"group2 {name}: ok, {red:.3f} of the slots are the centre")
    out["names"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(HERE, "group2.npz"), **out)


def gen_probe():
    """linear_classify: Final_FC (imported unmodified), the loop body of linercls.py:109-122 and its accuracy(),
    plus save_single_feature of extract_motion_feature.py:217-221 (lifted; the script itself needs a GPU + dataset)."""
    import importlib.util
    import tempfile
    from oracle import probe as oprobe
    spec = importlib.util.spec_from_file_location("ref_fc_model", "/root/reference/linear_classify/fc_model.py")
    fcm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fcm)
    text = open("/root/reference/linear_classify/linercls.py").read().split("\n")
    ns = {"torch": torch}
    exec("\n".join(text[157:172]), ns)                       # accuracy(), :158-172
    text = open(os.path.join(REF, "extract_motion_feature.py")).read().split("\n")
    ns2 = {"np": np}
    exec("\n".join(text[216:221]), ns2)                      # save_single_feature, :217-221
    out = {}
    # ---- feature files
    rng = np.random.default_rng(5)
    B, G = 3, 10
    feat = rng.standard_normal(((G + 1) * B, 512)).astype(np.float32)
    names = [f"S001C00{i}P001R001A00{i}" for i in range(B)]
    with tempfile.TemporaryDirectory() as d:
        ns2["save_single_feature"](feat, d + "/", names)
        files = [open(os.path.join(d, n + ".npy"), "rb").read() for n in names]
        rows = np.stack([np.load(os.path.join(d, n + ".npy")) for n in names])
    assert np.array_equal(rows, oprobe.feature_rows(feat, G + 1))
    out["feat"], out["feat_rows"] = feat, rows
    out["file0"] = np.frombuffer(files[0], dtype=np.uint8)
    # ---- probe training steps
    torch.manual_seed(3)
    net = fcm.Final_FC(input_dim=512, gost=2, num_class=20)         # 2 blocks / 20 classes instead of 22 / 120: small fixture
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    crit = torch.nn.CrossEntropyLoss()
    optim = torch.optim.Adam(net.parameters(), lr=0.005, betas=(0.5, 0.999), eps=1e-06)
    g = torch.Generator().manual_seed(9)
    xs = torch.randn(3, 16, 2 * 512, generator=g) * torch.rand(3, 16, 1, generator=g) * 3
    ys = torch.randint(0, 20, (3, 16), generator=g)
    xs[0, 5] = 0                                                # F.normalize's eps branch
    sd, state, losses, top1s = {k: v.clone() for k, v in sd0.items()}, {}, [], []
    for it in range(3):
        output = net(xs[it])
        loss = crit(output, ys[it])
        optim.zero_grad()
        loss.backward()
        if it == 0:
            out["grad_w0"], out["grad_b0"] = net.fc.weight.grad.clone().numpy(), net.fc.bias.grad.clone().numpy()
            out["logits0"] = output.detach().numpy()
        optim.step()
        acc1, _ = ns["accuracy"](output, ys[it], topk=(1, 1))
        losses.append(float(loss.detach()))
        top1s.append(float(acc1[0]))
        o = oprobe.probe_step(sd, xs[it], ys[it], state)
        assert abs(o["loss"] - float(loss)) <= 1e-5 and abs(o["top1"] - float(acc1[0])) < 1e-4, (o["loss"], float(loss))
    for k in sd:
        assert float((sd[k] - net.state_dict()[k]).abs().max()) <= 2e-6, k
    out["x"], out["y"] = xs.numpy(), ys.numpy()
    out["w0"], out["b0"] = sd0["fc.weight"].numpy(), sd0["fc.bias"].numpy()
    out["w3"], out["b3"] = net.state_dict()["fc.weight"].numpy(), net.state_dict()["fc.bias"].numpy()
    out["losses"], out["top1"] = np.array(losses), np.array(top1s)
    print("probe:", losses, top1s)
    np.savez_compressed(os.path.join(HERE, "probe.npz"), **out)


def lift_reference_augment():
    """A stand-in object carrying the reference's own augmentation methods (source text exec'd at run time)."""
    text = open(os.path.join(REF, "cn3D_data_set.py")).read().split("\n")
    ns = {"np": np, "NUM_POINT": 512}
    for a, b in [(285, 350), (654, 663), (708, 713), (734, 748), (765, 776)]:
        exec(textwrap.dedent("\n".join(text[a - 1:b])), ns)
    cls = type("RefAug", (), {k: ns[k] for k in ("get_data_train", "get_temporal_augment_data", "reverse_transform",
                                                  "rotate_trans", "jitter_point_cloud")})
    return cls()


def gen_augment():
    from oracle import augment as oaug
    ref = lift_reference_augment()
    out = {}
    cases = [(2048, 2048, 600, 200), (700, 512, 90, 33), (5000, 1500, 2048, 2048)]
    for ci, (P, Pk, P1, P2) in enumerate(cases):
        rng = np.random.default_rng(900 + ci)

        def cloud(n, ch):
            a = np.empty((n, ch), np.float32)
            a[:, 0] = 0.45 * rng.uniform(-0.5, 0.5, n)
            a[:, 1] = rng.uniform(-0.5, 0.5, n)
            a[:, 2] = 0.30 * rng.uniform(-0.5, 0.5, n)
            a[:, 3:] = rng.uniform(-0.5, 0.5, (n, ch - 3))
            return a
        points, key, res1, res2 = cloud(P, 8), cloud(Pk, 8), cloud(P1, 8), cloud(P2, 8)
        points[rng.uniform(size=P) < 0.4, 4] = 0.0          # rows get_temporal_augment_data must skip
        points[rng.uniform(size=P) < 0.7, 7] = 0.0
        srcs64 = [a.astype(np.float64) for a in (points, key, res1, res2)]   # the dataset stores float64 (.npy)
        np.random.seed(50 + ci)
        p = srcs64[0]
        t2 = ref.get_temporal_augment_data(p, 4)
        t4 = ref.get_temporal_augment_data(p, 7)
        views = ref.get_data_train(p[:, :4], srcs64[1][:, :4], t2[:, :4], t4[:, :4], srcs64[2][:, :4], srcs64[3][:, :4],
                                   num_crop=10)
        views = torch.from_numpy(views).type(torch.FloatTensor).numpy()      # cn3d_train_motion_GL.py:228
        draws = oaug.record_draws(np.random.RandomState(50 + ci), srcs64, 512)
        mine = oaug.make_views(srcs64, draws)
        rot = [g for g, r in enumerate(oaug.GET_DATA_TRAIN) if r[5]]
        exact = [g for g in range(10) if g not in rot]
        assert np.array_equal(mine[exact], views[exact]), f"oracle augmentation != reference (case {ci})"
        assert np.abs(mine[rot] - views[rot]).max() <= 6e-8, np.abs(mine[rot] - views[rot]).max()   # BLAS dot vs explicit sums
        for k, a in zip(("points", "key", "res1", "res2"), (points, key, res1, res2)):
            out[f"{k}_{ci}"] = a
        out[f"idx_{ci}"], out[f"noise_{ci}"], out[f"angle_{ci}"] = draws.idx, draws.noise.astype(np.float64), draws.angle_u
        out[f"views_{ci}"] = views
        print(f"augment case {ci}: ok, max rot diff {np.abs(mine[rot] - views[rot]).max():.2e}")
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "augment.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_augment()
    gen_group2()
    gen_probe()
    gen_fps()
    gen_group()
    gen_losses()
    gen_encoder_and_step()
    gen_fine_geometry()
    gen_heads()
