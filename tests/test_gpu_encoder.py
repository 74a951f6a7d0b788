"""GPU parity tests: encoder forward / backward and the contrastive losses, through the reference-shaped Python API
(facl_b200.cn3d_model_conbag / facl_b200.utils_my), i.e. through the C ABI.  Checker: oracle/ + tests/golden/."""
import os
import types

import numpy as np
import pytest
import torch

import oracle
from oracle.encoder import EncoderParams
from facl_b200 import cn3d_model_conbag as MODELL
from facl_b200 import losses as facl_losses
from facl_b200 import utils_my

pytestmark = pytest.mark.gpu
DEV = "cuda"

# Tolerances, measured as ||a-b||_2 / ||b||_2 per tensor against the fp64 oracle:
#   fp32 mode (bf16x3 split products, fp32 accumulate): features, loss, gradients <= 1e-3   (north_star: <= 1e-3)
#   bf16 mode: features, loss <= 2e-2 (north_star: <= 2e-2); gradients <= 2e-2 stage-wise (see below).  "bf16" is the MIXED
#     mode: single bf16 products for net3DV_3 layers 2-3, the split products for net3DV_1 and the first net3DV_3 layer.
#     tools/bf16_layers.py measured why: with every layer in single bf16 products ("bf16_fast") the embedding error is
#     3.1e-2 at the fixture size and 4.4e-2 at 8 x 20 x 2048, 71 % of its variance from net3DV_1 layers 1-2 and 25 % from
#     the 259-wide layer (their errors are amplified ~12x by the BatchNorms / max-pools downstream); keeping those three on
#     the split path gives 9e-3 / 8e-3.  A CPU emulation of the reference's modules with bf16-rounded operands (what
#     autocast computes) reproduces the 3.1e-2, so it is a property of 8-bit mantissas on this network, not of the kernels.
# Gradients: the gradient of this network is a DISCONTINUOUS function of the activations -- a near-tie in a max-pool
# moves the routed gradient to another row, an activation crossing zero flips a ReLU.  The reference's own fp32 run
# differs from its fp64 run by up to 3e-3 on these inputs for that reason alone (tests/golden `noise/*`).  Gradient
# parity is therefore asserted against the oracle evaluated with the SAME discrete decisions the CUDA forward took
# (oracle.encoder_forward(routing=...)), where the comparison is smooth; the decisions themselves are validated by
# the forward checks (a wrong winner or mask would show up in x / loss).
TOL = {"fp32": 1e-3, "bf16": 2e-2, "bf16_fast": 2e-2}
TOL_FEAT = {"fp32": 1e-3, "bf16": 2e-2, "bf16_fast": 5e-2}      # bf16_fast: outside the bound by design, see cn3d_model_conbag.py
TOL_GRAD = {"fp32": 1e-3, "bf16": 2e-2}


def make_opt(B, N, S=64, K=64):
    return types.SimpleNamespace(temperal_num=3, knn_K=K, ball_radius=0.16, ball_radius2=0.25, sample_num_level1=S,
                                 sample_num_level2=64, INPUT_FEATURE_NUM=4, Num_Class=512, batchSize=B,
                                 pooling="concatenation", SAMPLE_NUM=N)


def rel2(a, b):
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def load_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "train_step.npz"))
    sd = oracle.init_state_dict(seed=int(z["seed_sd"]))
    for k in list(sd):
        if "sd0/" + k in z.files:
            sd[k] = torch.from_numpy(z["sd0/" + k]).clone()
    return z, sd


# ------------------------------------------------------------------------------------------------- losses
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_losses_golden(golden_dir, prec):
    z = np.load(os.path.join(golden_dir, "losses.npz"))
    tol = TOL[prec]
    for i in range(int(z["n_cases"])):
        G, B, C = (int(v) for v in z[f"cfg_{i}"])
        x = torch.from_numpy(z[f"x_{i}"]).to(DEV).requires_grad_(True)
        xg = torch.from_numpy(z[f"xg_{i}"]).to(DEV).requires_grad_(True)
        lg, lc = facl_losses.contrast_losses(x, xg, G, B, order=z[f"order_{i}"], prec=prec)
        (lg + lc).backward()
        ref_g, ref_c = z[f"loss_{i}"]
        assert abs(float(lg) - ref_g) <= tol * abs(ref_g), (i, float(lg), ref_g)
        assert abs(float(lc) - ref_c) <= tol * abs(ref_c), (i, float(lc), ref_c)
        assert rel2(x.grad, z[f"dx_{i}"]) <= tol * 5, (i, rel2(x.grad, z[f"dx_{i}"]))
        assert rel2(xg.grad, z[f"dxg_{i}"]) <= tol * 5, (i, rel2(xg.grad, z[f"dxg_{i}"]))


def test_losses_reference_api(golden_dir):
    z = np.load(os.path.join(golden_dir, "losses.npz"))
    G, B, C = (int(v) for v in z["cfg_0"])
    x, xg = torch.from_numpy(z["x_0"]).to(DEV), torch.from_numpy(z["xg_0"]).to(DEV)
    opt = make_opt(B, 128)
    crit = torch.nn.CrossEntropyLoss()
    lg = utils_my.global_contrast(G, xg, x, opt, crit)
    np.random.seed(0)                                     # make_golden.py seeded numpy with the case index
    lc = utils_my.circle_contrast(G, x, B, crit)
    assert abs(float(lg) - z["loss_0"][0]) <= 1e-3 * z["loss_0"][0]
    assert abs(float(lc) - z["loss_0"][1]) <= 1e-3 * z["loss_0"][1]


# ------------------------------------------------------------------------------------------------- encoder
def _build(sd, B, G, N, S, K, prec):
    opt = make_opt(B, N, S, K)
    net = MODELL.PointNet_Plus_fine(opt, gost=G, sample_num_level1=S, knn_K=K)
    net.load_state_dict({k: v.clone() for k, v in sd.items()})
    net = net.to(DEV)
    net.precision = prec
    return net, opt


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_train_step_vs_golden_and_oracle(golden_dir, prec):
    z, sd0 = load_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    tol = TOL[prec]
    net, opt = _build(sd0, B, G, N, S, K, prec)
    net.fused_l1 = False          # per-layer schedule: keeps every pre-BN activation, so all discrete decisions can be read back
    net.train()
    pts = torch.from_numpy(z["points"])
    clouds = pts.permute(1, 0, 2, 3).reshape(-1, N, 4).type(torch.FloatTensor).to(DEV)
    xt, yt = utils_my.group_points_3DV(clouds, opt)
    assert xt.shape == (G * B, 4, S, K) and yt.shape == (G * B, 3, S, 1)
    x, code, x_nor, x_global = net(xt, yt, 1)
    x.retain_grad()
    x_global.retain_grad()
    lg, lc = facl_losses.contrast_losses(x, x_global, G, B, order=z["order"], prec=prec)
    loss = lc + lg
    loss.backward()
    torch.cuda.synchronize()

    from facl_b200.debug import routing_of_last_forward
    routing = routing_of_last_forward(net)
    sd64 = {k: (v.clone().double() if v.dtype.is_floating_point else v.clone()) for k, v in sd0.items()}
    o64 = oracle.train_step(sd64, pts, z["order"], S=S, K=K, r2=float(z["r2"]), apply_update=False, dtype=torch.float64)
    sdr = {k: (v.clone().double() if v.dtype.is_floating_point else v.clone()) for k, v in sd0.items()}
    # fp32 mode: the whole step is compared (loss gradient included).  bf16 mode: STAGE-WISE -- the oracle back-propagates the
    # upstream gradient the CUDA path actually used.  The reference loss has temperature 1 on un-normalised dot products
    # (logits of magnitude ~500, utils_my.py:72-82): a 1 % embedding error moves a logit by ~5 and a softmax weight by e^5, so
    # gradients of the two END-TO-END runs cannot agree to 2e-2 in ANY 8-bit-mantissa arithmetic; what bf16 mode is held to
    # is each stage within 2e-2 given identical inputs (the loss stage runs the split products in both modes and is checked
    # against the reference fixture in test_losses_golden).
    upstream = None if prec == "fp32" else (x.grad.detach().cpu(), x_global.grad.detach().cpu())
    ort = oracle.train_step(sdr, pts, z["order"], S=S, K=K, r2=float(z["r2"]), apply_update=False, dtype=torch.float64,
                            routing=routing, upstream=upstream)

    # forward: features and loss, against the reference fixture (fp64 run of the reference) and the fp64 oracle
    ft = TOL_FEAT[prec]
    assert rel2(x, z["x64"]) <= ft, ("x", rel2(x, z["x64"]))
    assert rel2(x_global, z["x_global64"]) <= ft * 2, ("x_global", rel2(x_global, z["x_global64"]))
    assert rel2(x_nor, z["x_nor"]) <= ft * 2
    assert rel2(code, z["code"]) <= ft * 2
    assert abs(float(loss) - z["loss64"][0]) <= tol * abs(z["loss64"][0]), (float(loss), z["loss64"][0])
    # the discrete decisions of the CUDA forward are near-optimal: imposing them on the fp64 oracle moves its output
    # by no more than the feature tolerance
    assert rel2(ort["x"], o64["x"]) <= ft and abs(ort["loss"] - o64["loss"]) <= tol * abs(o64["loss"])
    # running statistics after one training forward
    for name, ci, bi in [("net3DV_1", 0, 1), ("net3DV_1", 3, 4), ("net3DV_1", 6, 7), ("net3DV_3", 0, 1),
                         ("net3DV_3", 3, 4), ("net3DV_3", 6, 7), ("netR_FC", 0, 1)]:
        bn = getattr(net, name)[bi]
        key = f"{name}.{bi}"
        assert rel2(bn.running_mean, sd64[key + ".running_mean"]) <= max(ft, 1e-3), key
        assert rel2(bn.running_var, sd64[key + ".running_var"]) <= max(ft, 1e-3) * 2, key
        assert int(bn.num_batches_tracked) == int(sd64[key + ".num_batches_tracked"]), key
    # backward: every parameter gradient, same discrete decisions on both sides
    gscale = max(float(g.abs().max()) for g in o64["grads"].values())
    report, worst = [], 0.0
    for k, p in net.named_parameters():
        g64 = ort["grads"][k].reshape(p.shape)
        if k == "mapping.weight":
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        if float(g64.norm()) <= 1e-9 * gscale * g64.numel() ** 0.5:        # bias in front of a train-mode BN: exactly 0
            assert float(p.grad.abs().max()) <= 1e-4 * gscale, k
            continue
        err = rel2(p.grad, g64)
        rms_err = float((p.grad.detach().double().cpu().reshape(-1) - g64.double().reshape(-1)).norm()) / g64.numel() ** 0.5
        report.append((k, err, rel2(p.grad, o64["grads"][k].reshape(p.shape))))
        # a gradient that is structurally a near-zero residual (the BatchNorm beta in front of a layer whose own BatchNorm removes
        # constants: rms 1e-5 of the largest gradient) is judged on its absolute error, as in _check_fused_backward
        if rms_err / gscale > TOL_GRAD[prec] * 1e-3:
            worst = max(worst, err)
    print("\n" + "\n".join(f"{k:24s} matched-decisions err {e:.2e}   free-running err {f:.2e}" for k, e, f in report))
    assert worst <= TOL_GRAD[prec], report
    if prec != "fp32":
        return
    # FREE-RUNNING gradients (no decisions imposed) against the reference's own fp64 gradients, held to the rule the fixture
    # generator holds the CPU oracle to (tests/golden/make_golden.py: err < 6 * noise + 1e-5, noise = the reference's fp32-vs-
    # fp64 deviation of that tensor, stored as noise/*): (i) whole tensors against the fp64 oracle, which reproduces the
    # reference's fp64 gradients to 1e-8 (asserted when the fixture was written); (ii) the 256 sampled entries per tensor the
    # fixture stores from the reference run itself (grad64_val/*), sampled-entry rule of tests/test_oracle_golden.py (8 * noise)
    free = []
    for k, p in net.named_parameters():
        if k == "mapping.weight" or p.grad is None:
            continue
        noise = float(z["noise/" + k])
        got = p.grad.detach().double().cpu().reshape(-1)
        if noise > 1.0:                                    # mathematically-zero gradients (bias in front of a train-mode BN)
            assert float(got.abs().max()) <= 1e-4 * gscale, k
            continue
        e_full = rel2(p.grad, o64["grads"][k].reshape(p.shape))
        rms = float(z["grad64_norm/" + k]) / np.sqrt(got.numel())
        e_samp = float(np.sqrt(np.mean((got.numpy()[z["grad_pos/" + k]] - z["grad64_val/" + k]) ** 2)) / rms)
        free.append((k, e_full, e_samp, noise))
    print("\n" + "\n".join(f"{k:24s} free-running: vs fp64 oracle {a:.2e}  vs reference samples {b:.2e}  (reference fp32 noise {n:.2e})"
                           for k, a, b, n in free))
    for k, a, b, n in free:
        assert a < 6 * n + 1e-5, (k, a, n)
        assert b < 8 * n + 1e-5, (k, b, n)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_eval_forward_matches_golden(golden_dir, prec):
    z, sd0 = load_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    # post-step weights + running stats from the oracle (pinned to the reference by the CPU tests)
    sd = {k: v.clone() for k, v in sd0.items()}
    oracle.train_step(sd, torch.from_numpy(z["points"]), z["order"], S=S, K=K, r2=float(z["r2"]))
    net, opt = _build(sd, B, G, N, S, K, prec)
    net.eval()
    clouds = torch.from_numpy(z["points"]).permute(1, 0, 2, 3).reshape(-1, N, 4).to(DEV)
    with torch.no_grad():
        xt, yt = utils_my.group_points_3DV(clouds, opt)
        x, _, _, xg = net(xt, yt)
        feat = torch.cat((x, xg), dim=0)                       # extract_motion_feature.py:182
    assert rel2(feat, z["eval_feat"]) <= max(TOL_FEAT[prec], 2e-3)
    # eval must not touch the running statistics
    assert torch.equal(net.net3DV_1[1].running_mean.cpu(), sd["net3DV_1.1.running_mean"])
    assert int(net.net3DV_1[1].num_batches_tracked) == int(sd["net3DV_1.1.num_batches_tracked"])


def test_pointnet_plus_returns_x_only(golden_dir):
    z, sd0 = load_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    opt = make_opt(B, N, S, K)
    net = MODELL.PointNet_Plus(opt, gost=G)
    net.load_state_dict({k: v.clone() for k, v in sd0.items()})
    net = torch.nn.DataParallel(net.to(DEV), device_ids=[0])   # the scripts wrap it (cn3d_train_motion_GL.py:176)
    net.train()
    clouds = torch.from_numpy(z["points"]).permute(1, 0, 2, 3).reshape(-1, N, 4).to(DEV)
    xt, yt = utils_my.group_points_3DV(clouds, opt)
    x = net(xt, yt, 1)
    assert isinstance(x, torch.Tensor) and x.shape == (G * B, 512)
    assert rel2(x, z["x64"]) <= 1e-3
    assert sorted(net.module.state_dict().keys()) == sorted(oracle.STATE_KEYS)


# ------------------------------------------------------------------------------------------------- whole step
def test_fused_step_equals_api_step_and_oracle(golden_dir):
    """facl_train_step (one C-ABI call) == the reference-shaped call sequence through torch autograd == oracle."""
    from facl_b200.train import FusedTrainStep, TrainStep
    z, sd0 = load_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    pts = torch.from_numpy(z["points"])
    order = z["order"]

    def fresh():
        opt = make_opt(B, N, S, K)
        opt.learning_rate = 0.0003
        net = MODELL.PointNet_Plus_fine(opt, gost=G, sample_num_level1=S, knn_K=K)
        net.load_state_dict({k: v.clone() for k, v in sd0.items()})
        return TrainStep(opt, num_crop=G, precision="fp32", model=net)

    a = fresh()
    loss_a = float(a.step(pts.to(DEV), order=order))
    b = fresh()
    fb = FusedTrainStep(b, B, G, N, r2=0.06)
    loss_b = fb.step(pts.pin_memory(), order=order, want_host_loss=True)      # host batch in, loss out
    torch.cuda.synchronize()
    fb.flush_counters()
    assert abs(float(fb.loss_host[0]) - float(loss_b[2])) == 0.0
    assert abs(float(loss_b[2]) - loss_a) <= 1e-5 * abs(loss_a)
    assert abs(loss_a - z["loss64"][0]) <= 1e-3 * abs(z["loss64"][0])
    sa, sb = a.netR.state_dict(), b.netR.state_dict()
    for k in sa:
        if sa[k].dtype.is_floating_point:
            # identical kernels and inputs; only summation order differs (atomics, dx_global + dx_circle).  Adam's first
            # step moves an entry by lr * g / (|g| + eps), so entries with |g| ~ 0 may move by +-lr in either run
            if "running_" in k:
                assert rel2(sb[k], sa[k]) <= 1e-5, k
            else:
                assert float((sb[k] - sa[k]).abs().max()) <= 2.1 * 3e-4, k
        else:
            assert int(sa[k]) == int(sb[k]), k
    # weights after the Adam step vs the reference fixture (|delta| <= lr per entry on the first step)
    for k, v in sb.items():
        if "sd1_pos/" + k in z.files:
            got = v.reshape(-1).cpu().numpy()[z["sd1_pos/" + k]]
            assert np.allclose(got, z["sd1_val/" + k], rtol=0, atol=2.5 * 3e-4), k
    # a second step must run on the same buffers (persistent state, Adam step count 2)
    fb.step(pts.to(DEV), order=order)
    torch.cuda.synchronize()
    assert np.isfinite(float(fb.loss2[2]))
    # prefetched host batches (H2D of batch i+1 on a side stream during step i) give the same steps as device batches
    c, d = fresh(), fresh()
    fc, fd = FusedTrainStep(c, B, G, N, r2=0.06), FusedTrainStep(d, B, G, N, r2=0.06)
    hosts = [pts.pin_memory(), (pts * 0.5).pin_memory()]
    fc.prefetch(hosts[0])
    for i in range(4):
        lc = fc.step(hosts[i % 2], order=order, next_batch=hosts[(i + 1) % 2]).clone()
        ld = fd.step(hosts[i % 2].to(DEV), order=order).clone()
        # same kernels and inputs: the first step agrees to rounding; afterwards atomic summation order differs between the
        # two runs and Adam turns a sign flip of a ~0 gradient entry into a +-lr move, so later losses only agree loosely
        # (a wrong or stale batch would be off by far more: the two host batches differ by a factor 2)
        tol = 1e-5 if i == 0 else 2e-2
        assert float((lc - ld).abs().max()) <= tol * float(ld.abs().max()), (i, lc, ld)


@pytest.mark.parametrize("B,G,N,r2", [(4, 10, 512, 0.06), (3, 2, 256, 0.16), (5, 7, 1024, 0.06)])
def test_fused_step_other_shapes(B, G, N, r2):
    """The whole step (one C-ABI call, pinned host batch in, loss out) at the reference's own num_crop = 10 and at odd batch / view
    counts, two consecutive steps: losses against the fp32 CPU oracle run on the same weights, inputs and view orders."""
    from facl_b200 import synth
    from facl_b200.train import FusedTrainStep, TrainStep, default_opt
    tr = TrainStep(default_opt(batchSize=B, SAMPLE_NUM=N), num_crop=G, precision="fp32", radius2=r2, seed=2)
    sd = {k: v.detach().cpu().clone() for k, v in tr.netR.state_dict().items()}
    fused = FusedTrainStep(tr, B, G, N, r2=r2)
    state = {}
    for step in range(2):
        pts = torch.from_numpy(synth.make_sequences(B, G, N, seed=40 + step, skeleton=True))
        order = synth.view_order(G, 7 + step)
        got = fused.step(pts.pin_memory(), order=order, want_host_loss=True)
        torch.cuda.synchronize()
        ref = oracle.train_step(sd, pts, order, S=64, K=64, r2=r2, adam_state=state)
        state = ref["adam_state"]
        # step 0 is the parity check.  Step 1 runs on weights after one Adam step, whose first update is +-lr * sign(g): entries whose
        # gradient is ~0 move by +-3e-4 in either direction on rounding noise alone (see test_fused_step_equals_api_step_and_oracle), so
        # the second loss only has to show that the persistent buffers, the step count and the optimiser state carried over
        tol = 2e-3 if step == 0 else 3e-2
        assert abs(float(got[0]) - ref["loss_global"]) <= tol * abs(ref["loss_global"]) + 1e-4, (step, float(got[0]), ref["loss_global"])
        assert abs(float(got[1]) - ref["loss_circle"]) <= tol * abs(ref["loss_circle"]) + 1e-4, (step, float(got[1]), ref["loss_circle"])
        assert float(fused.loss_host[0]) == float(got[2])


def test_adam_state_dict_carries_step_and_fused_step_rebinds(golden_dir):
    """Checkpoint resume: facl_b200.optim.Adam keeps its step count in state_dict() like torch.optim.Adam (state[p]["step"]), and a
    FusedTrainStep bound to an optimiser whose state was replaced by load_state_dict() picks the new moment tensors up (their
    device pointers are baked into its table).  Two runs -- uninterrupted, and saved / reloaded after step 2 -- must end identical."""
    from facl_b200.train import FusedTrainStep, TrainStep
    z, sd0 = load_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    pts = torch.from_numpy(z["points"]).to(DEV)
    order = z["order"]

    def fresh():
        opt = make_opt(B, N, S, K)
        opt.learning_rate = 0.0003
        net = MODELL.PointNet_Plus_fine(opt, gost=G, sample_num_level1=S, knn_K=K)
        net.load_state_dict({k: v.clone() for k, v in sd0.items()})
        tr = TrainStep(opt, num_crop=G, precision="fp32", model=net)
        return tr, FusedTrainStep(tr, B, G, N, r2=0.06)

    ta, fa = fresh()
    for _ in range(2):
        fa.step(pts, order=order)
    torch.cuda.synchronize()
    osd = ta.optimizer.state_dict()
    steps = [int(st["step"]) for st in osd["state"].values()]
    assert steps and all(s == 2 for s in steps)
    # the state is interchangeable with torch.optim.Adam
    ref_opt = torch.optim.Adam(ta.netR.parameters(), lr=3e-4, betas=(0.5, 0.999), eps=1e-6)
    ref_opt.load_state_dict(osd)
    msd = {k: v.clone() for k, v in ta.netR.state_dict().items()}
    # resumed run: new model + optimiser objects, state loaded, then two more steps
    tb, fb = fresh()
    tb.netR.load_state_dict(msd)
    tb.optimizer.load_state_dict(osd)
    assert tb.optimizer._step == 2
    for _ in range(2):
        fa.step(pts, order=order)
        fb.step(pts, order=order)
    torch.cuda.synchronize()
    assert fb.args.step == 4 and tb.optimizer._step == 4
    sa, sb = ta.netR.state_dict(), tb.netR.state_dict()
    for k in sa:
        if sa[k].dtype.is_floating_point and "running_" not in k:
            # same kernels, same inputs, same optimiser state: only atomic summation order differs between the two runs
            assert float((sa[k] - sb[k]).abs().max()) <= 2.1 * 3e-4, k
    ma = ta.optimizer.state[ta.netR.netR_FC[3].weight]["exp_avg"]
    mb = tb.optimizer.state[tb.netR.netR_FC[3].weight]["exp_avg"]
    assert float((ma - mb).norm() / ma.norm()) <= 1e-2


def test_extract_features_layout(golden_dir):
    from facl_b200.train import extract_features
    z, sd0 = load_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    sd = {k: v.clone() for k, v in sd0.items()}
    oracle.train_step(sd, torch.from_numpy(z["points"]), z["order"], S=S, K=K, r2=float(z["r2"]))
    net, opt = _build(sd, B, G, N, S, K, "fp32")
    feat = extract_features(net, opt, torch.from_numpy(z["points"]).to(DEV))
    ref = z["eval_feat"].reshape(G + 1, B, 512).transpose(1, 0, 2).reshape(B, (G + 1) * 512)   # save_single_feature
    assert feat.shape == (B, (G + 1) * 512)
    assert rel2(feat, ref) <= 2e-3


def test_sharded_losses_sum_to_global(golden_dir):
    """Two emulated ranks (B=8 sharded 4+4) on one GPU: the per-rank loss shares and gradients of
    facl_contrast_losses(keys=all-gather) add up to the single-rank result on the same global batch."""
    import ctypes as C
    from facl_b200 import _lib
    from facl_b200.dist import key_index
    z = np.load(os.path.join(golden_dir, "losses.npz"))
    G, B, Cd = (int(v) for v in z["cfg_0"])
    x_ref, xg = torch.from_numpy(z["x_0"]).to(DEV), torch.from_numpy(z["xg_0"]).to(DEV)
    order = torch.as_tensor(np.asarray(z["order_0"], dtype=np.int32), device=DEV)
    L = _lib.lib()
    R, Bl = 2, B // 2
    Ml, Mk = G * Bl, G * B
    perm = torch.as_tensor([key_index(g, n, G, Bl) for g in range(G) for n in range(B)], device=DEV)   # ref row -> key row
    keys = torch.empty_like(x_ref)
    keys[perm] = x_ref
    tot_loss = torch.zeros(2, device=DEV)
    dkeys_sum = torch.zeros_like(keys)
    dx_anchor_all = torch.zeros_like(keys)
    dxg_all = torch.zeros_like(xg)
    for r in range(R):
        ws = torch.empty(L.facl_contrast_workspace_bytes(G, Bl, R, Cd), dtype=torch.uint8, device=DEV)
        x_loc = keys[r * Ml:(r + 1) * Ml].contiguous()
        xg_loc = xg[r * Bl:(r + 1) * Bl].contiguous()
        loss = torch.zeros(2, device=DEV)
        dxa, dxgl, dk = torch.empty_like(x_loc), torch.empty_like(xg_loc), torch.empty_like(keys)
        _lib.check(L.facl_contrast_losses(x_loc.data_ptr(), xg_loc.data_ptr(), keys.data_ptr(), G, B, Bl, r * Bl, Cd,
                                          order.data_ptr(), 1, 1, 3, ws.data_ptr(), loss.data_ptr(), dxa.data_ptr(),
                                          dxgl.data_ptr(), dk.data_ptr(), _lib.stream_ptr()))
        tot_loss += loss
        dkeys_sum += dk                                    # == reduce-scatter + sum
        dx_anchor_all[r * Ml:(r + 1) * Ml] = dxa
        dxg_all[r * Bl:(r + 1) * Bl] = dxgl
    dx_total = (dx_anchor_all + dkeys_sum)[perm]           # back to the reference row order
    assert abs(float(tot_loss[0]) - z["loss_0"][0]) <= 1e-4 * z["loss_0"][0]
    assert abs(float(tot_loss[1]) - z["loss_0"][1]) <= 1e-4 * z["loss_0"][1]
    assert rel2(dx_total, z["dx_0"]) <= 1e-3
    assert rel2(dxg_all, z["dxg_0"]) <= 1e-3


@pytest.mark.parametrize("prec", ["fp32", "bf16", "bf16_fast"])
def test_fused_l1_training_forward(golden_dir, prec):
    """Training-mode forward through the fused net3DV_1 kernels (closed-form BN1 statistics, pass A, pass B): features
    and running statistics against the fp64 reference fixture, and against the per-layer GEMM schedule."""
    z, sd0 = load_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    clouds = torch.from_numpy(z["points"]).permute(1, 0, 2, 3).reshape(-1, N, 4).to(DEV)
    outs = {}
    for fused in (True, False):
        net, opt = _build(sd0, B, G, N, S, K, prec)
        net.fused_l1 = fused
        net.train()
        with torch.no_grad():
            xt, yt = utils_my.group_points_3DV(clouds, opt)
            x, code, x_nor, xg = net(xt, yt, 1)
        outs[fused] = (x, xg, {k: v.clone() for k, v in net.state_dict().items()})
    x, xg, sd = outs[True]
    ft = TOL_FEAT[prec]
    assert rel2(x, z["x64"]) <= ft, rel2(x, z["x64"])
    assert rel2(xg, z["x_global64"]) <= 2 * ft
    assert rel2(x, outs[False][0]) <= ft
    sd64 = {k: (v.clone().double() if v.dtype.is_floating_point else v.clone()) for k, v in sd0.items()}
    oracle.train_step(sd64, torch.from_numpy(z["points"]), z["order"], S=S, K=K, r2=float(z["r2"]), apply_update=False,
                      dtype=torch.float64)
    for k, v in sd.items():
        if "running_" in k:
            assert rel2(v, sd64[k]) <= max(ft, 1e-3) * 2, (k, rel2(v, sd64[k]))
        if "num_batches" in k:
            assert int(v) == int(sd64[k])


def _check_fused_backward(pts, sd0, order, B, G, N, S, K, r2, tol, prec="fp32"):
    from facl_b200.debug import L1DecisionDump, routing_of_last_forward
    clouds = pts.permute(1, 0, 2, 3).reshape(-1, N, 4).to(DEV)
    net, opt = _build(sd0, B, G, N, S, K, prec)
    assert net.fused_l1 and (net._flags(True, True) & 1), "this geometry must run the fused net3DV_1 kernels"
    net.train()
    xt, yt = utils_my._group(clouds, S, K, r2)
    x, code, x_nor, xg = net(xt, yt, 1)
    x.retain_grad()
    xg.retain_grad()
    lg, lc = facl_losses.contrast_losses(x, xg, G, B, order=order, prec=prec)
    with L1DecisionDump(G * B, S, K) as dump:
        (lg + lc).backward()
    routing = routing_of_last_forward(net, l1_dump=dump)
    sdr = {k: (v.clone().double() if v.dtype.is_floating_point else v.clone()) for k, v in sd0.items()}
    # bf16 mode is checked stage-wise (see test_train_step_vs_golden_and_oracle): the oracle back-propagates the CUDA path's own dL/dx
    upstream = None if prec == "fp32" else (x.grad.detach().cpu(), xg.grad.detach().cpu())
    ort = oracle.train_step(sdr, pts, order, S=S, K=K, r2=r2, apply_update=False, dtype=torch.float64, routing=routing,
                            upstream=upstream)
    assert rel2(x, ort["x"]) <= TOL_FEAT[prec] and abs(float(lg + lc) - ort["loss"]) <= TOL[prec] * abs(ort["loss"])
    gscale = max(float(g.abs().max()) for g in ort["grads"].values())
    report = []
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        ref = ort["grads"][k].reshape(p.shape)
        if float(ref.norm()) <= 1e-9 * gscale * ref.numel() ** 0.5:
            assert float(p.grad.abs().max()) <= 1e-4 * gscale, k
            continue
        rms_ref = float(ref.double().norm()) / ref.numel() ** 0.5
        rms_err = float((p.grad.detach().double().cpu().reshape(-1) - ref.double().reshape(-1)).norm()) / ref.numel() ** 0.5
        report.append((k, rel2(p.grad, ref), rms_ref / gscale, rms_err / gscale))
    print("\n" + "\n".join(f"{k:24s} fused, matched decisions: err {a:.2e}  (rms/gscale: ref {r:.2e} err {e:.2e})" for k, a, r, e in report))
    for k, a, r, e in report:
        # A gradient that is structurally a near-zero residual (the BN beta in front of a layer whose own BatchNorm removes
        # constants: its rms is 1e-5 of the largest gradient) is judged on its absolute error, which is the smallest of all.
        assert a <= tol or e <= tol * 1e-3, (k, a, r, e)


def test_fused_l1_backward(golden_dir):
    """Gradients of the fused net3DV_1 path (activations recomputed in the backward, dW1 in closed form) against the
    fp64 oracle under the SAME discrete decisions: the recomputing backward records its ReLU patterns and max-pool
    winners through the facl_debug_l1_dump test hook."""
    z, sd0 = load_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    _check_fused_backward(torch.from_numpy(z["points"]), sd0, z["order"], B, G, N, S, K, float(z["r2"]), TOL_GRAD["fp32"])


@pytest.mark.parametrize("fused", [True, False])
def test_fine_default_geometry(golden_dir, fused):
    """PointNet_Plus_fine's DEFAULT geometry, sample_num_level1 = 32 and knn_K = 128 (cn3d_model_conbag.py:142): a max-pool group
    spans two 64-row tiles of the fused net3DV_1 backward.  Features / loss against the reference's fp64 run (fixture
    fine_geometry.npz), gradients against the fp64 oracle under the CUDA path's own discrete decisions; fused and per-layer."""
    from facl_b200.debug import L1DecisionDump, routing_of_last_forward
    z = np.load(os.path.join(golden_dir, "fine_geometry.npz"))
    sd0 = oracle.init_state_dict(seed=int(z["seed_sd"]))
    for k in list(sd0):
        if "sd0/" + k in z.files:
            sd0[k] = torch.from_numpy(z["sd0/" + k]).clone()
    B, G, N, S, K = (int(v) for v in z["cfg"])
    pts, order, r2 = torch.from_numpy(z["points"]), z["order"], float(z["r2"])
    opt = make_opt(B, N, S, K)
    net = MODELL.PointNet_Plus_fine(opt, gost=G)                               # default arguments: S1 = 32, K = 128
    assert (net.sample_num_level1, net.knn_K) == (S, K) == (32, 128)
    net.load_state_dict({k: v.clone() for k, v in sd0.items()})
    net = net.to(DEV)
    net.fused_l1 = fused
    assert bool(net._flags(True, True) & 1) == fused
    net.train()
    clouds = pts.permute(1, 0, 2, 3).reshape(-1, N, 4).to(DEV)
    xt, yt = utils_my.group_points_3DV_nums(clouds, opt, S, K)
    x, code, x_nor, xg = net(xt, yt, 1)
    lg, lc = facl_losses.contrast_losses(x, xg, G, B, order=order)
    if fused:
        with L1DecisionDump(G * B, S, K) as dump:
            (lg + lc).backward()
        routing = routing_of_last_forward(net, l1_dump=dump)
    else:
        (lg + lc).backward()
        torch.cuda.synchronize()
        routing = routing_of_last_forward(net)
    assert rel2(x, z["x64"]) <= 1e-3 and rel2(xg, z["x_global64"]) <= 2e-3
    assert abs(float(lg + lc) - float(z["loss64"])) <= 1e-3 * abs(float(z["loss64"]))
    sdr = {k: (v.clone().double() if v.dtype.is_floating_point else v.clone()) for k, v in sd0.items()}
    ort = oracle.train_step(sdr, pts, order, S=S, K=K, r2=r2, apply_update=False, dtype=torch.float64, routing=routing)
    gscale = max(float(g.abs().max()) for g in ort["grads"].values())
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        ref = ort["grads"][k].reshape(p.shape)
        if float(ref.norm()) <= 1e-9 * gscale * ref.numel() ** 0.5:
            assert float(p.grad.abs().max()) <= 1e-4 * gscale, k
            continue
        rms_err = float((p.grad.detach().double().cpu().reshape(-1) - ref.double().reshape(-1)).norm()) / ref.numel() ** 0.5
        assert rel2(p.grad, ref) <= 1e-3 or rms_err / gscale <= 1e-6, (k, rel2(p.grad, ref))


def test_fused_l1_backward_zero_bn_weight(golden_dir):
    """A BatchNorm weight that is EXACTLY zero in net3DV_1's second layer: the fused backward cannot derive that channel's
    sum dh2' z2 from the (constant) activation and recomputes it from the inputs (l1_gamma0_fix_kernel); a zero gamma in the first
    and third layer exercises the sign / max-vs-min selection of the pooled layer.  Same matched-decision check as above."""
    z, sd0 = load_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    sd0 = {k: v.clone() for k, v in sd0.items()}
    sd0["net3DV_1.4.weight"][5] = 0.0
    sd0["net3DV_1.4.weight"][40] = 0.0
    sd0["net3DV_1.4.bias"][40] = -0.3            # one degenerate channel inactive everywhere, one active everywhere
    sd0["net3DV_1.4.bias"][5] = 0.25
    sd0["net3DV_1.1.weight"][7] = 0.0
    sd0["net3DV_1.7.weight"][9] = 0.0
    _check_fused_backward(torch.from_numpy(z["points"]), sd0, z["order"], B, G, N, S, K, float(z["r2"]), TOL_GRAD["fp32"])


def test_fused_l1_backward_bf16_8x20x2048():
    """bf16 mode (the arithmetic of BASELINE configs[2]) at 8 sequences x 20 views x 2048 points: features and loss within 2e-2 of
    the fp64 oracle, every parameter gradient within 2e-2 of the oracle's back-propagation of the same upstream gradient under
    the same discrete decisions."""
    from facl_b200 import synth
    B, G, N, S, K = 8, 20, 2048, 64, 64
    pts = torch.from_numpy(synth.make_sequences(B, G, N, seed=31))
    _check_fused_backward(pts, oracle.init_state_dict(seed=12), synth.view_order(G, 5), B, G, N, S, K, 0.06, TOL_GRAD["bf16"], prec="bf16")


@pytest.mark.parametrize("B", [8, 64])
def test_fused_l1_backward_long_accumulation(B):
    """The same matched-decision check at B sequences x 20 views x 2048 points.  B = 64 is BASELINE configs[1] at FULL size
    (5 242 880 grouped rows, ~550 tiles per CTA: the TMEM-resident Gram / weight-gradient accumulators and the per-thread
    BatchNorm sums run over thousands of steps); the fp64 oracle needs ~40 s and ~100 GB of host memory for it, so that case is
    skipped on hosts with less.  Measured on the 196 GB B200 box: every gradient within 1.7e-4 of the oracle (typically 5e-5)."""
    import psutil
    from facl_b200 import synth
    if B > 8 and psutil.virtual_memory().available < 150 * 2 ** 30:
        pytest.skip("the fp64 oracle of the full-size step needs ~100 GB of host memory")
    G, N, S, K = 20, 2048, 64, 64
    pts = torch.from_numpy(synth.make_sequences(B, G, N, seed=31))
    _check_fused_backward(pts, oracle.init_state_dict(seed=12), synth.view_order(G, 5), B, G, N, S, K, 0.06, TOL_GRAD["fp32"])


def test_full_size_fused_vs_per_layer_schedule():
    """BASELINE configs[1] at FULL size (64 sequences x 20 views x 2048 points, fp32 mode): the fused net3DV_1 path (recompute,
    closed-form BN1 / dW1, 64x64 backward algebra, Gram matrices accumulated over ~550 tiles per CTA, TMEM-resident weights)
    against the per-layer GEMM schedule of the same library, which stores every activation and shares none of that algebra.
    Both see the same weights and batch; loss, features, running statistics and every parameter gradient must agree.
    (The oracle cannot run this size in seconds; it pins both schedules at the fixture size in the tests above.)"""
    B, G, N, S, K = 64, 20, 2048, 64, 64
    from facl_b200 import synth
    sd0 = oracle.init_state_dict(seed=11)
    clouds = torch.from_numpy(synth.make_sequences(B, G, N, seed=21)).permute(1, 0, 2, 3).reshape(-1, N, 4).contiguous().to(DEV)
    order = synth.view_order(G, 3)
    res = {}
    for fused in (True, False):
        net, opt = _build(sd0, B, G, N, S, K, "fp32")
        net.fused_l1 = fused
        net.train()
        xt, yt = utils_my.group_points_3DV_2048(clouds, K, S)
        x, code, x_nor, xg = net(xt, yt, 1)
        lg, lc = facl_losses.contrast_losses(x, xg, G, B, order=order, prec="fp32")
        (lg + lc).backward()
        torch.cuda.synchronize()
        res[fused] = dict(loss=float(lg + lc), x=x.detach().clone(), xg=xg.detach().clone(),
                          grads={k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None},
                          sd={k: v.detach().clone() for k, v in net.state_dict().items()})
        del net, xt, yt, x, xg, lg, lc
        torch.cuda.empty_cache()
    a, b = res[True], res[False]
    assert np.isfinite(a["loss"]) and abs(a["loss"] - b["loss"]) <= 1e-4 * abs(b["loss"]), (a["loss"], b["loss"])
    assert rel2(a["x"], b["x"]) <= 1e-3 and rel2(a["xg"], b["xg"]) <= 1e-3
    # Gradients: the two schedules round differently, so a few of the 5 M x 64 ReLU decisions and of the max-pool winners (over
    # 64 neighbours, 64 centres, 20 views) flip between them, and each flip re-routes a gradient.  Measured at 8 sequences, where
    # the fp64 oracle runs: fused vs oracle, per-layer vs oracle and fused vs per-layer ALL differ by 0.7-1.3e-2 (without
    # matched decisions), while with matched decisions the fused path is within 1.5e-4 of the oracle (test_fused_l1_backward).
    # This full-size check therefore bounds gross errors of scale (index overflow, accumulation blow-up), not rounding.
    gscale = max(float(g.abs().max()) for g in b["grads"].values())
    worst = []
    for k, g in b["grads"].items():
        if float(g.norm()) <= 1e-6 * gscale * g.numel() ** 0.5:            # zero by construction (conv bias in front of a BN)
            assert float(a["grads"][k].abs().max()) <= 1e-4 * gscale, k
            continue
        worst.append((rel2(a["grads"][k], g), k))
    print("\n" + "\n".join(f"{k:24s} fused vs per-layer, full size: {e:.2e}" for e, k in sorted(worst, reverse=True)[:8]))
    for e, k in worst:
        assert e <= 3e-2, (k, e)
    for k, v in b["sd"].items():
        if "running_" in k:
            assert rel2(a["sd"][k], v) <= 1e-3, k
