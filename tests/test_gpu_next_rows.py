"""GPU parity of the SURVEY section 8 (f) rows: device-side view augmentation (f1) and level-2 grouping (f2).
Checker: oracle/ (CPU) and the fixtures the unmodified reference produced (tests/golden/make_golden.py)."""
import os
import types

import numpy as np
import pytest
import torch

import oracle
from oracle import augment as oaug
from facl_b200 import cn3D_data_set as ds
from facl_b200 import ops, synth, utils_my

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sorted_cols(a):
    M, C, S, K = a.shape
    rows = np.ascontiguousarray(a.transpose(0, 2, 3, 1)).reshape(M * S, K, C)
    out = np.empty_like(rows)
    for i in range(rows.shape[0]):
        r = rows[i]
        out[i] = r[np.lexsort(r.T[::-1])]
    return out.reshape(M, S, K, C)


# ------------------------------------------------------------------------------------------- f1: augmentation
def _ragged(list_of_arrays):
    return ds.Ragged.from_list([torch.from_numpy(a) for a in list_of_arrays], DEV)


def _check_views(got, want):
    rot = [g for g, r in enumerate(ds.GET_DATA_TRAIN) if r[5]]
    exact = [g for g in range(len(ds.GET_DATA_TRAIN)) if g not in rot]
    for g in exact:                                                    # gather / jitter / mirror: IEEE f64, bit-exact
        assert np.array_equal(got[g], want[g]), f"view {g}: max diff {np.abs(got[g] - want[g]).max()}"
    # rotation: cos/sin of the device vs libm may differ in the last f64 bit -> at most one f32 ulp after rounding
    assert np.abs(got[rot] - want[rot]).max() <= 6e-8


def test_augment_golden(golden_dir):
    """The whole batch of golden sequences in ONE launch, explicit draws = the recorded numpy stream."""
    z = np.load(os.path.join(golden_dir, "augment.npz"))
    n = int(z["n_cases"])
    srcs = [_ragged([z[f"{k}_{i}"] for i in range(n)]) for k in ("points", "key", "res1", "res2")]
    draws = ds.Draws(torch.from_numpy(np.stack([z[f"idx_{i}"] for i in range(n)])).to(DEV),
                     torch.from_numpy(np.stack([z[f"noise_{i}"] for i in range(n)])).to(DEV),
                     torch.from_numpy(np.stack([z[f"angle_{i}"] for i in range(n)])).to(DEV))
    aug = ds.ViewAugmenter()
    bg = aug.get_data_train(srcs, draws, g_major=False).cpu().numpy()            # (B,G,N,4)
    gm = aug.get_data_train(srcs, draws, g_major=True).cpu().numpy()             # (G*B,N,4)
    assert np.array_equal(gm.reshape(10, n, 512, 4).transpose(1, 0, 2, 3), bg)   # cn3d_train_motion_GL.py:226
    for i in range(n):
        _check_views(bg[i], z[f"views_{i}"])


def test_augment_vs_oracle_ragged_batch():
    rng = np.random.default_rng(3)
    B, N = 7, 300
    lens = rng.integers(40, 3000, size=(4, B))

    def cloud(n):
        a = rng.uniform(-0.5, 0.5, (n, 8)).astype(np.float32)
        a[rng.uniform(size=n) < 0.5, 4] = 0
        a[rng.uniform(size=n) < 0.9, 7] = 0
        a[0, 7] = 0.25                                  # at least one drawable row
        return a
    src_np = [[cloud(int(lens[s, b])) for b in range(B)] for s in range(4)]
    draws_np = [oaug.record_draws(np.random.RandomState(b), [src_np[s][b] for s in range(4)], N) for b in range(B)]
    draws = ds.Draws(torch.from_numpy(np.stack([d.idx for d in draws_np])).to(DEV),
                     torch.from_numpy(np.stack([d.noise for d in draws_np])).to(DEV),
                     torch.from_numpy(np.stack([d.angle_u for d in draws_np])).to(DEV))
    aug = ds.ViewAugmenter(num_point=N)
    got, rows = aug.get_data_train([_ragged(s) for s in src_np], draws, g_major=False, want_rows=True)
    got, rows = got.cpu().numpy(), rows.cpu().numpy()
    for b in range(B):
        _check_views(got[b], oaug.make_views([src_np[s][b] for s in range(4)], draws_np[b]))
        # temporal views: the compaction keeps row order (np.where) and only non-zero rows
        nz = oaug.nonzero_rows(src_np[0][b], 4)
        assert np.array_equal(rows[b, 6], nz[draws_np[b].idx[6]])


def test_augment_device_rng():
    """No draws given: Philox on the device.  Same distributions as the reference's numpy draws, reproducible per
    (seed, step), and the transforms are those of the recipe."""
    rng = np.random.default_rng(4)
    B, N = 16, 2048
    src_np = [[rng.uniform(-0.5, 0.5, (2048, 8)).astype(np.float32) for _ in range(B)] for _ in range(4)]
    srcs = [_ragged(s) for s in src_np]
    a1, a2 = ds.ViewAugmenter(num_point=N, seed=11), ds.ViewAugmenter(num_point=N, seed=11)
    v1, r1 = a1.get_data_train(srcs, g_major=False, want_rows=True)
    v2, _ = a2.get_data_train(srcs, g_major=False, want_rows=True)
    assert torch.equal(v1, v2)                                           # same (seed, step) -> same views
    v3 = a1.get_data_train(srcs, g_major=False)
    assert not torch.equal(v1, v3)                                       # next step -> new draws
    v1, r1 = v1.cpu().numpy().astype(np.float64), r1.cpu().numpy()
    assert r1.min() >= 0 and r1.max() < 2048
    cnt = np.bincount(r1[:, 0].ravel(), minlength=2048)                  # resampling is uniform over the rows
    assert abs(cnt.mean() - B * N / 2048) < 1e-9 and cnt.std() < 2.0 * np.sqrt(B * N / 2048)
    for b in range(B):
        pts = src_np[0][b].astype(np.float64)
        assert np.array_equal(v1[b, 0], pts[r1[b, 0], :4])               # raw view: pure gather
        j = v1[b, 2, :, :3] - src_np[1][b][r1[b, 2], :3]                  # jitter only (key points)
        assert np.abs(j).max() <= 0.05 + 1e-6
        if b == 0:
            inside = j[np.abs(j) < 0.049]
            assert abs(inside.std() - 0.01) < 5e-4 and abs(inside.mean()) < 5e-4
        m = v1[b, 1, :, :3] * [-1, 1, 1] - pts[r1[b, 1], :3] * [1, 1, 1]  # mirrored view: |x' + x| within two jitters
        assert np.abs(m).max() <= 0.1 + 1e-6
        ro = v1[b, 4]                                                    # rotation about y keeps y and the xz radius
        base = pts[r1[b, 4], :3]
        assert np.abs(ro[:, 1] - base[:, 1]).max() <= 0.05 + 1e-6
        assert np.abs(np.hypot(ro[:, 0], ro[:, 2]) - np.hypot(base[:, 0], base[:, 2])).max() <= 0.08
    # rotation angles spread over (-0.4 pi, 0.4 pi)
    ang = []
    for b in range(B):
        base = src_np[0][b][r1[b, 4], :3].astype(np.float64)
        big = np.hypot(base[:, 0], base[:, 2]) > 0.3
        d = np.arctan2(v1[b, 4][big, 2], v1[b, 4][big, 0]) - np.arctan2(base[big, 2], base[big, 0])
        ang.append(np.median((d + np.pi) % (2 * np.pi) - np.pi))
    assert np.abs(ang).max() <= 0.4 * np.pi + 0.1 and np.std(ang) > 0.2


def test_augment_empty_temporal_view_is_nan():
    src = np.random.default_rng(5).uniform(-0.5, 0.5, (64, 8)).astype(np.float32)
    src[:, 7] = 0                                                        # numpy would raise in randint(0, 0)
    srcs = [_ragged([src])] * 4
    v = ds.ViewAugmenter(num_point=32).get_data_train(srcs, g_major=False)
    assert torch.isnan(v[0, 7]).all() and not torch.isnan(v[0, :7]).any()


def test_fps_sample_data_mirror():
    pts = synth.make_sequences(5, 1, 512, seed=8)[:, 0]
    starts = np.arange(5, dtype=np.int32) * 7
    got = ds.fps_sample_data(torch.from_numpy(pts).to(DEV), 64, 64, torch.from_numpy(starts).to(DEV)).cpu().numpy()
    assert np.array_equal(got, oracle.fps_sample_data(pts.copy(), 64, list(starts)))


# ------------------------------------------------------------------------------------------- f2: level-2 grouping
def test_group_level2_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "group2.npz"))
    for name in z["names"]:
        feats = torch.from_numpy(z[f"{name}_feats"]).to(DEV)
        S2, K = (int(v) for v in z[f"{name}_cfg"])
        out, _ = ops.group_level2(feats, S2, K, float(z[f"{name}_r2"]))
        assert np.array_equal(_sorted_cols(out.cpu().numpy()), z[f"{name}_sorted"]), name


@pytest.mark.parametrize("M,C,S1,S2,K,r2", [(6, 131, 512, 128, 64, 0.02), (3, 259, 64, 64, 32, 0.11), (2, 5, 1000, 10, 16, 0.001),
                                            (1, 3, 2048, 64, 128, 0.16)])
def test_group_level2_vs_oracle(M, C, S1, S2, K, r2):
    g = torch.Generator().manual_seed(S1 + C)
    xyz = torch.from_numpy(synth.make_sequences(M, 1, S1, seed=S1 + K, skeleton=True)[:, 0, :, 0:3]).permute(0, 2, 1)
    feats = torch.cat([xyz, torch.randn(M, C - 3, S1, generator=g)], 1).contiguous()
    out, idx = ops.group_level2(feats.to(DEV), S2, K, r2, want_idx=True)
    oout, ocentre, oidx = oracle.group_points_level2(feats, S2, K, r2)
    assert np.array_equal(idx.cpu().numpy(), oidx.numpy())
    assert np.array_equal(out.cpu().numpy(), oout.numpy())


def test_group_points_2_mirrors():
    """The reference-named entry points (their hard-coded K / radius included)."""
    g = torch.Generator().manual_seed(1)
    xyz = torch.from_numpy(synth.make_sequences(2, 1, 256, seed=3)[:, 0, :, 0:3]).permute(0, 2, 1)
    feats = torch.cat([xyz, torch.randn(2, 8, 256, generator=g)], 1).contiguous()
    a, ac = utils_my.group_points_2(feats.to(DEV), 256, 64, 999, torch.tensor(0.03))
    b, bc = utils_my.group_points_2_3DV(feats.to(DEV), 256, 64, 999, torch.tensor(123.0))
    oa, oc, _ = oracle.group_points_level2(feats, 64, 64, 0.03)
    ob, _, _ = oracle.group_points_level2(feats, 64, 32, 0.11)
    assert a.shape == (2, 11, 64, 64) and b.shape == (2, 11, 64, 32) and ac.shape == (2, 3, 64, 1)
    assert np.array_equal(a.cpu().numpy(), oa.numpy()) and np.array_equal(b.cpu().numpy(), ob.numpy())
    assert torch.equal(ac.cpu(), oc) and torch.equal(bc.cpu(), oc)


# ------------------------------------------------------------------------------------------- f3: linear probe + files
def test_probe_golden(golden_dir):
    """Final_FC + CrossEntropy + Adam against the fixture the reference's linear_classify code produced."""
    from facl_b200 import fc_model, linercls
    z = np.load(os.path.join(golden_dir, "probe.npz"))
    net = fc_model.Final_FC(input_dim=512, gost=2, num_class=20)
    assert list(net.state_dict().keys()) == ["fc.weight", "fc.bias"]
    net.load_state_dict({"fc.weight": torch.from_numpy(z["w0"]), "fc.bias": torch.from_numpy(z["b0"])})
    net = net.to(DEV)
    # module API (autograd) on the first batch
    x0, y0 = torch.from_numpy(z["x"][0]).to(DEV), torch.from_numpy(z["y"][0]).to(DEV)
    logits = net(x0)
    torch.nn.functional.cross_entropy(logits, y0).backward()
    assert np.abs(logits.detach().cpu().numpy() - z["logits0"]).max() <= 1e-3 * np.abs(z["logits0"]).max()
    gw = net.fc.weight.grad.cpu().numpy()
    assert np.abs(gw - z["grad_w0"]).max() <= 1e-3 * np.abs(z["grad_w0"]).max()
    assert np.abs(net.fc.bias.grad.cpu().numpy() - z["grad_b0"]).max() <= 1e-3 * np.abs(z["grad_b0"]).max()
    # fused loop body, three steps
    tr = linercls.ProbeTrainer(net)
    for it in range(3):
        tr.step(torch.from_numpy(z["x"][it]).to(DEV), torch.from_numpy(z["y"][it]).to(DEV))
        loss, top1 = tr.pop_meters()
        assert abs(loss - z["losses"][it]) <= 1e-3 * z["losses"][it], (it, loss, z["losses"][it])
        assert abs(top1 - z["top1"][it]) < 1e-3
    # Adam moves every weight by ~lr per step whatever the gradient size: compare where the reference moved clearly
    w3 = net.fc.weight.detach().cpu().numpy()
    assert np.abs(w3 - z["w3"]).mean() <= 2e-4 and np.mean(np.abs(w3 - z["w3"]) > 2e-3) < 1e-2


def test_probe_full_size_vs_oracle():
    from oracle import probe as oprobe
    from facl_b200 import linercls
    g = torch.Generator().manual_seed(2)
    x = torch.randn(64, 22 * 512, generator=g)
    y = torch.randint(0, 120, (64,), generator=g)
    tr = linercls.ProbeTrainer()
    sd = {k: v.detach().cpu().clone() for k, v in tr.netR.state_dict().items()}
    logits = tr.step(x.to(DEV), y.to(DEV))
    loss, top1 = tr.pop_meters()
    o = oprobe.probe_step(sd, x, y, {})
    assert float((logits.cpu() - o["logits"]).abs().max()) <= 1e-3 * float(o["logits"].abs().max())
    assert abs(loss - o["loss"]) <= 1e-3 * o["loss"] and abs(top1 - o["top1"]) < 1e-3
    gw = tr.netR.fc.weight.grad.cpu()
    assert float((gw - o["grads"]["fc.weight"]).abs().max()) <= 1e-3 * float(o["grads"]["fc.weight"].abs().max())
    ev = tr.evaluate(x.to(DEV), y.to(DEV))
    assert ev.shape == (64, 120) and tr.pop_meters()[1] >= top1


def test_feature_files_roundtrip(golden_dir, tmp_path):
    from facl_b200 import features
    z = np.load(os.path.join(golden_dir, "probe.npz"))
    names = ["a", "b", "c"]
    features.save_single_feature(torch.from_numpy(z["feat"]).to(DEV), str(tmp_path) + "/", names, num_crop=11)
    assert np.array_equal(np.frombuffer(open(tmp_path / "a.npy", "rb").read(), dtype=np.uint8), z["file0"])   # byte-identical file
    batch = features.load_batch([(str(tmp_path / "a.npy"), str(tmp_path / "b.npy")), (str(tmp_path / "c.npy"), str(tmp_path / "a.npy"))])
    assert batch.shape == (2, 22 * 512)
    assert np.array_equal(batch[0].numpy(), np.concatenate([z["feat_rows"][0], z["feat_rows"][1]]))


# ------------------------------------------------------------------------------------------- f4: disabled loss heads
def test_info_nce_golden_and_gradient(golden_dir):
    z = np.load(os.path.join(golden_dir, "train_step.npz"))
    B = int(z["cfg"][0])
    x = torch.from_numpy(z["x"])[: 2 * B]
    opt = types.SimpleNamespace(batchSize=B)
    xd = x.to(DEV).requires_grad_(True)
    logits, labels = utils_my.Info_NCE(xd, opt)
    assert logits.shape == (B, 1 + 4 * B) and labels.dtype == torch.long and int(labels.abs().sum()) == 0
    ref = z["info_nce_logits"]
    assert np.abs(logits.detach().cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max()
    torch.nn.functional.cross_entropy(logits, labels).backward()
    xo = x.clone().requires_grad_(True)
    ol, olab = oracle.info_nce_logits(xo, B)
    torch.nn.functional.cross_entropy(ol, olab).backward()
    assert float((xd.grad.cpu() - xo.grad).abs().max()) <= 1e-3 * float(xo.grad.abs().max())
    # a batch-sized case against the oracle
    B = 48
    x = torch.randn(2 * B, 512, generator=torch.Generator().manual_seed(4)) * 0.2
    xd, xo = x.to(DEV).requires_grad_(True), x.clone().requires_grad_(True)
    logits, labels = utils_my.Info_NCE(xd, types.SimpleNamespace(batchSize=B))
    ol, olab = oracle.info_nce_logits(xo, B)
    assert float((logits.detach().cpu() - ol.detach()).abs().max()) <= 1e-3 * float(ol.abs().max())
    torch.nn.functional.cross_entropy(logits, labels).backward()
    torch.nn.functional.cross_entropy(ol, olab).backward()
    assert float((xd.grad.cpu() - xo.grad).abs().max()) <= 1e-3 * float(xo.grad.abs().max())


# ------------------------------------------------------------------------------------------- f1 -> the training step
def test_augmented_views_feed_the_training_step():
    """Loader-free pipeline: resident source clouds -> facl_augment_views -> facl_train_step; the loss equals the oracle's
    step on the oracle's views (same recorded draws)."""
    from facl_b200 import cn3d_model_conbag as MODELL
    from facl_b200.train import FusedTrainStep, TrainStep
    rng = np.random.default_rng(12)
    B, G, N, S, K = 4, 10, 512, 64, 64

    def cloud(n):
        a = np.empty((n, 8), np.float32)
        a[:, 0], a[:, 1], a[:, 2] = 0.45 * rng.uniform(-0.5, 0.5, n), rng.uniform(-0.5, 0.5, n), 0.3 * rng.uniform(-0.5, 0.5, n)
        a[:, 3:] = rng.uniform(-0.5, 0.5, (n, 5))
        a[rng.uniform(size=n) < 0.3, 4] = 0
        a[rng.uniform(size=n) < 0.3, 7] = 0
        return a
    src_np = [[cloud(int(n)) for n in rng.integers(600, 2500, size=B)] for _ in range(4)]
    draws_np = [oaug.record_draws(np.random.RandomState(100 + b), [src_np[s][b] for s in range(4)], N) for b in range(B)]
    draws = ds.Draws(torch.from_numpy(np.stack([d.idx for d in draws_np])).to(DEV),
                     torch.from_numpy(np.stack([d.noise for d in draws_np])).to(DEV),
                     torch.from_numpy(np.stack([d.angle_u for d in draws_np])).to(DEV))
    views = ds.ViewAugmenter(num_point=N).get_data_train([_ragged(s) for s in src_np], draws, g_major=False)      # (B,G,N,4)
    oviews = np.stack([oaug.make_views([src_np[s][b] for s in range(4)], draws_np[b]) for b in range(B)])
    assert np.abs(views.cpu().numpy() - oviews).max() <= 6e-8

    sd0 = oracle.init_state_dict(seed=5)
    opt = types.SimpleNamespace(temperal_num=3, knn_K=K, ball_radius=0.16, ball_radius2=0.25, sample_num_level1=S,
                                sample_num_level2=64, INPUT_FEATURE_NUM=4, Num_Class=512, batchSize=B,
                                pooling="concatenation", SAMPLE_NUM=N, learning_rate=0.0003)
    net = MODELL.PointNet_Plus_fine(opt, gost=G, sample_num_level1=S, knn_K=K)
    net.load_state_dict({k: v.clone() for k, v in sd0.items()})
    step = FusedTrainStep(TrainStep(opt, num_crop=G, precision="fp32", model=net), B, G, N, r2=0.06)
    order = np.arange(G)[::-1].copy()
    loss = step.step(views, order=order)
    torch.cuda.synchronize()
    want = oracle.train_step({k: v.clone() for k, v in sd0.items()}, torch.from_numpy(oviews), order, S=S, K=K, r2=0.06,
                             apply_update=False)
    assert abs(float(loss[2]) - want["loss"]) <= 1e-3 * abs(want["loss"]), (float(loss[2]), want["loss"])


# ------------------------------------------------------------------------------------------------- disabled loss heads (f4)
def test_sinkhorn_swav_cld_heads_match_reference(golden_dir):
    """distributed_sinkhorn / shoot_infs (cn3d_model_conbag.py:391-425), the SwAV block of
    cn3d_train_motion_GL.py:236-262 and utils_my.CLD_Loss / grouping / KMeans (:152-198) on the GPU, against outputs of the reference
    functions themselves (heads.npz).  k-means labels are index work: bit-exact."""
    import types
    from facl_b200 import cn3d_model_conbag as MODELL, heads, utils_my
    z = np.load(os.path.join(golden_dir, "heads.npz"))

    def rel2(a, b):
        a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
        b = torch.as_tensor(b).double().reshape(-1)
        return float((a - b).norm() / b.norm())

    for i in range(int(z["sk_n"])):
        q = torch.from_numpy(z[f"sk_q_{i}"]).to(DEV)
        got = MODELL.distributed_sinkhorn(q, 3)
        assert np.allclose(got.cpu().numpy(), z[f"sk_out_{i}"], rtol=2e-5, atol=1e-7), i
    t = torch.tensor([1.0, float("inf"), 3.0, -float("inf")], device=DEV)
    assert MODELL.shoot_infs(t).tolist() == [1.0, 3.0, 3.0, 3.0]
    # SwAV: x -> normalize -> mapping -> Sinkhorn targets / soft cross-entropy, gradients to x and mapping.weight
    G, B = (int(v) for v in z["swav_cfg"])
    x = torch.from_numpy(z["swav_x"]).to(DEV).requires_grad_(True)
    w = torch.from_numpy(z["swav_w"]).to(DEV).requires_grad_(True)
    x_nor, code = heads.normalized_code(x, w)
    assert rel2(code, z["swav_code"]) <= 1e-4
    loss = heads.swav_loss(code, G, B)
    loss.backward()
    assert abs(float(loss) - float(z["swav_loss"])) <= 1e-3 * abs(float(z["swav_loss"]))
    assert rel2(x.grad, z["swav_dx"]) <= 2e-3 and rel2(w.grad, z["swav_dw"]) <= 2e-3
    # CLD
    G, B = (int(v) for v in z["cld_cfg"])
    f = torch.from_numpy(z["cld_x"]).to(DEV).requires_grad_(True)
    labels, cent = utils_my.KMeans(f.detach()[: 3 * B], 60, 5)
    assert np.array_equal(labels.cpu().numpy(), z["km_labels"])
    assert np.allclose(cent.cpu().numpy(), z["km_centroids"], rtol=1e-5, atol=1e-6)
    loss = utils_my.CLD_Loss(0, G, f, types.SimpleNamespace(batchSize=B))
    loss.backward()
    assert abs(float(loss) - float(z["cld_loss"])) <= 1e-3 * abs(float(z["cld_loss"]))
    assert rel2(f.grad, z["cld_dx"]) <= 2e-3
