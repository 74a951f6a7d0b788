"""CPU: the oracle must reproduce the fixtures the unmodified reference produced (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

import oracle
from oracle.encoder import EncoderParams


def _sorted_rows(a):
    M, S, K, D = a.shape
    flat = a.reshape(M * S, K, D)
    out = np.empty_like(flat)
    for i in range(flat.shape[0]):
        r = flat[i]
        out[i] = r[np.lexsort(r.T[::-1])]
    return out.reshape(M, S, K, D)


def test_fps_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "fps.npz"))
    for i in range(int(z["n_cases"])):
        got = oracle.farthest_point_sampling(z[f"pc_{i}"], int(z[f"m_{i}"]), int(z[f"start_{i}"]))
        assert np.array_equal(got, z[f"idx_{i}"]), f"case {i}"
        assert len(set(got.tolist())) == len(got)          # picks stay unique


def test_fps_reorder():
    pts = np.random.default_rng(0).random((2, 50, 4)).astype(np.float32)
    out = oracle.fps_sample_data(pts, 8, [3, 7])
    for v in range(2):
        picks = oracle.farthest_point_sampling(pts[v, :, :3], 8, [3, 7][v])
        assert np.array_equal(out[v, :8], pts[v, picks])
        rest = np.setdiff1d(np.arange(50), picks)
        assert np.array_equal(out[v, 8:], pts[v, rest])


def test_fps_reorder_fewer_distinct_points_than_picks():
    """A cloud resampled with replacement from 5 distinct points: FPS repeats an index after 5 picks, and the
    reference truncates concatenate(picks, setdiff1d(...)) to N rows (cn3D_data_set.py:669-671)."""
    rng = np.random.default_rng(1)
    base = rng.random((5, 4)).astype(np.float32)
    pts = base[rng.integers(0, 5, size=40)][None]
    picks = oracle.farthest_point_sampling(pts[0, :, :3], 16, 2)
    assert len(set(picks.tolist())) < 16
    idx = oracle.fps_reorder_indices(picks, 40)
    other = np.setdiff1d(np.arange(40), picks.ravel())                 # the reference expression, verbatim semantics
    want = np.concatenate((picks.ravel(), other))[:40]
    assert idx.shape == (40,) and np.array_equal(idx, want)
    out = oracle.fps_sample_data(pts, 16, [2])
    assert out.shape == pts.shape and np.array_equal(out[0], pts[0, want])


def test_grouping_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "group.npz"))
    for name in z["names"]:
        pts = torch.from_numpy(z[f"{name}_points"])
        S, K = (int(v) for v in z[f"{name}_cfg"])
        r2 = float(z[f"{name}_r2"])
        xt, yt, idx = oracle.group_points(pts, S, K, r2)
        rows = xt.permute(0, 2, 3, 1).contiguous().numpy()
        assert np.array_equal(_sorted_rows(rows), z[f"{name}_rows_sorted"]), name
        assert np.array_equal(np.sort(idx.numpy(), axis=2), z[f"{name}_idx_sorted"].astype(np.int32)), name
        assert xt.shape == (pts.shape[0], 4, S, K) and yt.shape == (pts.shape[0], 3, S, 1)
        assert torch.equal(yt[:, :, :, 0].permute(0, 2, 1), pts[:, :S, :3])


def test_losses_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "losses.npz"))
    for i in range(int(z["n_cases"])):
        G, B, C = (int(v) for v in z[f"cfg_{i}"])
        x = torch.from_numpy(z[f"x_{i}"]).requires_grad_(True)
        xg = torch.from_numpy(z[f"xg_{i}"]).requires_grad_(True)
        lg = oracle.global_contrast(G, xg, x, B)
        lc = oracle.circle_contrast(G, x, B, z[f"order_{i}"])
        (lg + lc).backward()
        ref_g, ref_c = z[f"loss_{i}"]
        assert abs(float(lg) - ref_g) <= 2e-5 * abs(ref_g) + 1e-4
        assert abs(float(lc) - ref_c) <= 2e-5 * abs(ref_c) + 1e-4
        for got, ref in ((x.grad, z[f"dx_{i}"]), (xg.grad, z[f"dxg_{i}"])):
            ref = torch.from_numpy(ref)
            assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-6


def _load_step_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "train_step.npz"))
    sd = oracle.init_state_dict(seed=int(z["seed_sd"]))
    for k in list(sd):
        if "sd0/" + k in z.files:
            sd[k] = torch.from_numpy(z["sd0/" + k]).clone()
    return z, sd


def test_train_step_matches_reference(golden_dir):
    z, sd = _load_step_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    res = oracle.train_step(sd, torch.from_numpy(z["points"]), z["order"], S=S, K=K, r2=float(z["r2"]))

    def rel2(a, b):
        a = torch.as_tensor(a).double().reshape(-1)
        b = torch.as_tensor(b).double().reshape(-1)
        return float((a - b).norm() / b.norm().clamp_min(1e-300))

    floor = rel2(z["x"], z["x64"])                         # the reference's own fp32-vs-fp64 deviation
    assert rel2(res["x"], z["x"]) < 4 * floor + 1e-6
    assert rel2(res["x_global"], z["x_global"]) < 4 * rel2(z["x_global"], z["x_global64"]) + 1e-6
    assert abs(res["loss"] - z["loss"][0]) < 4 * abs(z["loss"][0] - z["loss64"][0]) + 1e-5 * abs(z["loss64"][0])
    gscale = max(float(np.abs(z["grad64_val/" + k]).max()) for k in res["grads"])
    for k, g in res["grads"].items():
        pos = z["grad_pos/" + k]
        got = g.reshape(-1).numpy()[pos]
        ref64 = z["grad64_val/" + k]
        noise = float(z["noise/" + k])
        if noise > 1.0:                                    # mathematically-zero gradients (bias in front of a train-mode BN)
            assert np.abs(got).max() < 1e-4 * gscale, k
            continue
        # sampled entries: error relative to the tensor's RMS magnitude, within the reference's rounding sensitivity
        rms = float(z["grad64_norm/" + k]) / np.sqrt(g.numel())
        if rms == 0.0:                                     # mapping.weight: feeds only the unused `code` output
            assert np.abs(got).max() == 0.0, k
            continue
        err = np.sqrt(np.mean((got - ref64) ** 2)) / rms
        assert err < 8 * noise + 1e-5, (k, err, noise)
    # weights after the Adam step + BN running statistics
    for k, v in sd.items():
        if "sd1/" + k in z.files:
            ref = torch.from_numpy(z["sd1/" + k])
            if "running_" in k:
                assert rel2(v, ref) < 1e-3, k
            elif v.dtype.is_floating_point:
                # first Adam step moves every entry by at most lr = 3e-4, in the direction of sign(grad): entries whose
                # true gradient is 0 (biases in front of a train-mode BN) move by +-lr on rounding noise alone
                assert float((v - ref).abs().max()) <= 2.5 * 3e-4, k
            else:
                assert int(v) == int(ref), k
        elif "sd1_pos/" + k in z.files:
            got = v.reshape(-1).numpy()[z["sd1_pos/" + k]]
            assert np.allclose(got, z["sd1_val/" + k], rtol=0, atol=2.5 * 3e-4), k   # |delta| <= lr per Adam step


def test_fine_geometry_matches_reference(golden_dir):
    """PointNet_Plus_fine's default geometry (sample_num_level1=32, knn_K=128, cn3d_model_conbag.py:142): the oracle against the
    reference's fp64 run of one step (tests/golden/make_golden.py: gen_fine_geometry)."""
    z = np.load(os.path.join(golden_dir, "fine_geometry.npz"))
    sd = oracle.init_state_dict(seed=int(z["seed_sd"]))
    for k in list(sd):
        if "sd0/" + k in z.files:
            sd[k] = torch.from_numpy(z["sd0/" + k]).clone()
    B, G, N, S, K = (int(v) for v in z["cfg"])
    assert (S, K) == (32, 128)
    sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
    res = oracle.train_step(sd64, torch.from_numpy(z["points"]), z["order"], S=S, K=K, r2=float(z["r2"]), dtype=torch.float64,
                            apply_update=False)
    rel2 = lambda a, b: float((torch.as_tensor(a).double().reshape(-1) - torch.as_tensor(b).double().reshape(-1)).norm() /
                              torch.as_tensor(b).double().norm())
    assert rel2(res["x"], z["x64"]) < 1e-9 and rel2(res["x_global"], z["x_global64"]) < 1e-9
    assert abs(res["loss"] - float(z["loss64"])) < 1e-9 * abs(float(z["loss64"]))
    for k, g in res["grads"].items():
        ref = z["grad64_val/" + k]
        got = g.reshape(-1).numpy()[z["grad_pos/" + k]]
        scale = max(float(np.abs(ref).max()), 1e-30)
        if float(z["grad64_norm/" + k]) > 0:
            assert np.abs(got - ref).max() <= 1e-7 * scale + 1e-12, k


def test_disabled_heads_match_reference(golden_dir):
    """Sinkhorn, the SwAV block and the CLD k-means loss (disabled in the reference scripts by constants) against outputs of the
    reference functions (tests/golden/make_golden.py: gen_heads)."""
    from oracle import heads as oh
    z = np.load(os.path.join(golden_dir, "heads.npz"))
    for i in range(int(z["sk_n"])):
        got = oh.distributed_sinkhorn(torch.from_numpy(z[f"sk_q_{i}"]), 3)
        assert np.allclose(got.numpy(), z[f"sk_out_{i}"], rtol=1e-5, atol=1e-7), i
    G, B = (int(v) for v in z["swav_cfg"])
    x = torch.from_numpy(z["swav_x"]).requires_grad_(True)
    w = torch.from_numpy(z["swav_w"]).requires_grad_(True)
    loss = oh.swav_loss(torch.nn.functional.normalize(x, dim=1) @ w.t(), G, B)
    loss.backward()
    assert abs(float(loss) - float(z["swav_loss"])) < 1e-5 * abs(float(z["swav_loss"]))
    assert np.allclose(x.grad.numpy(), z["swav_dx"], rtol=1e-4, atol=1e-7) and np.allclose(w.grad.numpy(), z["swav_dw"], rtol=1e-4, atol=1e-6)
    G, B = (int(v) for v in z["cld_cfg"])
    f = torch.from_numpy(z["cld_x"]).requires_grad_(True)
    labels, cent = oh.kmeans(f.detach()[: 3 * B], 60, 5)
    assert np.array_equal(labels.numpy(), z["km_labels"]) and np.allclose(cent.numpy(), z["km_centroids"], rtol=1e-5, atol=1e-6)
    loss = oh.cld_loss(G, f, B)
    loss.backward()
    assert abs(float(loss) - float(z["cld_loss"])) < 1e-5 * abs(float(z["cld_loss"]))
    assert np.allclose(f.grad.numpy(), z["cld_dx"], rtol=1e-3, atol=1e-6)


def test_eval_forward_matches_reference(golden_dir):
    z, sd = _load_step_fixture(golden_dir)
    B, G, N, S, K = (int(v) for v in z["cfg"])
    # the fixture's eval features were computed with the post-step weights: redo the step, then eval
    oracle.train_step(sd, torch.from_numpy(z["points"]), z["order"], S=S, K=K, r2=float(z["r2"]))
    clouds = torch.from_numpy(z["points"]).permute(1, 0, 2, 3).reshape(G * B, N, 4)
    xt, yt, _ = oracle.group_points(clouds, S, K, float(z["r2"]))
    with torch.no_grad():
        x, _, _, xg = oracle.encoder_forward(EncoderParams(sd, training=False), xt, yt, gost=G)
    feat = torch.cat([x, xg], 0)
    ref = torch.from_numpy(z["eval_feat"])
    assert float((feat - ref).norm() / ref.norm()) < 2e-3


def test_info_nce_logits(golden_dir):
    z, _ = _load_step_fixture(golden_dir)
    B = int(z["cfg"][0])
    logits, labels = oracle.info_nce_logits(torch.from_numpy(z["x"])[: 2 * B], B)
    assert np.allclose(logits.numpy(), z["info_nce_logits"], rtol=1e-5, atol=1e-5)
    assert labels.shape == (B,) and int(labels.sum()) == 0


def test_augment_matches_reference(golden_dir):
    """get_data_train + get_temporal_augment_data of the unmodified reference, driven by np.random.seed, against the
    restatement fed with the same draws in the order the reference consumes them."""
    from oracle import augment as oaug
    z = np.load(os.path.join(golden_dir, "augment.npz"))
    for i in range(int(z["n_cases"])):
        srcs = [z[f"{k}_{i}"] for k in ("points", "key", "res1", "res2")]
        draws = oaug.Draws(z[f"idx_{i}"], z[f"noise_{i}"], z[f"angle_{i}"])
        # the recorded draws are what numpy's legacy RandomState yields in the reference's call order
        again = oaug.record_draws(np.random.RandomState(50 + i), srcs, 512)
        assert np.array_equal(again.idx, draws.idx) and np.array_equal(again.noise, draws.noise)
        views = oaug.make_views(srcs, draws)
        assert views.dtype == np.float32 and views.shape == (10, 512, 4)
        assert np.array_equal(views, z[f"views_{i}"]), f"case {i}"
        # temporal views only hold rows whose temporal channel is non-zero
        assert np.all(views[6, :, 3] != 0) and np.all(views[7, :, 3] != 0)


def _sorted_cols(a):
    M, C, S, K = a.shape
    rows = np.ascontiguousarray(a.transpose(0, 2, 3, 1)).reshape(M * S, K, C)
    out = np.empty_like(rows)
    for i in range(rows.shape[0]):
        r = rows[i]
        out[i] = r[np.lexsort(r.T[::-1])]
    return out.reshape(M, S, K, C)


def test_level2_grouping_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "group2.npz"))
    for name in z["names"]:
        feats = torch.from_numpy(z[f"{name}_feats"])
        S2, K = (int(v) for v in z[f"{name}_cfg"])
        out, centre, idx = oracle.group_points_level2(feats, S2, K, float(z[f"{name}_r2"]))
        assert np.array_equal(_sorted_cols(out.numpy()), z[f"{name}_sorted"]), name
        assert torch.equal(centre[..., 0], feats[:, 0:3, 0:S2])


def test_probe_and_feature_files_match_reference(golden_dir, tmp_path):
    """Final_FC + CrossEntropy + Adam of linear_classify/, and save_single_feature's per-video file."""
    from oracle import probe as oprobe
    z = np.load(os.path.join(golden_dir, "probe.npz"))
    rows = oprobe.feature_rows(z["feat"], 11)
    assert np.array_equal(rows, z["feat_rows"])
    np.save(tmp_path / "v.npy", rows[0])
    assert np.array_equal(np.frombuffer(open(tmp_path / "v.npy", "rb").read(), dtype=np.uint8), z["file0"])
    sd = {"fc.weight": torch.from_numpy(z["w0"]).clone(), "fc.bias": torch.from_numpy(z["b0"]).clone()}
    state = {}
    for it in range(3):
        o = oprobe.probe_step(sd, torch.from_numpy(z["x"][it]), torch.from_numpy(z["y"][it]), state)
        assert abs(o["loss"] - z["losses"][it]) <= 1e-5 and abs(o["top1"] - z["top1"][it]) < 1e-4
        if it == 0:
            assert np.abs(o["logits"].numpy() - z["logits0"]).max() <= 1e-6
            assert np.abs(o["grads"]["fc.weight"].numpy() - z["grad_w0"]).max() <= 1e-7
            assert np.abs(o["grads"]["fc.bias"].numpy() - z["grad_b0"]).max() <= 1e-7
    assert np.abs(sd["fc.weight"].numpy() - z["w3"]).max() <= 2e-6 and np.abs(sd["fc.bias"].numpy() - z["b3"]).max() <= 2e-6
