"""GPU parity tests for the stand-alone kernels, called through the C ABI (ctypes): FPS, grouping, tcgen05 GEMM.

The checker is `oracle/` (CPU) and the committed reference fixtures in tests/golden/.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from facl_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sorted_rows(a):
    M, S, K, D = a.shape
    flat = a.reshape(M * S, K, D)
    out = np.empty_like(flat)
    for i in range(flat.shape[0]):
        r = flat[i]
        out[i] = r[np.lexsort(r.T[::-1])]
    return out.reshape(M, S, K, D)


# ------------------------------------------------------------------------------------------------- FPS
def test_fps_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "fps.npz"))
    for i in range(int(z["n_cases"])):
        pc = z[f"pc_{i}"].astype(np.float32)
        m, start = int(z[f"m_{i}"]), int(z[f"start_{i}"])
        pts = torch.from_numpy(pc)[None].to(DEV)
        got = ops.fps(pts, m, torch.tensor([start], dtype=torch.int32, device=DEV)).cpu().numpy()[0]
        assert np.array_equal(got, z[f"idx_{i}"]), f"case {i}: first mismatch at {np.flatnonzero(got != z[f'idx_{i}'])[:4]}"


@pytest.mark.parametrize("N,m", [(1024, 64), (2048, 64), (4096, 128), (8192, 64), (16384, 64), (333, 50)])
def test_fps_vs_oracle_batched(N, m):
    V = 6
    pts = synth.make_sequences(V, 1, N, seed=N + m, skeleton=(N % 2048 == 0), resample=(N == 4096))[:, 0]
    starts = np.random.default_rng(N).integers(0, N, size=V).astype(np.int32)
    got = ops.fps(torch.from_numpy(pts).to(DEV), m, torch.from_numpy(starts).to(DEV)).cpu().numpy()
    for v in range(V):
        want = oracle.farthest_point_sampling(pts[v, :, :3], m, int(starts[v]))
        assert np.array_equal(got[v], want), f"cloud {v}"
    # reorder: picks first, remainder ascending (cn3D_data_set.py:669-671)
    re = ops.fps_reorder(torch.from_numpy(pts).to(DEV), torch.from_numpy(got).to(DEV)).cpu().numpy()
    for v in range(V):
        want = pts[v, oracle.fps_reorder_indices(got[v], N)]
        assert np.array_equal(re[v], want), f"reorder cloud {v}"


def test_fps_reorder_with_repeated_picks():
    """Clouds with fewer than m distinct points (resampled with replacement, cn3D_data_set.py:287): picks repeat, the tail is
    truncated at N rows like the reference's new_idx[:NUM_POINT] (:671); the LAST cloud of the batch must not write past the
    output buffer and the following cloud must stay intact."""
    rng = np.random.default_rng(5)
    V, N, m = 3, 200, 64
    pts = np.empty((V, N, 4), dtype=np.float32)
    for v in range(V):
        base = rng.random((7 + v, 4)).astype(np.float32)
        pts[v] = base[rng.integers(0, base.shape[0], size=N)]
    starts = np.array([0, 5, 199], dtype=np.int32)
    dev = torch.from_numpy(pts).to(DEV)
    picks = ops.fps(dev, m, torch.from_numpy(starts).to(DEV))
    p = picks.cpu().numpy()
    # guard rows behind the output: a tail overrun would land there
    out_buf = torch.full((V * N + 64, 4), -7.0, device=DEV)
    from facl_b200._lib import check, lib, ptr, stream_ptr
    check(lib().facl_fps_reorder(ptr(dev), V, N, 4, ptr(picks), m, ptr(out_buf), stream_ptr()), "facl_fps_reorder")
    got = out_buf.cpu().numpy()
    assert np.all(got[V * N:] == -7.0)
    for v in range(V):
        want_p = oracle.farthest_point_sampling(pts[v, :, :3], m, int(starts[v]))
        assert np.array_equal(p[v], want_p)
        assert len(set(want_p.tolist())) < m
        assert np.array_equal(got[v * N:(v + 1) * N], pts[v, oracle.fps_reorder_indices(want_p, N)]), f"cloud {v}"


# ------------------------------------------------------------------------------------------------- grouping
def test_group_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "group.npz"))
    for name in z["names"]:
        pts = torch.from_numpy(z[f"{name}_points"]).to(DEV)
        S, K = (int(v) for v in z[f"{name}_cfg"])
        rows, idx = ops.group_points_raw(pts, S, K, float(z[f"{name}_r2"]))
        rows, idx = rows.cpu().numpy(), idx.cpu().numpy()
        assert np.array_equal(_sorted_rows(rows), z[f"{name}_rows_sorted"]), f"{name}: gathered rows differ"
        if bool(z[f"{name}_tie_free"]):
            assert np.array_equal(np.sort(idx, axis=2), z[f"{name}_idx_sorted"].astype(np.int32)), f"{name}: index sets differ"


@pytest.mark.parametrize("N,S,K,r2,resample", [(2048, 64, 64, 0.16, False), (2048, 64, 64, 0.0025, False),
                                               (1024, 128, 32, 0.06, True), (8192, 64, 128, 0.01, False),
                                               (5000, 64, 16, 0.06, False), (200, 64, 64, 0.06, True),
                                               (16384, 64, 64, 0.16, False), (4096, 64, 64, 0.01, False),
                                               (1024, 64, 64, 0.06, False), (2048, 50, 64, 0.06, True)])
def test_group_vs_oracle(N, S, K, r2, resample):
    M = 4
    pts = synth.make_sequences(M, 1, N, seed=N + K, skeleton=True, resample=resample)[:, 0]
    rows, idx = ops.group_points_raw(torch.from_numpy(pts).to(DEV), S, K, r2)
    oxt, oyt, oidx = oracle.group_points(torch.from_numpy(pts), S, K, r2)
    # this implementation and the oracle share the (distance, index) order, so even slot order must agree
    assert np.array_equal(idx.cpu().numpy(), oidx.numpy())
    assert np.array_equal(rows.cpu().numpy(), oxt.permute(0, 2, 3, 1).contiguous().numpy())


def test_group_heavy_duplicates_slow_path():
    # 700 copies of the centre: more candidates than the per-warp list holds -> exercises the fallback
    N, S, K = 1024, 8, 64
    rng = np.random.default_rng(5)
    pts = (rng.random((2, N, 4)).astype(np.float32) - 0.5)
    pts[:, 100:800, :3] = pts[:, 0:1, :3]
    rows, idx = ops.group_points_raw(torch.from_numpy(pts).to(DEV), S, K, 0.06)
    oxt, _, oidx = oracle.group_points(torch.from_numpy(pts), S, K, 0.06)
    assert np.array_equal(idx.cpu().numpy(), oidx.numpy())
    assert np.array_equal(rows.cpu().numpy(), oxt.permute(0, 2, 3, 1).contiguous().numpy())


# ------------------------------------------------------------------------------------------------- tcgen05 GEMM
def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float64)


def _xform(src0, src1, s0, s1, s2, lo):
    v = src0.double() * (1.0 if s0 is None else s0.double())
    if src1 is not None:
        v = v + src1.double() * s1.double()
    if s2 is not None:
        v = v + s2.double()
    if lo is not None:
        v = torch.maximum(v, lo.double())
    return v


def _report(name, got, want, tol):
    err = (got.double().cpu() - want).abs()
    scale = want.abs().max().clamp_min(1e-30)
    rel = float(err.max() / scale)
    if rel > tol:
        bad = (err > tol * scale).nonzero()
        rows = sorted(set(bad[:, 0].tolist()))[:12]
        cols = sorted(set(bad[:, 1].tolist()))[:12] if bad.shape[1] > 1 else []
        pytest.fail(f"{name}: rel err {rel:.3e} > {tol:.1e}; {bad.shape[0]} bad of {err.numel()}; rows {rows} cols {cols}; "
                    f"got[0,:4]={got.flatten()[:4].tolist()} want[0,:4]={want.flatten()[:4].tolist()}")
    return rel


@pytest.mark.parametrize("nsplit", [1, 3])
@pytest.mark.parametrize("Md,Nd,Kd", [(128, 256, 64), (128, 256, 256), (256, 512, 128), (64, 100, 40), (300, 1000, 259)])
def test_gemm_packed_a_rowmajor_b(nsplit, Md, Nd, Kd):
    g = torch.Generator().manual_seed(Md + Nd + Kd)
    A = torch.randn(Md, Kd, generator=g)
    ldb = (Kd + 3) // 4 * 4
    Bm = torch.zeros(Nd, ldb)
    Bm[:, :Kd] = torch.randn(Nd, Kd, generator=g)
    bias = torch.randn(Md, generator=g)
    Ad, Bd = A.to(DEV), Bm.to(DEV)
    img = ops.pack_weight(Ad, Md, Kd, Kd, 1)
    out = torch.full((Md, Nd), float("nan"), device=DEV)
    out_t = torch.full((Nd, Md), float("nan"), device=DEV)
    for mode, o, ldo in ((ops.OUT_CHMAJOR, out, Nd), (ops.OUT_ROWMAJOR, out_t, Md)):
        ops.gemm_tc(Md, Nd, Kd, nsplit=nsplit, a_packed=img, b_mode=ops.B_ROWMAJOR, b=dict(src0=Bd, ld=ldb),
                    bias=bias.to(DEV), out_mode=mode, out=o, ldo=ldo)
    torch.cuda.synchronize()
    if nsplit == 1:
        want = _bf16_round(A) @ _bf16_round(Bm[:, :Kd]).t() + bias.double()[:, None]
        tol = 2e-5
    else:
        want = A.double() @ Bm[:, :Kd].double().t() + bias.double()[:, None]
        tol = 1e-4
    _report("chmajor", out, want, tol)
    _report("rowmajor", out_t, want.t(), tol)


@pytest.mark.parametrize("nsplit", [1, 3])
def test_gemm_chmajor_b_transform_stats_pool(nsplit):
    Md, Nd, Kd, pool = 256, 2048, 259, 64
    g = torch.Generator().manual_seed(7)
    W = torch.randn(Md, Kd, generator=g) / Kd ** 0.5
    Z = torch.randn(Kd, Nd, generator=g)                     # channel-major source
    s0 = 0.5 + torch.rand(Kd, generator=g)
    s2 = 0.3 * torch.randn(Kd, generator=g)
    lo = torch.zeros(Kd)
    lo[:3] = -float("inf")                                   # xyz channels: no ReLU
    bias = torch.randn(Md, generator=g)
    sign = torch.where(torch.rand(Md, generator=g) > 0.3, 1.0, -1.0)
    P = ops.stat_partials(Md, Nd)
    out = torch.zeros(Md, Nd, device=DEV)
    stats = torch.zeros(P, Md, 2, device=DEV)
    pout = torch.zeros(Md, Nd // pool, device=DEV)
    parg = torch.zeros(Md, Nd // pool, dtype=torch.uint8, device=DEV)
    img = ops.pack_weight(W.to(DEV), Md, Kd, Kd, 1)
    ops.gemm_tc(Md, Nd, Kd, nsplit=nsplit, a_packed=img, b_mode=ops.B_CHMAJOR,
                b=dict(src0=Z.to(DEV), ld=Nd, s0=s0.to(DEV), s2=s2.to(DEV), lo=lo.to(DEV)),
                bias=bias.to(DEV), out_mode=ops.OUT_CHMAJOR, out=out, ldo=Nd, stats=stats,
                pool=pool, pool_sign=sign.to(DEV), pool_out=pout, pool_arg=parg, ldp=Nd // pool)
    torch.cuda.synchronize()
    H = _xform(Z, None, s0[:, None], None, s2[:, None], lo[:, None])          # (Kd, Nd)
    if nsplit == 1:
        want = _bf16_round(W) @ _bf16_round(H.float()) + bias.double()[:, None]
        tol = 3e-3     # a transform result that lands on a bf16 rounding boundary may round the other way
    else:
        want = W.double() @ H + bias.double()[:, None]
        tol = 1e-4
    _report("out", out, want, tol)
    st = stats.double().sum(0).cpu()
    _report("sum", st[:, 0:1], want.sum(1, keepdim=True), 1e-4)
    _report("sumsq", st[:, 1:2], (want ** 2).sum(1, keepdim=True), 1e-4)
    got = out.double().cpu().reshape(Md, Nd // pool, pool)
    sel = torch.where(sign[:, None] >= 0, got.max(2).values, got.min(2).values)
    assert torch.equal(pout.double().cpu(), sel), "pooled value must be the selected element of the stored output"
    arg = torch.where(sign[:, None] >= 0, got.argmax(2), got.argmin(2))
    assert torch.equal(parg.long().cpu(), arg)


@pytest.mark.parametrize("nsplit", [1, 3])
def test_gemm_rowmajor_a_two_source_splitk(nsplit):
    # wgrad shape: dW[co,ci] = sum_r dz[co,r] * h[ci,r];  dz = c0*dh + c1*z + c2 (per co), h = relu(a*zin + b) (per ci)
    Md, Nd, Kd, ksplit = 200, 64, 5000, 7
    g = torch.Generator().manual_seed(9)
    dh, z = torch.randn(Md, Kd, generator=g), torch.randn(Md, Kd, generator=g)
    c0, c1, c2 = torch.rand(Md, generator=g) + 0.5, 0.1 * torch.randn(Md, generator=g), 0.1 * torch.randn(Md, generator=g)
    zin = torch.randn(Nd, Kd, generator=g)
    a, b = torch.rand(Nd, generator=g) + 0.5, 0.2 * torch.randn(Nd, generator=g)
    out = torch.zeros(Md, Nd, device=DEV)
    ops.gemm_tc(Md, Nd, Kd, nsplit=nsplit,
                a=dict(src0=dh.to(DEV), src1=z.to(DEV), ld=Kd, s0=c0.to(DEV), s1=c1.to(DEV), s2=c2.to(DEV)),
                b_mode=ops.B_ROWMAJOR, b=dict(src0=zin.to(DEV), ld=Kd, s0=a.to(DEV), s2=b.to(DEV), lo=torch.zeros(Nd, device=DEV)),
                ksplit=ksplit, out_mode=ops.OUT_ATOMIC, out=out, ldo=Nd)
    torch.cuda.synchronize()
    dz = _xform(dh, z, c0[:, None], c1[:, None], c2[:, None], None)
    h = _xform(zin, None, a[:, None], None, b[:, None], torch.zeros(Nd, 1))
    if nsplit == 1:
        want = _bf16_round(dz.float()) @ _bf16_round(h.float()).t()
        tol = 3e-3
    else:
        want = dz @ h.t()
        tol = 1e-4
    _report("wgrad", out, want, tol)


def test_gemm_zin_mask_and_second_stat():
    Md, Nd, Kd = 128, 768, 192
    g = torch.Generator().manual_seed(11)
    W, X = torch.randn(Md, Kd, generator=g), torch.randn(Nd, Kd, generator=g)
    zin = torch.randn(Md, Nd, generator=g)
    zs0, zs2 = torch.randn(Md, generator=g), 0.2 * torch.randn(Md, generator=g)
    P = ops.stat_partials(Md, Nd)
    out = torch.zeros(Md, Nd, device=DEV)
    stats = torch.zeros(P, Md, 2, device=DEV)
    ops.gemm_tc(Md, Nd, Kd, nsplit=3, a_packed=ops.pack_weight(W.to(DEV), Md, Kd, Kd, 1), b_mode=ops.B_ROWMAJOR,
                b=dict(src0=X.to(DEV), ld=Kd), out_mode=ops.OUT_CHMAJOR, out=out, ldo=Nd, zin=zin.to(DEV), ldz=Nd,
                zs0=zs0.to(DEV), zs2=zs2.to(DEV), stats=stats)
    torch.cuda.synchronize()
    mask = (zs0[:, None] * zin + zs2[:, None]) > 0
    want = (W.double() @ X.double().t()) * mask
    _report("masked", out, want, 1e-4)
    st = stats.double().sum(0).cpu()
    _report("sum", st[:, 0:1], want.sum(1, keepdim=True), 2e-4)
    _report("sum_vz", st[:, 1:2], (want * zin.double()).sum(1, keepdim=True), 2e-4)


@pytest.mark.parametrize("nsplit", [1, 3])
def test_gemm_xt4(nsplit):
    Md, Nd = 64, 4096
    g = torch.Generator().manual_seed(13)
    W, X = torch.randn(Md, 4, generator=g), torch.randn(Nd, 4, generator=g)
    out = torch.zeros(Md, Nd, device=DEV)
    ops.gemm_tc(Md, Nd, 4, nsplit=nsplit, a_packed=ops.pack_weight(W.to(DEV), Md, 4, 4, 1), b_mode=ops.B_XT4,
                b=dict(src0=X.to(DEV), ld=4), out_mode=ops.OUT_CHMAJOR, out=out, ldo=Nd)
    torch.cuda.synchronize()
    want = (_bf16_round(W) @ _bf16_round(X).t()) if nsplit == 1 else W.double() @ X.double().t()
    _report("xt4", out, want, 2e-5 if nsplit == 1 else 1e-4)


@pytest.mark.parametrize("nsplit", [1, 3])
@pytest.mark.parametrize("Cout,Cin,R", [(256, 259, 1024), (512, 256, 640), (128, 64, 200), (64, 128, 512), (96, 64, 256)])
def test_gemm_activation_images(nsplit, Cout, Cin, R):
    """TMA-only GEMMs over pre-converted activation images (gemm_img.cu): forward (reduction over channels, with the
    BN+ReLU transform, statistics and max-pool epilogue) and weight gradient (reduction over rows, split-K)."""
    g = torch.Generator().manual_seed(Cout + Cin + R + nsplit)
    W = torch.randn(Cout, Cin, generator=g) / Cin ** 0.5
    z = torch.randn(Cin, R, generator=g)
    s0, s2 = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g) * 0.3
    lo = torch.zeros(Cin)
    bias = torch.randn(Cout, generator=g)
    h = torch.maximum(z.double() * s0.double()[:, None] + s2.double()[:, None], lo.double()[:, None])
    hq = _bf16_round(h.float()) if nsplit == 1 else h
    Wq = _bf16_round(W) if nsplit == 1 else W.double()
    zd = z.to(DEV)
    himg = ops.act_image(Cin, R, nsplit, src0=zd, ld=R, s0=s0.to(DEV), s2=s2.to(DEV), lo=lo.to(DEV))
    wimg = ops.pack_weight(W.to(DEV), Cout, Cin, Cin, 1)
    pool = 8
    P = ops.stat_partials(Cout, R)
    out = torch.full((Cout, R), float("nan"), device=DEV)
    stats = torch.zeros((P, Cout, 2), device=DEV)
    pooled = torch.full((Cout, R // pool), float("nan"), device=DEV)
    parg = torch.zeros((Cout, R // pool), dtype=torch.uint8, device=DEV)
    sign = torch.ones(Cout, device=DEV)
    ops.gemm_tc(Cout, R, Cin, nsplit=nsplit, a_packed=wimg, b_mode=ops.B_IMAGE_MN, b_img=himg, bias=bias.to(DEV),
                out_mode=ops.OUT_CHMAJOR, out=out, ldo=R, stats=stats, pool=pool, pool_sign=sign, pool_out=pooled,
                pool_arg=parg, ldp=R // pool)
    torch.cuda.synchronize()
    want = Wq.double() @ hq.double() + bias.double()[:, None]
    tol = 2e-5 if nsplit == 1 else 1e-4
    _report("fwd", out, want, tol)
    _report("stats0", stats.sum(0)[:, 0:1], want.sum(1, keepdim=True), 1e-4)
    wp = want.reshape(Cout, R // pool, pool)
    _report("pooled", pooled, wp.max(2).values, tol)
    got_arg = parg.cpu().long()
    picked = torch.gather(out.cpu().double().reshape(Cout, R // pool, pool), 2, got_arg[..., None])[..., 0]
    assert torch.equal(picked.float(), pooled.cpu())
    # weight gradient: dW[co][ci] = sum_r dz[co][r] h[ci][r], dz = c0*dy + c1*zz + c2 (two-source transform)
    dy, zz = torch.randn(Cout, R, generator=g), torch.randn(Cout, R, generator=g)
    c0, c1, c2 = (torch.randn(Cout, generator=g) for _ in range(3))
    dz = dy.double() * c0.double()[:, None] + zz.double() * c1.double()[:, None] + c2.double()[:, None]
    dzimg = ops.act_image(Cout, R, nsplit, src0=dy.to(DEV), src1=zz.to(DEV), ld=R, s0=c0.to(DEV), s1=c1.to(DEV), s2=c2.to(DEV))
    dW = torch.zeros((Cout, Cin), device=DEV)
    ops.gemm_tc(Cout, Cin, R, nsplit=nsplit, a_img=dzimg, b_mode=ops.B_IMAGE_K, b_img=himg, ksplit=min(3, (R + 63) // 64),
                out_mode=ops.OUT_ATOMIC, out=dW, ldo=Cin)
    torch.cuda.synchronize()
    dzq = _bf16_round(dz.float()) if nsplit == 1 else dz
    _report("wgrad", dW, dzq.double() @ hq.double().t(), tol * 5)


def test_act_image_pooled_source():
    """dz image of a max-pooled gradient read through the argmax table == image of the dense scatter."""
    g = torch.Generator().manual_seed(5)
    Cc, groups, pool = 64, 24, 16
    R = groups * pool
    df = torch.randn(Cc, groups, generator=g)
    arg = torch.randint(0, pool, (Cc, groups), generator=g, dtype=torch.uint8)
    z = torch.randn(Cc, R, generator=g)
    c0, c1, c2 = (torch.randn(Cc, generator=g) for _ in range(3))
    dense = torch.zeros(Cc, groups, pool)
    dense.scatter_(2, arg.long()[..., None], df[..., None])
    dense = dense.reshape(Cc, R)
    kw = dict(s0=c0.to(DEV), s1=c1.to(DEV), s2=c2.to(DEV))
    a = ops.act_image(Cc, R, 3, src0=df.to(DEV), src1=z.to(DEV), ld=groups, ld1=R, pool_arg=arg.to(DEV), pool=pool, **kw)
    b = ops.act_image(Cc, R, 3, src0=dense.to(DEV), src1=z.to(DEV), ld=R, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a.buf, b.buf)
