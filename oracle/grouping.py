"""Oracle: kNN + ball-query grouping (test infrastructure; see oracle/__init__.py).

Restates `group_points_3DV` (reference training_code/utils_my.py:255-291) and its near copies
`group_points_3DV_2048` (:7-42), `group_points` (:217-253), `group_points_3DV_nums` (:293-328).
They differ only in where S, K and the squared radius come from.

Semantics (pinned by tests/golden/group_*.npz, generated from the reference functions):
  * centres are the first S rows of each cloud (utils_my.py:266);
  * d(s, n) = ((dx*dx + dy*dy) + dz*dz), dx = p_n.x - c_s.x, all fp32 (utils_my.py:265-268);
  * the K smallest d per centre are kept (torch.topk(sorted=False), :269).  The reference's order
    inside the K slots is arbitrary; this oracle fixes it to ascending (d, n) -- a stable sort --
    so index SETS are compared on tie-free inputs and gathered-value multisets otherwise;
  * slots with d > r2 (strict, on the squared distance, :272) are redirected to index s, the centre
    itself (:274-275);
  * the D channels of the chosen rows are gathered and the centre xyz subtracted (:277-282);
  * returned views: xt logical (M, D, S, K) over a physical [M][S][K][D] buffer, yt (M, 3, S, 1).
"""
import numpy as np
import torch


def knn_ball_indices(points, S, K, r2):
    """points (M,N,D) float32 -> idx (M,S,K) int32 (ascending (d,n) order, ball-redirected),
    dist (M,S,K) float32 (pre-redirect distances of the chosen neighbours)."""
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float32))
    M, N, _ = pts.shape
    r2 = np.float32(r2)
    idx = np.empty((M, S, K), dtype=np.int32)
    dsel = np.empty((M, S, K), dtype=np.float32)
    for m in range(M):
        p = pts[m, :, 0:3]
        c = p[:S]
        dx = p[None, :, 0] - c[:, None, 0]
        dy = p[None, :, 1] - c[:, None, 1]
        dz = p[None, :, 2] - c[:, None, 2]
        d = (dx * dx + dy * dy) + dz * dz                      # (S,N) fp32
        order = np.argsort(d, axis=1, kind="stable")[:, :K]     # ties -> lowest index first
        dk = np.take_along_axis(d, order, axis=1)
        centre = np.arange(S, dtype=np.int64)[:, None]
        idx[m] = np.where(dk > r2, centre, order).astype(np.int32)
        dsel[m] = dk
    return idx, dsel


def group_points(points, S, K, r2):
    """torch in / torch out restatement of the reference grouping functions.

    points (M,N,D) float32 tensor.  Returns (xt, yt, idx): xt (M,D,S,K) view, yt (M,3,S,1) view,
    idx (M,S,K) int32."""
    pts = points.detach().to(torch.float32).contiguous()
    M, N, D = pts.shape
    idx_np, _ = knn_ball_indices(pts.numpy(), S, K, r2)
    idx = torch.from_numpy(idx_np)
    flat = idx.view(M, S * K).to(torch.int64)
    rows = torch.gather(pts, 1, flat[:, :, None].expand(M, S * K, D)).view(M, S, K, D).clone()
    centre = pts[:, :S, 0:3]
    rows[..., 0:3] = rows[..., 0:3] - centre[:, :, None, :]
    xt = rows.permute(0, 3, 1, 2)                        # logical (M,D,S,K), physical [M][S][K][D]
    yt = centre.contiguous().permute(0, 2, 1)[..., None]  # (M,3,S,1)
    return xt, yt, idx


def group_points_level2(feats, S2, K, r2):
    """Level-2 set abstraction, reference utils_my.py:332-356 (`group_points_2`, K = 64 hard-coded at :335) and
    :358-381 (`group_points_2_3DV`, K = 32, radius 0.11 hard-coded at :361-362).

    feats (M, C, N1) float32 channel-first, channels 0..2 = xyz; centres = the first S2 points (:339, :353).
    Same selection rule as level 1 (K smallest squared distances, slots with d > r2 redirected to the centre,
    :342-348 -- note the reference compares the SQUARED distance with `ball_radius` as given), then every channel is
    gathered (:350-351) and the centre subtracted from xyz (:354).
    Returns (inputs_level2 (M,C,S2,K), centre (M,3,S2,1), idx (M,S2,K) int32 in ascending (d, n) order)."""
    f = feats.detach().to(torch.float32).contiguous()
    M, C, N1 = f.shape
    rows = f[:, 0:3, :].permute(0, 2, 1).contiguous()                     # (M,N1,3)
    idx_np, _ = knn_ball_indices(rows.numpy(), S2, K, r2)
    idx = torch.from_numpy(idx_np)
    flat = idx.view(M, 1, S2 * K).to(torch.int64).expand(M, C, S2 * K)
    out = torch.gather(f, 2, flat).view(M, C, S2, K).clone()
    centre = f[:, 0:3, 0:S2].unsqueeze(3)
    out[:, 0:3] = out[:, 0:3] - centre
    return out, centre, idx
