"""Oracle: one training step (test infrastructure; see oracle/__init__.py).

Restates the loop body of reference training_code/cn3d_train_motion_GL.py:225-335
(cn3d_train_apperance_GL.py is the same file up to three constants):
  G-major flatten (:225-226) -> grouping (:230) -> encoder (:234) -> global loss (:265-287)
  -> circle loss (:290-316) -> loss = circle + global (:329; the SwAV and CLD terms are disabled by
  the constants at :238,319) -> backward -> Adam(lr 3e-4, betas (0.5,0.999), eps 1e-6) (:180,330-332).
"""
import torch

from .encoder import EncoderParams, encoder_forward
from .grouping import group_points
from .losses import global_contrast, circle_contrast


def adam_update(sd, grads, state, lr=3e-4, betas=(0.5, 0.999), eps=1e-6):
    """torch.optim.Adam (no weight decay, no amsgrad) restated; `state` maps key -> (step, m, v)."""
    b1, b2 = betas
    with torch.no_grad():
        for k, g in grads.items():
            step, m, v = state.get(k, (0, torch.zeros_like(g), torch.zeros_like(g)))
            step += 1
            m = b1 * m + (1 - b1) * g
            v = b2 * v + (1 - b2) * g * g
            denom = (v.sqrt() / (1 - b2 ** step) ** 0.5) + eps
            sd[k] -= (lr / (1 - b1 ** step)) * (m / denom)
            state[k] = (step, m, v)
    return state


def train_step(sd, points_bgnd, order, S=64, K=64, r2=0.06, adam_state=None, lr=3e-4,
               apply_update=True, dtype=torch.float32, routing=None, upstream=None):
    """points_bgnd (B,G,N,D) fp32.  Returns dict(loss, loss_global, loss_circle, grads, x, x_global, dx, dx_global).
    `sd` is updated in place (BN running stats always; weights when apply_update).
    `upstream` (tests only): (dL/dx, dL/dx_global) to back-propagate through the encoder INSTEAD of the loss gradient -- the
    stage-wise check of an implementation whose embeddings differ by rounding (the loss has temperature 1 on un-normalised
    dot products of magnitude ~500, so its softmax weights amplify a 1 % embedding error by orders of magnitude; the
    encoder backward itself is linear in the upstream gradient)."""
    B, G, N, D = points_bgnd.shape
    clouds = points_bgnd.permute(1, 0, 2, 3).reshape(G * B, N, D).to(torch.float32)
    xt, yt, _ = group_points(clouds, S, K, r2)        # grouping is always fp32 (:228)
    xt, yt = xt.to(dtype), yt.to(dtype)                 # dtype=float64: rounding-noise yardstick for tests
    params = EncoderParams(sd, training=True).requires_grad_(True)
    leaves = params.trainable()
    for v in leaves.values():
        v.grad = None
    x, code, x_nor, x_global = encoder_forward(params, xt, yt, gost=G, routing=routing)
    lg = global_contrast(G, x_global, x, B)
    lc = circle_contrast(G, x, B, order)
    loss = lc + lg
    x.retain_grad()
    x_global.retain_grad()
    if upstream is None:
        loss.backward()
    else:
        torch.autograd.backward([x, x_global], [upstream[0].to(x.dtype).reshape(x.shape), upstream[1].to(x.dtype).reshape(x_global.shape)])
    grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v))
             for k, v in leaves.items()}
    for v in leaves.values():
        v.requires_grad_(False)
        v.grad = None
    out = dict(loss=float(loss.detach()), loss_global=float(lg.detach()), loss_circle=float(lc.detach()), grads=grads,
               x=x.detach(), x_global=x_global.detach(), dx=x.grad.detach().clone(), dx_global=x_global.grad.detach().clone())
    if apply_update:
        if adam_state is None:
            adam_state = {}
        # mapping.weight only feeds `code`, which no live loss uses -> grad None -> Adam skips it
        live = {k: g for k, g in grads.items() if k != "mapping.weight"}
        out["adam_state"] = adam_update(sd, live, adam_state, lr=lr)
    return out


def train_step_sharded(sd, points_bgnd, order, world, S=64, K=64, r2=0.06, lr=3e-4, apply_update=False, dtype=torch.float32,
                       routing=None):
    """The multi-GPU parity definition of SURVEY.md section 8e, on one CPU: the batch of B sequences is split into `world`
    contiguous shards; every shard is encoded on its own with ITS OWN BatchNorm batch statistics (what nn.DataParallel,
    cn3d_train_motion_GL.py:176, does per replica), the embeddings are re-interleaved into the reference's global G-major
    order (row g*B + r*B_loc + b), the reference losses run on the global batch (opt.batchSize = B), and the gradient flows
    back through the gather into the shared weights (= the sum of the per-rank gradients an all-reduce produces).
    `routing`: optional list of per-shard routing tables (see encoder_forward).  BatchNorm running statistics in `sd` see
    one update per shard and are not meaningful afterwards (they stay rank-local in the product)."""
    B, G, N, D = points_bgnd.shape
    assert B % world == 0
    Bl = B // world
    params = EncoderParams(sd, training=True).requires_grad_(True)
    leaves = params.trainable()
    for v in leaves.values():
        v.grad = None
    xs, xgs = [], []
    for r in range(world):
        shard = points_bgnd[r * Bl:(r + 1) * Bl]
        clouds = shard.permute(1, 0, 2, 3).reshape(G * Bl, N, D).to(torch.float32)
        xt, yt, _ = group_points(clouds, S, K, r2)
        x, _, _, xg = encoder_forward(params, xt.to(dtype), yt.to(dtype), gost=G, routing=None if routing is None else routing[r])
        xs.append(x.reshape(G, Bl, -1))
        xgs.append(xg)
    x_all = torch.cat(xs, dim=1).reshape(G * B, -1)
    xg_all = torch.cat(xgs, dim=0)
    lg = global_contrast(G, xg_all, x_all, B)
    lc = circle_contrast(G, x_all, B, order)
    loss = lc + lg
    loss.backward()
    grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    for v in leaves.values():
        v.requires_grad_(False)
        v.grad = None
    out = dict(loss=float(loss.detach()), loss_global=float(lg.detach()), loss_circle=float(lc.detach()), grads=grads,
               x=x_all.detach(), x_global=xg_all.detach())
    if apply_update:
        out["adam_state"] = adam_update(sd, {k: g for k, g in grads.items() if k != "mapping.weight"}, {}, lr=lr)
    return out
