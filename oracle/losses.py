"""Oracle: the two contrastive losses (test infrastructure; see oracle/__init__.py).

Restates the inline "global" loss of reference training_code/cn3d_train_motion_GL.py:265-287
(= utils_my.py:53-83 `global_contrast`) and the "circle" loss of :290-316 (= utils_my.py:85-116
`circle_contrast`) in closed form.  Both are cross-entropies with label 0 over un-normalised dot
products (temperature 1); "masked" negatives are multiplied by 0, i.e. they stay in the softmax
denominator as exp(0) (utils_my.py:72,106).  CrossEntropyLoss is the default mean reduction over
the B anchors; the per-view terms are summed (utils_my.py:82,115).

x is G-major: row g*B + n is view g of sample n (cn3d_train_motion_GL.py:225-226).
"""
import torch


def _same_sample_mask(B, cols, dtype):
    """(B, cols) matrix, 0 where column j belongs to the anchor's own sample (j mod B == n)."""
    n = torch.arange(B)[:, None]
    j = torch.arange(cols)[None, :]
    return ((j % B) != n).to(dtype)


def global_contrast(num_crop, x_global, x, batch_size, per_sample=False):
    """sum_g mean_n CE([x_global[n].x[gB+n], (x_global @ x^T)[n,:] * mask], 0).
    per_sample=True returns the (B,) vector of per-anchor-sample shares (they sum to the loss): the quantity a
    rank that owns a slice of the batch contributes in the sharded multi-GPU scheme."""
    G, B = num_crop, batch_size
    xv = x.reshape(G, B, -1)
    pos = torch.einsum("nc,gnc->gn", x_global, xv)                    # (G,B)
    neg = (x_global @ x.t()) * _same_sample_mask(B, G * B, x.dtype)    # (B,GB), shared by every g
    logits = torch.cat([pos[:, :, None], neg[None].expand(G, B, G * B)], dim=2)
    terms = torch.logsumexp(logits, dim=2) - pos                       # (G,B)
    if per_sample:
        return terms.sum(dim=0) / B
    return terms.mean(dim=1).sum()


def circle_contrast(num_crop, x, batch_size, order, per_sample=False):
    """order: permutation of range(G) (the reference draws it with np.random.shuffle, :297-298).
    sum_{i<G-1} mean_n CE([x[o_i B+n].x[o_{i+1} B+n], concat_i' (x[o_i' B+n] @ x^T) * mask], 0)."""
    G, B = num_crop, batch_size
    order = [int(o) for o in order]
    xv = x.reshape(G, B, -1)
    anchors = xv[order[:-1]]                                           # (G-1,B,C)
    nxt = xv[order[1:]]
    pos = (anchors * nxt).sum(dim=2)                                   # (G-1,B)
    sims = torch.einsum("inc,kc->nik", anchors, x).reshape(B, (G - 1) * G * B)
    neg = sims * _same_sample_mask(B, (G - 1) * G * B, x.dtype)
    logits = torch.cat([pos[:, :, None], neg[None].expand(G - 1, B, neg.shape[1])], dim=2)
    terms = torch.logsumexp(logits, dim=2) - pos                       # (G-1,B)
    if per_sample:
        return terms.sum(dim=0) / B
    return terms.mean(dim=1).sum()


def info_nce_logits(x, batch_size):
    """Two-view logits of utils_my.py:200-213 (`Info_NCE`, unused by the live scripts)."""
    B = batch_size
    a, b = x[0:B], x[B:2 * B]
    mask = _same_sample_mask(B, 2 * B, x.dtype)
    pos = (a * b).sum(dim=1, keepdim=True)
    logits = torch.cat([pos, (a @ x.t()) * mask, (b @ x.t()) * mask], dim=1)
    return logits, torch.zeros(B, dtype=torch.long)
