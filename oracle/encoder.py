"""Oracle: PointNet_Plus_fine encoder forward (test infrastructure; see oracle/__init__.py).

Restates `PointNet_Plus_fine.forward` (reference training_code/cn3d_model_conbag.py:213-234) and
the layers built at :162-210 (identical to `PointNet_Plus` :43-91) as plain fp32 matrix algebra on
a state dict with the reference's 52 keys.  torch-CPU ops are used so autograd provides the
gradients the CUDA backward is checked against.

  L1  net3DV_1 : [1x1 conv + bias -> BatchNorm2d -> ReLU] x3, 4->64->64->256, max over K   (:162-177)
  cat          : centre xyz first, then the 256 pooled channels -> 259                      (:219)
  L3  net3DV_3 : [1x1 conv + bias -> BatchNorm2d -> ReLU] x3, 259->256->512->1024           (:180-196)
  x            : max over the S centres of a cloud                                          (:222-223)
  x_global     : max over all G*S positions of a sequence, clouds are G-major (row g*B+b)   (:225-226)
  head netR_FC : Linear 1024->1024 -> BatchNorm1d -> ReLU -> Linear 1024->512, applied to x and
                 then to x_global (two separate BN batches, two running-stat updates)       (:201-207,228-229)
  x_nor, code  : L2-normalise (eps 1e-12), bias-free Linear 512->64                         (:231-232)

BatchNorm: eps 1e-5, momentum 0.1; training mode normalises with the biased batch variance and
updates running_var with the unbiased one (torch.nn.BatchNorm semantics, which is what the
reference instantiates).
"""
from dataclasses import dataclass
import math
import torch

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# (conv key, bn key, Cin, Cout)
L1_LAYERS = [("net3DV_1.0", "net3DV_1.1", 4, 64), ("net3DV_1.3", "net3DV_1.4", 64, 64),
             ("net3DV_1.6", "net3DV_1.7", 64, 256)]
L3_LAYERS = [("net3DV_3.0", "net3DV_3.1", 259, 256), ("net3DV_3.3", "net3DV_3.4", 256, 512),
             ("net3DV_3.6", "net3DV_3.7", 512, 1024)]

STATE_KEYS = []
for _c, _b, _ci, _co in L1_LAYERS + L3_LAYERS:
    STATE_KEYS += [_c + ".weight", _c + ".bias", _b + ".weight", _b + ".bias",
                   _b + ".running_mean", _b + ".running_var", _b + ".num_batches_tracked"]
STATE_KEYS += ["netR_FC.0.weight", "netR_FC.0.bias", "netR_FC.1.weight", "netR_FC.1.bias",
               "netR_FC.1.running_mean", "netR_FC.1.running_var", "netR_FC.1.num_batches_tracked",
               "netR_FC.3.weight", "netR_FC.3.bias", "mapping.weight"]


def init_state_dict(seed=1, in_channels=4, dim=512, num_clusters=64):
    """Random state dict with the reference's key names / shapes (torch default-init scale).
    Not bit-identical to `PointNet_Plus_fine(...)` initialisation -- parity tests that need the
    reference's own init load the state dict saved by tests/golden/make_golden.py instead."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    for conv, bn, ci, co in L1_LAYERS + L3_LAYERS:
        if conv == "net3DV_1.0":
            ci = in_channels
        bound = 1.0 / math.sqrt(ci)
        sd[conv + ".weight"] = uniform((co, ci, 1, 1), bound)
        sd[conv + ".bias"] = uniform((co,), bound)
        sd[bn + ".weight"] = torch.ones(co)
        sd[bn + ".bias"] = torch.zeros(co)
        sd[bn + ".running_mean"] = torch.zeros(co)
        sd[bn + ".running_var"] = torch.ones(co)
        sd[bn + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    b = 1.0 / math.sqrt(1024)
    sd["netR_FC.0.weight"] = uniform((1024, 1024), b)
    sd["netR_FC.0.bias"] = uniform((1024,), b)
    sd["netR_FC.1.weight"] = torch.ones(1024)
    sd["netR_FC.1.bias"] = torch.zeros(1024)
    sd["netR_FC.1.running_mean"] = torch.zeros(1024)
    sd["netR_FC.1.running_var"] = torch.ones(1024)
    sd["netR_FC.1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd["netR_FC.3.weight"] = uniform((dim, 1024), b)
    sd["netR_FC.3.bias"] = uniform((dim,), b)
    sd["mapping.weight"] = uniform((num_clusters, dim), 1.0 / math.sqrt(dim))
    return sd


@dataclass
class EncoderParams:
    """Trainable leaves + BN buffers, keyed like the reference state dict."""
    sd: dict
    training: bool = True

    def trainable(self):
        return {k: v for k, v in self.sd.items() if v.dtype.is_floating_point and "running_" not in k}

    def requires_grad_(self, flag=True):
        for v in self.trainable().values():
            v.requires_grad_(flag)
        return self


def _batchnorm(z, sd, key, training):
    """z (rows, C).  Returns normalised+affine output; updates running stats in place when training."""
    gamma, beta = sd[key + ".weight"], sd[key + ".bias"]
    if training:
        n = z.shape[0]
        mean = z.mean(dim=0)
        var_b = ((z - mean) ** 2).mean(dim=0)
        with torch.no_grad():
            rm, rv = sd[key + ".running_mean"], sd[key + ".running_var"]
            rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
            rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var_b * (n / max(n - 1, 1)))
            sd[key + ".num_batches_tracked"] += 1
    else:
        mean, var_b = sd[key + ".running_mean"], sd[key + ".running_var"]
    return (z - mean) / torch.sqrt(var_b + BN_EPS) * gamma + beta


def _relu(y, routing, key):
    """ReLU; with a routing table the activity pattern is imposed instead (see encoder_forward)."""
    if routing is None or key not in routing:
        return torch.relu(y)
    return y * routing[key].to(y.dtype)


def _shared_mlp(rows, sd, layers, training, taps=None, routing=None):
    h = rows
    for conv, bn, ci, co in layers:
        w = sd[conv + ".weight"].reshape(co, -1)
        z = h @ w.t() + sd[conv + ".bias"]
        h = _relu(_batchnorm(z, sd, bn, training), routing, "relu:" + conv)
        if taps is not None:
            taps[conv] = z
    return h


def _head(feat, sd, training, routing=None, key=None):
    z = feat @ sd["netR_FC.0.weight"].t() + sd["netR_FC.0.bias"]
    h = _relu(_batchnorm(z, sd, "netR_FC.1", training), routing, key)
    return h @ sd["netR_FC.3.weight"].t() + sd["netR_FC.3.bias"]


def _pool(h, dim, routing, key):
    """max over `dim`; with a routing table the winner index is imposed instead (see encoder_forward)."""
    if routing is None or key not in routing:
        return h.max(dim=dim).values
    idx = routing[key].to(torch.int64).unsqueeze(dim)
    return h.gather(dim, idx).squeeze(dim)


def encoder_forward(params, xt, yt, gost, taps=None, routing=None):
    """xt (M,D,S,K), yt (M,3,S,1) fp32; M = gost*B clouds, G-major.
    Returns (x (M,512), code (M,64), x_nor (M,512), x_global (B,512)).

    `routing` (tests only) imposes the network's discrete decisions instead of recomputing them, so that gradients
    can be compared between two implementations whose forward values differ by rounding: the gradient of this
    network is a discontinuous function of the activations (a near-tie in a max-pool flips where the gradient is
    routed, an activation crossing 0 flips a ReLU), while for FIXED decisions it is smooth.  Keys: "k" (M*S,256)
    winner among the K neighbours, "s" (M,1024) winner among the S centres, "g" (B,1024) winning view,
    "relu:<conv key>" / "relu:head.x" / "relu:head.g" boolean activity patterns of the ReLUs."""
    sd, training = params.sd, params.training
    M, D, S, K = xt.shape
    rows = xt.permute(0, 2, 3, 1).reshape(M * S * K, D)
    h = _shared_mlp(rows, sd, L1_LAYERS, training, taps, routing)
    pooled = _pool(h.reshape(M * S, K, -1), 1, routing, "k")               # max over the K neighbours
    centre = yt.reshape(M, 3, S).permute(0, 2, 1).reshape(M * S, 3)
    h = _shared_mlp(torch.cat([centre, pooled], dim=1), sd, L3_LAYERS, training, taps, routing)
    local = h.reshape(M, S, -1)                                        # xt_local, (M,S,1024)
    feat = _pool(local, 1, routing, "s")                               # (M,1024)
    B = M // gost
    # max over all G*S positions of a sequence == max over the G per-cloud maxima
    feat_seq = _pool(feat.reshape(gost, B, -1), 0, routing, "g")       # (B,1024)
    x = _head(feat, sd, training, routing, "relu:head.x")
    x_global = _head(feat_seq, sd, training, routing, "relu:head.g")
    x_nor = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)
    code = x_nor @ sd["mapping.weight"].t()
    return x, code, x_nor, x_global
