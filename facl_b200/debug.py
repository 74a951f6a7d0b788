"""Test / debugging helpers: read the discrete decisions (max-pool winners, ReLU activity) a training-mode
forward of the CUDA encoder took, out of its work buffers."""
import torch

_CONVS = ["net3DV_1.0", "net3DV_1.3", "net3DV_1.6", "net3DV_3.0", "net3DV_3.3", "net3DV_3.6"]
_COUT = [64, 64, 256, 256, 512, 1024]


def routing_of_last_forward(net):
    """-> dict understood by oracle.encoder_forward(routing=...); CPU tensors."""
    ws = net._ws
    M, S, K, G, _ = ws.key
    B = M // G
    R3, R1, MB = M * S, M * S * K, M + B
    bn = ws.view("bn", (8, 7, 1024))
    vec = ws.view("vec", (2048,))
    out = dict(k=ws.view("arg3", (256, R3), torch.uint8).t().long().cpu(),
               s=ws.view("arg6", (1024, MB), torch.uint8)[:, :M].t().long().cpu(),
               g=ws.view("argg", (1024, B), torch.uint8).t().long().cpu())
    for l, (conv, C) in enumerate(zip(_CONVS, _COUT)):
        R = R1 if l < 3 else R3
        z = ws.view(f"z{l + 1}", (C, R))
        if l == 2:
            scale, shift = vec[3:3 + C], vec[323:323 + C]
        else:
            scale, shift = bn[l, 2, :C], bn[l, 3, :C]
        out["relu:" + conv] = ((z * scale[:, None] + shift[:, None]) > 0).t().cpu()
    z7 = ws.view("z7", (1024, MB))
    out["relu:head.x"] = ((z7[:, :M] * bn[6, 2][:, None] + bn[6, 3][:, None]) > 0).t().cpu()
    out["relu:head.g"] = ((z7[:, M:] * bn[7, 2][:, None] + bn[7, 3][:, None]) > 0).t().cpu()
    return out
