"""Test / debugging helpers: read the discrete decisions (max-pool winners, ReLU activity) a training-mode
forward of the CUDA encoder took, out of its work buffers."""
import torch

_CONVS = ["net3DV_1.0", "net3DV_1.3", "net3DV_1.6", "net3DV_3.0", "net3DV_3.3", "net3DV_3.6"]
_COUT = [64, 64, 256, 256, 512, 1024]


class L1DecisionDump:
    """Context manager: asks the fused net3DV_1 backward to record its discrete decisions (see facl_debug_l1_dump)."""

    def __init__(self, M, S, K, device="cuda"):
        R1, R3 = M * S * K, M * S
        self.shape = (M, S, K)
        self.mask1 = torch.zeros(R1 * 64, dtype=torch.uint8, device=device)
        self.mask2 = torch.zeros(R1 * 64, dtype=torch.uint8, device=device)

    def __enter__(self):
        from ._lib import lib
        lib().facl_debug_l1_dump(self.mask1.data_ptr(), self.mask2.data_ptr())
        return self

    def __exit__(self, *exc):
        from ._lib import lib
        torch.cuda.synchronize()
        lib().facl_debug_l1_dump(None, None)

    def masks(self):
        M, S, K = self.shape
        R1 = M * S * K
        m1 = self.mask1.view(R1 // 64, 64, 64).permute(0, 2, 1).reshape(R1, 64).bool().cpu()     # (row, channel)
        m2 = self.mask2.view(R1 // 64, 64, 64).permute(0, 2, 1).reshape(R1, 64).bool().cpu()
        return m1, m2


def routing_of_last_forward(net, l1_dump=None):
    """-> dict understood by oracle.encoder_forward(routing=...); CPU tensors.  With the fused net3DV_1 kernels no
    per-row activation exists in memory: pass the L1DecisionDump that was active during the backward."""
    ws = net._ws
    M, S, K, G = ws.key[:4]
    B = M // G
    R3, R1, MB = M * S, M * S * K, M + B
    bn = ws.view("bn", (8, 7, 1024))
    vec = ws.view("vec", (2048,))
    out = dict(s=ws.view("arg6", (1024, MB), torch.uint8)[:, :M].t().long().cpu(),
               g=ws.view("argg", (1024, B), torch.uint8).t().long().cpu())
    out["k"] = ws.view("arg3", (256, R3), torch.uint8).t().long().cpu()          # written by pass B when fused
    fused = l1_dump is not None
    if fused:
        m1, m2 = l1_dump.masks()
        out["relu:net3DV_1.0"], out["relu:net3DV_1.3"] = m1, m2
        # ReLU3 only matters on the winner rows: impose the pooled activation's sign on the whole group
        pooled = ws.view("pcat", (259, R3))[3:]
        act = ((pooled * vec[3:259][:, None] + vec[323:579][:, None]) > 0).t().cpu()            # (M*S, 256)
        out["relu:net3DV_1.6"] = act[:, None, :].expand(R3, K, 256).reshape(R1, 256)
    for l, (conv, C) in enumerate(zip(_CONVS, _COUT)):
        if fused and l < 3:
            continue
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
