"""GPU debugging aid: per-layer activation / gradient error of the CUDA encoder against the fp64 oracle."""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from oracle.encoder import EncoderParams, encoder_forward
from facl_b200 import cn3d_model_conbag as MODELL, losses as FL, utils_my, synth

def rel2(a, b):
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1); b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))

def main(B=4, G=3, N=128, prec="fp32", seed=300):
    S = K = 64
    opt = types.SimpleNamespace(temperal_num=3, knn_K=K, ball_radius=0.16, ball_radius2=0.25, sample_num_level1=S, sample_num_level2=64,
                                INPUT_FEATURE_NUM=4, Num_Class=512, batchSize=B, pooling="concatenation", SAMPLE_NUM=N)
    pts = torch.from_numpy(synth.make_sequences(B, G, N, seed=seed, skeleton=True))
    order = synth.view_order(G, 1)
    sd0 = oracle.init_state_dict(seed=11)
    net = MODELL.PointNet_Plus_fine(opt, gost=G, sample_num_level1=S, knn_K=K)
    net.load_state_dict({k: v.clone() for k, v in sd0.items()}); net = net.cuda(); net.precision = prec; net.train()
    clouds = pts.permute(1, 0, 2, 3).reshape(-1, N, 4).float().cuda()
    xt, yt = utils_my.group_points_3DV(clouds, opt)
    x, code, xn, xg = net(xt, yt, 1)
    lg, lc = FL.contrast_losses(x, xg, G, B, order=order, prec=prec)
    (lg + lc).backward(); torch.cuda.synchronize()
    M = G * B; R3 = M * S; R1 = R3 * K
    ws = net._ws
    from facl_b200.debug import routing_of_last_forward
    routing = routing_of_last_forward(net)
    res = {}
    for dt in (torch.float32, torch.float64, "routed"):
        rt = routing if dt == "routed" else None
        key = dt
        if dt == "routed": dt = torch.float64
        sd = {k: (v.clone().to(dt) if v.dtype.is_floating_point else v.clone()) for k, v in sd0.items()}
        taps = {}
        oxt, oyt, _ = oracle.group_points(clouds.cpu(), S, K, 0.06)
        params = EncoderParams(sd, training=True).requires_grad_(True)
        ox, _, _, oxg = encoder_forward(params, oxt.to(dt), oyt.to(dt), gost=G, taps=taps, routing=rt)
        l = oracle.global_contrast(G, oxg, ox, B) + oracle.circle_contrast(G, ox, B, order)
        for t in taps.values(): t.retain_grad()
        l.backward()
        res[key] = dict(taps=taps, x=ox.detach(), xg=oxg.detach(), loss=float(l), grads={k: v.grad for k, v in params.trainable().items()})
    t64, t32, trt = res[torch.float64], res[torch.float32], res['routed']
    print(f'routed-oracle loss {trt["loss"]:.6f}  x rel {rel2(trt["x"], t64["x"]):.2e}')
    print(f"loss cuda {float(lg+lc):.6f} o32 {t32['loss']:.6f} o64 {t64['loss']:.6f}")
    names = [("net3DV_1.0", "z1", 64, R1), ("net3DV_1.3", "z2", 64, R1), ("net3DV_1.6", "z3", 256, R1),
             ("net3DV_3.0", "z4", 256, R3), ("net3DV_3.3", "z5", 512, R3), ("net3DV_3.6", "z6", 1024, R3)]
    for key, buf, C, R in names:
        z = ws.view(buf, (C, R)).t()
        print(f"{buf}: cuda-vs-o64 {rel2(z, t64['taps'][key]):.2e}   o32-vs-o64 {rel2(t32['taps'][key], t64['taps'][key]):.2e}")
    print(f"x : {rel2(x, t64['x']):.2e} (o32 {rel2(t32['x'], t64['x']):.2e});  xg: {rel2(xg, t64['xg']):.2e} (o32 {rel2(t32['xg'], t64['xg']):.2e})")
    # activation gradients (dz = grad wrt pre-BN z) are not stored by the CUDA path; compare parameter grads
    for k, p in net.named_parameters():
        if p.grad is None: continue
        g64 = t64['grads'][k]
        if g64 is None or float(g64.norm()) < 1e-12: continue
        print(f"{k:22s} cuda {rel2(p.grad, g64.reshape(p.shape)):.2e}   o32 {rel2(t32['grads'][k], g64):.2e}   routed {rel2(p.grad, trt['grads'][k].reshape(p.shape)):.2e}  |g| {float(g64.norm()):.2e}")

if __name__ == "__main__":
    a = sys.argv[1:]
    main(B=int(a[0]) if a else 4, G=int(a[1]) if len(a) > 1 else 3, N=int(a[2]) if len(a) > 2 else 128, prec=a[3] if len(a) > 3 else "fp32")
