"""CPU: the operand packing that puts layer 0 (K = 4) on the tensor pipe (facl_b200/csrc/l1_fused.cu, passes A / C / D).

One K = 16 bf16 instruction per tile computes  z1' = s1 (W1 x + b1) + t1  from operand rows

    A[ch]  = [ Wh0..3 | Wh0..3 | Wl0..3 | bh bl 0 0 ]        W = s1 W1, b = s1 b1 + t1, each split into bf16 hi + bf16 lo
    B[row] = [ xh0..3 | xl0..3 | xh0..3 | 1  1  0 0 ]

so that A.B = Wh.xh + Wh.xl + Wl.xh + bh + bl with fp32 accumulation.  This test restates the packing in torch and checks (1) that it
is the bf16x3 product it claims to be, (2) its error against fp64: the same 2^-16 class as every other layer's split products, far
inside the 1e-3 bound of the fp32 mode, (3) that the ReLU decisions differ from the exact ones only where |z1'| is at rounding level.
"""
import torch


def split(t):
    hi = t.to(torch.bfloat16)
    lo = (t - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


def pack_rows(W, b, x):
    """W (C, 4), b (C,), x (R, 4) fp32 -> A (C, 16), B (R, 16) bf16, laid out as the kernels write them."""
    Wh, Wl = split(W)
    bh, bl = split(b)
    xh, xl = split(x)
    C, R = W.shape[0], x.shape[0]
    z2 = torch.zeros((C, 2), dtype=torch.bfloat16)
    A = torch.cat([Wh, Wh, Wl, bh[:, None], bl[:, None], z2], dim=1)
    one = torch.ones((R, 2), dtype=torch.bfloat16)
    B = torch.cat([xh, xl, xh, one, torch.zeros((R, 2), dtype=torch.bfloat16)], dim=1)
    return A, B


def test_packing_is_the_three_term_product_and_meets_the_bound():
    g = torch.Generator().manual_seed(7)
    C, R = 64, 4096
    W1 = torch.randn(C, 4, generator=g) * 0.5
    b1 = torch.randn(C, generator=g) * 0.1
    s1 = torch.rand(C, generator=g) + 0.5                       # BatchNorm-1 scale / shift folded into the layer
    t1 = torch.randn(C, generator=g) * 0.3
    x = torch.randn(R, 4, generator=g)
    x[:, 3] = torch.rand(R, generator=g)                        # the motion channel is in [0, 1]
    W = s1[:, None] * W1
    b = s1 * b1 + t1
    A, B = pack_rows(W, b, x)
    assert A.shape == (C, 16) and B.shape == (R, 16)
    # the instruction: exact bf16 x bf16 products, fp32 accumulation (emulated in fp64 and rounded: upper bound on the fp32 path)
    z_tc = (A.double() @ B.double().t()).float()                                   # (C, R)
    Wh, Wl = split(W)
    bh, bl = split(b)
    xh, xl = split(x)
    three = (Wh.double() @ xh.double().t() + Wh.double() @ xl.double().t() + Wl.double() @ xh.double().t()
             + (bh.double() + bl.double())[:, None]).float()
    assert torch.equal(z_tc, three)                                                # (1) the K slots line up
    exact = W.double() @ x.double().t() + b.double()[:, None]
    scale = float(exact.abs().max())
    err = float((z_tc.double() - exact).abs().max()) / scale
    assert err <= 2.0 ** -15, err                                                  # (2) dropped term Wl.xl ~ 2^-18, roundings 2^-17
    flips = (z_tc > 0) != (exact > 0)
    assert float(flips.float().mean()) <= 1e-4                                     # (3) ReLU decisions: only at rounding level
    assert float(exact[flips].abs().max()) <= 2.0 ** -14 * scale if bool(flips.any()) else True
