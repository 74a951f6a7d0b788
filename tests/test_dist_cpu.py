"""CPU, world_size 2 over gloo: the sharding scheme of facl_b200/dist.py (rank-major key order, per-rank loss shares,
reduce of the key-side gradients) reproduces the global losses and gradients.  The arithmetic is the CPU oracle's --
this checks the host-side distribution logic, not the CUDA kernels."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from facl_b200.dist import key_index, reference_order_from_keys

G, B, C, WORLD = 4, 6, 32, 2


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        Bl = B // WORLD
        g = torch.Generator().manual_seed(100 + rank)
        x_loc = torch.randn(G * Bl, C, generator=g)              # this rank's views, G-major inside the rank
        xg_loc = torch.randn(Bl, C, generator=g)
        order = np.random.RandomState(3).permutation(G)           # same seed on every rank
        # forward exchange
        gathered = [torch.empty_like(x_loc) for _ in range(WORLD)]
        dist.all_gather(gathered, x_loc)
        keys = torch.cat(gathered, 0).requires_grad_(True)         # rank-major
        xgs = [torch.empty_like(xg_loc) for _ in range(WORLD)]
        dist.all_gather(xgs, xg_loc)
        xg_all = torch.cat(xgs, 0)
        # this rank's share: the per-sample terms of ITS samples, evaluated against all keys
        x_ref = reference_order_from_keys(keys, G, B, Bl)
        xg_in = xg_all.clone().requires_grad_(True)
        mine = slice(rank * Bl, (rank + 1) * Bl)
        share = oracle.global_contrast(G, xg_in, x_ref, B, per_sample=True)[mine].sum() + \
            oracle.circle_contrast(G, x_ref, B, order, per_sample=True)[mine].sum()
        share.backward()
        # backward exchange: key gradients are summed over ranks; the loss value too
        dkeys = keys.grad.clone()
        dist.all_reduce(dkeys)
        loss = share.detach().clone()
        dist.all_reduce(loss)
        assert float(xg_in.grad[: rank * Bl].abs().sum() + xg_in.grad[(rank + 1) * Bl:].abs().sum()) == 0.0
        if rank == 0:
            # single-process truth on the same global batch
            xr = reference_order_from_keys(keys.detach(), G, B, Bl).requires_grad_(True)
            xgr = xg_all.clone().requires_grad_(True)
            full = oracle.global_contrast(G, xgr, xr, B) + oracle.circle_contrast(G, xr, B, order)
            full.backward()
            perm = torch.as_tensor([key_index(gg, n, G, Bl) for gg in range(G) for n in range(B)])
            out["loss_err"] = abs(float(loss) - float(full)) / abs(float(full))
            out["grad_err"] = float((dkeys[perm] - xr.grad).abs().max() / xr.grad.abs().max())
            out["xg_err"] = float((xg_in.grad[mine] - xgr.grad[mine]).abs().max() / xgr.grad.abs().max())
    finally:
        dist.destroy_process_group()


def test_sharded_scheme_matches_global_loss():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(port, out), nprocs=WORLD, join=True)
    assert out["loss_err"] < 1e-5, dict(out)
    assert out["grad_err"] < 1e-4, dict(out)
    assert out["xg_err"] < 1e-4, dict(out)


def test_key_index_roundtrip():
    Bl = 3
    seen = set()
    for n in range(B):
        for g in range(G):
            j = key_index(g, n, G, Bl)
            assert 0 <= j < G * B and j not in seen
            seen.add(j)
            r, rem = divmod(j, G * Bl)
            assert r * Bl + rem % Bl == n and rem // Bl == g
