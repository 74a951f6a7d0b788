"""CPU reference arm: one training step of the UNMODIFIED reference modules (test / bench infrastructure).

`oracle/_ref/` holds byte-identical copies of the reference's `training_code/utils_my.py` and
`training_code/cn3d_model_conbag.py` (placed there by tools/vendor_reference.sh; git-ignored, shipped to the GPU
box like a built .so).  This module imports them AS THEY ARE and re-issues the loop body of
`training_code/cn3d_train_motion_GL.py:224-335` on a synthetic batch:

    permute / reshape (:225-226) -> .type(FloatTensor).cuda() (:228) -> group_points_3DV_2048 (utils_my.py:7-42;
    the r2 = 0.16 variant BASELINE configs[1] names) -> PointNet_Plus_fine (cn3d_model_conbag.py:141-234)
    -> global_contrast (utils_my.py:53-83) -> circle_contrast (:85-116) -> loss = circle + global (:329)
    -> zero_grad / backward / Adam(3e-4, (0.5, 0.999), 1e-6).step (:180, 330-332) -> loss.item() (:335)

The reference hard-codes `.cuda()`; on a host without the reference's GPU setup it runs on the CPU after
`torch.Tensor.cuda` / `nn.Module.cuda` are patched to identity -- the same patch tests/golden/make_golden.py uses.
Because of that patch this module must only be imported in a process that does not use the GPU
(`bench.py --impl reference`, or a subprocess of the cpu_baseline leg).
"""
import os
import sys
import time
import types

import numpy as np
import torch

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("utils_my.py", "cn3d_model_conbag.py"))


def load():
    """-> (utils_my, cn3d_model_conbag) reference modules, `.cuda()` patched to identity."""
    if not available():
        raise FileNotFoundError(f"{REF_DIR} is empty: run tools/vendor_reference.sh where /root/reference exists")
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import utils_my as ref_utils                      # noqa: E402  (the reference's own files)
    import cn3d_model_conbag as ref_model             # noqa: E402
    assert os.path.dirname(os.path.abspath(ref_utils.__file__)) == REF_DIR, ref_utils.__file__
    assert os.path.dirname(os.path.abspath(ref_model.__file__)) == REF_DIR, ref_model.__file__
    return ref_utils, ref_model


def make_opt(B, N, S=64, K=64):
    """The argparse defaults of cn3d_train_motion_GL.py:77-135 that the hot path reads."""
    return types.SimpleNamespace(temperal_num=3, knn_K=K, ball_radius=0.16, ball_radius2=0.25, sample_num_level1=S,
                                 sample_num_level2=64, INPUT_FEATURE_NUM=4, Num_Class=512, batchSize=B,
                                 pooling="concatenation", SAMPLE_NUM=N, learning_rate=0.0003)


class ReferenceTrainer:
    """Model + optimiser + criterion as cn3d_train_motion_GL.py:173-181 builds them (without the DataParallel wrapper,
    which is single-device there)."""

    def __init__(self, B, G, N, S=64, K=64, seed=1, threads=None):
        self.utils, self.model = load()
        if threads:
            torch.set_num_threads(threads)
        torch.manual_seed(seed)                                              # :142-144
        np.random.seed(seed)
        self.opt = make_opt(B, N, S, K)
        self.G, self.S, self.K = G, S, K
        self.netR = self.model.PointNet_Plus_fine(self.opt, gost=G, sample_num_level1=S, knn_K=K)
        self.netR.train()
        self.optimizer = torch.optim.Adam(self.netR.parameters(), lr=self.opt.learning_rate, betas=(0.5, 0.999), eps=1e-06)
        self.criterion = torch.nn.CrossEntropyLoss()

    def step(self, out_points):
        """out_points: the DataLoader tensor (B, G, N, 4).  Returns (loss value, dict of phase seconds)."""
        t = [time.perf_counter()]
        B, G, N, D = out_points.shape
        out_points = out_points.permute(1, 0, 2, 3).reshape(-1, N, D)                       # :225-226
        data1 = out_points.type(torch.FloatTensor).cuda()                                   # :228
        xt, yt = self.utils.group_points_3DV_2048(data1, self.K, self.S, SAMPLE_NUM=N)      # :230 (r2 = 0.16 variant)
        t.append(time.perf_counter())
        x, code, x_nor, x_global = self.netR(xt, yt, 1)                                     # :234
        t.append(time.perf_counter())
        loss_c = self.utils.global_contrast(G, x_global, x, self.opt, self.criterion)       # :265-287
        loss_circle = self.utils.circle_contrast(G, x, B, self.criterion)                   # :290-316
        loss = loss_circle + loss_c                                                         # :329
        t.append(time.perf_counter())
        self.optimizer.zero_grad()
        loss.backward()
        t.append(time.perf_counter())
        self.optimizer.step()
        val = loss.item()                                                                   # :335
        t.append(time.perf_counter())
        names = ["group", "forward", "loss", "backward", "adam"]
        return val, {n: t[i + 1] - t[i] for i, n in enumerate(names)}
