"""CPU oracle for the FACL hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, in plain numpy / torch-CPU fp32, the algorithms of the reference's
point-cloud-sequence encoder + contrastive-loss path (tangent-T/FACL, `training_code/`).
Every function cites the reference file:line it follows.

Who may import it: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs -- and there only as the checker / the timed CPU baseline.  Nothing under
`facl_b200/` imports this package; the product path is CUDA-only and raises when `libfacl_b200.so`
is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures of its own (SURVEY.md
section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced by importing the
unmodified reference modules from /root/reference in the authoring container
(`tests/golden/make_golden.py`, committed together with the fixtures it wrote to `tests/golden/*.npz`).
`tests/test_oracle_golden.py` re-checks the oracle against those fixtures on every CPU test run.
"""
from .fps import farthest_point_sampling, fps_reorder_indices, fps_sample_data  # noqa: F401
from .grouping import knn_ball_indices, group_points, group_points_level2  # noqa: F401
from .encoder import EncoderParams, encoder_forward, init_state_dict, STATE_KEYS  # noqa: F401
from .losses import global_contrast, circle_contrast, info_nce_logits  # noqa: F401
from . import heads  # noqa: F401
from .train_step import train_step, train_step_sharded, adam_update  # noqa: F401
