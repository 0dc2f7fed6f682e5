#!/usr/bin/env python
"""Same-box A/B of whole-step time (CUDA-graph replay, B = 24) under the library's tuning switches.  Box-to-box
variance of the pool is several percent, so optimisations are judged only by runs interleaved on one box."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendernet_b200 import layer_util  # noqa: E402
from rendernet_b200._lib import lib  # noqa: E402
from rendernet_b200.engine import RenderEngine  # noqa: E402

B = 24
rng = np.random.default_rng(0)
vox = (rng.random((B, 64, 64, 64, 1)) < 0.10).astype(np.float32)
pose = np.stack([rng.uniform(0, 2 * np.pi, B), (90 - rng.uniform(10, 170, B)) * np.pi / 180,
                 3.3 / rng.uniform(2.5, 4.5, B)], 1).astype(np.float32)

configs = [("epilogue groups=1", 0, 1, True, 1), ("epilogue groups=2", 0, 1, True, 2),
           ("msub=1 res_prefetch=0 unfused groups=1 (round-1 mid state)", 1, 0, False, 1)]
engines = []
for name, msub, pre, fused, groups in configs:
    lib.rn_set_epilogue_groups(groups)
    lib.rn_set_default_msub(msub)
    lib.rn_set_res_prefetch(pre)
    layer_util.USE_FUSED_RESAMPLE_CONV1 = fused
    eng = RenderEngine(None, B, seed=0)
    eng.upload(vox, pose, non_blocking=False)
    torch.cuda.synchronize()
    engines.append((name, eng))
lib.rn_set_default_msub(0)
lib.rn_set_res_prefetch(1)
lib.rn_set_epilogue_groups(2)
layer_util.USE_FUSED_RESAMPLE_CONV1 = True

ref = None
for name, eng in engines:
    out = eng.step_device().clone()
    torch.cuda.synchronize()
    if ref is None:
        ref = out
    print(f"[ab] {name}: launches/step {eng.launches_per_step}, bit-identical to first config: {torch.equal(out, ref)}", flush=True)

for rnd in range(3):
    for name, eng in engines:
        for _ in range(3):
            eng.step_device()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            eng.step_device()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"[ab] round {rnd} {name}: {ms:.2f} ms/step  {B / ms * 1e3:.1f} renders/s", flush=True)
