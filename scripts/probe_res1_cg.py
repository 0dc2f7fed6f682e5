#!/usr/bin/env python
"""Banded 3^3 conv (res1): CTA-pair mode (cta_group::2) vs multicast clusters, with the residual epilogue on / off."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendernet_b200 import ops  # noqa: E402
from rendernet_b200._lib import lib  # noqa: E402
from scripts.tune_conv import timeit  # noqa: E402

dev = "cuda"
B = 24
torch.manual_seed(0)
x = torch.randn(B, 64, 64, 32, 32, device=dev).half()
res = torch.randn(B, 64, 64, 32, 32, device=dev).half()
w = torch.randn(3, 3, 3, 32, 32, device=dev) / (27 * 32) ** 0.5
Lb = ops.BandedConv3d(w, torch.zeros(32))
al = torch.rand(32, device=dev) * 0.3
out = torch.empty_like(x)
for rnd in range(3):
    for label, kw in (("prelu", dict(act="prelu", alpha=al)), ("residual", dict(residual=res))):
        for cl, cg in ((2, 2), (2, 1), (1, 1), (4, 1)):
            for pre in ((1, 0) if label == "residual" else (1,)):
                lib.rn_set_default_cluster(cl)
                lib.rn_set_default_cta_group(cg)
                lib.rn_set_res_prefetch(pre)
                ms = timeit(lambda: ops.conv3d_banded(x, Lb, out16=out, **kw), iters=40, warm=5)
                print(f"[cg] round {rnd} {label:8s} CL={cl} CG={cg} prefetch={pre}: {ms:.3f} ms", flush=True)
lib.rn_set_default_cluster(2)
lib.rn_set_default_cta_group(2)
lib.rn_set_res_prefetch(1)
