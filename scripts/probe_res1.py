#!/usr/bin/env python
"""Where does the banded 3^3 conv (res1, 17 % of the step) lose time?  Same main loop, different epilogues /
cluster modes; plus the fused resampler+e_conv1 kernel against the unfused pair."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendernet_b200 import ops  # noqa: E402
from rendernet_b200._lib import lib  # noqa: E402
from rendernet_b200.engine import RenderEngine  # noqa: E402
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
useful = 2.0 * B * 64 * 64 * 32 * 27 * 32 * 32
for cl, cg, msub in ((2, 2, 1), (2, 2, 2), (2, 1, 2), (1, 1, 2)):
    lib.rn_set_default_cluster(cl)
    lib.rn_set_default_cta_group(cg)
    lib.rn_set_default_msub(msub)
    for ts in (1,):
        lib.rn_set_tma_store(ts)
        for label, kw in (("plain", {}), ("prelu", dict(act="prelu", alpha=al)), ("residual", dict(residual=res))):
            try:
                ms = timeit(lambda: ops.conv3d_banded(x, Lb, out16=out, **kw))
            except Exception as e:  # noqa: BLE001
                print(f"[res1] CL={cl} CG={cg} msub={msub} {label}: FAILED {e}", flush=True)
                torch.cuda.synchronize()
                continue
            print(f"[res1] CL={cl} CG={cg} msub={msub} {label:8s}: {ms:.3f} ms  {useful / ms / 1e9:6.1f} TFLOP/s useful",
                  flush=True)
lib.rn_set_default_cluster(2)
lib.rn_set_default_cta_group(2)
lib.rn_set_tma_store(1)
lib.rn_set_default_msub(0)
for rnd in range(2):
    for pre in (0, 1):
        lib.rn_set_res_prefetch(pre)
        ms = timeit(lambda: ops.conv3d_banded(x, Lb, out16=out, residual=res), iters=40)
        print(f"[res1] round {rnd} residual conv, residual prefetch (registers + L2) {'on ' if pre else 'off'}: {ms:.3f} ms", flush=True)
lib.rn_set_res_prefetch(1)

# other BN <= 128 layers of the decoder / encoder with and without M sub-tiles
def layer_probe():
    xs = {}
    specs = [("e_conv7 T s2 256->128 @64", "merged", 64, 256, 128), ("e_conv8 T s2 128->64 @128", "merged", 128, 128, 64),
             ("e_conv9 T s2 64->32 @256", "merged", 256, 64, 32), ("e_conv10 T s1 32->16 @512 xfold", "xfold", 512, 32, 16),
             ("e_conv11 T s1 16->3 @512 xfold", "xfold", 512, 16, 3), ("e_conv7_1 T s1 128->128 @128", "tconv", 128, 128, 128)]
    for name, kind, hw, cin, cout in specs:
        xx = torch.randn(B, hw, hw, cin, device=dev).half()
        ww = torch.randn(4, 4, cout, cin, device=dev) / (16 * cin) ** 0.5
        alp = torch.rand(cout, device=dev) * 0.3
        if kind == "merged":
            L = ops.MergedConvT2(ww, torch.zeros(cout))
            fn = lambda: ops.conv2d_transpose_s2_merged(xx, L, act="prelu", alpha=alp, out16=oo)
            oo = torch.empty(B, 2 * hw, 2 * hw, cout, device=dev, dtype=torch.float16)
        elif kind == "xfold":
            L = ops.XFoldConvT(ww, torch.zeros(cout), ops.XFoldConvT.factor(cin, hw))
            oo = torch.empty(B, hw, hw, cout, device=dev, dtype=torch.float16)
            fn = lambda: ops.conv2d_transpose_xfold(xx, L, act="prelu", alpha=alp, out16=oo)
        else:
            L = ops.pack_conv("conv2d_transpose", ww, torch.zeros(cout), alp, stride=1)
            oo = torch.empty(B, hw, hw, cout, device=dev, dtype=torch.float16)
            fn = lambda: ops.conv2d_transpose(xx, L, act="prelu", out16=oo)
        t = {}
        for msub in (1, 2):
            lib.rn_set_default_msub(msub)
            try:
                t[msub] = timeit(fn)
            except Exception as e:  # noqa: BLE001
                t[msub] = float("nan")
                print(f"[msub] {name} msub={msub} FAILED {e}", flush=True)
                torch.cuda.synchronize()
        print(f"[msub] {name}: msub=1 {t[1]:.3f} ms, msub=2 {t[2]:.3f} ms", flush=True)
        del xx, oo
    lib.rn_set_default_msub(0)


try:
    layer_probe()
except Exception as e:  # noqa: BLE001
    print("[msub] layer probe failed:", e, flush=True)
lib.rn_set_default_msub(0)

# fused resampler + e_conv1
rng = np.random.default_rng(0)
vox = torch.from_numpy((rng.random((B, 64, 64, 64, 1)) < 0.10).astype(np.float32)).to(dev)
pose = np.stack([rng.uniform(0, 2 * np.pi, B), (90 - rng.uniform(10, 170, B)) * np.pi / 180,
                 3.3 / rng.uniform(2.5, 4.5, B)], 1).astype(np.float32)
minv = torch.from_numpy(RenderEngine.pose_to_matrix(pose)).to(dev)
w1 = (torch.randn(5, 5, 5, 1, 8, device=dev) / 125 ** 0.5).contiguous()
b1 = torch.full((8,), 0.001, device=dev)
a1 = torch.rand(8, device=dev) * 0.3
t_f = timeit(lambda: ops.resample_conv1(vox, minv, 128, w1, b1, a1))
t_r = timeit(lambda: ops.resample(vox, minv, 128, True))
grid = ops.resample(vox, minv, 128, True)
t_c = timeit(lambda: ops.conv3d_direct(grid, w1, b1, a1, (2, 2, 2)))
same = torch.equal(ops.resample_conv1(vox, minv, 128, w1, b1, a1), ops.conv3d_direct(grid, w1, b1, a1, (2, 2, 2)))
print(f"[f1] fused resample+e_conv1 {t_f:.3f} ms vs resample {t_r:.3f} + direct conv {t_c:.3f} ms; bit-identical={same}; "
      f"non-zero grid fraction {float((grid != 0).float().mean()):.3f}", flush=True)
