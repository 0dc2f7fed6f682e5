#!/usr/bin/env python
"""One vs two epilogue warp groups on the layers whose tiles are short in K (same box, interleaved)."""
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
cases = []
x3 = torch.randn(B, 64, 64, 32, 32, device=dev).half()
r3 = torch.randn(B, 64, 64, 32, 32, device=dev).half()
w3 = torch.randn(3, 3, 3, 32, 32, device=dev) / (27 * 32) ** 0.5
Lb = ops.BandedConv3d(w3, torch.zeros(32))
al = torch.rand(32, device=dev) * 0.3
o3 = torch.empty_like(x3)
cases.append(("res1 banded prelu", lambda: ops.conv3d_banded(x3, Lb, act="prelu", alpha=al, out16=o3)))
cases.append(("res1 banded residual", lambda: ops.conv3d_banded(x3, Lb, residual=r3, out16=o3)))
for name, hw, Cin, Cout, k, use_res in (("projection 1x1 1024", 64, 1024, 1024, 1, False), ("res3 3x3 512 prelu", 32, 512, 512, 3, False),
                                        ("res3 3x3 512 residual", 32, 512, 512, 3, True), ("e_conv6 4x4 512->256", 32, 512, 256, 4, False),
                                        ("res2 3x3 1024 residual", 64, 1024, 1024, 3, True)):
    x = torch.randn(B, hw, hw, Cin, device=dev).half()
    res = torch.randn(B, hw, hw, Cout, device=dev).half() if use_res else None
    w = torch.randn(k, k, Cin, Cout, device=dev) / (k * k * Cin) ** 0.5
    L = ops.pack_conv("conv2d", w, torch.zeros(Cout), torch.rand(Cout) * 0.3)
    out = torch.empty(B, hw, hw, Cout, device=dev, dtype=torch.float16)
    cases.append((name, (lambda x=x, L=L, res=res, out=out, use_res=use_res:
                         ops.conv2d(x, L, act=None if use_res else "prelu", residual=res, out16=out))))
xt = torch.randn(B, 128, 128, 128, device=dev).half()
wt = torch.randn(4, 4, 128, 128, device=dev) / (16 * 128) ** 0.5
Lt = ops.pack_conv("conv2d_transpose", wt, torch.zeros(128), torch.rand(128) * 0.3, stride=1)
ot = torch.empty(B, 128, 128, 128, device=dev, dtype=torch.float16)
cases.append(("e_conv7_1 T s1 128->128", lambda: ops.conv2d_transpose(xt, Lt, act="prelu", out16=ot)))
for nm, cin, cout in (("e_conv10 T s1 32->16 @512 xfold", 32, 16), ("e_conv11 T s1 16->3 @512 xfold", 16, 3)):
    xx = torch.randn(B, 512, 512, cin, device=dev).half()
    ww = torch.randn(4, 4, cout, cin, device=dev) / (16 * cin) ** 0.5
    Lx = ops.XFoldConvT(ww, torch.zeros(cout), ops.XFoldConvT.factor(cin, 512))
    ox = torch.empty(B, 512, 512, cout, device=dev, dtype=torch.float16)
    ax = torch.rand(cout, device=dev) * 0.3
    cases.append((nm, (lambda xx=xx, Lx=Lx, ox=ox, ax=ax: ops.conv2d_transpose_xfold(xx, Lx, act="prelu", alpha=ax, out16=ox))))
cases = cases[-2:] + cases[:2]
for rnd in range(2):
    for name, fn in cases:
        t = {}
        for g in (1, 2):
            lib.rn_set_epilogue_groups(g)
            t[g] = timeit(fn, iters=40, warm=5)
        print(f"[eg] round {rnd} {name}: groups=1 {t[1]:.4f} ms, groups=2 {t[2]:.4f} ms ({(t[1] / t[2] - 1) * 100:+.1f} %)", flush=True)
lib.rn_set_epilogue_groups(2)
