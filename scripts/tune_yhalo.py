#!/usr/bin/env python
"""y-halo sharing on/off (and M-tile width) on the BASELINE-size 3x3 trunk conv and the banded 3^3 conv."""
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
for name, Cin, Cout in (("res2 3x3 1024->1024", 1024, 1024), ("res3 3x3 512->512", 512, 512)):
    x = torch.randn(B, 64, 64, Cin, device=dev).half()
    w = torch.randn(3, 3, Cin, Cout, device=dev) / (9 * Cin) ** 0.5
    L = ops.pack_conv("conv2d", w, torch.zeros(Cout), torch.rand(Cout) * 0.3)
    out = torch.empty(B, 64, 64, Cout, device=dev, dtype=torch.float16)
    taps = [(kx - 1, ky - 1, 0) for ky in range(3) for kx in range(3)]
    fl = 2.0 * B * 64 * 64 * Cin * Cout * 9
    ref = None
    for ny, tw in ((0, 0), (3, 16), (3, 32), (3, 8), (3, 64)):
        def run():
            ops.conv_igemm_raw(x, L.w, L.bias, taps, 2, B, 64, 64, 1, Cin, Cout, L.cout_pad, out16=out, alpha=L.alpha, act=1,
                               ny=ny, tile_w=tw)
        try:
            ms = timeit(run)
        except Exception as e:
            print(f"[yhalo] {name} ny={ny} tile_w={tw}: FAILED {e}")
            torch.cuda.synchronize()
            continue
        if ref is None:
            ref = out.float().clone()
        err = float((out.float() - ref).abs().max())
        print(f"[yhalo] {name} ny={ny} tile_w={tw or 'auto'}: {ms:.3f} ms {fl / ms / 1e9:7.1f} TFLOP/s  max|diff vs ny=0|={err:.2e}", flush=True)
x = torch.randn(B, 64, 64, 32, 32, device=dev).half()
w = torch.randn(3, 3, 3, 32, 32, device=dev) / (27 * 32) ** 0.5
Lb = ops.BandedConv3d(w, torch.zeros(32))
al = torch.rand(32, device=dev) * 0.3
out = torch.empty_like(x)
ref = None
for on in (0, 1):
    lib.rn_set_yhalo(on)
    for kps in (0, 2):
        lib.rn_set_default_kps(kps)
        ms = timeit(lambda: ops.conv3d_banded(x, Lb, act="prelu", alpha=al, out16=out))
        if ref is None:
            ref = out.float().clone()
        print(f"[yhalo] res1 3^3 banded yhalo={on} kps={kps}: {ms:.3f} ms {2.0 * B * 64 * 64 * 32 * 27 * 32 * 32 / ms / 1e9:.1f} TFLOP/s (useful) "
              f"max|diff|={float((out.float() - ref).abs().max()):.2e}", flush=True)
lib.rn_set_yhalo(1)
lib.rn_set_default_kps(0)
