#!/usr/bin/env python
"""Timing sweep of igemm_kernel variants (cluster size, N tile, k-iters per stage) on the BASELINE-size layers."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendernet_b200 import ops  # noqa: E402
from rendernet_b200._lib import lib  # noqa: E402

dev = "cuda"


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = 24
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0))
    layers = [("res2 3x3 1024->1024", 1024, 1024, 3), ("res3 3x3 512->512", 512, 512, 3),
              ("e_conv5 4x4 1024->512", 1024, 512, 4), ("projection 1x1", 1024, 1024, 1), ("e_conv6 4x4 512->256", 512, 256, 4)]
    for name, Cin, Cout, k in layers:
        x = torch.randn(B, 64, 64, Cin, device=dev).half()
        w = torch.randn(k, k, Cin, Cout, device=dev) / (k * k * Cin) ** 0.5
        L = ops.pack_conv("conv2d", w, torch.zeros(Cout), torch.rand(Cout) * 0.3)
        out = torch.empty(B, 64, 64, Cout, device=dev, dtype=torch.float16)
        taps = [(kx - (k - 1) // 2, ky - (k - 1) // 2, 0) for ky in range(k) for kx in range(k)]
        fl = 2.0 * B * 64 * 64 * Cin * Cout * k * k
        ref = None
        for bn in (256, 128):
            if Cout % bn or (bn == 128 and "--all" not in sys.argv):
                continue
            for cl, cg in ((1, 1), (2, 1), (2, 2)):
                for kps in (0, 2, 3):
                    if (kps >= 2 and bn == 256 and cg == 1) or (kps == 3 and bn == 256):
                        continue
                    def run():
                        ops.conv_igemm_raw(x, L.w, L.bias, taps, 2, B, 64, 64, 1, Cin, Cout, L.cout_pad, out16=out,
                                           alpha=L.alpha, act=1, force_bn=bn, force_kps=kps, cluster=cl, cta_group=cg)
                    try:
                        ms = timeit(run)
                    except Exception as e:
                        print(f"[tune] {name} BN={bn} CL={cl} CG={cg} kps={kps}: FAILED {e}")
                        torch.cuda.synchronize()
                        continue
                    if ref is None:
                        ref = out.clone()
                    same = torch.equal(out, ref)
                    print(f"[tune] {name} BN={bn} CL={cl} CG={cg} kps={kps or 'auto'}: {ms:.3f} ms {fl / ms / 1e9:7.1f} TFLOP/s same={same}",
                          flush=True)
    # banded conv3d
    x = torch.randn(B, 64, 64, 32, 32, device=dev).half()
    w = torch.randn(3, 3, 3, 32, 32, device=dev) / (27 * 32) ** 0.5
    Lb = ops.BandedConv3d(w, torch.zeros(32))
    al = torch.rand(32, device=dev) * 0.3
    out = torch.empty_like(x)
    ref = None
    for cl, cg, kps in ((1, 1, 0), (1, 1, 2), (1, 1, 3), (2, 1, 2), (2, 2, 0), (2, 2, 2), (2, 2, 3), (2, 2, 4)):
        lib.rn_set_default_cluster(cl)
        lib.rn_set_default_cta_group(cg)
        lib.rn_set_default_kps(kps)
        ms = timeit(lambda: ops.conv3d_banded(x, Lb, act="prelu", alpha=al, out16=out))
        if ref is None:
            ref = out.clone()
        print(f"[tune] res1 3^3 banded CL={cl} CG={cg} kps={kps}: {ms:.3f} ms {2.0 * B * 64 * 64 * 32 * 27 * 32 * 32 / ms / 1e9:.1f} TFLOP/s (useful) "
              f"same={torch.equal(out, ref)}", flush=True)
    lib.rn_set_default_cluster(2)
    lib.rn_set_default_cta_group(2)
    lib.rn_set_default_kps(0)


if __name__ == "__main__":
    main()
