#!/usr/bin/env python
"""Same-box A/B of two builds of the library (gpurun_ab/librn_old.so vs the in-tree one) on the BN = 256 layers."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendernet_b200 import ops  # noqa: E402
from rendernet_b200._lib import lib as new  # noqa: E402

vp, i = C.c_void_p, C.c_int
sig = [vp, vp, vp, vp, i, vp, i, vp, vp, i, i, i, i, i, i, i, i, i, vp]
libs = [("new", new)]
for tag in ("old", "v2"):
    path = os.path.join(ROOT, "gpurun_ab", f"librn_{tag}.so")
    if os.path.exists(path):
        l = C.CDLL(path)
        l.rn_conv2d_same.argtypes = sig
        l.rn_conv2d_same.restype = i
        libs.insert(0, (tag, l))
dev = "cuda"
B = 24
torch.manual_seed(0)


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


for name, hw, Cin, Cout, k in (("res2 3x3 1024", 64, 1024, 1024, 3), ("res3 3x3 512", 32, 512, 512, 3),
                               ("e_conv5 4x4 1024->512", 64, 1024, 512, 4), ("projection 1x1", 64, 1024, 1024, 1)):
    x = torch.randn(B, hw, hw, Cin, device=dev).half()
    res = torch.randn(B, hw, hw, Cout, device=dev).half()
    w = torch.randn(k, k, Cin, Cout, device=dev) / (k * k * Cin) ** 0.5
    L = ops.pack_conv("conv2d", w, torch.zeros(Cout), torch.rand(Cout) * 0.3)
    out = torch.empty(B, hw, hw, Cout, device=dev, dtype=torch.float16)
    st = torch.cuda.current_stream().cuda_stream

    def call(lib, with_res):
        rc = lib.rn_conv2d_same(x.data_ptr(), L.w.data_ptr(), L.bias.data_ptr(), L.alpha.data_ptr(), 0 if with_res else 1,
                                res.data_ptr() if with_res else None, 0, out.data_ptr(), None, B, hw, hw, Cin, Cout,
                                L.cout_pad, k, k, 0, st)
        assert rc == 0, rc

    outs = {}
    for rnd in range(2):
        for tag, lib in libs:
            for with_res in (False, True):
                ms = timeit(lambda: call(lib, with_res))
                outs[(tag, with_res)] = out.clone()
                print(f"[oldnew] round {rnd} {name} {'residual' if with_res else 'prelu   '} {tag}: {ms:.4f} ms", flush=True)
    for with_res in (False, True):
        print(f"[oldnew] {name} residual={with_res} bit-identical across builds: "
              f"{all(torch.equal(outs[(libs[0][0], with_res)], outs[(t, with_res)]) for t, _ in libs)}", flush=True)
