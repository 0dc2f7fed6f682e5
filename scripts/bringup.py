#!/usr/bin/env python
"""GPU bring-up: each kernel vs the CPU oracle on small seeded inputs, then timing of the big layers.
Run on the B200 box:  timeout 600 python scripts/bringup.py [--quick]
Every case is isolated in try/except so one failure does not hide the rest."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rendernet_oracle as orc  # noqa: E402
from rendernet_b200 import ops  # noqa: E402

dev = "cuda"
RESULTS = []


def report(name, got, want, tol, absolute=False):
    got = got.float().cpu().numpy() if isinstance(got, torch.Tensor) else got
    want = want.float().cpu().numpy() if isinstance(want, torch.Tensor) else want
    err = float(np.abs(got - want).max())
    scale = float(np.abs(want).max())
    ok = err <= tol if absolute else err <= tol * max(scale, 1e-6)
    RESULTS.append((name, ok, err, scale))
    print(f"[{'OK' if ok else 'FAIL'}] {name}: max_abs_err={err:.3e} ref_max={scale:.3e}", flush=True)
    return ok


def case(fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception:
        RESULTS.append((fn.__name__, False, float("nan"), 0.0))
        print(f"[EXC] {fn.__name__}\n{traceback.format_exc()}", flush=True)
        try:
            torch.cuda.synchronize()
        except Exception as e:  # sticky error: nothing else can run
            print("CUDA context is broken:", e, flush=True)
            summary()
            sys.exit(2)


def summary():
    bad = [r for r in RESULTS if not r[1]]
    print(f"\n==== {len(RESULTS) - len(bad)}/{len(RESULTS)} passed ====")
    for r in bad:
        print("   FAILED:", r[0], r[2])


def q16(a):
    return torch.from_numpy(np.asarray(a, np.float32)).half().float().numpy()


def t_resample():
    g = np.load(os.path.join(ROOT, "tests/golden/resample.npz"))
    vox, pose = g["small_vox"], g["small_pose"]
    R, S = orc.rotation_around_grid_centroid(pose)
    minv = orc.inverse_total_matrix(R, S, 16, 32)
    for tr in (0, 1):
        out = ops.resample(torch.from_numpy(vox).to(dev), torch.from_numpy(minv).to(dev), 32, bool(tr))
        report(f"resample small C=2 transform={tr}", out, g["small_net_in"] if tr else g["small_out"], 2e-4, absolute=True)
    bv = np.load(os.path.join(ROOT, "tests/golden/binvox.npz"))
    chair = np.unpackbits(bv["chair_bits"]).reshape(1, 64, 64, 64, 1).astype(np.float32)
    R, S = orc.rotation_around_grid_centroid(g["chair_pose"])
    minv = orc.inverse_total_matrix(R, S, 64, 128)
    out = ops.resample(torch.from_numpy(chair).to(dev), torch.from_numpy(minv).to(dev), 128, True)
    ref = np.zeros(128 ** 3, np.float32)
    ref[g["chair_nz_idx"]] = g["chair_nz_val"]
    report("resample chair 64->128 (+axis transform)", out.reshape(-1), ref, 2e-4, absolute=True)


def _conv2d_case(B, H, W, Cin, Cout, k, act, with_res, seed=0, res32=False):
    rng = np.random.default_rng(seed)
    x = q16(rng.standard_normal((B, H, W, Cin)))
    w = q16(rng.standard_normal((k, k, Cin, Cout)) / np.sqrt(k * k * Cin))
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    a = rng.uniform(0, 0.3, Cout).astype(np.float32)
    res = q16(rng.standard_normal((B, H, W, Cout))) if with_res else None
    L = ops.pack_conv("conv2d", torch.from_numpy(w), torch.from_numpy(b), torch.from_numpy(a))
    rt = None
    if with_res:
        rt = torch.from_numpy(res).to(dev)
        rt = rt if res32 else rt.half()
    y16, y32 = ops.conv2d(torch.from_numpy(x).to(dev).half(), L, act=act, residual=rt, want16=True, want32=True)
    ref = orc.conv2d(x, w, b)
    if act == "prelu":
        ref = orc.prelu(ref, a)
    elif act == "sigmoid":
        ref = torch.sigmoid(ref)
    if with_res:
        ref = ref + torch.from_numpy(res)
    nm = f"conv2d k{k} B{B} {H}x{W} {Cin}->{Cout} act={act} res={with_res}/{'f32' if res32 else 'f16'}"
    report(nm + " [f32 out]", y32, ref, 2e-3)
    report(nm + " [f16 out]", y16, ref, 3e-3)


def t_gemm_1x1():
    for Cin, Cout in [(64, 64), (128, 256), (16, 16), (32, 32), (64, 128), (256, 512), (64, 3), (64, 24)]:
        _conv2d_case(1, 16, 16, Cin, Cout, 1, None, False, seed=Cin + Cout)


def t_conv2d_3x3():
    _conv2d_case(2, 16, 16, 64, 64, 3, "prelu", False, seed=1)
    _conv2d_case(1, 16, 16, 128, 128, 3, None, True, seed=2)
    _conv2d_case(1, 16, 16, 128, 128, 3, None, True, seed=2, res32=True)
    _conv2d_case(1, 64, 64, 64, 256, 3, "prelu", False, seed=3)
    _conv2d_case(1, 24, 20, 32, 64, 3, "prelu", False, seed=4)     # ragged tiles
    _conv2d_case(1, 8, 8, 64, 64, 3, "sigmoid", False, seed=5)     # box taller than the image
    _conv2d_case(1, 128, 128, 32, 16, 3, "prelu", False, seed=6)
    _conv2d_case(1, 256, 256, 16, 16, 3, None, False, seed=7)


def t_conv2d_4x4():
    _conv2d_case(1, 16, 16, 64, 128, 4, "prelu", False, seed=8)
    _conv2d_case(2, 32, 32, 128, 64, 4, None, False, seed=9)


def t_conv3d():
    rng = np.random.default_rng(11)
    for (B, H, W, D, Cin, Cout) in [(1, 8, 8, 32, 32, 32), (1, 8, 8, 32, 16, 32), (2, 4, 8, 16, 16, 16), (1, 6, 5, 32, 32, 32)]:
        x = q16(rng.standard_normal((B, H, W, D, Cin)))
        w = q16(rng.standard_normal((3, 3, 3, Cin, Cout)) / np.sqrt(27 * Cin))
        b = rng.standard_normal(Cout).astype(np.float32) * 0.1
        a = rng.uniform(0, 0.3, Cout).astype(np.float32)
        res = q16(rng.standard_normal((B, H, W, D, Cout)))
        L = ops.pack_conv("conv3d", torch.from_numpy(w), torch.from_numpy(b), torch.from_numpy(a))
        xt = torch.from_numpy(x).to(dev).half()
        y = ops.conv3d(xt, L, act="prelu")
        report(f"conv3d 3^3 B{B} {H}x{W}x{D} {Cin}->{Cout} prelu", y, orc.prelu(orc.conv3d(x, w, b), a), 3e-3)
        y = ops.conv3d(xt, L, act=None, residual=torch.from_numpy(res).to(dev).half())
        report(f"conv3d 3^3 B{B} {H}x{W}x{D} {Cin}->{Cout} +res", y, orc.conv3d(x, w, b) + torch.from_numpy(res), 3e-3)


def t_conv3d_banded():
    rng = np.random.default_rng(21)
    for (B, H, W, D, Cin, Cout) in [(1, 8, 8, 32, 32, 32), (1, 8, 16, 32, 16, 32), (2, 4, 8, 16, 16, 16), (1, 6, 5, 8, 32, 32),
                                     (1, 16, 16, 4, 32, 32)]:
        x = q16(rng.standard_normal((B, H, W, D, Cin)))
        w = q16(rng.standard_normal((3, 3, 3, Cin, Cout)) / np.sqrt(27 * Cin))
        b = rng.standard_normal(Cout).astype(np.float32) * 0.1
        a = rng.uniform(0, 0.3, Cout).astype(np.float32)
        res = q16(rng.standard_normal((B, H, W, D, Cout)))
        L = ops.BandedConv3d(torch.from_numpy(w), torch.from_numpy(b))
        xt = torch.from_numpy(x).to(dev).half()
        y = ops.conv3d_banded(xt, L, act="prelu", alpha=torch.from_numpy(a).to(dev))
        report(f"conv3d banded B{B} {H}x{W}x{D} {Cin}->{Cout} prelu", y, orc.prelu(orc.conv3d(x, w, b), a), 3e-3)
        y = ops.conv3d_banded(xt, L, act=None, residual=torch.from_numpy(res).to(dev).half())
        report(f"conv3d banded B{B} {H}x{W}x{D} {Cin}->{Cout} +res", y, orc.conv3d(x, w, b) + torch.from_numpy(res), 3e-3)


def t_conv2d_transpose():
    rng = np.random.default_rng(12)
    for (B, H, W, Cin, Cout, s) in [(1, 16, 16, 64, 32, 2), (1, 16, 16, 64, 64, 1), (2, 8, 8, 256, 128, 2),
                                     (1, 32, 32, 32, 16, 1), (1, 32, 32, 16, 3, 1), (1, 64, 64, 64, 32, 2)]:
        x = q16(rng.standard_normal((B, H, W, Cin)))
        w = q16(rng.standard_normal((4, 4, Cout, Cin)) / np.sqrt(16 * Cin / (s * s)))
        b = rng.standard_normal(Cout).astype(np.float32) * 0.1
        a = rng.uniform(0, 0.3, Cout).astype(np.float32)
        L = ops.pack_conv("conv2d_transpose", torch.from_numpy(w), torch.from_numpy(b), torch.from_numpy(a), stride=s)
        y16, y32 = ops.conv2d_transpose(torch.from_numpy(x).to(dev).half(), L, act="prelu", want32=True)
        ref = orc.prelu(orc.conv2d_transpose(x, w, b, (s, s)), a)
        report(f"conv2d_transpose k4 s{s} B{B} {H}x{W} {Cin}->{Cout}", y32, ref, 2e-3)
        report(f"conv2d_transpose k4 s{s} B{B} {H}x{W} {Cin}->{Cout} [f16]", y16, ref, 3e-3)


def t_conv3d_direct():
    rng = np.random.default_rng(13)
    x = rng.random((1, 16, 16, 32, 1)).astype(np.float32)
    w = (rng.standard_normal((5, 5, 5, 1, 8)) / np.sqrt(125)).astype(np.float32)
    b = (rng.standard_normal(8) * 0.1).astype(np.float32)
    a = rng.uniform(0, 0.3, 8).astype(np.float32)
    y = ops.conv3d_direct(torch.from_numpy(x).to(dev), torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev),
                          torch.from_numpy(a).to(dev), (2, 2, 2))
    report("conv3d_direct e_conv1 5^3 s2 1->8", y, orc.prelu(orc.conv3d(x, w, b, (2, 2, 2)), a), 2e-3)
    x = q16(rng.standard_normal((1, 8, 8, 32, 8)))
    w = (rng.standard_normal((3, 3, 3, 8, 16)) / np.sqrt(27 * 8)).astype(np.float32)
    b = (rng.standard_normal(16) * 0.1).astype(np.float32)
    a = rng.uniform(0, 0.3, 16).astype(np.float32)
    y = ops.conv3d_direct(torch.from_numpy(x).to(dev).half(), torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev),
                          torch.from_numpy(a).to(dev), (1, 1, 2))
    report("conv3d_direct e_conv2 3^3 s(1,1,2) 8->16", y, orc.prelu(orc.conv3d(x, w, b, (1, 1, 2)), a), 2e-3)


def t_phong():
    g = np.load(os.path.join(ROOT, "tests/golden/phong.npz"))
    lc = torch.ones(2, 3)
    out, u8 = ops.phong_composite(torch.from_numpy(g["normal_map"]).to(dev), torch.from_numpy(g["light"]).float(), lc,
                                  0.1, 0.9, want_u8=True)
    report("phong composite (black bg)", out, g["composite"].astype(np.float32), 1e-4, absolute=True)
    d = np.abs(u8[0].cpu().numpy().astype(int) - g["uint8_first"].astype(int)).max()
    print("   uint8 max diff:", d)
    out = ops.phong_composite(torch.from_numpy(g["normal_map"]).to(dev), torch.from_numpy(g["light"]).float(), lc,
                              0.1, 0.9, background_white=True)
    report("phong composite (white bg)", out, g["composite_white"].astype(np.float32), 1e-4, absolute=True)


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


def t_big_layers():
    torch.manual_seed(0)
    B = 24
    for (name, Cin, Cout, k) in [("projection 1x1 1024->1024", 1024, 1024, 1), ("res2 3x3 1024->1024", 1024, 1024, 3),
                                 ("e_conv5 4x4 1024->512", 1024, 512, 4), ("res3 3x3 512->512", 512, 512, 3)]:
        x = torch.randn(B, 64, 64, Cin, device=dev).half()
        w = torch.randn(k, k, Cin, Cout, device=dev) / (k * k * Cin) ** 0.5
        L = ops.pack_conv("conv2d", w, torch.zeros(Cout), torch.rand(Cout) * 0.3)
        out = torch.empty(B, 64, 64, Cout, device=dev, dtype=torch.float16)
        ms = timeit(lambda: ops.conv2d(x, L, act="prelu", out16=out))
        fl = 2.0 * B * 64 * 64 * Cin * Cout * k * k
        print(f"[perf] {name}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        # check a slice against torch (cuDNN) on the same fp16 operands
        ref = torch.nn.functional.conv2d(torch.nn.functional.pad(x[:2].permute(0, 3, 1, 2).float(),
                                                                   ((k - 1) // 2, k - 1 - (k - 1) // 2) * 2),
                                         w.half().float().permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
        ref = torch.clamp(ref, min=0) + L.alpha[:Cout] * torch.clamp(ref, max=0)
        report(name + " vs torch", out[:2], ref, 3e-3)
    x = torch.randn(B, 64, 64, 32, 32, device=dev).half()
    w = torch.randn(3, 3, 3, 32, 32, device=dev) / (27 * 32) ** 0.5
    L = ops.pack_conv("conv3d", w, torch.zeros(32), torch.rand(32) * 0.3)
    out = torch.empty_like(x)
    ms = timeit(lambda: ops.conv3d(x, L, act="prelu", out16=out))
    print(f"[perf] res1 3^3 32->32: {ms:.3f} ms  {2.0 * B * 64 * 64 * 32 * 27 * 32 * 32 / ms / 1e9:.1f} TFLOP/s", flush=True)
    Lb = ops.BandedConv3d(w, torch.zeros(32))
    al = torch.rand(32, device=dev) * 0.3
    ms = timeit(lambda: ops.conv3d_banded(x, Lb, act="prelu", alpha=al, out16=out))
    print(f"[perf] res1 3^3 32->32 banded: {ms:.3f} ms  {2.0 * B * 64 * 64 * 32 * 27 * 32 * 32 / ms / 1e9:.1f} TFLOP/s (useful)", flush=True)
    vox = (torch.rand(B, 64, 64, 64, 1, device=dev) < 0.1).float()
    minv = torch.eye(4, device=dev)[:3].repeat(B, 1, 1).contiguous()
    minv[:, :, 3] = -32
    ms = timeit(lambda: ops.resample(vox, minv, 128, True))
    print(f"[perf] resample B={B}: {ms:.3f} ms  {B * 9.44e6 / ms / 1e6:.1f} GB/s (algorithmic)", flush=True)
    r = ops.resample(vox, minv, 128, True)
    w1 = torch.randn(5, 5, 5, 1, 8, device=dev) / 125 ** 0.5
    ms = timeit(lambda: ops.conv3d_direct(r, w1, torch.zeros(8, device=dev), torch.rand(8, device=dev), (2, 2, 2)))
    print(f"[perf] e_conv1 direct: {ms:.3f} ms", flush=True)
    e1 = ops.conv3d_direct(r, w1, torch.zeros(8, device=dev), torch.rand(8, device=dev), (2, 2, 2))
    w2 = torch.randn(3, 3, 3, 8, 16, device=dev) / 216 ** 0.5
    ms = timeit(lambda: ops.conv3d_direct(e1, w2, torch.zeros(16, device=dev), torch.rand(16, device=dev), (1, 1, 2)))
    print(f"[perf] e_conv2 direct: {ms:.3f} ms", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    for fn in (t_resample, t_phong, t_conv3d_direct, t_gemm_1x1, t_conv2d_3x3, t_conv2d_4x4, t_conv3d, t_conv3d_banded, t_conv2d_transpose):
        case(fn)
    if "--quick" not in sys.argv:
        case(t_big_layers)
    summary()
    print(f"elapsed {time.time() - t0:.1f}s")
    sys.exit(0 if all(r[1] for r in RESULTS) else 1)
