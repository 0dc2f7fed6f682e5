#!/usr/bin/env python
"""Full-size end-to-end parity + first timing on the B200 box.
  python scripts/e2e_check.py [--gain 1.1] [--batch 24] [--skip-oracle]
1) chair.binvox-equivalent input (from the golden bit-packed fixture), demo pose, B=1: CUDA path vs the CPU
   oracle, per stage and on the final image;  2) B=24 synthetic batch timing (graph replay and end to end)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rendernet_oracle as orc  # noqa: E402
from rendernet_b200 import ops, tfcompat as tf  # noqa: E402
from rendernet_b200.RenderNet_Shader import RenderNet  # noqa: E402
from rendernet_b200.engine import RenderEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gain", type=float, default=1.1)
    ap.add_argument("--batch", type=int, default=24)
    ap.add_argument("--skip-oracle", action="store_true")
    ap.add_argument("--skip-timing", action="store_true")
    args = ap.parse_args()
    print(torch.cuda.get_device_name(0), "cpu threads:", torch.get_num_threads(), flush=True)
    bv = np.load(os.path.join(ROOT, "tests/golden/binvox.npz"))
    chair = np.unpackbits(bv["chair_bits"]).reshape(1, 64, 64, 64, 1).astype(np.float32)
    pose = orc.compute_pose_param(250.0, 60.0, 3.3)
    t0 = time.time()
    W = orc.init_shader_weights(seed=1, alpha_range=(0.05, 0.3), gain=args.gain, bias_jitter=0.02)
    print(f"weights generated in {time.time() - t0:.1f}s", flush=True)

    if not args.skip_oracle:
        t0 = time.time()
        ref_img, st = orc.render_forward(chair, pose, W, return_stages=True)
        t_cpu = time.time() - t0
        print(f"oracle B=1 forward: {t_cpu:.2f}s on {torch.get_num_threads()} threads", flush=True)
        tf.reset_default_graph()
        tf.load_weight_dict(W)
        minv = RenderEngine.pose_to_matrix(pose)
        grid = ops.resample(torch.from_numpy(chair).cuda(), torch.from_numpy(minv).cuda(), 128, True)
        stages = {}
        img = RenderNet(grid, is_training=False, stages=stages)
        torch.cuda.synchronize()
        stages["rotated"] = grid
        stages["logits"] = torch.log(img.double() / (1 - img.double())).float()
        for k in ("rotated", "enc3", "enc3_skip", "enc4", "enc4_skip", "enc5_skip", "enc10", "logits"):
            a = stages[k].float().cpu().numpy()
            b = st[k].float().numpy()
            err = np.abs(a - b)
            print(f"  stage {k:10s} max_abs_err={err.max():.3e}  rms_err={np.sqrt((err ** 2).mean()):.3e}  "
                  f"ref_absmax={np.abs(b).max():.3e} ref_rms={np.sqrt((b ** 2).mean()):.3e}", flush=True)
        e = np.abs(img.cpu().numpy() - ref_img.numpy())
        print(f"IMAGE max_abs_err={e.max():.3e} mean_abs_err={e.mean():.3e}  (bar 1e-3)  "
              f"image range [{ref_img.min():.3f},{ref_img.max():.3f}]", flush=True)

    if args.skip_timing:
        return
    B = args.batch
    rng0, rng1 = np.random.default_rng(0), np.random.default_rng(1)
    vox = (rng0.random((B, 64, 64, 64, 1)) < 0.10).astype(np.float32)
    poses = np.stack([rng1.uniform(0, 2 * np.pi, B), (90 - rng1.uniform(10, 170, B)) * np.pi / 180,
                      3.3 / rng1.uniform(2.5, 4.5, B)], axis=1).astype(np.float32)
    t0 = time.time()
    eng = RenderEngine(W, B)
    print(f"engine built (pack + warm-up + capture) in {time.time() - t0:.1f}s; "
          f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    eng.upload(vox, poses)
    for _ in range(3):
        eng.step_device()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    K = 10
    e0.record()
    for _ in range(K):
        eng.step_device()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"[perf] graph replay B={B}: {ms:.2f} ms/step -> {B / ms * 1e3:.1f} renders/s "
          f"({B * 2.114 / ms:.1f} TFLOP/s algorithmic)", flush=True)
    t0 = time.time()
    for _ in range(5):
        out = eng.render(vox, poses)
    dt = (time.time() - t0) / 5
    print(f"[perf] end-to-end (H2D + graph + D2H) B={B}: {dt * 1e3:.2f} ms/step -> {B / dt:.1f} renders/s", flush=True)
    print("output", tuple(out.shape), out.dtype, float(out.min()), float(out.max()))


if __name__ == "__main__":
    main()
