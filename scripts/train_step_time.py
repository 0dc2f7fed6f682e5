#!/usr/bin/env python
"""Time of one training step of the Shader network (rendernet_b200.training.ShaderTrainer: forward with tape, loss, backward
with every weight gradient, Adam, re-pack) on one GPU, B = 1 whole frames, CUDA events around whole steps, both precisions.
  python scripts/train_step_time.py [--steps 5] [--batch 1] [--keep-prob 0.75]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rendernet_b200.training import ShaderTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--keep-prob", type=float, default=0.75)
args = ap.parse_args()
rng = np.random.default_rng(0)
B = args.batch
vox = (rng.random((B, 64, 64, 64, 1)) < 0.1).astype(np.float32)
vox[:, 16:48, 16:48, 16:48] = 1.0
poses = np.tile(np.array([[4.36, 1.05, 3.3]], np.float32), (B, 1))
target = rng.random((B, 512, 512, 3)).astype(np.float32)
for precision in ("exact", "fast"):
    tr = ShaderTrainer(None, B, precision=precision, keep_prob=args.keep_prob, learning_rate=1e-5, seed=1)
    losses = [tr.step(vox, poses, target) for _ in range(2)]                     # warm-up: allocator, first packs
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_f = t_b = t_a = 0.0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e[0].record()
        loss, grads = tr.loss_and_gradients(vox, poses, target)              # forward + loss + backward (syncs on loss.item())
        e[1].record()
        tr.apply_gradients(grads)
        e[2].record()
        torch.cuda.synchronize()
        t_f += e[0].elapsed_time(e[1])
        t_a += e[1].elapsed_time(e[2])
        losses.append(loss)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(f"[train_step_time] {precision} B={B} keep_prob={args.keep_prob}: {ms:.1f} ms/step ({B * 1000 / ms:.2f} frames/s) = forward+loss+backward "
          f"{t_f / args.steps:.1f} ms + Adam {t_a / args.steps:.1f} ms; variables {sum(v.numel() for v in tr.store.vars.values()) / 1e6:.1f} M; "
          f"loss {losses[0]:.5f} -> {losses[-1]:.5f}; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del tr
    torch.cuda.empty_cache()
