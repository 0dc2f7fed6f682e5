#!/usr/bin/env python
"""Secondary BASELINE configs on the B200 box (the headline config 2/3 is bench.py):
   --config 4 : Texture+Normal face render, B=24, 1 GPU (RenderNet_Texture_Face_Normal)
   --config 5 : 360-frame (and the reference's own 72-frame) azimuth turntable of one voxel grid, frames sharded over
                the ranks (run under torchrun for N > 1); voxel uploaded once
   --config 1 : single 64^3 -> 512^2 Phong render (B=1) latency through Session.run (RenderNet_demo path)
Prints one JSON line per config."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ev_time(fn, iters, warm=3):
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


def config4(args):
    from rendernet_b200.engine import TextureRenderEngine
    B = 24
    rng0, rng1, rng2 = np.random.default_rng(0), np.random.default_rng(1), np.random.default_rng(2)
    vox = (rng0.random((B, 64, 64, 64, 1)) < 0.10).astype(np.float32)
    poses = np.stack([rng1.uniform(0, 2 * np.pi, B), (90 - rng1.uniform(10, 170, B)) * np.pi / 180,
                      3.3 / rng1.uniform(2.5, 4.5, B)], axis=1).astype(np.float32)
    tex = rng2.standard_normal((B, 199)).astype(np.float32)
    eng = TextureRenderEngine(None, B)
    eng.upload(vox, tex, poses)
    ms = ev_time(eng.step_device, args.steps)
    t0 = time.perf_counter()
    for _ in range(5):
        eng.render(vox, tex, poses)
    e2e = (time.perf_counter() - t0) / 5
    print(json.dumps({"config": 4, "workload": "Texture+Normal face render B=24, 1xB200", "renders_per_sec": B / ms * 1e3,
                      "ms_per_step": ms, "e2e_renders_per_sec": B / e2e, "tflops_algorithmic": B * 0.539 / ms,
                      "gpu_launches_per_step": eng.launches_per_step}), flush=True)


def config5(args):
    import torch.distributed as dist
    from rendernet_b200.engine import RenderEngine
    from rendernet_b200.parallel import all_gather_images, shard_bounds, turntable_poses
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bv = np.load(os.path.join(ROOT, "tests/golden/binvox.npz"))
    bunny = np.unpackbits(bv["bunny_bits"]).reshape(1, 64, 64, 64, 1).astype(np.float32)
    res = {}
    for nframes in (360, 72):
        poses = turntable_poses(nframes, 60.0, 3.3)
        lo, hi = shard_bounds(nframes, world, rank)
        mine = poses[lo:hi]
        B = hi - lo if (hi - lo) <= 48 else 24
        eng = RenderEngine(None, B, device=f"cuda:{local}")
        eng.vox.copy_(torch.from_numpy(np.repeat(bunny, B, 0)).to(eng.vox.device))       # voxel uploaded once
        nchunk = -(-len(mine) // B)

        def run():
            outs = []
            for c in range(nchunk):
                p = mine[c * B:(c + 1) * B]
                if len(p) < B:
                    p = np.concatenate([p, np.repeat(p[-1:], B - len(p), 0)])
                eng.minv_host.copy_(torch.from_numpy(eng.pose_to_matrix(p)))
                eng.minv.copy_(eng.minv_host, non_blocking=True)
                outs.append(eng.step_device().clone())
            frames = torch.cat(outs)[: len(mine)]
            return all_gather_images(frames, nframes) if world > 1 else frames

        frames = run()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            frames = run()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        res[nframes] = {"frames_per_sec": nframes / dt, "seconds_per_sweep": dt, "per_gpu_batch": B, "chunks": nchunk,
                        "gathered_shape": list(frames.shape)}
        del eng
    if rank == 0:
        print(json.dumps({"config": 5, "workload": "azimuth turntable bunny.binvox el=60 r=3.3", "n_gpus": world,
                          "sweeps": res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def config1(args):
    from rendernet_b200 import Phong_shading
    from rendernet_b200.RenderNet_demo import (AMBIENT_IN, K_DIFFUSE, LIGHT_COL, Session, compute_pose_param, load_graph)
    bv = np.load(os.path.join(ROOT, "tests/golden/binvox.npz"))
    chair = np.unpackbits(bv["chair_bits"]).reshape(1, 64, 64, 64, 1).astype(np.float32)
    light = Phong_shading.generate_light_pos(60.0, 250.0)
    with Session(load_graph(None)) as sess:
        def one():
            img = sess.run("encoder/output:0", {"real_model_in:0": chair, "view_name:0": compute_pose_param(250.0, 60.0, 3.3),
                                                "patch_size:0": 128, "is_training:0": False})
            return Phong_shading.np_phong_composite_uint8(img, light, LIGHT_COL, AMBIENT_IN, K_DIFFUSE)
        one()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            u8 = one()
        dt = (time.perf_counter() - t0) / args.steps
    print(json.dumps({"config": 1, "workload": "single chair.binvox 64^3 -> 512^2 Phong render (Session.run + Phong + uint8)",
                      "seconds_per_render": dt, "renders_per_sec": 1 / dt, "out": list(u8.shape)}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    {1: config1, 4: config4, 5: config5}[a.config](a)
