#!/usr/bin/env python
"""Programmatic dependent launch must not change a single bit: renders one seeded B = 4 batch through RenderEngine (CUDA graph,
20 replays) in two fresh processes, RN_TUNE=pdl=0 and pdl=1 (the tuning is read once per process), and compares the SHA-256 of
the images, for both precisions; also prints the launch attribute's effect on the replay time.
  python scripts/pdl_check.py            # parent: spawns the four children
"""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(precision):
    import numpy as np
    import torch
    from rendernet_b200.engine import RenderEngine
    rng = np.random.default_rng(2)
    vox = (rng.random((4, 64, 64, 64, 1)) < 0.08).astype(np.float32)
    vox[:, 20:44, 20:44, 20:44] = 1.0
    poses = np.array([[30, 20, 3.3], [120, 40, 3.0], [250, 10, 3.6], [0, 0, 3.3]], np.float32)
    eng = RenderEngine(None, batch=4, precision=precision, seed=11)      # the reference's initialisers, seeded
    img = eng.render(vox, poses).clone()
    h = hashlib.sha256(img.numpy().tobytes()).hexdigest()
    for _ in range(20):                                # replays racing each other would show up as a changed image
        eng.step_device()
    torch.cuda.synchronize()
    h2 = hashlib.sha256(eng.out.cpu().numpy().tobytes()).hexdigest()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.step_device()
    e1.record()
    torch.cuda.synchronize()
    print(f"[pdl_check] {precision} RN_TUNE={os.environ.get('RN_TUNE', '')} sha {h[:16]} replay-sha {h2[:16]} "
          f"{e0.elapsed_time(e1) / 20:.3f} ms/step (B=4)", flush=True)
    print("SHA", h, h2)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    bad = 0
    for prec in ("exact", "fast"):
        got = []
        for pdl in ("0", "1"):
            env = dict(os.environ, RN_TUNE=f"pdl={pdl}")
            p = subprocess.run([sys.executable, os.path.abspath(__file__), prec], env=env, capture_output=True, text=True)
            print("\n".join(ln for ln in p.stdout.splitlines() if ln.startswith("[pdl_check]")) or f"FAILED rc={p.returncode} {p.stderr[-400:]}")
            got.append([ln for ln in p.stdout.splitlines() if ln.startswith("SHA")])
        same = bool(got[0]) and got[0] == got[1] and got[0][0].split()[1] == got[0][0].split()[2]
        print(f"[pdl_check] {prec}: pdl=1 bit-identical to pdl=0 and stable over replays: {same}")
        bad += 0 if same else 1
    sys.exit(bad)
