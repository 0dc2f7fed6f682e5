"""GPU parity: the mirrored model function / layer ops (through the C ABI) vs the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import rendernet_oracle as orc

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = a.float().cpu().numpy() if isinstance(a, torch.Tensor) else a
    b = b.float().cpu().numpy() if isinstance(b, torch.Tensor) else b
    return float(np.abs(a - b).max()), float(np.abs(b).max())


def test_shader_patch_matches_reference_golden(golden_dir):
    """RenderNet() on the golden patch with seeded weights: CUDA path vs the fixture frozen from the
    reference's own Python, and vs the oracle's stage tensors.  Tolerance: 1e-3 max-abs on the image
    (north_star), 2e-2 relative on the unsaturated logits."""
    from rendernet_b200 import tfcompat as tf
    from rendernet_b200.RenderNet_Shader import RenderNet
    g = np.load(os.path.join(golden_dir, "shader_patch.npz"))
    r = np.load(os.path.join(golden_dir, "resample.npz"))
    bv = np.load(os.path.join(golden_dir, "binvox.npz"))
    chair = np.unpackbits(bv["chair_bits"]).reshape(1, 64, 64, 64, 1).astype(np.float32)
    n = np.ascontiguousarray(orc.transform_voxel_to_match_image(orc.rotation_resampling(chair, r["chair_pose"])))
    a, b, c, d = g["patch_slice"]
    patch = np.ascontiguousarray(n[:, a:b, c:d])
    W = orc.init_shader_weights(seed=int(g["weight_seed"]), alpha_range=tuple(g["alpha_range"]),
                                gain=float(g["gain"]), bias_jitter=float(g["bias_jitter"]))
    tf.reset_default_graph()
    tf.load_weight_dict(W)
    img = RenderNet(torch.from_numpy(patch).cuda(), is_training=False)
    assert img.dtype == torch.float32 and tuple(img.shape) == (1, 64, 64, 3)
    err, _ = _rel(img, g["image"])
    print("image max-abs err vs reference golden:", err)
    assert err < 1e-3
    ref_img, st = orc.rendernet_shader(patch, W, return_stages=True)
    logit = torch.log(img / (1 - img))
    e, s = _rel(logit, st["logits"])
    print("logits err", e, "scale", s)
    assert e < 2e-2 * s
