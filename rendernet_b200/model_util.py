"""Mirror of the hot-path part of the reference's tools/model_util.py: the voxel->image axis transform
(:41-49) and the npz weight-directory loader (:10-39)."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from .resampling_voxel_grid import ResampledGrid


def get_weight(weight_name, weight_dict):
    """:13-24."""
    if weight_dict is None:
        return None
    return weight_dict.get(weight_name)


def load_weights(weight_dir):
    """:26-39: every `<layer>.txt.npz` in the directory -> {layer_name: arr_0}."""
    out = {}
    for path in glob.glob(os.path.join(weight_dir, "*.txt.npz")):
        with np.load(path) as data:
            out[os.path.basename(path).split('.')[0]] = data['arr_0']
    return out


def tf_transform_voxel_to_match_image(tensor_voxel):
    """:41-49: N[b,p,q,r,c] = T[b,q,P-1-p,r,c].  Fused into the resampling kernel when applied to its
    (deferred) output -- which stays deferred, so that e_conv1 can absorb it too; a plain strided copy otherwise
    (pure data movement)."""
    if isinstance(tensor_voxel, ResampledGrid):
        return tensor_voxel.transformed()
    t = tensor_voxel if isinstance(tensor_voxel, torch.Tensor) else torch.as_tensor(np.asarray(tensor_voxel))
    return torch.flip(t.permute(0, 2, 1, 3, 4), dims=(1,)).contiguous()
