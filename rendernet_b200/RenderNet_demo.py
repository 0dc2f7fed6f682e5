"""Mirror of the reference's RenderNet_demo.py: same CLI flags (:72-108), `compute_pose_param` (:33-38),
`render` (:41-66), `load_graph` (:23-30) and the feed/fetch tensor names of the frozen graph
(`real_model_in:0`, `view_name:0`, `patch_size:0`, `is_training:0` -> `encoder/output:0`).

There is no TensorFlow here: `load_graph` reads weights (frozen `.pb`, checkpoint prefix, a directory of `*.txt.npz`
in the tools/model_util.py:26-39 convention, an .npz file, or nothing -> the reference's initialisers with a fixed
seed; `tf_import.py`) and `Session.run` executes the CUDA engine.
"""
from __future__ import annotations

import argparse
import math
import os

import numpy as np

from . import Phong_shading, binvox_rw

# Phong shading parameters (RenderNet_demo.py:18-20)
AMBIENT_IN = (0.1)
K_DIFFUSE = .9
LIGHT_COL = np.array([[1., 1., 1.]])


class Graph:
    def __init__(self, weights, is_greyscale=False):
        self.weights = weights
        self.is_greyscale = is_greyscale


def load_graph(frozen_graph_filename=None, is_greyscale=None):
    """:23-30.  Accepts what the reference's tooling produces — a frozen GraphDef `.pb`
    (demo/RenderNet_converter.py), a checkpoint prefix, an npz directory (tools/model_util.py:26-39) — or a plain
    `.npz`; None -> the reference's initialisers with a fixed seed.  `is_greyscale` defaults to what the stored
    e_conv11 filter says (RenderNet_Shader.py:125-131)."""
    weights = None
    if frozen_graph_filename:
        from .tf_import import load_variables
        weights = load_variables(frozen_graph_filename)
        last = [v for k, v in weights.items() if k.replace("/", "_").endswith("e_conv11_weights")]   # encoder/e_conv11/weights
        if not last:
            raise ValueError(f"{frozen_graph_filename}: no RenderNet variables found (looked for e_conv11/weights)")
        if is_greyscale is None:
            is_greyscale = int(last[0].shape[2]) == 1      # transposed filter [kh,kw,Cout,Cin]
    return Graph(weights, bool(is_greyscale))


class Session:
    """`tf.Session(graph=graph)` look-alike for the demo's single fetch."""

    def __init__(self, graph: Graph = None, precision: str = "exact"):
        """precision: "exact" (default here: matches the reference's fp32 graph to ~1e-5 on the image whatever the
        weights) or "fast" (fp16 operands, 3x the throughput; see engine.py)."""
        self.graph = graph or Graph(None)
        self.precision = precision
        self._engines = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self._engines.clear()
        return False

    def run(self, fetches, feed_dict):
        if fetches not in ("encoder/output:0", "encoder/output"):
            raise KeyError(f"unknown fetch {fetches!r}; the frozen graph exposes 'encoder/output:0'")
        voxel = np.asarray(feed_dict["real_model_in:0"], np.float32)
        param = np.asarray(feed_dict["view_name:0"], np.float32)
        if bool(feed_dict.get("is_training:0", False)):
            raise NotImplementedError("inference only (is_training must be False)")
        B = voxel.shape[0]
        eng = self._engines.get(B)
        if eng is None:
            from .engine import RenderEngine
            eng = RenderEngine(self.graph.weights, B, is_greyscale=self.graph.is_greyscale, precision=self.precision)
            self._engines[B] = eng
        return eng.render(voxel, param).numpy().copy()


def compute_pose_param(azimuth, elevation, radius):
    """:33-38."""
    phi = azimuth * math.pi / 180.0
    theta = (90 - elevation) * math.pi / 180
    param = np.array([phi, theta, 3.3 / radius])
    param = np.expand_dims(param, axis=0)
    return param


def render(azimuth, elevation, radius, sess, voxel, light_dir, render_dir, count, light_azimuth, light_elevation,
           model_name):
    """:41-66."""
    param = compute_pose_param(azimuth, elevation, radius)
    rendered_samples = sess.run("encoder/output:0",
                                feed_dict={"real_model_in:0": voxel, "view_name:0": param, "patch_size:0": 128,
                                           "is_training:0": False})
    img_phong = Phong_shading.np_phong_composite(rendered_samples, light_dir, LIGHT_COL, AMBIENT_IN, K_DIFFUSE)
    image_out = np.clip(255. * img_phong[0], 0, 255).astype(np.uint8)
    save_path = os.path.join(render_dir, str(count).zfill(3) + "_" + model_name +
                             "_pose_%f_%f_%f_light_%f_%f.png" % (azimuth, elevation, radius, light_azimuth,
                                                                   light_elevation))
    print(save_path)
    from PIL import Image
    Image.fromarray(image_out).save(save_path)
    return image_out


def main(argv=None):
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--voxel_path', type=str, default="./voxel/Misc/bunny.binvox", help="Path to the input voxel.")
    parser.add_argument('--azimuth', type=float, default=250, help="Value of azimuth, between (0,360)")
    parser.add_argument('--elevation', type=float, default=60, help="Value of elevation, between (0,360)")
    parser.add_argument('--light_azimuth', type=float, default=250, help="Value of azimuth for light, between (0,360)")
    parser.add_argument('--light_elevation', type=float, default=60,
                        help="Value of elevation for light, between (0,360)")
    parser.add_argument('--radius', type=float, default=3.3, help="Value of radius, between (2.5, 4.5)")
    parser.add_argument('--render_dir', type=str, default='./render', help='Path to the rendered images.')
    parser.add_argument('--rotate', type=bool, default=False,
                        help='Flag rotate and render an object by 360 degree in azimuth. '
                             'Overwrites early settings in azimuth.')
    parser.add_argument('--model', type=str, default="./model/3d2d_renderer.pb",
                        help='Weights: frozen .pb (the reference hard-codes this path, RenderNet_demo.py:111), '
                             'checkpoint prefix, directory of *.txt.npz, or .npz; seeded random weights if missing.')
    args = parser.parse_args(argv)

    have_model = os.path.exists(args.model) or os.path.exists(args.model + ".index")
    graph = load_graph(args.model if have_model else None)
    with Session(graph=graph) as sess:
        os.makedirs(args.render_dir, exist_ok=True)
        light_dir = Phong_shading.generate_light_pos(args.light_elevation, args.light_azimuth)
        with open(args.voxel_path, 'rb') as f:
            voxel = np.reshape(binvox_rw.read_as_3d_array(f).data.astype(np.float32), (1, 64, 64, 64, 1))
            model_name = os.path.basename(args.voxel_path).split('.binvox')[0]
        if args.rotate:
            count = 0
            for azimuth in np.arange(0.0, 360.0, 5.0):
                render(azimuth, args.elevation, args.radius, sess, voxel, light_dir, args.render_dir, count,
                       args.light_azimuth, args.light_elevation, model_name)
                count = count + 1
        else:
            render(args.azimuth, args.elevation, args.radius, sess, voxel, light_dir, args.render_dir, 0,
                   args.light_azimuth, args.light_elevation, model_name)


if __name__ == "__main__":
    main()
