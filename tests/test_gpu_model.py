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


def test_texture_model_matches_reference_golden(golden_dir):
    """BASELINE config 4 (reduced H,W): texture decoder + resample (C=1 and C=4) + concat + Texture/Normal RenderNet
    through the CUDA path vs the fixture frozen from the reference's own Python.  Bar: 1e-3 max-abs on both images."""
    from rendernet_b200 import ops, tfcompat as tf
    from rendernet_b200.RenderNet_Texture_Face_Normal import RenderNet, decoder_texture
    from rendernet_b200.model_util import tf_transform_voxel_to_match_image
    from rendernet_b200.resampling_voxel_grid import tf_rotation_resampling
    g = np.load(os.path.join(golden_dir, "texture_patch.npz"))
    r = np.load(os.path.join(golden_dir, "resample.npz"))
    bv = np.load(os.path.join(golden_dir, "binvox.npz"))
    chair = np.unpackbits(bv["chair_bits"]).reshape(1, 64, 64, 64, 1).astype(np.float32)
    W = orc.init_texture_weights(seed=int(g["weight_seed"]), alpha_range=tuple(g["alpha_range"]),
                                 gain=float(g["gain"]), bias_jitter=float(g["bias_jitter"]))
    tf.reset_default_graph()
    tf.load_weight_dict(W)
    tex = decoder_texture(torch.from_numpy(g["z_in"]).cuda())
    tex_t = tf.realize(tex)
    assert tuple(tex_t.shape) == (1, 64, 64, 64, 4)
    err, _ = _rel(tex_t[:, ::4, ::4, ::4], g["decoder_sub"])
    print("texture decoder max-abs err:", err)
    assert err < 1e-5
    pose = r["chair_pose"]
    n = tf_transform_voxel_to_match_image(tf_rotation_resampling(chair, pose))
    tr = tf_transform_voxel_to_match_image(tf_rotation_resampling(tex_t, pose))
    x5 = ops.concat_channels(n, tr)
    a, b, c, d = g["patch_slice"]
    img, nrm = RenderNet(x5[:, a:b, c:d].contiguous())      # x5: concat realises the two deferred grids
    e1, _ = _rel(img, g["image"]); e2, _ = _rel(nrm, g["normal"])
    print("texture net image/normal max-abs err:", e1, e2)
    assert e1 < 1e-3 and e2 < 1e-3


def test_shader_greyscale_variant_matches_oracle(golden_dir):
    """cfg['is_greyscale'] (RenderNet_Shader.py:125-131): e_conv11 has ONE output channel.  Patch-level parity vs the
    oracle with the reference's initialisers; bar 1e-3 on the sigmoid image."""
    from rendernet_b200 import tfcompat as tf
    from rendernet_b200.RenderNet_Shader import RenderNet
    r = np.load(os.path.join(golden_dir, "resample.npz"))
    bv = np.load(os.path.join(golden_dir, "binvox.npz"))
    chair = np.unpackbits(bv["chair_bits"]).reshape(1, 64, 64, 64, 1).astype(np.float32)
    n = np.ascontiguousarray(orc.transform_voxel_to_match_image(orc.rotation_resampling(chair, r["chair_pose"])))
    patch = np.ascontiguousarray(n[:, 56:72, 52:68])
    W = orc.init_shader_weights(seed=11, is_greyscale=True, alpha_range=(0.05, 0.3), bias_jitter=0.02)
    assert W["encoder/e_conv11/weights"].shape == (4, 4, 1, 16)
    tf.reset_default_graph()
    tf.load_weight_dict(W)
    img = RenderNet(torch.from_numpy(patch).cuda(), is_training=False, is_greyscale=True)
    assert img.dtype == torch.float32 and tuple(img.shape) == (1, 64, 64, 1)
    ref = orc.rendernet_shader(patch, W).numpy()
    err, _ = _rel(img, ref)
    print("greyscale image max-abs err:", err)
    assert err < 1e-3


def test_demo_render_end_to_end_png(golden_dir, tmp_path):
    """RenderNet_demo.render (:41-66) through Session.run + NumPy-convention Phong + uint8 + PNG, full size (64^3 -> 512^2),
    against the oracle's restatement of the same pipeline: the float network output must meet the 1e-3 bar, the file name
    follows :60-64, the PNG holds exactly the returned array."""
    from PIL import Image
    from rendernet_b200 import Phong_shading
    from rendernet_b200.RenderNet_demo import AMBIENT_IN, K_DIFFUSE, LIGHT_COL, Session, compute_pose_param, load_graph, render
    bv = np.load(os.path.join(golden_dir, "binvox.npz"))
    chair = np.unpackbits(bv["chair_bits"]).reshape(1, 64, 64, 64, 1).astype(np.float32)
    W = orc.init_shader_weights(seed=0, alpha_range=(0.05, 0.3))
    wfile = tmp_path / "w.npz"
    np.savez(str(wfile), **W)
    light = Phong_shading.generate_light_pos(60.0, 250.0)
    with Session(graph=load_graph(str(wfile))) as sess:
        out = render(250.0, 60.0, 3.3, sess, chair, light.copy(), str(tmp_path), 0, 250.0, 60.0, "chair")
        net = sess.run("encoder/output:0", {"real_model_in:0": chair, "view_name:0": compute_pose_param(250.0, 60.0, 3.3),
                                            "patch_size:0": 128, "is_training:0": False})
    pngs = [f for f in os.listdir(tmp_path) if f.endswith(".png")]
    assert pngs == ["000_chair_pose_250.000000_60.000000_3.300000_light_250.000000_60.000000.png"]   # RenderNet_demo.py:60-64
    assert np.array_equal(np.asarray(Image.open(tmp_path / pngs[0])), out)
    ref_net = orc.render_forward(chair, orc.compute_pose_param(250.0, 60.0, 3.3), W).numpy()
    err = float(np.abs(net - ref_net).max())
    print("demo path network output max-abs err:", err)
    assert net.shape == (1, 512, 512, 3) and err < 1e-3
    # Phong + uint8 + PNG plumbing: the oracle's restatement applied to the SAME network output must give the same picture
    # (the normalisation (img - 0.5)/|img - 0.5| amplifies the 1e-3 network tolerance, so the two halves are checked apart)
    ref_u8 = orc.to_uint8(orc.np_phong_composite(net.astype(np.float32), orc.generate_light_pos(60.0, 250.0), LIGHT_COL,
                                                  AMBIENT_IN, K_DIFFUSE))[0]
    assert out.dtype == np.uint8 and out.shape == (512, 512, 3)
    diff = np.abs(out.astype(np.int16) - ref_u8.astype(np.int16))
    print("uint8 image: max level diff", int(diff.max()), "pixels differing", int((diff > 0).sum()))
    assert int(diff.max()) <= 1


def test_engine_pipelined_submit_matches_blocking_render():
    """RenderEngine.submit/result (overlapped H2D / compute / D2H, 2 steps in flight) returns exactly what the
    blocking render() returns, for different inputs in consecutive steps; graph replay == eager."""
    from rendernet_b200.engine import RenderEngine
    rng = np.random.default_rng(0)
    B = 2
    eng = RenderEngine(None, B, seed=0)
    eager = RenderEngine.__new__(RenderEngine)
    voxs = [(rng.random((B, 64, 64, 64, 1)) < 0.1).astype(np.float32) for _ in range(3)]
    poses = [np.stack([rng.uniform(0, 6.28, B), rng.uniform(-1.3, 1.3, B), rng.uniform(0.8, 1.3, B)], 1).astype(np.float32)
             for _ in range(3)]
    want = [eng.render(v, p).clone() for v, p in zip(voxs, poses)]
    tickets, got = [], []
    for v, p in zip(voxs, poses):
        tickets.append(eng.submit(v, p))
        if len(tickets) >= 2:
            got.append(eng.result(tickets[-2]).clone())
    got.append(eng.result(tickets[-1]).clone())
    for w, g in zip(want, got):
        assert torch.equal(w, g)
    assert not torch.equal(want[0], want[1])
    assert eng.launches_per_step == 65      # 1 fused resample+e_conv1 + 64 tensor-core launches


def test_two_engines_coexist():
    """A second RenderEngine (e.g. Session.run with another batch size) resets the process-wide variable store; the
    first engine's captured graph must keep its own packed weights alive and still render the same image."""
    from rendernet_b200.engine import RenderEngine
    rng = np.random.default_rng(3)
    vox = (rng.random((1, 64, 64, 64, 1)) < 0.1).astype(np.float32)
    pose = np.array([[4.36, 0.52, 1.0]], np.float32)
    a = RenderEngine(None, 1, seed=0)
    first = a.render(vox, pose).clone()
    b = RenderEngine(None, 2, seed=1)                 # different weights, different shapes: reuses freed memory if any
    junk = [torch.randn(64 << 20, device="cuda") for _ in range(4)]      # churn the allocator
    other = b.render(np.concatenate([vox, vox]), np.concatenate([pose, pose])).clone()
    again = a.render(vox, pose).clone()
    del junk
    assert torch.equal(first, again)
    assert not torch.equal(first, other[:1])


def test_full_size_batch_independence_and_sharding():
    """BASELINE-size properties (64^3 -> 512^2, full network): (i) every item of a B=6 batch equals the same item
    rendered in a B=2 engine (renders are independent: SURVEY 8e), bit for bit -- which is also what makes batch sharding
    over ranks exact: concatenating the shard outputs reproduces the unsharded batch; (ii) eager == CUDA-graph replay;
    (iii) image values are probabilities."""
    from rendernet_b200.engine import RenderEngine
    from rendernet_b200.parallel import shard_bounds
    rng = np.random.default_rng(3)
    B = 6
    vox = (rng.random((B, 64, 64, 64, 1)) < 0.1).astype(np.float32)
    poses = np.stack([rng.uniform(0, 6.28, B), rng.uniform(-1.3, 1.3, B), rng.uniform(0.8, 1.3, B)], 1).astype(np.float32)
    full = RenderEngine(None, B, seed=0)
    ref = full.render(vox, poses).clone()
    assert tuple(ref.shape) == (B, 512, 512, 3) and float(ref.min()) >= 0.0 and float(ref.max()) <= 1.0
    eager = RenderEngine(None, B, seed=0, use_graph=False)
    assert torch.equal(eager.render(vox, poses), ref)
    del full, eager
    small = RenderEngine(None, 2, seed=0)
    parts = []
    for r in range(3):                                   # 3 "ranks" of a world-size-3 sharding
        lo, hi = shard_bounds(B, 3, r)
        parts.append(small.render(vox[lo:hi], poses[lo:hi]).clone())
    assert torch.equal(torch.cat(parts), ref)


def test_rendernet_pretrained_matches_oracle(golden_dir):
    """Reconstruct_RenderNet_Face.RenderNet_pretrained (:113-302): the Texture/Normal network built from an
    npz-keyed weight dict, ReLU residual blocks (layer_util.py:76,109), projection written as reshape + 1x1 conv `e_conv4`.
    Equivalent to the oracle's Texture net with zero residual-block alphas; patch-level parity, bar 1e-3 on both images."""
    from rendernet_b200 import tfcompat as tf
    from rendernet_b200.Reconstruct_RenderNet_Face import RenderNet_pretrained, pretrained_dict_from_texture_weights
    rng = np.random.default_rng(7)
    W = orc.init_texture_weights(seed=4, alpha_range=(0.05, 0.3), bias_jitter=0.02)
    for k in list(W):
        if k.endswith("/alpha") and "/res" in k:
            W[k] = np.zeros_like(W[k])                       # the weight_dict branch of res_block_* is ReLU
    wd = pretrained_dict_from_texture_weights(W)
    assert "e_conv4_e_conv4_weights" in wd and "Image_e_conv11_1_e_conv11_1_biases" in wd and "res2_7_con1_3X3_weights" in wd
    x5 = (rng.random((1, 16, 16, 128, 5)) * (rng.random((1, 16, 16, 128, 1)) < 0.2)).astype(np.float32)
    ref_img, ref_nrm = orc.rendernet_texture(x5, W)
    for prec in ("exact", "fast"):
        with tf.use_store(tf.VariableStore(precision=prec)):
            img, nrm = RenderNet_pretrained(torch.from_numpy(x5).cuda(), wd)
        e1, _ = _rel(img, ref_img.numpy()); e2, _ = _rel(nrm, ref_nrm.numpy())
        print(f"RenderNet_pretrained [{prec}]: albedo err {e1:.2e}, normal err {e2:.2e}")
        assert tuple(img.shape) == (1, 64, 64, 3) and e1 < 1e-3 and e2 < 1e-3


def test_tf_interpolate_standalone_bit_faithful():
    """tf_interpolate(voxel, x, y, z, out_size) (tools/resampling_voxel_grid.py:381-486) at arbitrary coordinates, including
    points outside the cube where the reference's clamped-corner weights cancel only up to fp32 rounding: the kernel
    reproduces the NumPy restatement bit for bit (same operation order, no FMA contraction)."""
    from rendernet_b200.resampling_voxel_grid import tf_interpolate
    rng = np.random.default_rng(11)
    B, S, C, n = 2, 16, 4, 5000
    vox = (rng.standard_normal((B, S, S, S, C)) * 50).astype(np.float32)       # unbounded texture-like values
    x, y, z = (rng.uniform(-6, S + 5, B * n).astype(np.float32) for _ in range(3))
    x[:64] = np.round(x[:64]); y[64:128] = np.round(y[64:128]); z[128:192] = 0.0; x[192:256] = S - 1.0    # knife edges
    out = tf_interpolate(vox, x, y, z, [B, 10, 10, n // 100, C])
    ref = orc.interpolate(vox, x, y, z)
    assert tuple(out.shape) == (B * n, C)
    assert np.array_equal(out.cpu().numpy(), ref)
    with pytest.raises(ValueError):
        tf_interpolate(vox, x[:-1], y[:-1], z[:-1], [B, 10, 10, n // 100, C])


def test_resampler_zero_outside_deviation_is_bounded_for_large_values():
    """rn_resample_f32 writes exact zeros for points outside the cube where the reference's arithmetic
    leaves cancellation noise proportional to |v|.  On a 4-channel grid with |v| ~ 100 (texture volumes are unbounded) the
    deviation from the faithful restatement stays below 2e-4 x max|v|, and is exactly zero inside the cube."""
    from rendernet_b200 import ops
    rng = np.random.default_rng(5)
    B = 2
    vox = (rng.standard_normal((B, 64, 64, 64, 4)) * 100).astype(np.float32)
    poses = np.stack([rng.uniform(0, 6.28, B), rng.uniform(-1.0, 1.0, B), rng.uniform(0.8, 1.3, B)], 1).astype(np.float32)
    ref = orc.rotation_resampling(vox, poses)                                  # faithful: noise outside the cube
    R, S = orc.rotation_around_grid_centroid(poses)
    minv = torch.from_numpy(orc.inverse_total_matrix(R, S, 64, 128)).cuda()
    out = ops.resample(torch.from_numpy(vox).cuda(), minv, 128, False).cpu().numpy()
    vmax = float(np.abs(vox).max())
    dev_ = np.abs(out - ref)
    print(f"max deviation {dev_.max():.3e} = {dev_.max() / vmax:.2e} x max|v|; nonzero outputs {int((out != 0).sum())}")
    assert dev_.max() <= 2e-4 * vmax
    assert np.array_equal(out[out != 0], ref[out != 0])                        # inside the cube: bit-identical


@pytest.mark.parametrize("fmt", [0, 2])
def test_texture_input_fusion_bit_identical(fmt):
    """rn_resample5_conv1_fused (resample C=1 + resample C=4 + axis transform + concat + e_conv1 5^3 s2 5->8 + bias + PReLU in
    one kernel, RenderNet_Texture_Face_Normal.py:155-179 + :50-53) == the unfused chain, bit for bit, in both 16-bit formats;
    and a TextureRenderEngine with / without the fusion renders identical images."""
    from rendernet_b200 import ops
    rng = np.random.default_rng(8)
    B = 2
    vox = torch.from_numpy((rng.random((B, 64, 64, 64, 1)) < 0.15).astype(np.float32)).cuda()
    tex = torch.from_numpy(rng.standard_normal((B, 64, 64, 64, 4)).astype(np.float32)).cuda()
    poses = np.stack([rng.uniform(0, 6.28, B), rng.uniform(-1.0, 1.0, B), rng.uniform(0.8, 1.3, B)], 1).astype(np.float32)
    R, S = orc.rotation_around_grid_centroid(poses)
    minv = torch.from_numpy(orc.inverse_total_matrix(R, S, 64, 128)).cuda()
    w = torch.from_numpy((rng.uniform(-1, 1, (5, 5, 5, 5, 8)) * 0.1).astype(np.float32)).cuda()
    b = torch.from_numpy(rng.uniform(-0.1, 0.1, 8).astype(np.float32)).cuda()
    al = torch.from_numpy(rng.uniform(0.05, 0.3, 8).astype(np.float32)).cuda()
    x5 = ops.concat_channels(ops.resample(vox, minv, 128, True), ops.resample(tex, minv, 128, True))
    want = ops.conv3d_direct(x5, w, b, al, (2, 2, 2), fmt=fmt)
    got = ops.resample5_conv1(vox, tex, minv, 128, w, b, al, fmt=fmt)
    a = got.planes if fmt == 2 else got
    c = want.planes if fmt == 2 else want
    assert torch.equal(a, c)
    ref = orc.prelu(orc.conv3d(x5.cpu().numpy(), w.cpu().numpy(), b.cpu().numpy(), (2, 2, 2)), al.cpu().numpy()).numpy()
    err = float(np.abs((got.float() if fmt == 2 else got.float()).cpu().numpy() - ref).max())
    assert err < (2e-5 if fmt == 2 else 5e-3) * max(1.0, float(np.abs(ref).max()))


def test_texture_engine_fused_input_matches_unfused():
    from rendernet_b200.engine import TextureRenderEngine
    rng = np.random.default_rng(9)
    B = 2
    vox = (rng.random((B, 64, 64, 64, 1)) < 0.1).astype(np.float32)
    tex = rng.standard_normal((B, 199)).astype(np.float32)
    poses = np.stack([rng.uniform(0, 6.28, B), rng.uniform(-1.0, 1.0, B), rng.uniform(0.8, 1.3, B)], 1).astype(np.float32)
    for prec in ("exact", "fast"):
        a = TextureRenderEngine(None, B, seed=0, precision=prec, fuse_input=True)
        b = TextureRenderEngine(None, B, seed=0, precision=prec, fuse_input=False)
        ia, na = (t.clone() for t in a.render(vox, tex, poses))
        ib, nb = b.render(vox, tex, poses)
        assert torch.equal(ia, ib) and torch.equal(na, nb)
        assert a.launches_per_step == b.launches_per_step - 3          # 2 resamplings + concat + conv -> 1 kernel
        del a, b


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_fused_phong_epilogue_is_bit_identical_to_the_separate_pass(precision):
    """The Phong composite + uint8 quantisation applied inside the output layer's sigmoid epilogue (RenderEngine default) ==
    the same network followed by rn_phong_composite, bit for bit (fp32 shaded image and uint8 image), B = 2, two lights,
    and == the oracle's np_phong_composite (tools/Phong_shading.py:202-228) of that normal map."""
    from rendernet_b200.engine import RenderEngine
    rng = np.random.default_rng(5)
    W = orc.init_shader_weights(seed=3, gain=1.0)
    vox = np.zeros((2, 64, 64, 64, 1), np.float32)
    vox[0, 16:48, 20:44, 12:52] = 1.0
    vox[1, 24:40, 8:56, 24:40] = 1.0
    poses = np.array([[40.0, 20.0, 3.3], [200.0, 35.0, 2.9]], np.float32)
    phong = dict(light_dir=np.concatenate([orc.generate_light_pos(60.0, 250.0), orc.generate_light_pos(30.0, 90.0)]).astype(np.float32),
                 light_col=np.array([[1.0, 1.0, 1.0], [0.9, 0.8, 0.7]], np.float32), ambient=0.3, k_diffuse=0.7)
    outs = {}
    for fuse in (True, False):
        eng = RenderEngine(W, batch=2, precision=precision, phong=phong, fuse_phong=fuse, use_graph=fuse)
        shaded, u8 = eng.render(vox, poses)
        outs[fuse] = (shaded.clone().numpy(), u8.clone().numpy(), eng.launches_per_step)
    assert outs[True][2] == outs[False][2] - 1, (outs[True][2], outs[False][2])     # one launch fewer
    assert np.array_equal(outs[True][0], outs[False][0])
    assert np.array_equal(outs[True][1], outs[False][1])
    assert outs[True][1].std() > 5          # not a constant image
    plain = RenderEngine(W, batch=2, precision=precision, use_graph=False).render(vox, poses).numpy()
    for b in range(2):
        ref = orc.np_phong_composite(plain[b:b + 1], phong["light_dir"][b:b + 1], phong["light_col"][b:b + 1], 0.3, 0.7)
        assert np.abs(ref - outs[True][0][b:b + 1]).max() < 2e-6
