"""GPU tests of the "exact" precision mode (RN_FMT_F16X2: fp16 hi/lo operand pairs, three tensor-core products per tap)
and full-size oracle parity for every BASELINE config.

The reference computes every convolution in fp32 (tools/layer_util.py:171,212,253; slim.conv2d RenderNet_Shader.py:83-129).
Kernel-level tests here compare against float64 PyTorch-CPU convolutions of the UNROUNDED fp32 inputs -- unlike
tests/test_gpu_kernels.py, which rounds both sides to fp16 and therefore proves kernel correctness, not precision.
Tolerances are stated per test: exact mode ~1e-5 of the output scale per layer (fp32-accumulation level), image bar 1e-3
(north_star); the fast mode is asserted at its measured bound and reported.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import rendernet_oracle as orc

pytestmark = pytest.mark.gpu
dev = "cuda"


def _ops():
    from rendernet_b200 import ops
    return ops


def _wide(rng, shape, scale=1.0):
    """fp32 values spanning several binades (incl. magnitudes whose LO half lands in fp16's subnormal range)."""
    return (rng.standard_normal(shape) * np.exp(rng.uniform(-6.0, 1.0, shape)) * scale).astype(np.float32)


def _err(got, want):
    got = got.double().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got, np.float64)
    want = want.double().cpu().numpy() if isinstance(want, torch.Tensor) else np.asarray(want, np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    return float(np.abs(got - want).max()), float(np.abs(want).max())


def _prelu64(x, a):
    return torch.clamp(x, min=0) + a * torch.clamp(x, max=0)


def _conv2d_f64(x, w, b, pads):
    """x [B,H,W,Ci], w [kh,kw,Ci,Co] (TF), SAME with explicit (top, bottom, left, right) pads -> [B,H,W,Co] float64."""
    xt = torch.from_numpy(x).double().permute(0, 3, 1, 2)
    xt = F.pad(xt, (pads[2], pads[3], pads[0], pads[1]))
    y = F.conv2d(xt, torch.from_numpy(w).double().permute(3, 2, 0, 1), torch.from_numpy(b).double())
    return y.permute(0, 2, 3, 1).contiguous()


# ----------------------------------------------------------------------------------------- representation
def test_split_pair_roundtrip_precision():
    """cast fp32 -> (hi, lo) -> fp32: relative error <= 2^-21 for normal-range values, absolute <= 2^-24 below."""
    ops = _ops()
    rng = np.random.default_rng(0)
    x = _wide(rng, (1 << 16,), 4.0)
    x[:8] = [0.0, 1.0, -1.0, 65504.0, 1e-7, -3e-5, 0.1, 1.0 / 3.0]
    xt = torch.from_numpy(x).to(dev)
    s = ops.cast_to_16(xt, fmt=2)
    assert isinstance(s, ops.Split16) and tuple(s.shape) == tuple(xt.shape) and tuple(s.planes.shape) == (2,) + tuple(xt.shape)
    assert torch.equal(s.planes[0], xt.half())                                  # hi plane == the fast mode's rounding
    back = ops.cast_to_f32(s)
    err = (back.double() - xt.double()).abs()
    bound = torch.maximum(xt.double().abs() * 2.0 ** -21, torch.full_like(err, 2.0 ** -24))
    assert bool((err <= bound).all()), float((err / bound).max())
    assert torch.equal(back, s.float())


# ----------------------------------------------------------------------------------------- kernels vs float64
@pytest.mark.parametrize("k,cin,cout,hw,B", [(3, 128, 256, 32, 2),      # 3x3 trunk shape class: BN 256, CTA pairs, y-halo
                                              (1, 1024, 1024, 64, 1),    # projection unit: 1x1, K = 1024
                                              (4, 64, 64, 16, 2),        # 4x4 (e_conv5/6 tap set): 16 taps x 3 = 48 pseudo-taps
                                              (3, 64, 128, 20, 1),       # ragged tiles (20 % 8 != 0), BN 128, M sub-tiles
                                              (3, 32, 16, 16, 3)])       # thin N tile
def test_exact_conv2d_matches_float64(k, cin, cout, hw, B):
    """SAME conv + bias + PReLU + residual in exact mode vs a float64 convolution of the unrounded fp32 tensors.
    Tolerance 2e-5 of the output scale on the fp32 output (fp32 accumulation of K <= 1152..1024 terms); the stored hi/lo
    pair carries it to 2e-5 too; the fast mode on the same inputs is > 20x worse."""
    ops = _ops()
    rng = np.random.default_rng(k * 1000 + cin)
    x = _wide(rng, (B, hw, hw, cin))
    w = (rng.uniform(-1, 1, (k, k, cin, cout)) * np.sqrt(6.0 / (k * k * (cin + cout)))).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, cout).astype(np.float32)
    al = rng.uniform(0.05, 0.3, cout).astype(np.float32)
    res = _wide(rng, (B, hw, hw, cout))
    pb = (k - 1) // 2
    ref = _prelu64(_conv2d_f64(x, w, b, (pb, k - 1 - pb, pb, k - 1 - pb)), torch.from_numpy(al).double()) + torch.from_numpy(res).double()
    xt, rt = torch.from_numpy(x).to(dev), torch.from_numpy(res).to(dev)
    out = {}
    for name, fmt in (("exact", 2), ("fast", 0)):
        L = ops.pack_conv("conv2d", torch.from_numpy(w), torch.from_numpy(b), torch.from_numpy(al), device=dev, fmt=fmt)
        xs, rs = ops.cast_to_16(xt, fmt=fmt), ops.cast_to_16(rt, fmt=fmt)
        y16, y32 = ops.conv2d(xs, L, act="prelu", residual=rs, want16=True, want32=True)
        out[name] = (_err(y32, ref), _err(y16.float(), ref))
    (e32, s), (e16, _) = out["exact"]
    (f32, _), _ = out["fast"]
    print(f"k={k} {cin}->{cout}: exact fp32-out err {e32:.2e}, pair-out err {e16:.2e}, fast err {f32:.2e} (scale {s:.2e})")
    assert e32 <= 2e-5 * s and e16 <= 2e-5 * s
    assert f32 > 20 * e32


def test_exact_conv2d_f32_residual_and_sigmoid():
    """fp32 residual + fp32-only output and the sigmoid epilogue in exact mode."""
    ops = _ops()
    rng = np.random.default_rng(5)
    x = _wide(rng, (1, 16, 16, 64))
    w = (rng.uniform(-1, 1, (3, 3, 64, 32)) * 0.05).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, 32).astype(np.float32)
    res = rng.standard_normal((1, 16, 16, 32)).astype(np.float32)
    L = ops.pack_conv("conv2d", torch.from_numpy(w), torch.from_numpy(b), None, device=dev, fmt=2)
    xs = ops.cast_to_16(torch.from_numpy(x).to(dev), fmt=2)
    y = ops.conv2d(xs, L, act=None, residual=torch.from_numpy(res).to(dev), want16=False, want32=True)
    ref = _conv2d_f64(x, w, b, (1, 1, 1, 1)) + torch.from_numpy(res).double()
    e, s = _err(y, ref)
    assert e <= 2e-5 * s
    y = ops.conv2d(xs, L, act="sigmoid", want16=False, want32=True)
    e, _ = _err(y, torch.sigmoid(_conv2d_f64(x, w, b, (1, 1, 1, 1))))
    assert e <= 3e-6                                  # __expf-based sigmoid in the epilogue


@pytest.mark.parametrize("cin,cout,sz,D", [(32, 32, 1, 32), (16, 32, 1, 32), (8, 16, 2, 64), (16, 16, 1, 32)])
def test_exact_conv3d_banded_matches_float64(cin, cout, sz, D):
    """Depth-folded 3^3 conv3d (res_block_3d / e_conv2 / e_conv3 shapes) in exact mode, PReLU then hi/lo residual."""
    ops = _ops()
    rng = np.random.default_rng(cin * 7 + sz)
    B, H, W = 1, 16, 16
    x = _wide(rng, (B, H, W, D, cin))
    w = (rng.uniform(-1, 1, (3, 3, 3, cin, cout)) * np.sqrt(6.0 / (27 * (cin + cout)))).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, cout).astype(np.float32)
    al = rng.uniform(0.05, 0.3, cout).astype(np.float32)
    Do = -(-D // sz)
    res = _wide(rng, (B, H, W, Do, cout))
    xt = torch.from_numpy(x).double().permute(0, 4, 1, 2, 3)
    pz = orc.same_pads(D, 3, sz)
    xt = F.pad(xt, (pz[0], pz[1], 1, 1, 1, 1))
    y = F.conv3d(xt, torch.from_numpy(w).double().permute(4, 3, 0, 1, 2), torch.from_numpy(b).double(), stride=(1, 1, sz))
    ref = _prelu64(y.permute(0, 2, 3, 4, 1), torch.from_numpy(al).double()) + torch.from_numpy(res).double()
    L = ops.BandedConv3d(torch.from_numpy(w), torch.from_numpy(b), device=dev, sz=sz, fmt=2)
    xs = ops.cast_to_16(torch.from_numpy(x).to(dev), fmt=2)
    rs = ops.cast_to_16(torch.from_numpy(res).to(dev), fmt=2)
    out = ops.conv3d_banded(xs, L, act="prelu", residual=rs, alpha=torch.from_numpy(al).to(dev))
    assert isinstance(out, ops.Split16) and tuple(out.shape) == (B, H, W, Do, cout)
    e, s = _err(out.float(), ref)
    print(f"banded {cin}->{cout} sz={sz}: err {e:.2e} scale {s:.2e}")
    assert e <= 2e-5 * s


@pytest.mark.parametrize("cin,cout,stride,hw,merged", [(64, 32, 2, 16, True), (64, 32, 2, 16, False), (128, 128, 1, 16, False)])
def test_exact_conv2d_transpose_matches_float64(cin, cout, stride, hw, merged):
    """k=4 SAME transposed convs (e_conv7..9 stride 2 in the merged-phase and the 4-launch phase form; e_conv7_1 stride 1)."""
    ops = _ops()
    rng = np.random.default_rng(cin + stride)
    B = 2
    x = _wide(rng, (B, hw, hw, cin))
    w = (rng.uniform(-1, 1, (4, 4, cout, cin)) * np.sqrt(6.0 / (16 * (cin + cout)))).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, cout).astype(np.float32)
    al = rng.uniform(0.05, 0.3, cout).astype(np.float32)
    xt = torch.from_numpy(x).double().permute(0, 3, 1, 2)
    wt = torch.from_numpy(w).double().permute(3, 2, 0, 1)                         # [Cin, Cout, kh, kw]
    y = F.conv_transpose2d(xt, wt, torch.from_numpy(b).double(), stride=stride, padding=1)
    if stride == 1:
        y = y[:, :, :hw, :hw]                                                    # TF SAME crops to in*stride
    ref = _prelu64(y.permute(0, 2, 3, 1), torch.from_numpy(al).double())
    xs = ops.cast_to_16(torch.from_numpy(x).to(dev), fmt=2)
    if merged:
        L = ops.MergedConvT2(torch.from_numpy(w), torch.from_numpy(b), device=dev, fmt=2)
        out = ops.conv2d_transpose_s2_merged(xs, L, act="prelu", alpha=torch.from_numpy(al).to(dev))
    else:
        L = ops.pack_conv("conv2d_transpose", torch.from_numpy(w), torch.from_numpy(b), torch.from_numpy(al), stride=stride,
                          device=dev, fmt=2)
        out = ops.conv2d_transpose(xs, L, act="prelu")
    e, s = _err(out.float(), ref)
    print(f"tconv {cin}->{cout} s{stride} merged={merged}: err {e:.2e} scale {s:.2e}")
    assert e <= 2e-5 * s


@pytest.mark.parametrize("cin,cout", [(32, 16), (16, 3)])
def test_exact_xfold_transposed_conv_matches_float64(cin, cout):
    """e_conv10 / e_conv11: x-folded stride-1 transposed conv, PReLU -> pair output, sigmoid -> fp32 output."""
    ops = _ops()
    rng = np.random.default_rng(cin)
    B, H, W = 1, 24, 64
    x = _wide(rng, (B, H, W, cin))
    w = (rng.uniform(-1, 1, (4, 4, cout, cin)) * np.sqrt(6.0 / (16 * (cin + cout)))).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, cout).astype(np.float32)
    al = rng.uniform(0.05, 0.3, cout).astype(np.float32)
    y = F.conv_transpose2d(torch.from_numpy(x).double().permute(0, 3, 1, 2), torch.from_numpy(w).double().permute(3, 2, 0, 1),
                           torch.from_numpy(b).double(), stride=1, padding=1)[:, :, :H, :W].permute(0, 2, 3, 1)
    Fx = ops.XFoldConvT.factor(cin, W)
    assert Fx > 1
    L = ops.XFoldConvT(torch.from_numpy(w), torch.from_numpy(b), Fx, device=dev, fmt=2)
    xs = ops.cast_to_16(torch.from_numpy(x).to(dev), fmt=2)
    if cout == 3:
        out = ops.conv2d_transpose_xfold(xs, L, act="sigmoid", want16=False, want32=True)
        e, _ = _err(out, torch.sigmoid(y))
        assert e <= 3e-6
    else:
        out = ops.conv2d_transpose_xfold(xs, L, act="prelu", alpha=torch.from_numpy(al).to(dev))
        e, s = _err(out.float(), _prelu64(y, torch.from_numpy(al).double()))
        assert e <= 2e-5 * s


def test_exact_fused_resample_conv1_planes():
    """rn_resample_conv1_fused / rn_conv3d_direct with fmt 2: the HI plane is bit-identical to the fast mode's output and
    hi + lo reproduces the fp32 value to 2^-21."""
    ops = _ops()
    rng = np.random.default_rng(2)
    B = 2
    vox = torch.from_numpy((rng.random((B, 64, 64, 64, 1)) < 0.2).astype(np.float32)).to(dev)
    poses = np.stack([rng.uniform(0, 6.28, B), rng.uniform(-1.0, 1.0, B), rng.uniform(0.8, 1.3, B)], 1).astype(np.float32)
    R, S = orc.rotation_around_grid_centroid(poses)
    minv = torch.from_numpy(orc.inverse_total_matrix(R, S, 64, 128)).to(dev)
    w = torch.from_numpy((rng.uniform(-1, 1, (5, 5, 5, 1, 8)) * 0.2).astype(np.float32)).to(dev)
    b = torch.from_numpy(rng.uniform(-0.1, 0.1, 8).astype(np.float32)).to(dev)
    al = torch.from_numpy(rng.uniform(0.05, 0.3, 8).astype(np.float32)).to(dev)
    fast = ops.resample_conv1(vox, minv, 128, w, b, al, fmt=0)
    pair = ops.resample_conv1(vox, minv, 128, w, b, al, fmt=2)
    assert torch.equal(pair.planes[0], fast)
    grid = ops.resample(vox, minv, 128, True)
    direct = ops.conv3d_direct(grid, w, b, al, (2, 2, 2), fmt=2)
    assert torch.equal(direct.planes, pair.planes)                  # fused == unfused, both planes
    ref = orc.prelu(orc.conv3d(grid.cpu().numpy(), w.cpu().numpy(), b.cpu().numpy(), (2, 2, 2)), al.cpu().numpy())
    e, s = _err(pair.float(), ref)
    ef, _ = _err(fast.float(), ref)
    print(f"e_conv1 exact err {e:.2e}, fast err {ef:.2e}, scale {s:.2e}")
    assert e <= 2e-6 * s and ef > 20 * e


def test_exact_and_fast_engines_coexist_and_store_is_strict():
    """Engines own their variable stores: an exact and a fast engine (different weights) interleave without disturbing each
    other (process-wide store), eager == graph in both; a weight dict with a missing variable raises instead of
    silently using a random initialiser."""
    from rendernet_b200.engine import RenderEngine
    rng = np.random.default_rng(3)
    vox = (rng.random((1, 64, 64, 64, 1)) < 0.1).astype(np.float32)
    pose = np.array([[4.36, 0.52, 1.0]], np.float32)
    a = RenderEngine(None, 1, seed=0, precision="exact")
    first = a.render(vox, pose).clone()
    b = RenderEngine(None, 1, seed=1, precision="fast", use_graph=False)      # eager: re-runs RenderNet() on ITS store
    other = b.render(vox, pose).clone()
    c = RenderEngine(None, 1, seed=0, precision="exact", use_graph=False)
    assert torch.equal(a.render(vox, pose), first) and torch.equal(c.render(vox, pose), first)
    assert torch.equal(b.render(vox, pose), other) and not torch.equal(first, other)
    assert a.launches_per_step == c.launches_per_step
    W = orc.init_shader_weights(seed=0)
    del W["encoder/res2_4/alpha"]
    with pytest.raises(KeyError, match="res2_4/alpha"):
        RenderEngine(W, 1)


# ----------------------------------------------------------------------------------------- whole network, stress weights
def _chair(golden_dir):
    bv = np.load(os.path.join(golden_dir, "binvox.npz"))
    return np.unpackbits(bv["chair_bits"]).reshape(1, 64, 64, 64, 1).astype(np.float32)


def _stage_report(stages, st, keys):
    for k in keys:
        a = stages[k].float().cpu().numpy() if not isinstance(stages[k], np.ndarray) else stages[k]
        b = st[k].float().numpy()
        err = np.abs(a - b)
        print(f"    stage {k:10s} max_abs_err={err.max():.3e} rel_rms_err={np.sqrt((err ** 2).mean()) / max(np.sqrt((b ** 2).mean()), 1e-30):.3e} "
              f"ref_absmax={np.abs(b).max():.3e}")


def test_full_size_stress_weights_exact_meets_bar_fast_at_its_bound(golden_dir):
    """THE precision gate: full-size Shader network (chair, demo pose, 64^3 -> 512^2) with gain-1.1 weights --
    logits span +-11, the image spans [0,1], what a trained network produces.  Exact mode must meet the north_star's 1e-3
    max-abs bar against the fp32 oracle; the fast mode (fp16 operands) is asserted at its measured bound (8e-3 in r01;
    asserted <= 2e-2) and reported."""
    from rendernet_b200 import ops, tfcompat as tf
    from rendernet_b200.RenderNet_Shader import RenderNet
    from rendernet_b200.engine import pose_to_matrix
    chair = _chair(golden_dir)
    pose = orc.compute_pose_param(250.0, 60.0, 3.3)
    W = orc.init_shader_weights(seed=1, alpha_range=(0.05, 0.3), gain=1.1, bias_jitter=0.02)
    ref_img, st = orc.render_forward(chair, pose, W, return_stages=True)
    ref = ref_img.numpy()
    assert ref.min() < 0.02 and ref.max() > 0.98, "stress weights must saturate the sigmoid both ways"
    minv = torch.from_numpy(pose_to_matrix(pose)).to(dev)
    errs = {}
    for prec in ("exact", "fast"):
        store = tf.VariableStore(precision=prec)
        with tf.use_store(store):
            tf.load_weight_dict(W)
            grid = ops.resample(torch.from_numpy(chair).to(dev), minv, 128, True)
            stages = {}
            img = RenderNet(grid, is_training=False, stages=stages)
            torch.cuda.synchronize()
        stages["logits"] = torch.log(img.double() / (1 - img.double())).float()
        e = np.abs(img.cpu().numpy() - ref)
        errs[prec] = float(e.max())
        print(f"  [{prec}] IMAGE max_abs_err={e.max():.3e} mean_abs_err={e.mean():.3e} (bar 1e-3); image range [{ref.min():.3f},{ref.max():.3f}]")
        _stage_report(stages, st, ("enc3", "enc3_skip", "enc4", "enc4_skip", "enc5_skip", "enc10", "logits"))
        del store, stages, img, grid
        torch.cuda.empty_cache()
    assert errs["exact"] <= 1e-3, errs
    assert errs["fast"] <= 2e-2, errs


# ----------------------------------------------------------------------------------------- BASELINE configs, full size
def _synthetic_batch(B):
    import bench
    return bench.synthetic_batch(B)


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_config2_random_batch_full_size_vs_oracle(precision):
    """BASELINE config 2 (batch of random 10 %-occupancy 64^3 voxels, random poses, Shader net, reference initialisers): the
    first 2 items of the bench's synthetic batch through RenderEngine vs orc.render_forward; bar 1e-3 max-abs, both modes."""
    from rendernet_b200.engine import RenderEngine
    vox, poses = _synthetic_batch(24)
    vox, poses = vox[:2], poses[:2]
    W = orc.init_shader_weights(seed=0, alpha_range=(0.05, 0.3))
    ref = orc.render_forward(vox, poses, W).numpy()
    eng = RenderEngine(W, 2, precision=precision)
    img = eng.render(vox, poses).numpy()
    e = np.abs(img - ref)
    print(f"config 2 [{precision}]: image max_abs_err={e.max():.3e} mean={e.mean():.3e}; image range [{ref.min():.3f},{ref.max():.3f}]")
    assert img.shape == (2, 512, 512, 3) and e.max() <= 1e-3


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_config4_texture_full_size_vs_oracle(golden_dir, precision):
    """BASELINE config 4: texture+normal face render, full size, B=1, through TextureRenderEngine (texture decoder, two
    resamplings, concat, Texture/Normal RenderNet) vs orc.render_forward_texture -- both outputs, bar 1e-3."""
    from rendernet_b200.engine import TextureRenderEngine
    rng = np.random.default_rng(2)
    vox = _chair(golden_dir)
    tex = rng.standard_normal((1, 199)).astype(np.float32)
    pose = orc.compute_pose_param(250.0, 60.0, 3.3).astype(np.float32)
    W = orc.init_texture_weights(seed=3, alpha_range=(0.05, 0.3), bias_jitter=0.02)
    ref_img, ref_nrm = orc.render_forward_texture(vox, tex, pose, W)
    eng = TextureRenderEngine(W, 1, precision=precision)
    img, nrm = eng.render(vox, tex, pose)
    e1 = float(np.abs(img.numpy() - ref_img.numpy()).max())
    e2 = float(np.abs(nrm.numpy() - ref_nrm.numpy()).max())
    print(f"config 4 [{precision}]: albedo max_abs_err={e1:.3e} normal max_abs_err={e2:.3e}")
    # exact: the north_star bar.  fast (fp16 operands): asserted at its measured bound on these weights (1.85e-3, r02) --
    # it does NOT meet 1e-3 here, which is why "exact" is the default precision of the engines and of bench.py
    bar = 1e-3 if precision == "exact" else 5e-3
    assert tuple(img.shape) == (1, 512, 512, 3) and e1 <= bar and e2 <= bar
    # pipelined API returns the same images
    t0 = eng.submit(vox, tex, pose)
    got = eng.result(t0)
    assert torch.equal(got[0], img) and torch.equal(got[1], nrm)


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_config5_turntable_frames_full_size_vs_oracle(golden_dir, precision):
    """BASELINE config 5: bunny turntable frames at azimuth 0 / 90 / 133 / 270 degrees (el 60, r 3.3) -- the axis-aligned
    ones sit exactly on the resampler's clamp discontinuity -- vs the oracle, full size; bar 1e-3."""
    from rendernet_b200.engine import RenderEngine
    bv = np.load(os.path.join(golden_dir, "binvox.npz"))
    bunny = np.unpackbits(bv["bunny_bits"]).reshape(1, 64, 64, 64, 1).astype(np.float32)
    assert int(bunny.sum()) == int(bv["bunny_count"])
    az = [0.0, 90.0, 133.0, 270.0]
    poses = np.concatenate([orc.compute_pose_param(a, 60.0, 3.3) for a in az]).astype(np.float32)
    vox = np.repeat(bunny, len(az), axis=0)
    W = orc.init_shader_weights(seed=0, alpha_range=(0.05, 0.3))
    ref = orc.render_forward(vox, poses, W).numpy()
    eng = RenderEngine(W, len(az), precision=precision)
    img = eng.render(vox, poses).numpy()
    for i, a in enumerate(az):
        print(f"config 5 [{precision}] az={a:5.1f}: max_abs_err={np.abs(img[i] - ref[i]).max():.3e}")
    # The resampler computes its sample coordinates in the oracle's fp32 arithmetic order (rn_ops.cu sample_coord), so even
    # the axis-aligned frames, whose sample points sit ON the clamp discontinuity, pick the same voxels as the oracle.
    assert np.abs(img - ref).max() <= (1e-3 if precision == "exact" else 2e-3)


def test_exact_trunk_conv_k9216_accumulation_error():
    """The 3x3 1024->1024 trunk convolution (K = 9216, the longest accumulation chain of the network) in exact mode vs float64.
    The tensor core truncates its fp32 accumulator once per MMA step, so the error grows with the number of steps taken while
    the accumulator is large; the exact mode therefore sums all 2^-11-sized correction products first (rn_igemm.cu,
    plan_conv).  Measured (r02): max 1.3e-5 of the output scale, relative rms 1.05e-5, zero mean bias (the truncation pulls every
    output toward zero by ~1e-5 of its magnitude); asserted at 2.5e-5.  The fast mode is ~25x above."""
    ops = _ops()
    rng = np.random.default_rng(42)
    B, H, W, C = 1, 16, 16, 1024
    x = (rng.standard_normal((B, H, W, C)) * 3).astype(np.float32)
    w = (rng.uniform(-1, 1, (3, 3, C, C)) * np.sqrt(6.0 / (9 * 2 * C))).astype(np.float32)
    b = np.zeros(C, np.float32)
    ref = _conv2d_f64(x, w, b, (1, 1, 1, 1))
    res = {}
    for name, fmt in (("exact", 2), ("fast", 0)):
        L = ops.pack_conv("conv2d", torch.from_numpy(w), torch.from_numpy(b), None, device=dev, fmt=fmt)
        y = ops.conv2d(ops.cast_to_16(torch.from_numpy(x).to(dev), fmt=fmt), L, want16=False, want32=True)
        d = (y.double().cpu() - ref)
        res[name] = (float(d.abs().max()), float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()), float(d.mean() / ref.abs().mean()))
    s = float(ref.abs().max())
    print(f"K=9216 trunk conv: exact max {res['exact'][0]:.2e} rel-rms {res['exact'][1]:.2e} mean-bias {res['exact'][2]:+.2e}; "
          f"fast max {res['fast'][0]:.2e} rel-rms {res['fast'][1]:.2e} (scale {s:.2e})")
    assert res["exact"][0] <= 2.5e-5 * s and res["exact"][1] <= 2e-5
    assert res["fast"][0] > 15 * res["exact"][0]
