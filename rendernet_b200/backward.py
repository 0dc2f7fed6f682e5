"""Backward pass of the forward rendering path with respect to its INPUTS (first stage).

This is what the reference's inverse rendering differentiates (Reconstruct_RenderNet_Face.py:383-412: `tf.gradients` of an
image loss w.r.t. the latent shape / texture / pose, through the frozen RenderNet and through the trilinear weights of
tools/resampling_voxel_grid.py:465-485).  With `want_weight_grads` the same walk also produces dL/d(every variable) -- filters,
biases, PReLU slopes -- which rendernet_b200/training.py turns into the training step of RenderNet_Shader.py:154-167.

How it works: the model function (RenderNet_Shader.RenderNet) is run once with a TAPE attached to the variable store; every
realised layer appends (kind, input, filter, activation, residual, output).  `backward()` walks the tape in reverse:

  * activations: PReLU / sigmoid derivatives from the STORED post-activation tensors (rn_prelu_backward_16, rn_sigmoid_backward);
  * data gradient of a stride-1 SAME convolution = a stride-1 convolution of the output gradient with the spatially mirrored,
    channel-transposed filter -> the SAME tcgen05 implicit-GEMM kernel as the forward pass (rn_conv_igemm) with a mirrored
    tap list; 3^3 convs through the depth-folded (banded) form; the residual adds of the forward graph become the fused
    `residual` input of the gradient convolution (gradient accumulation at a fan-out costs no extra pass);
  * data gradient of a stride-1 transposed conv = a forward SAME conv with the very same filter array;
  * data gradient of a stride-2 transposed conv (k = 4) = space-to-depth of the output gradient ([B,2H,2W,C] -> [B,H,W,4C])
    followed by ONE 3x3 convolution whose filter holds the 16 taps at their (phase, offset) slots (the transpose of the
    forward "merged-phase" trick);
  * e_conv2 / e_conv1 (thin, strided, 1.7 % of the MACs): CUDA-core gather kernels (rn_conv3d_backward_data_direct);
  * resampler: scatter-add to the voxel grid and the 3x4 matrix gradient (rn_resample_backward_f32); the 12 matrix entries
    are mapped to (azimuth, elevation, scale) on the host by differentiating the reference's matrix construction
    (tools/resampling_voxel_grid.py:515-602) -- a 3 -> 12 map, float64.

Gradients travel in the activations' 16-bit format (fp16, or fp16 hi/lo pairs in the exact mode) with a loss scale to keep
them inside fp16's range; they are un-scaled when they leave the tensor-core part (fp32 from there on).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from . import tfcompat as tf
from .RenderNet_Shader import RenderNet
from .engine import pose_to_matrix
from .resampling_voxel_grid import ResampledGrid


def _key(t) -> int:
    """Identity of an activation on the tape: its storage address (reshapes / views share it)."""
    return t.data_ptr()


def pose_matrix_jacobian_vjp(view_params: np.ndarray, dminv: np.ndarray, size: int = 64, new_size: int = 128) -> np.ndarray:
    """dL/d(view_params) [B,3] from dL/d(Minv[:, :3, :]) [B,3,4]: differentiates M = T(+new/2).S.R.T(-size/2), R = RotZ(el).RotY(az-pi/2)
    (tools/resampling_voxel_grid.py:515-602) and the matrix inverse, in float64 on the host (a 3 -> 12 map per item)."""
    vp = torch.tensor(np.asarray(view_params, np.float64), requires_grad=True)
    B = vp.shape[0]
    az = vp[:, 0] - math.pi * 0.5
    el, sc = vp[:, 1], vp[:, 2]
    ca, sa, ce, se = torch.cos(az), torch.sin(az), torch.cos(el), torch.sin(el)
    z, o = torch.zeros_like(ca), torch.ones_like(ca)
    rot_y = torch.stack([torch.stack([ca, z, -sa, z], 1), torch.stack([z, o, z, z], 1), torch.stack([sa, z, ca, z], 1),
                         torch.stack([z, z, z, o], 1)], 1)
    rot_z = torch.stack([torch.stack([ce, se, z, z], 1), torch.stack([-se, ce, z, z], 1), torch.stack([z, z, o, z], 1),
                         torch.stack([z, z, z, o], 1)], 1)
    R = rot_z @ rot_y
    S = torch.stack([torch.stack([sc, z, z, z], 1), torch.stack([z, sc, z, z], 1), torch.stack([z, z, sc, z], 1),
                     torch.stack([z, z, z, o], 1)], 1)
    T = torch.eye(4, dtype=torch.float64).repeat(B, 1, 1).clone()
    T[:, :3, 3] = -size * 0.5
    Tn = torch.eye(4, dtype=torch.float64).repeat(B, 1, 1).clone()
    Tn[:, :3, 3] = new_size * 0.5
    minv = torch.linalg.inv(Tn @ S @ R @ T)[:, :3, :]
    (minv * torch.as_tensor(np.asarray(dminv, np.float64))).sum().backward()
    return vp.grad.numpy()


class ShaderInputGradients:
    """Forward + input-gradient backward of the Shader network for a fixed batch size.

        ig = ShaderInputGradients(weights, batch=1, precision="exact")
        img = ig.forward(voxels, view_params)                 # [B,512,512,3] fp32 on the device (tape recorded)
        dvox, dpose = ig.backward(dL_dimg)                    # [B,64,64,64,1], [B,3] fp32 (NumPy)

    `weights`: {tf variable name: array} or None (seeded reference initialisers)."""

    def __init__(self, weights: Optional[Dict[str, np.ndarray]], batch: int, precision: str = "exact", is_greyscale: bool = False,
                 size: int = 64, new_size: int = 128, loss_scale: float = 4096.0, seed: int = 0, device: str = "cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("ShaderInputGradients needs a CUDA device (no CPU fallback)")
        self.B, self.size, self.new_size = batch, size, new_size
        self.is_greyscale = is_greyscale
        self.loss_scale = float(loss_scale)
        self.device = torch.device(device)
        self.store = tf.VariableStore(precision=precision, device=str(self.device) if self.device.index is not None else "cuda",
                                      seed=seed)
        if weights is not None:
            with tf.use_store(self.store):
                tf.load_weight_dict(weights)
        self._dgrad_cache: Dict[object, object] = {}
        self.tape = None
        self.img = None

    # ------------------------------------------------------------------------------------------- forward
    def forward(self, voxels, view_params) -> torch.Tensor:
        dev = self.store.device
        self.view_params = np.asarray(view_params, np.float32)
        self.vox = torch.as_tensor(np.asarray(voxels, np.float32)).reshape(self.B, self.size, self.size, self.size, 1).to(dev)
        self.minv = torch.from_numpy(pose_to_matrix(self.view_params, self.size, self.new_size)).to(dev)
        self.tape = []
        self.store.tape = self.tape
        try:
            with tf.use_store(self.store):
                grid = ResampledGrid(self.vox, self.minv, self.new_size, transform=True)
                self.img = RenderNet(grid, is_training=False, is_greyscale=self.is_greyscale)
        finally:
            self.store.tape = None
        return self.img

    # ------------------------------------------------------------------------------------------- packed gradient filters
    def _zeros(self, n):
        z = self._dgrad_cache.get(("zeros", n))
        if z is None:
            z = torch.zeros(n, device=self.store.device, dtype=torch.float32)
            self._dgrad_cache[("zeros", n)] = z
        return z

    def _dgrad_layer(self, rec):
        """Kernel-ready filter of the data-gradient convolution of one recorded layer (cached per weight)."""
        w, kind, stride, fmt = rec["w"], rec["kind"], rec["stride"], self.store.fmt
        key = (w._rn_name, kind, stride, fmt)
        L = self._dgrad_cache.get(key)
        if L is not None:
            return L
        dev = self.store.device
        wt = w.to(dev)
        if kind == "conv2d":                       # [kh,kw,Ci,Co] -> mirrored, channel roles swapped: [kh,kw,Co,Ci]
            k = int(w.shape[0])
            wd = torch.flip(wt, dims=(0, 1)).permute(0, 1, 3, 2).contiguous()
            L = ops.pack_conv("conv2d", wd, None, None, device=dev, fmt=fmt)
            pb = (k - 1) // 2
            L.taps = [(kx - (k - 1 - pb), ky - (k - 1 - pb)) for ky in range(k) for kx in range(k)]   # offset of mirrored tap
        elif kind == "conv3d":                     # 3^3, stride 1: depth-folded gradient conv
            wd = torch.flip(wt, dims=(0, 1, 2)).permute(0, 1, 2, 4, 3).contiguous()
            L = ops.BandedConv3d(wd, None, device=dev, sz=1, fmt=fmt)
        elif kind == "conv2d_transpose" and stride == 1:
            # y[o] = sum_k x[o - k + pb] w[k][co][ci]  =>  dx[i] = sum_k g[i + k - pb] w[k][co][ci]: a forward SAME conv whose TF filter
            # [kh,kw,Cin'=Co,Cout'=Ci] IS the transposed-conv filter array; Co is zero padded to a multiple of 16 (e_conv11: 3)
            co = int(w.shape[2])
            cp = ops.round_up(co, 16)
            wd = torch.zeros((w.shape[0], w.shape[1], cp, w.shape[3]), device=dev, dtype=torch.float32)
            wd[:, :, :co] = wt
            L = ops.pack_conv("conv2d", wd, None, None, device=dev, fmt=fmt)
            L.taps = None
        elif kind == "conv2d_transpose" and stride == 2:
            # o = 2i + k - 1.  With g2[i,j,(ay,ax,co)] = g[2i+ay, 2j+ax, co]:  dx[i] = sum_{dy,dx in -1..1} g2[i+dy, j+dx] . Wd[dy+1][dx+1]
            # where Wd[dy+1][dx+1][(ay,ax,co)][ci] = w[2dy+ay+1][2dx+ax+1][co][ci] when both indices lie in [0,4), else 0
            assert tuple(w.shape[:2]) == (4, 4)
            co, ci = int(w.shape[2]), int(w.shape[3])
            wd = torch.zeros((3, 3, 2, 2, co, ci), device=dev, dtype=torch.float32)
            for dy in (-1, 0, 1):
                for ay in (0, 1):
                    ky = 2 * dy + ay + 1
                    if not 0 <= ky < 4:
                        continue
                    for dx in (-1, 0, 1):
                        for ax in (0, 1):
                            kx = 2 * dx + ax + 1
                            if 0 <= kx < 4:
                                wd[dy + 1, dx + 1, ay, ax] = wt[ky, kx]
            L = ops.pack_conv("conv2d", wd.reshape(3, 3, 4 * co, ci), None, None, device=dev, fmt=fmt)
            L.taps = None
        else:
            raise NotImplementedError(f"no data-gradient path for {kind} stride {stride}")
        self._dgrad_cache[key] = L
        return L

    @staticmethod
    def _space_to_depth(g):
        """[B,2H,2W,C] -> [B,H,W,(ay,ax,C)] (pure data movement; hi/lo planes alike)."""
        t = g.planes if isinstance(g, ops.Split16) else g
        lead = t.shape[:-3]
        H2, W2, Cc = t.shape[-3:]
        t = t.reshape(*lead, H2 // 2, 2, W2 // 2, 2, Cc).permute(*range(len(lead)), len(lead), len(lead) + 2, len(lead) + 1,
                                                                  len(lead) + 3, len(lead) + 4)
        t = t.reshape(*lead, H2 // 2, W2 // 2, 4 * Cc).contiguous()
        return ops.Split16(t) if isinstance(g, ops.Split16) else t

    def _has_negative_slope(self, alpha, a_dev) -> bool:
        key = ("alpha<0", alpha._rn_name)
        v = self._dgrad_cache.get(key)
        if v is None:
            v = bool((a_dev < 0).any().item())
            self._dgrad_cache[key] = v
        return v

    def _alpha(self, alpha, n):
        if isinstance(alpha, str):                 # tf.nn.relu branch
            return self._zeros(n)
        key = ("alpha", alpha._rn_name)
        a = self._dgrad_cache.get(key)
        if a is None:
            a = alpha.to(device=self.store.device, dtype=torch.float32).reshape(-1).contiguous()
            self._dgrad_cache[key] = a
        return a

    # ------------------------------------------------------------------------------------------- backward
    def backward(self, dimg, want_dvox: bool = True, want_dpose: bool = True, want_weight_grads: bool = False,
                 tensor_core_wgrad: bool = True):
        """dimg: dL/dimg [B,512,512,3|1] (NumPy or tensor).  Returns (dL/dvoxels [B,S,S,S,1] or None, dL/dview_params [B,3] or None).
        want_weight_grads: also fills `self.weight_grads` {variable name: fp32 device tensor in the variable's TF layout} for EVERY
        variable the forward pass used -- filters (tcgen05 weight-gradient kernel for the wide stride-1 2-D layers and, depth-folded,
        the 3^3 layers; the strided-correlation kernel rn_conv_weight_grad_direct for the thin / strided / transposed ones), biases
        and PReLU slopes (pre-activation recomputed: alpha starts at 0, tools/layer_util.py:38).  tensor_core_wgrad=False sends
        every filter through the direct kernel (cross-check)."""
        if self.tape is None:
            raise RuntimeError("call forward() first")
        dev = self.store.device
        fmt = self.store.fmt
        dimg = torch.as_tensor(np.asarray(dimg, np.float32) if not isinstance(dimg, torch.Tensor) else dimg).to(dev).float()
        if tuple(dimg.shape) != tuple(self.img.shape):
            raise ValueError(f"dimg shape {tuple(dimg.shape)} != image shape {tuple(self.img.shape)}")
        grads = {_key(self.img): dimg.contiguous()}
        dgrid = None
        self.weight_grads = {}
        inv = 1.0 / self.loss_scale
        need_inputs = want_dvox or want_dpose
        with torch.cuda.device(self.device), tf.use_store(self.store):
            for rec in reversed(self.tape):
                y = rec["y"]
                g = grads.pop(_key(y), None)
                if g is None:
                    continue                                   # no gradient reaches this layer
                if tuple(g.shape) != tuple(y.shape):           # the consumer saw a reshaped view (projection unit: [..,D,C] -> [..,D*C])
                    g = g.reshape(tuple(y.shape))
                if rec["op"] == "dropout":                     # d(x * mask / keep) = g * mask / keep: the same stateless kernel
                    grads[_key(rec["x"])] = ops.dropout(g, rec["keep"], rec["seed"], rec["salt"])
                    continue
                act = rec["act"]
                if act == "sigmoid":
                    co = int(y.shape[-1])
                    g = ops.sigmoid_backward(g, y, ops.round_up(co, 16), self.loss_scale, fmt)
                elif act == "prelu":
                    alpha = rec["alpha"]
                    a_dev = self._alpha(alpha, int(y.shape[-1]))
                    sign_src = y                    # the stored output tells the side of the kink as long as every slope is >= 0
                    if not isinstance(alpha, str) and (want_weight_grads or self._has_negative_slope(alpha, a_dev)):
                        # Re-run the layer without its PReLU to get the pre-activation z: dL/dalpha = sum_{z<0} g*z needs it (alpha
                        # starts at 0), and so does the derivative itself once a slope is negative (y = alpha*z > 0 for z < 0;
                        # Adam's first step already makes half of the slopes negative).
                        sign_src = rec["rerun"]()
                        if want_weight_grads:
                            self.weight_grads[alpha._rn_name] = ops.prelu_alpha_grad(g, sign_src, inv)
                    g = ops.prelu_backward(g, sign_src, a_dev)
                if want_weight_grads:
                    self._weight_grads_of(rec, g, inv, tensor_core_wgrad)
                if rec["op"] == "resample_conv1":              # e_conv1 (5^3 s2, 1 -> 8) fused with the resampler in the forward pass
                    if need_inputs:
                        N = self.new_size
                        w32 = rec["w"].to(dev).float().contiguous()
                        dgrid = ops.conv3d_backward_data_direct(g, w32, (self.B, N, N, N, 1), rec["stride"], want32=True,
                                                                out_scale=inv)
                    continue
                res = rec.get("residual")
                if res is not None:                            # y = act(conv(x) + res): the gradient flows to res unchanged
                    k = _key(res)
                    grads[k] = ops.bias_act(g, None, None, None, residual=grads[k]) if k in grads else g
                x, kind, stride = rec["x"], rec["kind"], rec["stride"]
                acc = grads.pop(_key(x), None)                 # gradient already collected for x (fan-out): fused as `residual`
                if kind == "conv3d" and stride == 2:           # e_conv2: thin, z-strided -> CUDA cores
                    w32 = rec["w"].to(dev).float().contiguous()
                    gx = ops.conv3d_backward_data_direct(g, w32, tuple(x.shape), (1, 1, 2))
                    if acc is not None:
                        gx = ops.bias_act(gx, None, None, None, residual=acc)
                else:
                    L = self._dgrad_layer(rec)
                    if kind == "conv3d":
                        gx = ops.conv3d_banded(g, L, residual=acc)
                    elif kind == "conv2d":
                        k = int(rec["w"].shape[0])
                        if k % 2 == 1:
                            gx = ops.conv2d(g, L, residual=acc)
                        else:
                            gx = ops.conv2d_taps(g, L.w, L.bias, L.taps, L.cout, L.cout_pad, fmt, residual=acc,
                                                 ny=k if L.cin % 64 == 0 else 0)
                    elif stride == 1:                          # transposed conv, stride 1
                        gx = ops.conv2d(g, L, residual=acc)
                    else:                                      # transposed conv, stride 2
                        gx = ops.conv2d(self._space_to_depth(g), L, residual=acc)
                grads[_key(x)] = gx
            dvox = dminv = None
            if need_inputs:
                if dgrid is None:
                    raise RuntimeError("the tape holds no fused resample + e_conv1 record (is this the Shader path?)")
                dvox, dminv = ops.resample_backward(self.vox, self.minv, dgrid, True, want_dvox, want_dpose)
            torch.cuda.synchronize()
        self.last_dgrid = dgrid
        dpose = None
        if want_dpose:
            dpose = pose_matrix_jacobian_vjp(self.view_params, dminv.cpu().numpy(), self.size, self.new_size)
        return (dvox.cpu().numpy() if dvox is not None else None), dpose

    # ------------------------------------------------------------------------------------------- weight gradients
    def _weight_grads_of(self, rec, g, inv: float, tensor_core: bool):
        """dL/dW and dL/db of one recorded layer from g = dL/d(pre-activation) (16-bit, carrying the loss scale)."""
        w, b = rec["w"], rec["b"]
        wg = self.weight_grads
        cout = int(w.shape[2]) if rec.get("kind") == "conv2d_transpose" else int(w.shape[-1])
        if b is not None:
            wg[b._rn_name] = ops.bias_grad(g)[:cout] * inv          # g may carry zero-padded channels (e_conv11: 3 of 16)
        if rec["op"] == "resample_conv1":
            # e_conv1 read the resampled grid straight out of the fused kernel: materialise it once for the correlation
            q = ops.resample(rec["grid"].voxel, rec["grid"].minv, self.new_size, True)
            st = tuple(int(v) for v in rec["stride"])
            ks = tuple(int(v) for v in w.shape[:3])
            pad = tuple(ops.same_pad_before(self.new_size, ks[i], st[i]) for i in range(3))
            d = ops.conv_weight_grad_direct(g, q, ks, st, pad, Ca=cout, Cb=int(w.shape[3]), scale=inv)      # [k,k,k,co,ci]
            wg[w._rn_name] = d.permute(0, 1, 2, 4, 3).contiguous()
            return
        x, kind, stride = rec["x"], rec["kind"], rec["stride"]
        if kind == "conv2d":
            kh, kw, ci, co = (int(v) for v in w.shape)
            if tensor_core and ci % 128 == 0 and co % 128 == 0 and kh * kw <= 16:
                wg[w._rn_name] = ops.conv2d_weight_grad(x, g, kh, kw) * inv
            else:
                pad = (ops.same_pad_before(int(x.shape[1]), kh, 1), ops.same_pad_before(int(x.shape[2]), kw, 1))
                d = ops.conv_weight_grad_direct(g, x, (kh, kw), (1, 1), pad, Ca=co, Cb=ci, scale=inv)           # [kh,kw,co,ci]
                wg[w._rn_name] = d.permute(0, 1, 3, 2).contiguous()
        elif kind == "conv3d":
            k1, k2, k3, ci, co = (int(v) for v in w.shape)
            st = (1, 1, int(stride))                                   # stride 2 = the z-strided e_conv2 ([1,1,2])
            D = int(x.shape[3])
            if (tensor_core and st == (1, 1, 1) and (k1, k2, k3) == (3, 3, 3) and (D * ci) % 128 == 0 and (D * co) % 128 == 0):
                # depth-folded: a 3x3 conv2d over [B,H,W,D*ci] -> [B,H,W,D*co] whose filter is block-banded; its tensor-core
                # weight gradient holds dW[k1][k2][k3] on the (k3-1)-th block diagonal, summed over the D depth slices
                B_, H, Wd = (int(v) for v in x.shape[:3])
                full = ops.conv2d_weight_grad(x.reshape(B_, H, Wd, D * ci), g.reshape(B_, H, Wd, D * co), 3, 3)
                full = full.view(3, 3, D, ci, D, co)
                d = torch.stack([torch.diagonal(full, offset=-(k - 1), dim1=2, dim2=4).sum(-1) for k in range(3)], dim=2)
                wg[w._rn_name] = (d * inv).contiguous()                # [3,3,3,ci,co]
            else:
                pad = tuple(ops.same_pad_before(int(x.shape[1 + i]), (k1, k2, k3)[i], st[i]) for i in range(3))
                d = ops.conv_weight_grad_direct(g, x, (k1, k2, k3), st, pad, Ca=co, Cb=ci, scale=inv)           # [k,k,k,co,ci]
                wg[w._rn_name] = d.permute(0, 1, 2, 4, 3).contiguous()
        elif kind == "conv2d_transpose":
            kh, kw, co, ci = (int(v) for v in w.shape)                 # TF transposed-conv filter [kh,kw,Cout,Cin]
            s = int(stride)
            pad = (ops.same_pad_before(int(x.shape[1]) * s, kh, s), ops.same_pad_before(int(x.shape[2]) * s, kw, s))
            d = ops.conv_weight_grad_direct(x, g, (kh, kw), (s, s), pad, Ca=ci, Cb=co, scale=inv)               # [kh,kw,ci,co]
            wg[w._rn_name] = d.permute(0, 1, 3, 2).contiguous()
        else:
            raise NotImplementedError(f"no weight-gradient path for {kind}")
