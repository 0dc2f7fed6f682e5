"""Torch-facing wrappers of the C-ABI kernels.  PyTorch is plumbing here (device memory, streams);
all arithmetic happens in librendernet_b200.so.  Everything requires CUDA tensors; there is no
CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from ._lib import check, lib, rn_conv_desc

ACT_NONE, ACT_PRELU, ACT_SIGMOID = 0, 1, 2
_ACT = {None: ACT_NONE, "none": ACT_NONE, "prelu": ACT_PRELU, "sigmoid": ACT_SIGMOID}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _cuda(t: torch.Tensor, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) and hasattr(t, "realize"):
        t = t.realize()            # deferred conv output / deferred resampled grid
    if not t.is_cuda:
        raise RuntimeError("rendernet_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def fmt_of(dtype: torch.dtype) -> int:
    if dtype == torch.float16:
        return 0
    if dtype == torch.bfloat16:
        return 1
    raise TypeError(f"16-bit dtype expected, got {dtype}")


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


# --------------------------------------------------------------------------------------- resampler
def resample(vox: torch.Tensor, minv: torch.Tensor, new_size: int, transform: bool) -> torch.Tensor:
    """vox [B,S,S,S,C] fp32, minv [B,3,4] fp32 -> [B,N,N,N,C] fp32 (rn_resample_f32)."""
    vox = _cuda(vox, torch.float32)
    minv = _cuda(minv, torch.float32)
    B, S, _, _, Cc = vox.shape
    out = torch.empty((B, new_size, new_size, new_size, Cc), device=vox.device, dtype=torch.float32)
    check(lib.rn_resample_f32(vox.data_ptr(), minv.data_ptr(), out.data_ptr(), B, Cc, S, new_size,
                              1 if transform else 0, _stream()), "rn_resample_f32")
    return out


# --------------------------------------------------------------------------------------- packed layers
@dataclass
class PackedConv:
    """Device-resident, kernel-ready parameters of one convolution layer."""
    kind: str                      # conv2d | conv3d | conv2d_transpose
    ksize: Sequence[int]
    stride: int
    cin: int
    cout: int
    cout_pad: int
    w: torch.Tensor                # 16-bit packed filter
    bias: torch.Tensor             # fp32 [cout_pad]
    alpha: Optional[torch.Tensor]  # fp32 [cout_pad]
    dtype: torch.dtype


def _pad_vec(v: Optional[torch.Tensor], n: int, n_pad: int, device) -> Optional[torch.Tensor]:
    if v is None:
        return None
    out = torch.zeros(n_pad, device=device, dtype=torch.float32)
    out[:n] = v.to(device=device, dtype=torch.float32).reshape(-1)
    return out


def pack_conv(kind: str, w_tf: torch.Tensor, bias: Optional[torch.Tensor], alpha: Optional[torch.Tensor],
              stride: int = 1, dtype: torch.dtype = torch.float16, device="cuda") -> PackedConv:
    """w_tf in TF filter layout: conv2d [kh,kw,Cin,Cout]; conv3d [k,k,k,Cin,Cout];
    conv2d_transpose [kh,kw,Cout,Cin]."""
    w_tf = torch.as_tensor(w_tf, dtype=torch.float32).to(device).contiguous()
    fmt = fmt_of(dtype)
    if kind == "conv2d_transpose":
        kh, kw, cout, cin = w_tf.shape
        ks = (kh, kw)
    elif kind == "conv2d":
        kh, kw, cin, cout = w_tf.shape
        ks = (kh, kw)
    elif kind == "conv3d":
        k0, k1, k2, cin, cout = w_tf.shape
        ks = (k0, k1, k2)
    else:
        raise ValueError(kind)
    cout_pad = round_up(cout, 16)
    ntaps = 1
    for k in ks:
        ntaps *= k
    packed = torch.empty((ntaps, cout_pad, cin), device=device, dtype=dtype)
    if kind == "conv2d_transpose":
        check(lib.rn_pack_conv2d_transpose_weights(w_tf.data_ptr(), packed.data_ptr(), ks[0], ks[1], cin, cout,
                                                   cout_pad, stride, fmt, _stream()), "pack transpose")
    else:
        if stride != 1:
            raise ValueError("tensor-core conv path is stride 1")
        check(lib.rn_pack_conv_weights(w_tf.data_ptr(), packed.data_ptr(), ntaps, cin, cout, cout_pad, 0, None, 0,
                                       fmt, _stream()), "pack")
    b = _pad_vec(bias if bias is not None else torch.zeros(cout), cout, cout_pad, device)
    a = _pad_vec(alpha, cout, cout_pad, device)
    return PackedConv(kind, ks, stride, cin, cout, cout_pad, packed, b, a, dtype)


def _out_buffers(shape, dtype, device, want16, want32, out16, out32):
    if want16 and out16 is None:
        out16 = torch.empty(shape, device=device, dtype=dtype)
    if want32 and out32 is None:
        out32 = torch.empty(shape, device=device, dtype=torch.float32)
    return out16, out32


def conv2d(x: torch.Tensor, L: PackedConv, act: Optional[str] = None, residual: Optional[torch.Tensor] = None,
           want16: bool = True, want32: bool = False, out16=None, out32=None, alpha=None):
    """SAME stride-1 conv2d + bias (+PReLU/sigmoid) (+residual).  x [B,H,W,Cin] 16-bit."""
    x = _cuda(x, L.dtype)
    B, H, W, Cin = x.shape
    assert Cin == L.cin and L.kind == "conv2d"
    out16, out32 = _out_buffers((B, H, W, L.cout), L.dtype, x.device, want16, want32, out16, out32)
    res_f32 = 0
    if residual is not None:
        residual = _cuda(residual)
        res_f32 = 1 if residual.dtype == torch.float32 else 0
        assert tuple(residual.shape) == (B, H, W, L.cout)
    a = _ACT[act]
    check(lib.rn_conv2d_same(x.data_ptr(), L.w.data_ptr(), L.bias.data_ptr(),
                             _ptr(alpha if alpha is not None else L.alpha) if a == ACT_PRELU else None, a, _ptr(residual), res_f32,
                             _ptr(out16), _ptr(out32), B, H, W, Cin, L.cout, L.cout_pad, L.ksize[0], L.ksize[1],
                             fmt_of(L.dtype), _stream()), "rn_conv2d_same")
    return out16 if not want32 else ((out16, out32) if want16 else out32)


def conv3d(x: torch.Tensor, L: PackedConv, act: Optional[str] = None, residual: Optional[torch.Tensor] = None,
           want16: bool = True, want32: bool = False, out16=None, out32=None, alpha=None):
    """SAME stride-1 k^3 conv3d on the tensor pipe.  x [B,H,W,D,Cin] 16-bit."""
    x = _cuda(x, L.dtype)
    B, H, W, D, Cin = x.shape
    assert Cin == L.cin and L.kind == "conv3d"
    out16, out32 = _out_buffers((B, H, W, D, L.cout), L.dtype, x.device, want16, want32, out16, out32)
    res_f32 = 0
    if residual is not None:
        residual = _cuda(residual)
        res_f32 = 1 if residual.dtype == torch.float32 else 0
    a = _ACT[act]
    check(lib.rn_conv3d_same(x.data_ptr(), L.w.data_ptr(), L.bias.data_ptr(),
                             _ptr(alpha if alpha is not None else L.alpha) if a == ACT_PRELU else None, a, _ptr(residual), res_f32,
                             _ptr(out16), _ptr(out32), B, H, W, D, Cin, L.cout, L.cout_pad, L.ksize[0],
                             fmt_of(L.dtype), _stream()), "rn_conv3d_same")
    return out16 if not want32 else ((out16, out32) if want16 else out32)


class BandedConv3d:
    """Depth-folded 3^3 SAME conv3d (rn_conv3d_banded_same): kernel-ready banded filter + per-depth expanded
    bias / alpha vectors (cached per D)."""

    def __init__(self, w_tf: torch.Tensor, bias: Optional[torch.Tensor], dtype=torch.float16, device="cuda", sz: int = 1):
        w_tf = torch.as_tensor(w_tf, dtype=torch.float32).to(device).contiguous()
        assert tuple(w_tf.shape[:3]) == (3, 3, 3)
        self.cin, self.cout, self.sz = int(w_tf.shape[3]), int(w_tf.shape[4]), int(sz)
        nbytes = lib.rn_conv3d_banded_bytes(self.cin, self.cout, self.sz)
        if nbytes < 0:
            raise ValueError("banded conv3d needs Cin | 64 and Cout | 128")
        self.dtype = dtype
        self.w = torch.empty(nbytes // 2, device=device, dtype=dtype)
        check(lib.rn_pack_conv3d_banded(w_tf.data_ptr(), self.w.data_ptr(), self.cin, self.cout, self.sz, fmt_of(dtype),
                                        _stream()), "rn_pack_conv3d_banded")
        self.bias = (bias if bias is not None else torch.zeros(self.cout)).to(device=device, dtype=torch.float32)
        self._full = {}

    @staticmethod
    def eligible(cin: int, cout: int, D: int, sz: int = 1) -> bool:
        return (cin >= 8 and cout >= 8 and 64 % cin == 0 and 128 % cout == 0 and sz in (1, 2)
                and (-(-D // sz) * cout) % 128 == 0 and (D * cin) % 8 == 0)

    def expanded(self, v: torch.Tensor, D: int, tag) -> torch.Tensor:
        key = (tag, D)
        f = self._full.get(key)
        if f is None:
            v = _cuda(v.to(device=self.w.device, dtype=torch.float32))
            f = torch.empty(D * self.cout, device=self.w.device, dtype=torch.float32)
            check(lib.rn_expand_channels(v.data_ptr(), f.data_ptr(), self.cout, D, _stream()), "rn_expand_channels")
            self._full[key] = f
        return f


def conv3d_banded(x: torch.Tensor, L: BandedConv3d, act: Optional[str] = None,
                  residual: Optional[torch.Tensor] = None, alpha: Optional[torch.Tensor] = None, alpha_tag=None,
                  want16: bool = True, want32: bool = False, out16=None, out32=None):
    """x [B,H,W,D,Cin] 16-bit -> [B,H,W,D,Cout]; alpha is the per-Cout PReLU slope (length Cout)."""
    x = _cuda(x, L.dtype)
    B, H, W, D, Cin = x.shape
    assert Cin == L.cin
    Do = -(-D // L.sz)
    out16, out32 = _out_buffers((B, H, W, Do, L.cout), L.dtype, x.device, want16, want32, out16, out32)
    res_f32 = 0
    if residual is not None:
        residual = _cuda(residual)
        res_f32 = 1 if residual.dtype == torch.float32 else 0
    a = _ACT[act]
    bias_full = L.expanded(L.bias, Do, "bias")
    alpha_full = None
    if a == ACT_PRELU:
        alpha_full = L.expanded(alpha[:L.cout], Do, ("alpha", alpha_tag if alpha_tag is not None else alpha.data_ptr()))
    check(lib.rn_conv3d_banded_same(x.data_ptr(), L.w.data_ptr(), bias_full.data_ptr(), _ptr(alpha_full), a,
                                    _ptr(residual), res_f32, _ptr(out16), _ptr(out32), B, H, W, D, Cin, L.cout, L.sz,
                                    fmt_of(L.dtype), _stream()), "rn_conv3d_banded_same")
    return out16 if not want32 else ((out16, out32) if want16 else out32)


def conv2d_transpose(x: torch.Tensor, L: PackedConv, act: Optional[str] = None, want16: bool = True,
                     want32: bool = False, out16=None, out32=None, alpha=None):
    """SAME transposed conv, out = in*stride.  x [B,H,W,Cin] 16-bit."""
    x = _cuda(x, L.dtype)
    B, H, W, Cin = x.shape
    assert Cin == L.cin and L.kind == "conv2d_transpose"
    s = L.stride
    out16, out32 = _out_buffers((B, H * s, W * s, L.cout), L.dtype, x.device, want16, want32, out16, out32)
    a = _ACT[act]
    check(lib.rn_conv2d_transpose_same(x.data_ptr(), L.w.data_ptr(), L.bias.data_ptr(),
                                       _ptr(alpha if alpha is not None else L.alpha) if a == ACT_PRELU else None, a, _ptr(out16), _ptr(out32),
                                       B, H, W, Cin, L.cout, L.cout_pad, L.ksize[0], L.ksize[1], s,
                                       fmt_of(L.dtype), _stream()), "rn_conv2d_transpose_same")
    return out16 if not want32 else ((out16, out32) if want16 else out32)


class MergedConvT2:
    """k=4 stride-2 SAME transposed conv as one launch (rn_conv2d_transpose_s2_merged)."""

    def __init__(self, w_tf: torch.Tensor, bias: Optional[torch.Tensor], dtype=torch.float16, device="cuda"):
        w_tf = torch.as_tensor(w_tf, dtype=torch.float32).to(device).contiguous()
        assert tuple(w_tf.shape[:2]) == (4, 4)
        self.cout, self.cin, self.dtype = int(w_tf.shape[2]), int(w_tf.shape[3]), dtype
        self.w = torch.empty((9, 4 * self.cout, self.cin), device=device, dtype=dtype)
        check(lib.rn_pack_conv2d_transpose_s2_merged(w_tf.data_ptr(), self.w.data_ptr(), self.cin, self.cout, fmt_of(dtype),
                                                     _stream()), "pack merged tconv")
        b = bias if bias is not None else torch.zeros(self.cout)
        self.bias = b.to(device=device, dtype=torch.float32).reshape(-1).repeat(4).contiguous()
        self._alpha = {}

    @staticmethod
    def eligible(cin: int, cout: int) -> bool:
        return cin % 16 == 0 and cout % 16 == 0


def conv2d_transpose_s2_merged(x: torch.Tensor, L: MergedConvT2, act: Optional[str] = None,
                               alpha: Optional[torch.Tensor] = None, alpha_tag=None, want16: bool = True,
                               want32: bool = False, out16=None, out32=None):
    x = _cuda(x, L.dtype)
    B, H, W, Cin = x.shape
    assert Cin == L.cin
    out16, out32 = _out_buffers((B, 2 * H, 2 * W, L.cout), L.dtype, x.device, want16, want32, out16, out32)
    a = _ACT[act]
    alpha4 = None
    if a == ACT_PRELU:
        key = alpha_tag if alpha_tag is not None else alpha.data_ptr()
        alpha4 = L._alpha.get(key)
        if alpha4 is None:
            alpha4 = alpha.to(device=x.device, dtype=torch.float32).reshape(-1)[: L.cout].repeat(4).contiguous()
            L._alpha[key] = alpha4
    check(lib.rn_conv2d_transpose_s2_merged(x.data_ptr(), L.w.data_ptr(), L.bias.data_ptr(), _ptr(alpha4), a, _ptr(out16),
                                            _ptr(out32), B, H, W, Cin, L.cout, fmt_of(L.dtype), _stream()),
          "rn_conv2d_transpose_s2_merged")
    return out16 if not want32 else ((out16, out32) if want16 else out32)


class XFoldConvT:
    """Stride-1 transposed conv with thin channels, x-folded (rn_conv2d_transpose_s1_xfold)."""

    def __init__(self, w_tf: torch.Tensor, bias: Optional[torch.Tensor], F: int, dtype=torch.float16, device="cuda"):
        w_tf = torch.as_tensor(w_tf, dtype=torch.float32).to(device).contiguous()
        self.kh, self.kw, self.cout, self.cin = (int(v) for v in w_tf.shape)
        self.F, self.dtype = F, dtype
        self.cout_pad = round_up(F * self.cout, 16)
        self.w = torch.empty((self.kh * 3, self.cout_pad, F * self.cin), device=device, dtype=dtype)
        check(lib.rn_pack_conv2d_transpose_xfold(w_tf.data_ptr(), self.w.data_ptr(), self.kh, self.kw, self.cin, self.cout,
                                                 F, self.cout_pad, fmt_of(dtype), _stream()), "pack xfold")
        b = bias if bias is not None else torch.zeros(self.cout)
        self.bias = self.tiled(b, device)
        self._alpha = {}

    def tiled(self, v: torch.Tensor, device) -> torch.Tensor:
        out = torch.zeros(self.cout_pad, device=device, dtype=torch.float32)
        out[: self.F * self.cout] = v.to(device=device, dtype=torch.float32).reshape(-1)[: self.cout].repeat(self.F)
        return out

    @staticmethod
    def factor(cin: int, W: int) -> int:
        return int(lib.rn_xfold_factor(cin, W))


def conv2d_transpose_xfold(x: torch.Tensor, L: XFoldConvT, act: Optional[str] = None, alpha: Optional[torch.Tensor] = None,
                           alpha_tag=None, want16: bool = True, want32: bool = False, out16=None, out32=None):
    x = _cuda(x, L.dtype)
    B, H, W, Cin = x.shape
    assert Cin == L.cin and W % L.F == 0
    out16, out32 = _out_buffers((B, H, W, L.cout), L.dtype, x.device, want16, want32, out16, out32)
    a = _ACT[act]
    alpha_x = None
    if a == ACT_PRELU:
        key = alpha_tag if alpha_tag is not None else alpha.data_ptr()
        alpha_x = L._alpha.get(key)
        if alpha_x is None:
            alpha_x = L.tiled(alpha, x.device)
            L._alpha[key] = alpha_x
    check(lib.rn_conv2d_transpose_s1_xfold(x.data_ptr(), L.w.data_ptr(), L.bias.data_ptr(), _ptr(alpha_x), a, _ptr(out16),
                                           _ptr(out32), B, H, W, Cin, L.cout, L.kh, L.kw, L.F, L.cout_pad, fmt_of(L.dtype),
                                           _stream()), "rn_conv2d_transpose_s1_xfold")
    return out16 if not want32 else ((out16, out32) if want16 else out32)


def conv_igemm_raw(x, w_packed, bias, taps, ndim, B, H, W, D, Cin, Cout, cout_pad, out16=None, out32=None,
                   alpha=None, act=ACT_NONE, residual=None, o=None, fmt=0, force_bn=0, force_kps=0, max_ctas=0,
                   cluster=0, cta_group=0, ny=0, tile_w=0, msub=0):
    """Direct access to rn_conv_igemm for tests / tuning.  taps: list of (dx,dy,dz)."""
    n = len(taps)
    arr = (C.c_int8 * (3 * n))(*[v for t in taps for v in t])
    d = rn_conv_desc()
    d.ndim, d.B, d.H, d.W, d.D = ndim, B, H, W, D
    d.Cin, d.Cout, d.cout_pad, d.ntaps = Cin, Cout, cout_pad, n
    d.taps = C.cast(arr, C.c_void_p)
    d.x, d.w_packed, d.bias, d.alpha = x.data_ptr(), w_packed.data_ptr(), bias.data_ptr(), _ptr(alpha)
    d.act = act
    d.residual = _ptr(residual)
    d.residual_is_f32 = 1 if (residual is not None and residual.dtype == torch.float32) else 0
    d.out16, d.out32 = _ptr(out16), _ptr(out32)
    if o is None:
        if ndim == 2:
            o = (0, H * W * Cout, W * Cout, Cout, 0)
        else:
            o = (0, H * W * D * Cout, W * D * Cout, D * Cout, Cout)
    d.o_base, d.o_b, d.o_y, d.o_x, d.o_z = o
    d.fmt, d.force_bn, d.force_kps, d.max_ctas = fmt, force_bn, force_kps, max_ctas
    d.cluster = cluster
    d.cta_group = cta_group
    d.ny, d.tile_w, d.msub = ny, tile_w, msub
    check(lib.rn_conv_igemm(C.byref(d), _stream()), "rn_conv_igemm")


# --------------------------------------------------------------------------------------- thin conv3d
def conv3d_direct(x: torch.Tensor, w_tf: torch.Tensor, bias: torch.Tensor, alpha: Optional[torch.Tensor],
                  stride: Sequence[int], dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """CUDA-core SAME conv3d + bias + PReLU for the thin first layers.  x fp32 or 16-bit
    [B,H,W,D,Cin]; w_tf fp32 [k,k,k,Cin,Cout]; returns 16-bit."""
    x = _cuda(x)
    w_tf = _cuda(w_tf, torch.float32)
    bias = _cuda(bias, torch.float32)
    B, H, W, D, Cin = x.shape
    k = w_tf.shape[0]
    Cout = w_tf.shape[4]
    sy, sx, sz = stride
    Ho, Wo, Do = -(-H // sy), -(-W // sx), -(-D // sz)
    out = torch.empty((B, Ho, Wo, Do, Cout), device=x.device, dtype=dtype)
    check(lib.rn_conv3d_direct(x.data_ptr(), 1 if x.dtype == torch.float32 else 0, w_tf.data_ptr(),
                               bias.data_ptr(), _ptr(alpha), out.data_ptr(), B, H, W, D, Cin, Cout, k, sy, sx, sz,
                               fmt_of(dtype), _stream()), "rn_conv3d_direct")
    return out


def resample_conv1(vox: torch.Tensor, minv: torch.Tensor, new_size: int, w_tf: torch.Tensor, bias: torch.Tensor,
                   alpha: Optional[torch.Tensor], dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """Fused resampler + axis transform + e_conv1 (5^3 s2, 1->8) + bias + PReLU; empty tiles skip the conv.
    vox fp32 [B,S,S,S,1], minv fp32 [B,3,4], w_tf fp32 [5,5,5,1,8] -> 16-bit [B,new/2,new/2,new/2,8]."""
    vox = _cuda(vox, torch.float32)
    minv = _cuda(minv, torch.float32)
    w_tf = _cuda(w_tf, torch.float32)
    bias = _cuda(bias, torch.float32)
    B, S = vox.shape[0], vox.shape[1]
    if tuple(vox.shape) != (B, S, S, S, 1) or tuple(w_tf.shape) != (5, 5, 5, 1, 8) or tuple(minv.shape) != (B, 3, 4):
        raise ValueError(f"resample_conv1: unsupported shapes {tuple(vox.shape)}, {tuple(w_tf.shape)}")
    No = new_size // 2
    out = torch.empty((B, No, No, No, 8), device=vox.device, dtype=dtype)
    check(lib.rn_resample_conv1_fused(vox.data_ptr(), minv.data_ptr(), w_tf.data_ptr(), bias.data_ptr(), _ptr(alpha),
                                      out.data_ptr(), B, S, new_size, fmt_of(dtype), _stream()),
          "rn_resample_conv1_fused")
    return out


def binvox_decode(pairs_list, dims, fix_coords: bool = True, device="cuda") -> torch.Tensor:
    """Run-length (value, count) byte pairs of n binvox payloads -> float32 [n, d0, d2, d1, 1] on the device."""
    import numpy as np
    d0, d1, d2 = (int(v) for v in dims)
    kept, starts, first = [], [], [0]
    for pairs in pairs_list:
        pairs = np.asarray(pairs, np.uint8).reshape(-1, 2)
        pairs = pairs[pairs[:, 1] != 0]                       # zero-length runs carry no voxels
        cnt = pairs[:, 1].astype(np.int64)
        if int(cnt.sum()) != d0 * d1 * d2:
            raise ValueError(f"binvox payload decodes to {int(cnt.sum())} voxels, dims say {d0 * d1 * d2}")
        kept.append(pairs)
        starts.append(np.cumsum(cnt) - cnt)
        first.append(first[-1] + pairs.shape[0])
    pairs_d = torch.from_numpy(np.ascontiguousarray(np.concatenate(kept))).to(device)
    starts_d = torch.from_numpy(np.concatenate(starts).astype(np.int32)).to(device)
    first_d = torch.tensor(first, dtype=torch.int32, device=device)
    n = len(kept)
    shape = (n, d0, d2, d1, 1) if fix_coords else (n, d0, d1, d2, 1)
    out = torch.empty(shape, device=device, dtype=torch.float32)
    check(lib.rn_binvox_decode(pairs_d.data_ptr(), starts_d.data_ptr(), first_d.data_ptr(), out.data_ptr(), n, d0, d1, d2,
                               1 if fix_coords else 0, _stream()), "rn_binvox_decode")
    return out


# --------------------------------------------------------------------------------------- misc
def cast_to_16(x: torch.Tensor, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    x = _cuda(x, torch.float32)
    out = torch.empty(x.shape, device=x.device, dtype=dtype)
    check(lib.rn_cast_f32_to_16(x.data_ptr(), out.data_ptr(), x.numel(), x.numel(), fmt_of(dtype), _stream()), "cast")
    return out


def cast_to_f32(x: torch.Tensor) -> torch.Tensor:
    x = _cuda(x)
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    check(lib.rn_cast_16_to_f32(x.data_ptr(), out.data_ptr(), x.numel(), fmt_of(x.dtype), _stream()), "cast")
    return out


def phong_composite(img: torch.Tensor, light_dir: torch.Tensor, light_col: torch.Tensor, ambient: float,
                    k_diffuse: float, background_white: bool = False, with_mask: bool = True,
                    want_u8: bool = False):
    img = _cuda(img, torch.float32)
    B, H, W, _ = img.shape
    light_dir = _cuda(light_dir.to(device=img.device, dtype=torch.float32))
    light_col = _cuda(light_col.to(device=img.device, dtype=torch.float32))
    if light_dir.shape[0] == 1 and B > 1:
        light_dir = light_dir.expand(B, 3).contiguous()
    if light_col.shape[0] == 1 and B > 1:
        light_col = light_col.expand(B, 3).contiguous()
    out = torch.empty_like(img)
    u8 = torch.empty(img.shape, device=img.device, dtype=torch.uint8) if want_u8 else None
    check(lib.rn_phong_composite(img.data_ptr(), light_dir.data_ptr(), light_col.data_ptr(), float(ambient),
                                 float(k_diffuse), 1 if background_white else 0, 1 if with_mask else 0,
                                 out.data_ptr(), _ptr(u8), B, H, W, _stream()), "rn_phong_composite")
    return (out, u8) if want_u8 else out


def bias_act(x: torch.Tensor, bias: Optional[torch.Tensor], alpha: Optional[torch.Tensor], act: Optional[str],
             residual: Optional[torch.Tensor] = None, want32: bool = False):
    """y = act(x + bias[c]) + residual on a 16-bit channel-last tensor (rn_bias_act_16)."""
    x = _cuda(x)
    if x.dtype == torch.float32:
        x = cast_to_16(x)
    C_ = x.shape[-1]
    a = _ACT[act]
    dev = x.device
    if bias is not None:
        bias = _cuda(bias.to(device=dev, dtype=torch.float32))
    if alpha is not None:
        alpha = _cuda(alpha.to(device=dev, dtype=torch.float32))
    if residual is not None:
        residual = _cuda(residual)
        if residual.dtype != x.dtype:
            residual = cast_to_16(residual.float(), x.dtype) if residual.dtype == torch.float32 else residual.to(x.dtype)
    out16 = None if want32 else torch.empty_like(x)
    out32 = torch.empty(x.shape, device=dev, dtype=torch.float32) if want32 else None
    check(lib.rn_bias_act_16(x.data_ptr(), _ptr(bias), _ptr(alpha), a, _ptr(residual), _ptr(out16), _ptr(out32),
                             x.numel(), C_, fmt_of(x.dtype), _stream()), "rn_bias_act_16")
    return out32 if want32 else out16


# --------------------------------------------------------------------------------------- texture decoder (config 4)
def fully_connected(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], alpha: Optional[torch.Tensor],
                    want32: bool = True, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """y = prelu(x @ w + bias; alpha).  x [B,K] fp32, w [K,N] fp32 (TF layout) -> [B,N] fp32 (or 16-bit)."""
    x = _cuda(x, torch.float32)
    w = _cuda(w, torch.float32)
    B, K = x.shape
    N = w.shape[1]
    out = torch.empty((B, N), device=x.device, dtype=torch.float32 if want32 else dtype)
    check(lib.rn_fully_connected(x.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(alpha), None if want32 else out.data_ptr(),
                                 out.data_ptr() if want32 else None, B, K, N, fmt_of(dtype), _stream()),
          "rn_fully_connected")
    return out


def conv3d_small(x: torch.Tensor, w_tf: torch.Tensor, bias: Optional[torch.Tensor], alpha: Optional[torch.Tensor],
                 stride: int, transposed: bool, want32: bool = True, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """Thin (<= 8 channel) conv3d / conv3d_transpose, TF SAME, + bias + PReLU.  x [B,H,W,D,Cin] fp32 or 16-bit;
    w_tf fp32 [k,k,k,Cin,Cout] (forward) or [k,k,k,Cout,Cin] (transposed)."""
    x = _cuda(x)
    w_tf = _cuda(w_tf, torch.float32)
    B, H, W, D, Cin = x.shape
    k = w_tf.shape[0]
    Cout = w_tf.shape[3] if transposed else w_tf.shape[4]
    if transposed:
        oshape = (B, H * stride, W * stride, D * stride, Cout)
    else:
        oshape = (B, -(-H // stride), -(-W // stride), -(-D // stride), Cout)
    out = torch.empty(oshape, device=x.device, dtype=torch.float32 if want32 else dtype)
    check(lib.rn_conv3d_small(x.data_ptr(), 1 if x.dtype == torch.float32 else 0, w_tf.data_ptr(), _ptr(bias),
                              _ptr(alpha), None if want32 else out.data_ptr(), out.data_ptr() if want32 else None,
                              B, H, W, D, Cin, Cout, k, stride, 1 if transposed else 0,
                              fmt_of(dtype if x.dtype == torch.float32 else x.dtype), _stream()), "rn_conv3d_small")
    return out


def concat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """tf.concat([a, b], axis=-1) for fp32 channel-last tensors."""
    a = _cuda(a, torch.float32)
    b = _cuda(b, torch.float32)
    assert a.shape[:-1] == b.shape[:-1]
    Ca, Cb = a.shape[-1], b.shape[-1]
    out = torch.empty(tuple(a.shape[:-1]) + (Ca + Cb,), device=a.device, dtype=torch.float32)
    check(lib.rn_concat_channels_f32(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel() // Ca, Ca, Cb, _stream()),
          "rn_concat_channels_f32")
    return out
