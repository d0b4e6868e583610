"""Thin torch-tensor front end of the network-layer entry points of libmdctgan_b200.so
(include/mdctgan_b200.h, "network layers").  Everything here launches hand-written CUDA on the current
torch stream; nothing calls a torch / cuDNN compute op.

A value travelling through a network is a `Feat`: a raw NHWC fp32 tensor plus the normalisation /
activation its consumer still has to apply (`scale/shift/per_sample/act`).  InstanceNorm2d / BatchNorm2d
never run as kernels of their own: the producing convolution accumulates (sum, sumsq) in its epilogue,
`finalize_norm` turns them into per-(sample, channel) scale & shift, and the consuming convolution applies
them (and the ReLU / LeakyReLU) while it gathers its input.
"""
from __future__ import annotations

import ctypes
from ctypes import c_double, c_float, c_int, c_int64, c_void_p
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

import os

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH = 0, 1, 2, 3
PAD_ZERO, PAD_REFLECT = 0, 1

# Which convolution kernel runs the tensor-core-shaped layers (Cin % 4 == 0, Cout % 32 == 0):
#   "umma"   tcgen05 implicit GEMM, 3xTF32 split (fp32-class results)            [default]
#   "tf32"   tcgen05 implicit GEMM, single TF32 pass (what cuDNN does under torch's default allow_tf32)
#   "direct" fp32 FFMA kernel of nn_kernels.cuh (the on-device cross-check of the two above)
CONV_ENGINE = os.environ.get("MDCTGAN_CONV_ENGINE", "umma")

_bound = False


def _L():
    global _bound
    L = _lib.lib()
    if not _bound:
        L.mdctgan_conv2d_nhwc.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                          c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_double,
                                          c_float, c_int, c_void_p, c_void_p]
        L.mdctgan_conv2d_umma.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                          c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_double,
                                          c_float, c_int, c_void_p, c_int, c_void_p]
        L.mdctgan_conv2d_umma_pack_weight.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p]
        L.mdctgan_conv2d_umma_packed_floats.argtypes = [c_int, c_int]
        L.mdctgan_conv2d_umma_packed_floats.restype = c_int64
        L.mdctgan_conv2d_umma_supported.argtypes = [c_int, c_int]
        L.mdctgan_norm_finalize.argtypes = [c_void_p, c_int, c_int, c_double, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_float, c_void_p, c_void_p, c_void_p]
        L.mdctgan_norm_apply.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                         c_void_p, c_double, c_float, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
        L.mdctgan_attention_abs_pos.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                                c_void_p]
        L.mdctgan_residual_scale_add.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_void_p]
        L.mdctgan_avgpool3s2_nhwc.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
        L.mdctgan_nchw_to_nhwc.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
        L.mdctgan_nhwc_to_nchw.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
        _bound = True
    return L


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor, got device={t.device}; mdctgan_b200 has no CPU path")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError(f"{what}: expected a contiguous float32 tensor, got {t.dtype}, contiguous={t.is_contiguous()}")


@dataclass
class Feat:
    """Raw NHWC activation + what its consumer must apply: v = act(x*scale + shift)."""
    x: torch.Tensor                          # [B, H, W, C] fp32
    scale: Optional[torch.Tensor] = None     # [B, C] (per_sample) or [C]
    shift: Optional[torch.Tensor] = None
    per_sample: bool = True
    act: int = ACT_NONE
    stats: Optional[torch.Tensor] = None     # [B, C, 2] float64 (sum, sumsq) left by the producer, not yet finalised
    # InstanceNorm2d(affine=False) still in raw form: the consumer derives scale = rstd, shift = -mean*rstd from the
    # producer's statistics itself (no finalize launch); mutually exclusive with scale / shift
    norm_stats: Optional[torch.Tensor] = None
    norm_count: float = 0.0
    norm_eps: float = 1e-5

    @property
    def shape(self):
        return self.x.shape

    @property
    def has_norm(self) -> bool:
        return self.scale is not None or self.norm_stats is not None

    @property
    def is_plain(self) -> bool:
        return not self.has_norm and self.act == ACT_NONE


class _StatsArena:
    """Zeroed float64 scratch for the (sum, sumsq) statistics of one pass.  Inside `stats_pass(device)` every
    statistics buffer is a slice of one arena that is cleared with a single memset at the start of the pass
    (instead of one torch.zeros launch per normalised layer)."""

    def __init__(self, device, capacity=1 << 21):
        self.buf = torch.zeros(capacity, dtype=torch.float64, device=device)
        self.used = 0
        self.dirty = 0
        self.depth = 0

    def take(self, n):
        n = (n + 15) // 16 * 16
        if self.depth == 0 or self.used + n > self.buf.numel():
            return None
        out = self.buf[self.used:self.used + n]
        self.used += n
        self.dirty = max(self.dirty, self.used)
        return out


_arenas = {}


class stats_pass:
    """Context manager around one forward pass (inference, or a whole training step); nests."""

    def __init__(self, device):
        device = torch.device(device)
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        if key not in _arenas:
            _arenas[key] = _StatsArena(device)
        self.arena = _arenas[key]

    def __enter__(self):
        a = self.arena
        if a.depth == 0:
            if a.dirty:
                a.buf[:a.dirty].zero_()
            a.used = 0
        a.depth += 1
        return self

    def __exit__(self, *exc):
        self.arena.depth -= 1


def _new_stats(B, C, device):
    device = torch.device(device)
    a = _arenas.get((device.type, device.index if device.index is not None else torch.cuda.current_device()))
    if a is not None:
        t = a.take(B * C * 2)
        if t is not None:
            return t[:B * C * 2].view(B, C, 2)
    return torch.zeros((B, C, 2), dtype=torch.float64, device=device)


def to_nhwc(x_nchw: torch.Tensor) -> Feat:
    _req(x_nchw, "to_nhwc")
    B, C, H, W = x_nchw.shape
    y = torch.empty((B, H, W, C), dtype=torch.float32, device=x_nchw.device)
    if y.numel():
        with torch.cuda.device(x_nchw.device):
            _lib.check(_L().mdctgan_nchw_to_nhwc(x_nchw.data_ptr(), y.data_ptr(), B, C, H * W, _stream(y)))
    return Feat(y)


def to_nchw(f: Feat) -> torch.Tensor:
    f = materialize(f)
    B, H, W, C = f.x.shape
    if C == 1:
        return f.x.reshape(B, 1, H, W)
    y = torch.empty((B, C, H, W), dtype=torch.float32, device=f.x.device)
    if y.numel():
        with torch.cuda.device(y.device):
            _lib.check(_L().mdctgan_nhwc_to_nchw(f.x.data_ptr(), y.data_ptr(), B, C, H * W, _stream(y)))
    return y


def pack_conv_weight(w: torch.Tensor, transposed: bool = False) -> torch.Tensor:
    """[Cout,Cin,kh,kw] (Conv2d) or [Cin,Cout,kh,kw] (ConvTranspose2d) -> [kh*kw*Cin, Cout] contiguous.
    Host-side re-layout done once per weight version (not on the per-step path)."""
    w = w.detach().to(torch.float32)
    w = w.permute(2, 3, 0, 1) if transposed else w.permute(2, 3, 1, 0)
    kh, kw, cin, cout = w.shape
    return w.reshape(kh * kw * cin, cout).contiguous()


def umma_supported(cin: int, cout: int) -> bool:
    return CONV_ENGINE != "direct" and bool(_L().mdctgan_conv2d_umma_supported(int(cin), int(cout)))


def pack_conv_weight_umma(w_kn: torch.Tensor) -> torch.Tensor:
    """[K, Cout] fp32 (pack_conv_weight) -> the tcgen05 kernel's shared-memory image
    [ceil(K/32)][hi|lo][Cout][32] with TF32 hi / lo parts (conv_umma.cuh).  One small launch per weight version."""
    _req(w_kn, "pack_conv_weight_umma")
    K, cout = w_kn.shape
    out = torch.empty(int(_L().mdctgan_conv2d_umma_packed_floats(K, cout)), dtype=torch.float32, device=w_kn.device)
    with torch.cuda.device(w_kn.device):
        _lib.check(_L().mdctgan_conv2d_umma_pack_weight(w_kn.data_ptr(), K, cout, out.data_ptr(), _stream(w_kn)))
    return out


def conv2d(f: Feat, w_packed: torch.Tensor, bias: Optional[torch.Tensor], *, kh: int, kw: int, stride: int = 1, pad: int = 0,
           pad_mode: int = PAD_ZERO, transposed: bool = False, output_padding: int = 0, act: int = ACT_NONE,
           want_stats: bool = False, w_umma: Optional[torch.Tensor] = None) -> Feat:
    x = f.x
    _req(x, "conv2d input")
    B, H, W, Cin = x.shape
    K, Cout = w_packed.shape
    if K != kh * kw * Cin:
        raise RuntimeError(f"conv2d: packed weight has K={K}, expected {kh}*{kw}*{Cin}")
    if transposed:
        Ho = (H - 1) * stride - 2 * pad + kh + output_padding
        Wo = (W - 1) * stride - 2 * pad + kw + output_padding
    else:
        Ho = (H + 2 * pad - kh) // stride + 1
        Wo = (W + 2 * pad - kw) // stride + 1
    if f.stats is not None:
        raise RuntimeError("conv2d: input still carries un-finalised statistics (missing norm layer?)")
    y = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    stats = _new_stats(B, Cout, x.device) if want_stats else None
    in_act = f.act
    if not f.has_norm and f.act != ACT_NONE:
        f = materialize(f)      # an activation without an affine in front of it: apply it for real
        x, in_act = f.x, ACT_NONE
    use_umma = w_umma is not None and CONV_ENGINE != "direct"
    if use_umma and transposed and stride > 1 and (Ho % stride or Wo % stride or Cin % 32 or kh < stride or kw < stride):
        use_umma = False        # the tensor-core kernel runs strided transposed convolutions per output parity class only
    if B:
        with torch.cuda.device(x.device):
            if use_umma:
                _lib.check(_L().mdctgan_conv2d_umma(x.data_ptr(), B, H, W, Cin, w_umma.data_ptr(), _ptr(bias), y.data_ptr(), Ho, Wo, Cout,
                                                    kh, kw, stride, pad, pad_mode, 1 if transposed else 0, _ptr(f.scale), _ptr(f.shift),
                                                    1 if f.per_sample else 0, in_act, _ptr(f.norm_stats), f.norm_count, f.norm_eps,
                                                    act, _ptr(stats), 1 if CONV_ENGINE == "tf32" else 0, _stream(x)))
            else:
                _lib.check(_L().mdctgan_conv2d_nhwc(x.data_ptr(), B, H, W, Cin, w_packed.data_ptr(), _ptr(bias), y.data_ptr(), Ho, Wo, Cout,
                                                    kh, kw, stride, pad, pad_mode, 1 if transposed else 0, _ptr(f.scale), _ptr(f.shift),
                                                    1 if f.per_sample else 0, in_act, _ptr(f.norm_stats), f.norm_count, f.norm_eps,
                                                    act, _ptr(stats), _stream(x)))
    return Feat(y, stats=stats)


def finalize_norm(f: Feat, *, eps: float = 1e-5, mode: int = 0, gamma=None, beta=None, running_mean=None, running_var=None,
                  momentum: float = 0.1, eager: bool = False) -> Feat:
    """mode 0: InstanceNorm2d(affine=False); 1: BatchNorm2d training (batch statistics, updates the running
    buffers); 2: BatchNorm2d eval (running statistics).  Returns the same raw tensor with the normalisation
    attached: BatchNorm as scale / shift (one small launch); InstanceNorm as the raw statistics (no launch -- every
    consumer kernel derives rstd / mean itself) unless `eager`."""
    B, H, W, C = f.x.shape
    if mode != 2 and f.stats is None:
        raise RuntimeError("finalize_norm: the producer did not record statistics")
    if mode == 0 and not eager:
        return Feat(f.x, per_sample=True, act=ACT_NONE, norm_stats=f.stats, norm_count=float(H * W), norm_eps=float(eps))
    n = B * C if mode == 0 else C
    scale = torch.empty(n, dtype=torch.float32, device=f.x.device)
    shift = torch.empty(n, dtype=torch.float32, device=f.x.device)
    if n:
        with torch.cuda.device(f.x.device):
            _lib.check(_L().mdctgan_norm_finalize(_ptr(f.stats), B, C, float(H * W), eps, mode, _ptr(gamma), _ptr(beta),
                                                  _ptr(running_mean), _ptr(running_var), momentum, scale.data_ptr(), shift.data_ptr(),
                                                  _stream(f.x)))
    return Feat(f.x, scale=scale, shift=shift, per_sample=(mode == 0), act=ACT_NONE)


def resolve_norm(f: Feat) -> Feat:
    """Turn a deferred InstanceNorm (raw statistics) into explicit scale / shift with the finalize kernel."""
    if f.norm_stats is None:
        return f
    B, H, W, C = f.x.shape
    g = finalize_norm(Feat(f.x, stats=f.norm_stats), eps=f.norm_eps, mode=0, eager=True)
    return Feat(f.x, g.scale, g.shift, True, f.act)


def with_act(f: Feat, act: int) -> Feat:
    if f.act != ACT_NONE:
        f = materialize(f)
    return Feat(f.x, f.scale, f.shift, f.per_sample, act, f.stats, f.norm_stats, f.norm_count, f.norm_eps)


def combine(a: Feat, b: Optional[Feat] = None, act_out: int = ACT_NONE) -> Feat:
    """act_out( act_a(a*sa+ta) [+ act_b(b*sb+tb)] ) materialised as a plain NHWC tensor."""
    B, H, W, C = a.x.shape
    if b is not None and b.x.shape != a.x.shape:
        raise RuntimeError(f"combine: shape mismatch {tuple(a.x.shape)} vs {tuple(b.x.shape)}")
    if a.stats is not None or (b is not None and b.stats is not None):
        raise RuntimeError("combine: input carries un-finalised statistics")
    y = torch.empty_like(a.x)
    if b is not None and a.norm_stats is not None and b.norm_stats is not None and (a.norm_count, a.norm_eps) != (b.norm_count, b.norm_eps):
        b = resolve_norm(b)
    count, eps = (a.norm_count, a.norm_eps) if a.norm_stats is not None else ((b.norm_count, b.norm_eps) if b is not None else (0.0, 1e-5))
    if y.numel():
        with torch.cuda.device(y.device):
            _lib.check(_L().mdctgan_norm_apply(a.x.data_ptr(), _ptr(a.scale), _ptr(a.shift), 1 if a.per_sample else 0, a.act,
                                               _ptr(a.norm_stats), _ptr(b.x) if b is not None else None,
                                               _ptr(b.scale) if b is not None else None, _ptr(b.shift) if b is not None else None,
                                               1 if (b is not None and b.per_sample) else 0, b.act if b is not None else 0,
                                               _ptr(b.norm_stats) if b is not None else None, count, eps, y.data_ptr(), B, H * W, C, act_out,
                                               _stream(y)))
    return Feat(y)


def materialize(f: Feat) -> Feat:
    if f.stats is not None:
        raise RuntimeError("materialize: un-finalised statistics")
    return f if f.is_plain else combine(f)


def avgpool3s2(f: Feat) -> Feat:
    f = materialize(f)
    B, H, W, C = f.x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((B, Ho, Wo, C), dtype=torch.float32, device=f.x.device)
    if y.numel():
        with torch.cuda.device(y.device):
            _lib.check(_L().mdctgan_avgpool3s2_nhwc(f.x.data_ptr(), y.data_ptr(), B, H, W, C, _stream(y)))
    return Feat(y)


def attention(qkv: Feat, emb_h: torch.Tensor, emb_w: torch.Tensor, heads: int, dim_head: int, scale: float,
              want_stats: bool = True) -> Feat:
    """BoTNet attention on the fused q|k|v projection; records statistics for the BatchNorm2d that follows."""
    qkv = materialize(qkv)
    B, H, W, C3 = qkv.x.shape
    C = heads * dim_head
    if C3 != 3 * C:
        raise RuntimeError(f"attention: qkv has {C3} channels, expected 3*{heads}*{dim_head}")
    out = torch.empty((B, H, W, C), dtype=torch.float32, device=qkv.x.device)
    stats = _new_stats(B, C, out.device) if want_stats else None
    if B:
        with torch.cuda.device(out.device):
            _lib.check(_L().mdctgan_attention_abs_pos(qkv.x.data_ptr(), emb_h.data_ptr(), emb_w.data_ptr(), out.data_ptr(), B, H, W, heads,
                                                      dim_head, scale, _ptr(stats), _stream(out)))
    return Feat(out, stats=stats)


def residual_scale_add(sr: torch.Tensor, lr: torch.Tensor, lr_bins: int, low_scale: float = 1e-3) -> torch.Tensor:
    """sr[..., :lr_bins] *= low_scale; sr + lr   (pix2pixHD_model.py:631-635).  sr: [B,1,F,N] contiguous;
    lr: [B,1,F,N] or the first channel of a [B,2,F,N] tensor (rows of N with a uniform row stride)."""
    _req(sr, "residual sr")
    B, C, Fr, N = sr.shape
    assert C == 1 and lr.shape[0] == B and lr.shape[-2:] == (Fr, N) and lr.stride(-1) == 1 and lr.stride(-2) == N
    if B > 1 and lr.stride(0) != Fr * N:
        lr = lr[:, :1].contiguous()
    y = torch.empty_like(sr)
    if y.numel():
        with torch.cuda.device(sr.device):
            _lib.check(_L().mdctgan_residual_scale_add(sr.data_ptr(), lr.data_ptr(), N, y.data_ptr(), B * Fr, N, lr_bins, low_scale, _stream(sr)))
    return y
