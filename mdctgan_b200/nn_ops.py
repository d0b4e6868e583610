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

import contextlib
import ctypes
from ctypes import c_double, c_float, c_int, c_int64, c_void_p
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

import os

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
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
        # train step
        L.mdctgan_conv2d_wgrad.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                           c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_double, c_float, c_void_p, c_int64,
                                           c_int64, c_int64, c_void_p, c_int, c_void_p]
        L.mdctgan_conv2d_wgrad_umma_supported.argtypes = [c_int, c_int]
        L.mdctgan_norm_act_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_float, c_int, c_void_p, c_void_p, c_int,
                                           c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
        L.mdctgan_norm_act_bwd_folded.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_float, c_int, c_void_p, c_void_p, c_int,
                                                  c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
        L.mdctgan_reflect_pad_bwd_add.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
        L.mdctgan_act_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]
        L.mdctgan_add.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
        L.mdctgan_reflect_pad_bwd.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
        L.mdctgan_avgpool3s2_bwd.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
        L.mdctgan_attention_abs_pos_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                                    c_int, c_int, c_float, c_void_p]
        L.mdctgan_mse_const_fwd.argtypes = [c_void_p, c_int64, c_float, c_double, c_void_p, c_void_p]
        L.mdctgan_mse_const_bwd.argtypes = [c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p, c_int, c_void_p]
        L.mdctgan_bce_const_fwd.argtypes = [c_void_p, c_int64, c_float, c_double, c_void_p, c_void_p]
        L.mdctgan_multi_loss_fwd.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p]
        L.mdctgan_multi_loss_bwd.argtypes = [c_void_p, c_int, c_void_p]
        L.mdctgan_bce_const_bwd.argtypes = [c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p, c_int, c_void_p]
        L.mdctgan_l1_pair_fwd.argtypes = [c_void_p, c_void_p, c_int64, c_double, c_void_p, c_void_p]
        L.mdctgan_l1_pair_bwd.argtypes = [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_int, c_void_p]
        L.mdctgan_f64_to_f32.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
        L.mdctgan_disc_input_fwd.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p]
        L.mdctgan_disc_input_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
        L.mdctgan_adam_flat.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_float, c_int64,
                                        c_void_p, c_void_p]
        L.mdctgan_counter_inc.argtypes = [c_void_p, c_void_p]
        L.mdctgan_plane_stats.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
        L.mdctgan_upsample_nearest2x.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
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
    needs_grad: bool = True                  # False on data leaves: the tape skips the input gradient

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
    return Feat(y, needs_grad=False)


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


# development switches: run the forward / input-gradient convolutions of the tensor-core layers on the fp32 kernels ("direct")
_ROLE_ENGINE = {"fwd": os.environ.get("MDCTGAN_FWD_ENGINE", ""), "dgrad": os.environ.get("MDCTGAN_DGRAD_ENGINE", "")}
WGRAD_ENGINE = os.environ.get("MDCTGAN_WGRAD_ENGINE", "")      # "" = follow CONV_ENGINE; "direct" forces the fp32 FFMA weight gradient


def wgrad_engine(cin: int, cout: int) -> int:
    """Engine code of mdctgan_conv2d_wgrad for a layer: 1 / 2 = tcgen05 MN-major implicit GEMM (3xTF32 / single-pass TF32) where the
    channel counts are tensor-core shaped (Cin % 32 == 0, Cout % 32 == 0), else 0 = the fp32 FFMA kernel."""
    eng = WGRAD_ENGINE or CONV_ENGINE
    if eng == "direct" or not _L().mdctgan_conv2d_wgrad_umma_supported(int(cin), int(cout)):
        return 0
    return 2 if eng == "tf32" else 1


def pack_conv_weight_umma(w_kn: torch.Tensor) -> torch.Tensor:
    """[K, Cout] fp32 (pack_conv_weight) -> the tcgen05 kernel's shared-memory image
    [ceil(K/32)][hi|lo][Cout][32] with TF32 hi / lo parts (conv_umma.cuh).  One small launch per weight version."""
    _req(w_kn, "pack_conv_weight_umma")
    K, cout = w_kn.shape
    out = torch.empty(int(_L().mdctgan_conv2d_umma_packed_floats(K, cout)), dtype=torch.float32, device=w_kn.device)
    with torch.cuda.device(w_kn.device):
        _lib.check(_L().mdctgan_conv2d_umma_pack_weight(w_kn.data_ptr(), K, cout, out.data_ptr(), _stream(w_kn)))
    return out


def _conv_launch(f: Feat, w_packed: torch.Tensor, bias: Optional[torch.Tensor], *, kh: int, kw: int, stride: int = 1, pad: int = 0,
                 pad_mode: int = PAD_ZERO, transposed: bool = False, output_padding: int = 0, act: int = ACT_NONE,
                 want_stats: bool = False, w_umma: Optional[torch.Tensor] = None, out_hw=None, role: str = "fwd") -> Feat:
    """One convolution launch (no tape).  `out_hw` forces the output size of a transposed convolution (dgrad)."""
    x = f.x
    _req(x, "conv2d input")
    B, H, W, Cin = x.shape
    K, Cout = w_packed.shape
    if K != kh * kw * Cin:
        raise RuntimeError(f"conv2d: packed weight has K={K}, expected {kh}*{kw}*{Cin}")
    if out_hw is not None:
        Ho, Wo = out_hw
    elif transposed:
        Ho = (H - 1) * stride - 2 * pad + kh + output_padding
        Wo = (W - 1) * stride - 2 * pad + kw + output_padding
    else:
        Ho = (H + 2 * pad - kh) // stride + 1
        Wo = (W + 2 * pad - kw) // stride + 1
    if f.stats is not None:
        raise RuntimeError("conv2d: input still carries un-finalised statistics (missing norm layer?)")
    y = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    stats = _new_stats(B, Cout, x.device) if want_stats else None
    use_umma = w_umma is not None and CONV_ENGINE != "direct" and _ROLE_ENGINE.get(role, "") != "direct"
    if use_umma and transposed and stride > 1 and (Ho < stride or Wo < stride or Cin % 32 or kh < stride or kw < stride):
        use_umma = False        # the tensor-core kernel runs strided transposed convolutions per output parity class only
    if not use_umma and w_packed.stride(0) == 0:
        raise RuntimeError("conv2d: this layer only has a tcgen05 weight image (packing.WeightPacker) but the direct kernel was selected")
    if B:
        with torch.cuda.device(x.device):
            if use_umma:
                _lib.check(_L().mdctgan_conv2d_umma(x.data_ptr(), B, H, W, Cin, w_umma.data_ptr(), _ptr(bias), y.data_ptr(), Ho, Wo, Cout,
                                                    kh, kw, stride, pad, pad_mode, 1 if transposed else 0, _ptr(f.scale), _ptr(f.shift),
                                                    1 if f.per_sample else 0, f.act, _ptr(f.norm_stats), f.norm_count, f.norm_eps,
                                                    act, _ptr(stats), 1 if CONV_ENGINE == "tf32" else 0, _stream(x)))
            else:
                _lib.check(_L().mdctgan_conv2d_nhwc(x.data_ptr(), B, H, W, Cin, w_packed.data_ptr(), _ptr(bias), y.data_ptr(), Ho, Wo, Cout,
                                                    kh, kw, stride, pad, pad_mode, 1 if transposed else 0, _ptr(f.scale), _ptr(f.shift),
                                                    1 if f.per_sample else 0, f.act, _ptr(f.norm_stats), f.norm_count, f.norm_eps,
                                                    act, _ptr(stats), _stream(x)))
    return Feat(y, stats=stats)


def conv2d(f: Feat, w_packed: torch.Tensor, bias: Optional[torch.Tensor], *, kh: int, kw: int, stride: int = 1, pad: int = 0,
           pad_mode: int = PAD_ZERO, transposed: bool = False, output_padding: int = 0, act: int = ACT_NONE,
           want_stats: bool = False, w_umma: Optional[torch.Tensor] = None, owner=None) -> Feat:
    """nn.Conv2d / nn.ConvTranspose2d forward on a Feat.  Under `recording(tape)` the launch is also put on the tape
    (`owner` = the parameter-holding module, needed for dgrad / wgrad)."""
    if not f.has_norm and f.act != ACT_NONE:
        f = materialize(f)      # an activation without an affine in front of it: apply it for real
    out = _conv_launch(f, w_packed, bias, kh=kh, kw=kw, stride=stride, pad=pad, pad_mode=pad_mode, transposed=transposed,
                       output_padding=output_padding, act=act, want_stats=want_stats, w_umma=w_umma)
    if _tape is not None:
        if owner is None:
            raise RuntimeError("conv2d: recording a tape needs the owning module (dgrad / wgrad)")
        _tape.ops.append(_ConvOp(f, out, owner, kh, kw, stride, pad, pad_mode, transposed, act))
    return out


def finalize_norm(f: Feat, *, eps: float = 1e-5, mode: int = 0, gamma=None, beta=None, running_mean=None, running_var=None,
                  momentum: float = 0.1, eager: bool = False) -> Feat:
    """mode 0: InstanceNorm2d(affine=False); 1: BatchNorm2d training (batch statistics, updates the running
    buffers); 2: BatchNorm2d eval (running statistics).  Returns the same raw tensor with the normalisation
    attached: BatchNorm as scale / shift (one small launch); InstanceNorm as the raw statistics (no launch -- every
    consumer kernel derives rstd / mean itself) unless `eager`."""
    B, H, W, C = f.x.shape
    if mode != 2 and f.stats is None:
        raise RuntimeError("finalize_norm: the producer did not record statistics")
    if _tape is not None and (mode == 2 or eager):
        raise NotImplementedError("finalize_norm on a tape: eval-mode BatchNorm2d / eager InstanceNorm2d have no backward here "
                                  "(call model.train() before a train step)")
    if mode == 0 and not eager:
        out = Feat(f.x, per_sample=True, act=ACT_NONE, norm_stats=f.stats, norm_count=float(H * W), norm_eps=float(eps))
        if _tape is not None:
            _tape.ops.append(_ViewOp(f, out, "in", ACT_NONE, stats=f.stats, count=float(H * W), eps=float(eps)))
        return out
    n = B * C if mode == 0 else C
    scale = torch.empty(n, dtype=torch.float32, device=f.x.device)
    shift = torch.empty(n, dtype=torch.float32, device=f.x.device)
    if n:
        with torch.cuda.device(f.x.device):
            _lib.check(_L().mdctgan_norm_finalize(_ptr(f.stats), B, C, float(H * W), eps, mode, _ptr(gamma), _ptr(beta),
                                                  _ptr(running_mean), _ptr(running_var), momentum, scale.data_ptr(), shift.data_ptr(),
                                                  _stream(f.x)))
    out = Feat(f.x, scale=scale, shift=shift, per_sample=(mode == 0), act=ACT_NONE)
    if _tape is not None:
        _tape.ops.append(_ViewOp(f, out, "bn", ACT_NONE, stats=f.stats, count=float(H * W), eps=float(eps), gamma=gamma, beta=beta))
    return out


def resolve_norm(f: Feat) -> Feat:
    """Turn a deferred InstanceNorm (raw statistics) into explicit scale / shift with the finalize kernel."""
    if f.norm_stats is None:
        return f
    if _tape is not None:
        raise NotImplementedError("resolve_norm on a tape")
    B, H, W, C = f.x.shape
    g = finalize_norm(Feat(f.x, stats=f.norm_stats), eps=f.norm_eps, mode=0, eager=True)
    return Feat(f.x, g.scale, g.shift, True, f.act)


def with_act(f: Feat, act: int) -> Feat:
    if f.act != ACT_NONE:
        f = materialize(f)
    out = Feat(f.x, f.scale, f.shift, f.per_sample, act, f.stats, f.norm_stats, f.norm_count, f.norm_eps, f.needs_grad)
    if _tape is not None and act != ACT_NONE:
        last = _tape.ops[-1] if _tape.ops else None
        if isinstance(last, _ViewOp) and last.out is f and last.act == ACT_NONE:
            last.out, last.act = out, act            # norm + activation: one backward
        else:
            if f.has_norm:
                raise NotImplementedError("with_act on a tape: activation on a normalised view that is not the latest tape entry")
            if act == ACT_TANH:
                raise NotImplementedError("with_act on a tape: stand-alone tanh")
            _tape.ops.append(_ViewOp(f, out, "act", act))
    return out


def combine(a: Feat, b: Optional[Feat] = None, act_out: int = ACT_NONE, want_stats: bool = False) -> Feat:
    """act_out( act_a(a*sa+ta) [+ act_b(b*sb+tb)] ) materialised as a plain NHWC tensor.  `want_stats`: also take the
    per-(sample, channel) (sum, sumsq) of the result for a normalisation layer that follows (one more small launch)."""
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
    stats = None
    if want_stats:
        stats = _new_stats(B, C, y.device)
        if y.numel():
            with torch.cuda.device(y.device):
                _lib.check(_L().mdctgan_plane_stats(y.data_ptr(), B, H * W, C, stats.data_ptr(), _stream(y)))
    out = Feat(y, stats=stats)
    if _tape is not None:
        _tape.ops.append(_CombineOp(a, b, out, act_out))
    return out


def upsample_nearest2x(f: Feat) -> Feat:
    """F.interpolate(scale_factor=2.0, mode="nearest") (networks.py:396) on the materialised value of f."""
    f = materialize(f)
    B, H, W, C = f.x.shape
    y = torch.empty((B, 2 * H, 2 * W, C), dtype=torch.float32, device=f.x.device)
    if y.numel():
        with torch.cuda.device(y.device):
            _lib.check(_L().mdctgan_upsample_nearest2x(f.x.data_ptr(), y.data_ptr(), B, H, W, C, 0, _stream(y)))
    out = Feat(y, needs_grad=f.needs_grad)
    if _tape is not None and f.needs_grad:
        _tape.ops.append(_UpsampleOp(f, out))
    return out


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
    out = Feat(y, needs_grad=f.needs_grad)
    if _tape is not None and f.needs_grad:
        _tape.ops.append(_PoolOp(f, out))
    return out


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
    res = Feat(out, stats=stats)
    if _tape is not None:
        _tape.ops.append(_AttnOp(qkv, res, emb_h, emb_w, heads, dim_head, scale))
    return res


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


# =====================================================================================================================
# Tape: what torch autograd does for the reference's train step (train.py:175-202), written out over our own kernels.
# Forward ops append an entry while `recording(tape)` is active; `Tape.backward(grads, ...)` walks the entries in
# reverse.  Gradients are keyed by Feat object and are taken with respect to the VALUE a Feat stands for
# (act(norm(x)) for a deferred view, x for a raw tensor).
# =====================================================================================================================
_tape = None


class _SideWork:
    """Side CUDA streams for work whose results are only needed at the end of the step (weight gradients and the norm_finalize they
    read).  Several streams, used round-robin: the weight-gradient kernels of a step are short and latency-bound, on ONE stream they
    serialise into a chain about as long as the sweep itself (cfg4: 2.2 ms of kernels next to a 2.4 ms sweep) and the update tail
    waits for its end."""

    def __init__(self, device):
        self.device = device
        self.streams = [torch.cuda.Stream(device, priority=int(os.environ.get("MDCTGAN_SIDE_PRIORITY", "0"))) for _ in range(N_SIDE_STREAMS)]
        self.next = 0
        self.keep = []          # tensors read on the side streams: kept alive until join() so the allocator cannot recycle them
        self.active = False

    @property
    def stream(self):           # (single-stream view: the opt-in bucketed exchange runs with N_SIDE_STREAMS = 1)
        return self.streams[0]


_side = {}
_wgrad_hooks = {}       # id(module) -> callable(stream the module's weight-gradient kernel was enqueued on)
SIDE_STREAM_WGRAD = os.environ.get("MDCTGAN_SIDE_WGRAD", "1") != "0"
# r02: 2 / 3 / 4 side streams measured 5.07 / 5.07 / 5.11 ms per cfg4 step against 4.91 with one -- the step is bound by SM contention, and one
# stream throttles the weight gradients to what the dependent chain leaves free
N_SIDE_STREAMS = 1 if os.environ.get("MDCTGAN_BUCKETED_ALLREDUCE", "0") == "1" else max(1, int(os.environ.get("MDCTGAN_SIDE_STREAMS", "1")))


def _side_stream(dy: torch.Tensor):
    """Fork: make a side stream wait for everything issued so far on the current stream; returns it (or None when disabled)."""
    if not SIDE_STREAM_WGRAD:
        return None
    key = dy.device.index
    sw = _side.get(key)
    if sw is None:
        sw = _side[key] = _SideWork(dy.device)
    main = torch.cuda.current_stream(dy.device)
    if any(main == st for st in sw.streams):
        return None
    st = sw.streams[sw.next % len(sw.streams)]
    sw.next += 1
    ev = torch.cuda.Event()
    ev.record(main)
    st.wait_event(ev)
    sw.keep.append(dy)
    sw.active = True
    return st


def side_streams(device):
    """The side streams that carry work of the current step (for joins / events)."""
    sw = _side.get(torch.device(device).index)
    return list(sw.streams) if (sw is not None and sw.active) else []


def join_side_work(device) -> None:
    """Join: the current stream waits for the side streams (call before anything reads the weight gradients)."""
    sw = _side.get(torch.device(device).index)
    if sw is None or not sw.active:
        return
    cur = torch.cuda.current_stream(sw.device)
    for st in sw.streams:
        cur.wait_stream(st)
    sw.keep.clear()
    sw.active = False
    sw.next = 0


_branch_streams = {}
# MDCTGAN_STREAM_PRIORITY=1: the streams of the dependent chain (sweeps, branches) get a higher CUDA priority than the weight-gradient
# side stream and the update stream, whose kernels only have to finish by the end of the step
STREAM_PRIORITY = os.environ.get("MDCTGAN_STREAM_PRIORITY", "1") == "1"
_HI = int(os.environ.get("MDCTGAN_CHAIN_PRIORITY", "-2")) if STREAM_PRIORITY else 0
_MID = set(os.environ.get("MDCTGAN_MID_PRIORITY_STREAMS", "").split(","))      # streams between the chain and the rest (priority -1)
# r02 (cfg4 step): no priorities 4.91 ms; chain high / update low 4.71; discriminator sweep low as well 4.69 (it hides behind the generator's)
_LOW_PRIORITY = set(os.environ.get("MDCTGAN_LOW_PRIORITY_STREAMS", "update,sweep_D,comm").split(","))
PARALLEL_BRANCHES = os.environ.get("MDCTGAN_PARALLEL_BRANCHES", "1") != "0"


def _branch_stream(device, i: int, parent=None):
    """Stream of branch i forked from `parent` (default: the current stream): every parent has its own set, so two sweeps
    running on different streams do not serialise on shared branch streams."""
    parent = parent if parent is not None else torch.cuda.current_stream(device)
    key = (torch.device(device).index, parent.cuda_stream, i)
    if key not in _branch_streams:
        _branch_streams[key] = torch.cuda.Stream(device, priority=getattr(parent, "priority", 0) if STREAM_PRIORITY else 0)      # a branch inherits its parent's priority
    return _branch_streams[key]


def aux_stream(device, name: str):
    """A named long-lived stream (e.g. the discriminator sweep of the train step)."""
    key = (torch.device(device).index, name)
    if key not in _branch_streams:
        prio = 0 if name in _LOW_PRIORITY else _HI
        if STREAM_PRIORITY and name in _MID:
            prio = -1
        _branch_streams[key] = torch.cuda.Stream(device, priority=prio)
    return _branch_streams[key]


def run_branches(fns, device):
    """Independent sub-graphs (the PatchGAN scales of the multiscale discriminator) on their own CUDA streams: each branch is
    forked from the current stream, records into its own tape, and the current stream joins them all afterwards.  The backward
    of the group forks / joins the same way (`_ParallelOp`).  Most kernels of one branch fill a fraction of the 148 SMs."""
    main = torch.cuda.current_stream(device)
    use_streams = PARALLEL_BRANCHES and len(fns) > 1
    results, tapes = [], []
    ev = torch.cuda.Event()
    ev.record(main)
    outer = _tape
    for i, fn in enumerate(fns):
        sub = Tape() if outer is not None else None
        st = _branch_stream(device, i, main) if use_streams else main
        if use_streams:
            st.wait_event(ev)
        with torch.cuda.stream(st), (recording(sub) if sub is not None else contextlib.nullcontext()):
            results.append(fn())
        tapes.append(sub)
    if use_streams:
        for i in range(len(fns)):
            main.wait_stream(_branch_stream(device, i, main))
    if outer is not None:
        outer.ops.append(_ParallelOp(tapes, device))
    return results


class _ParallelOp:
    def __init__(self, tapes, device):
        self.tapes, self.device = tapes, device

    def backward(self, G, wgrad, nb):
        main = torch.cuda.current_stream(self.device)
        use_streams = PARALLEL_BRANCHES and len(self.tapes) > 1
        ev = torch.cuda.Event()
        ev.record(main)
        for i, t in enumerate(self.tapes):
            st = _branch_stream(self.device, i, main) if use_streams else main
            if use_streams:
                st.wait_event(ev)
            with torch.cuda.stream(st):
                t.backward(G, wgrad, nb, join=False)
        if use_streams:
            for i in range(len(self.tapes)):
                main.wait_stream(_branch_stream(self.device, i, main))


class _MarkOp:
    """A position on the tape: when the backward sweep (with weight gradients) reaches it, every op recorded after it has been
    back-propagated -- e.g. "all gradients of this update bucket have been enqueued" (the pipelined optimiser step)."""

    def __init__(self, fn):
        self.fn = fn

    def backward(self, G, wgrad, nb):
        if wgrad:
            self.fn()


_marks = {}             # id(module) -> callable, consulted by models.networks.run_layers before it runs the module


def mark(module) -> None:
    fn = _marks.get(id(module))
    if fn is not None and _tape is not None:
        _tape.ops.append(_MarkOp(fn))


class Tape:
    def __init__(self):
        self.ops = []

    def backward(self, grads: "GradMap", wgrad: bool = True, nb: Optional[int] = None, join: bool = True):
        """`nb`: only the first nb samples of every saved tensor take part (the generator-loss sweep through the
        discriminator runs on the fake half of the [fake ; real] batch).  `wgrad=False`: input gradients only.
        `join=False` leaves the weight-gradient kernels running on the side stream (the caller joins later)."""
        for op in reversed(self.ops):
            op.backward(grads, wgrad, nb)
        if join:
            for sw in list(_side.values()):
                join_side_work(sw.device)


class recording:
    def __init__(self, tape: Tape):
        self.tape = tape

    def __enter__(self):
        global _tape
        self.prev, _tape = _tape, self.tape
        return self.tape

    def __exit__(self, *exc):
        global _tape
        _tape = self.prev


class _Folded:
    """A gradient that still sits in the reflection-padded geometry of its consumer: `dpad` [B, H + 2p, W + 2p, C] = the raw output of an
    input-gradient convolution whose forward had nn.ReflectionPad2d(p) in front.  The fold back onto [B, H, W, C] is deferred so that the
    next kernel of the chain does it while loading (InstanceNorm / BatchNorm backward), or fuses it with the accumulation into a gradient
    that already exists (residual skip): one launch less per convolution on the dependent chain of the sweep."""

    def __init__(self, dpad: torch.Tensor, pad: int):
        self.dpad, self.pad = dpad, pad
        B, Hp, Wp, C = dpad.shape
        self.shape = (B, Hp - 2 * pad, Wp - 2 * pad, C)

    def materialize(self, other: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, H, W, C = self.shape
        dx = torch.empty(self.shape, dtype=torch.float32, device=self.dpad.device)
        with torch.cuda.device(dx.device):
            if other is not None and C % 4 == 0:
                _lib.check(_L().mdctgan_reflect_pad_bwd_add(self.dpad.data_ptr(), other.data_ptr(), dx.data_ptr(), B, H, W, C, self.pad, _stream(dx)))
                return dx
            _lib.check(_L().mdctgan_reflect_pad_bwd(self.dpad.data_ptr(), dx.data_ptr(), B, H, W, C, self.pad, _stream(dx)))
        return dx if other is None else add(dx, other)


FOLD_FUSION = os.environ.get("MDCTGAN_FOLD_FUSION", "1") != "0"


class GradMap:
    def __init__(self):
        self.g = {}

    def add(self, f: Feat, t):
        """t: a gradient tensor, or a _Folded one (kept lazy until something needs the plain tensor)."""
        if not f.needs_grad:
            return
        cur = self.g.get(id(f))
        if cur is None:
            self.g[id(f)] = t
        elif isinstance(t, _Folded):
            self.g[id(f)] = t.materialize(cur.materialize() if isinstance(cur, _Folded) else cur)
        elif isinstance(cur, _Folded):
            self.g[id(f)] = cur.materialize(t)
        else:
            self.g[id(f)] = add(cur, t)

    def pop(self, f: Feat, lazy_ok: bool = False):
        t = self.g.pop(id(f), None)
        return t.materialize() if (isinstance(t, _Folded) and not lazy_ok) else t

    def get(self, f: Feat):
        t = self.g.get(id(f))
        if isinstance(t, _Folded):
            t = self.g[id(f)] = t.materialize()
        return t


def _sl(t: Optional[torch.Tensor], nb: Optional[int]):
    return t if (t is None or nb is None) else t[:nb]


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(a)
    if y.numel():
        with torch.cuda.device(a.device):
            _lib.check(_L().mdctgan_add(a.data_ptr(), b.data_ptr(), y.data_ptr(), y.numel(), _stream(a)))
    return y


def act_bwd(dy: torch.Tensor, y: torch.Tensor, act: int) -> torch.Tensor:
    """dy * act'(.) evaluated from the activated value y (ReLU / LeakyReLU: the pre-activation works as well)."""
    g = torch.empty_like(dy)
    if g.numel():
        with torch.cuda.device(dy.device):
            _lib.check(_L().mdctgan_act_bwd(dy.data_ptr(), y.data_ptr(), g.data_ptr(), g.numel(), act, _stream(dy)))
    return g


def grad_of(p: torch.Tensor) -> torch.Tensor:
    """The gradient buffer a backward kernel accumulates into (a view of the flat bucket once optim.flatten ran)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


def _view_feat(f: Feat, nb: Optional[int]) -> Feat:
    """The first nb samples of a Feat (contiguous slices, no copies)."""
    if nb is None:
        return f
    C = f.x.shape[-1]
    sc = f.scale[:nb * C] if (f.scale is not None and f.per_sample) else f.scale
    sh = f.shift[:nb * C] if (f.shift is not None and f.per_sample) else f.shift
    return Feat(f.x[:nb], sc, sh, f.per_sample, f.act, None, _sl(f.norm_stats, nb), f.norm_count, f.norm_eps, f.needs_grad)


def _explicit_norm(f: Feat, cache_on: Optional[Feat]) -> Feat:
    """A deferred InstanceNorm (raw statistics) as explicit fp32 scale / shift: one tiny finalize launch instead of every CTA
    of the weight-gradient kernel re-deriving rstd in fp64 for every sample it visits.  Cached on the tape's Feat."""
    if f.norm_stats is None:
        return f
    cached = getattr(cache_on, "_explicit", None) if cache_on is not None else None
    if cached is None:
        B, H, W, C = f.x.shape
        scale = torch.empty(B * C, dtype=torch.float32, device=f.x.device)
        shift = torch.empty(B * C, dtype=torch.float32, device=f.x.device)
        with torch.cuda.device(f.x.device):
            _lib.check(_L().mdctgan_norm_finalize(f.norm_stats.data_ptr(), B, C, f.norm_count, f.norm_eps, 0, None, None, None, None, 0.0,
                                                  scale.data_ptr(), shift.data_ptr(), _stream(f.x)))
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(f.x.device))
        cached = (scale, shift, ev)
        if cache_on is not None:
            cache_on._explicit = cached
    else:
        torch.cuda.current_stream(f.x.device).wait_event(cached[2])      # computed on another (side) stream
    return Feat(f.x, cached[0], cached[1], True, f.act, None, None, 0.0, f.norm_eps, f.needs_grad)


class _ConvOp:
    def __init__(self, f, out, owner, kh, kw, stride, pad, pad_mode, transposed, act):
        self.f, self.out, self.owner = f, out, owner
        self.kh, self.kw, self.stride, self.pad, self.pad_mode, self.transposed, self.act = kh, kw, stride, pad, pad_mode, transposed, act

    def backward(self, G: GradMap, wgrad: bool, nb):
        dy = G.pop(self.out)
        if dy is None:
            return
        if self.act != ACT_NONE:
            dy = act_bwd(dy, _sl(self.out.x, nb), self.act)
        fv = _view_feat(self.f, nb)
        own = self.owner
        B, H, W, Cin = fv.x.shape
        _, Ho, Wo, Cout = dy.shape
        L = _L()
        if wgrad and own.weight.requires_grad:
            dw = grad_of(own.weight)
            db = grad_of(own.bias) if (own.bias is not None and own.bias.requires_grad) else None
            # element strides of dW[co][ci][tap] in the gradient tensor's own layout (the parameter's: contiguous reference layout, or
            # the [kh][kw][Cin][Cout] storage order of optim.FlatBucket, where co is contiguous)
            st_ = dw.stride()
            s_ci, s_co = (st_[0], st_[1]) if self.transposed else (st_[1], st_[0])
            s_tap = st_[3]
            if self.kh > 1 and st_[2] != self.kw * s_tap:
                raise RuntimeError(f"conv2d wgrad: unsupported weight-gradient strides {st_}")
            # nothing reads a weight gradient before the optimiser step: the wgrad kernels run on a side stream, concurrently with
            # the dgrad / norm-backward chain of the main stream (most kernels of this step fill a fraction of the 148 SMs)
            side = _side_stream(dy)
            engine = wgrad_engine(Cin, Cout)
            with torch.cuda.device(dy.device), (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                # the explicit (scale, shift) of a deferred InstanceNorm is only read by this weight-gradient launch: its finalize
                # launch goes to the side stream with it (44 launches per cfg4 step off the dgrad chain of the main stream)
                f = _explicit_norm(fv, self.f if nb is None else None)
                _lib.check(L.mdctgan_conv2d_wgrad(f.x.data_ptr(), B, H, W, Cin, dy.data_ptr(), Ho, Wo, Cout, self.kh, self.kw, self.stride,
                                                  self.pad, self.pad_mode, 1 if self.transposed else 0, _ptr(f.scale), _ptr(f.shift),
                                                  1 if f.per_sample else 0, f.act, _ptr(f.norm_stats), f.norm_count, f.norm_eps,
                                                  dw.data_ptr(), s_co, s_ci, s_tap, _ptr(db), engine, _stream(dy)))
            hook = _wgrad_hooks.get(id(own))
            if hook is not None:      # e.g. "every gradient of this all-reduce bucket has been enqueued" (bucketed exchange)
                hook(side if side is not None else torch.cuda.current_stream(dy.device))
        if not self.f.needs_grad:
            return
        # input gradient = the forward convolution on the transposed geometry with the same weights
        wd, wd_umma, flipped = own.packed_dgrad()
        g = Feat(dy)
        if self.transposed:          # ConvTranspose2d: a plain strided convolution of dy
            dv = _conv_launch(g, wd, None, kh=self.kh, kw=self.kw, stride=self.stride, pad=self.pad, w_umma=wd_umma, role="dgrad").x
            assert dv.shape[1:3] == (H, W)
        elif flipped:                # stride 1: plain convolution with the taps flipped, pad' = k - 1 - pad
            p_eff = 0 if self.pad_mode == PAD_REFLECT else self.pad
            dv = _conv_launch(g, wd, None, kh=self.kh, kw=self.kw, stride=1, pad=self.kh - 1 - p_eff, w_umma=wd_umma, role="dgrad").x
        else:                        # strided nn.Conv2d: transposed convolution of dy, output size forced to the input's
            p_eff = 0 if self.pad_mode == PAD_REFLECT else self.pad
            hw = (H + 2 * self.pad, W + 2 * self.pad) if self.pad_mode == PAD_REFLECT else (H, W)
            dv = _conv_launch(g, wd, None, kh=self.kh, kw=self.kw, stride=self.stride, pad=p_eff, transposed=True, w_umma=wd_umma,
                              out_hw=hw, role="dgrad").x
        if self.pad_mode == PAD_REFLECT:
            assert dv.shape[1:3] == (H + 2 * self.pad, W + 2 * self.pad), (dv.shape, H, W, self.pad)
            dv = _Folded(dv, self.pad)                # nn.ReflectionPad2d backward: deferred into the next kernel of the chain
            if not FOLD_FUSION:
                dv = dv.materialize()
        assert tuple(dv.shape) == tuple(fv.x.shape), (dv.shape, fv.x.shape)
        G.add(self.f, dv)


class _ViewOp:
    """out = act(norm(src)) as a deferred view: kind 'in' (InstanceNorm2d), 'bn' (train-mode BatchNorm2d), 'act'."""

    def __init__(self, src, out, kind, act, stats=None, count=0.0, eps=1e-5, gamma=None, beta=None):
        self.src, self.out, self.kind, self.act = src, out, kind, act
        self.stats, self.count, self.eps, self.gamma, self.beta = stats, count, eps, gamma, beta

    def backward(self, G: GradMap, wgrad: bool, nb):
        dv = G.pop(self.out, lazy_ok=self.kind != "act")
        if dv is None:
            return
        x = _sl(self.src.x, nb)
        if self.kind == "act":
            G.add(self.src, act_bwd(dv, x, self.act))
            return
        fold = 0
        if isinstance(dv, _Folded):
            if x.shape[-1] % 4 == 0:
                fold, dv = dv.pad, dv.dpad            # folded while the normalisation backward loads it
            else:
                dv = dv.materialize()
        if self.kind == "bn" and nb is not None:
            raise NotImplementedError("BatchNorm2d backward on a batch slice")
        B, H, W, C = x.shape
        dx = torch.empty_like(x)
        red = _new_stats(B, C, x.device)     # zeroed (sum g, sum g*xhat) scratch
        gam = self.gamma if self.kind == "bn" else None
        bet = self.beta if self.kind == "bn" else None
        dgam = grad_of(gam) if (gam is not None and wgrad and gam.requires_grad) else None
        dbet = grad_of(bet) if (bet is not None and wgrad and bet.requires_grad) else None
        with torch.cuda.device(x.device):
            _lib.check(_L().mdctgan_norm_act_bwd_folded(x.data_ptr(), dv.data_ptr(), dx.data_ptr(), _sl(self.stats, nb).data_ptr(), self.count,
                                                        self.eps, 0 if self.kind == "in" else 1, _ptr(gam), _ptr(bet), self.act, red.data_ptr(),
                                                        _ptr(dgam), _ptr(dbet), B, H, W, C, fold, _stream(x)))
        G.add(self.src, dx)


class _CombineOp:
    def __init__(self, a, b, out, act_out):
        self.a, self.b, self.out, self.act_out = a, b, out, act_out

    def backward(self, G: GradMap, wgrad: bool, nb):
        dy = G.pop(self.out)
        if dy is None:
            return
        if self.act_out != ACT_NONE:
            dy = act_bwd(dy, _sl(self.out.x, nb), self.act_out)
        G.add(self.a, dy)
        if self.b is not None:
            G.add(self.b, dy)


class _PoolOp:
    def __init__(self, f, out):
        self.f, self.out = f, out

    def backward(self, G: GradMap, wgrad: bool, nb):
        dy = G.pop(self.out)
        if dy is None:
            return
        B, H, W, C = _sl(self.f.x, nb).shape
        dx = torch.empty((B, H, W, C), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            _lib.check(_L().mdctgan_avgpool3s2_bwd(dy.data_ptr(), dx.data_ptr(), B, H, W, C, _stream(dy)))
        G.add(self.f, dx)


class _UpsampleOp:
    def __init__(self, f, out):
        self.f, self.out = f, out

    def backward(self, G: GradMap, wgrad: bool, nb):
        dy = G.pop(self.out)
        if dy is None:
            return
        B, H, W, C = _sl(self.f.x, nb).shape
        dx = torch.empty((B, H, W, C), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            _lib.check(_L().mdctgan_upsample_nearest2x(dy.data_ptr(), dx.data_ptr(), B, H, W, C, 1, _stream(dy)))
        G.add(self.f, dx)


class _AttnOp:
    def __init__(self, qkv, out, emb_h, emb_w, heads, d, scale):
        self.qkv, self.out, self.emb_h, self.emb_w, self.heads, self.d, self.scale = qkv, out, emb_h, emb_w, heads, d, scale

    def backward(self, G: GradMap, wgrad: bool, nb):
        dout = G.pop(self.out)
        if dout is None:
            return
        qkv = _sl(self.qkv.x, nb)
        B, H, W, _ = qkv.shape
        dqkv = torch.empty_like(qkv)
        dh = grad_of(self.emb_h) if (wgrad and self.emb_h.requires_grad) else None
        dw = grad_of(self.emb_w) if (wgrad and self.emb_w.requires_grad) else None
        with torch.cuda.device(qkv.device):
            _lib.check(_L().mdctgan_attention_abs_pos_bwd(qkv.data_ptr(), self.emb_h.data_ptr(), self.emb_w.data_ptr(), dout.data_ptr(),
                                                          dqkv.data_ptr(), _ptr(dh), _ptr(dw), B, H, W, self.heads, self.d, self.scale,
                                                          _stream(qkv)))
        G.add(self.qkv, dqkv)
