"""Audio <-> normalised MDCT spectrogram front / back end with the reference's names and signatures
(models/pix2pixHD_model.py:14-200), executed by the fused kernels of libmdctgan_b200.so:

  to_spectro  = MDCT4 + arcsinh/raw compress + abs-norm affine (+ optional second channel |s|*2+lo)
                in ONE kernel / one HBM pass   (reference: :32-81 -> mdct.py:392-425, normalize :83-125)
  to_audio    = denormalise + sinh expand + IMDCT4 + window + overlap-add + crop in ONE kernel
                (reference: :139-163 -> denormalize :127-137, mdct.py:457-489)

CUDA tensors only; there is no CPU / PyTorch fallback.
"""
from __future__ import annotations

from argparse import Namespace
from typing import Dict, Optional

import torch

from .. import _lib
from ..util.util import kbdwin
from .mdct import IMDCT4, MDCT4, _require_cuda, _stream_ptr


def default_audio_opt(**overrides) -> Namespace:
    """The `opt` fields Audio2MDCT reads, at the reference's defaults (options/audio_config.py:1-12,
    options/base_options.py:24-47,85-89, options/train_options.py:63-72)."""
    opt = Namespace(
        n_fft=512, hop_length=256, win_length=512, bins=128, segment_length=32512,
        lr_sampling_rate=12000, hr_sampling_rate=48000, sr_sampling_rate=48000,
        arcsinh_transform=True, arcsinh_gain=500.0, raw_mdct=False, explicit_encoding=False, alpha=0.6,
        min_value=1e-7, abs_norm=True, src_range=(-5.0, 5.0), norm_range=(0.0, 1.0),
        mask=False, mask_hr=False, fit_residual=False, gpu_ids=[0],
    )
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt


class Audio2MDCT(torch.nn.Module):
    """Drop-in for the reference's Audio2MDCT (pix2pixHD_model.py:14-200).

    Differences, all recorded in DESIGN.md: `pha` and the unused `mean/std/frames` entries are not
    materialised in arcsinh / raw mode (nothing downstream reads them); the
    reference's throw-away `randn` draw (:49-54) is skipped.  `precision='mixed'` (default) runs fp64
    butterflies on fp32 tensors with the fast fp32 compress / expand; `'fp32'` is the all-fp32 flavour;
    `'fp64'` runs fp64 butterflies + library asinh/sinh and returns fp64 audio like the reference.
    """

    def __init__(self, opt, device=None, precision: str = "mixed") -> None:
        super().__init__()
        for k, v in vars(opt).items():
            setattr(self, k, v)
        if device is None:
            if not len(self.gpu_ids):
                raise RuntimeError("Audio2MDCT: gpu_ids is empty; mdctgan_b200 has no CPU path")
            device = torch.device("cuda", self.gpu_ids[0])
        self.device = torch.device(device)
        self.up_ratio = self.hr_sampling_rate / self.lr_sampling_rate
        self.min_value = opt.min_value
        self.window = kbdwin(self.win_length)
        prec64 = precision in ("fp64", "float64", torch.float64)
        if not prec64 and precision not in ("mixed", "fp32", "float32", torch.float32):
            raise ValueError(f"Audio2MDCT: precision must be 'mixed', 'fp32' or 'fp64', got {precision!r}")
        sub = "fp64" if prec64 else ("mixed" if precision == "mixed" else "fp32")
        self.precision = {"fp64": _lib.F64, "mixed": _lib.MIXED, "fp32": _lib.F32}[sub]
        # raw-coefficient transforms keep the reference dtypes (fp64 in / out) in the fp64 flavour
        self._mdct = MDCT4(n_fft=self.n_fft, hop_length=self.hop_length, win_length=self.win_length, window=self.window,
                           device=self.device, precision=sub)
        self._imdct = IMDCT4(n_fft=self.n_fft, hop_length=self.hop_length, win_length=self.win_length, window=self.window,
                             device=self.device, precision=sub)
        # Which encoding (pix2pixHD_model.py:83-106, in the reference's order of precedence) and which path: the arcsinh / raw encodings
        # with --abs_norm (every shipped script) are fused into the transform kernels; dB, --explicit_encoding and the per-sample
        # min / max normalisation run as element-wise kernels around the raw transform (csrc/spectro_codec.cuh).
        if self.explicit_encoding:
            self._mode = _lib.MODE_EXPLICIT
        elif self.arcsinh_transform:
            self._mode = _lib.MODE_ARCSINH
        elif self.raw_mdct:
            self._mode = _lib.MODE_RAW
        else:
            self._mode = _lib.MODE_DB
        self._fused = self._mode in (_lib.MODE_ARCSINH, _lib.MODE_RAW) and bool(self.abs_norm)
        self._norm = _lib.NormSpec(self._mode if self._fused else _lib.MODE_RAW, float(self.arcsinh_gain), tuple(float(v) for v in self.src_range),
                                   tuple(float(v) for v in self.norm_range))
        self._cnorm = self._norm.c()
        self._minmax: Dict[int, tuple] = {}

    # ------------------------------------------------------------------ helpers
    def _src_minmax(self, device):
        key = device.index or 0
        if key not in self._minmax:
            lo = torch.tensor([self.src_range[0]], device=device, dtype=torch.float32)[None, None, None, :]
            hi = torch.tensor([self.src_range[1]], device=device, dtype=torch.float32)[None, None, None, :]
            self._minmax[key] = (lo, hi)
        return self._minmax[key]

    def _check_norm_param(self, norm_param) -> None:
        lo, hi = norm_param["min"], norm_param["max"]
        if torch.is_tensor(lo) and lo.numel() != 1:
            raise ValueError("to_audio: per-sample min / max parameters on an --abs_norm model")

    # ------------------------------------------------------------------ forward transform
    @torch.no_grad()
    def to_spectro(self, audio: torch.Tensor, mask: bool = False, mask_size: int = -1, channels: int = 1,
                   out: Optional[torch.Tensor] = None):
        """audio [B, T] (or [T]) fp32 -> (log_spectro fp32 [B, channels, F, n_fft/2], pha, norm_param).

        `channels=2` additionally writes the generator's second input channel `abs(s)*2 + norm_range[0]`
        (pix2pixHD_model.py:400-402) from the same kernel.
        """
        _require_cuda(audio, "Audio2MDCT.to_spectro")
        dim0 = audio.shape[0]   # len(signal): samples for 1-D input, batch size otherwise (mdct.py:394)
        orig = audio
        if audio.dim() == 1:
            audio = audio[None]
        x = audio.to(torch.float32)
        if x.dim() != 2:
            x = x.reshape(-1, x.shape[-1])
        if x.stride(-1) != 1:
            x = x.contiguous()
        B, T = x.shape
        nb = self.n_fft // 2
        F = _lib.frame_count(T, dim0, self.hop_length, self.win_length, True)
        if not self._fused:
            return self._to_spectro_generic(orig, B, F, nb, mask, mask_size, channels, out)
        if out is None:
            out = torch.empty((B, channels, F, nb), dtype=torch.float32, device=x.device)
        else:
            assert out.shape == (B, channels, F, nb) and out.dtype == torch.float32 and out.is_contiguous()
        if B and F:
            with torch.cuda.device(x.device):
                _lib.check(_lib.lib().mdctgan_audio2mdct_forward(
                    self._mdct._plan(x.device).handle, x.data_ptr(), B, T, x.stride(0) if B > 1 else T, F, self._cnorm,
                    out.data_ptr(), channels, channels * F * nb, F * nb, self.precision, _stream_ptr(x.device)))
        if mask:
            self._apply_mask(out, B, F, nb, mask_size, channels)
        lo, hi = self._src_minmax(x.device)
        return out, None, {"max": hi, "min": lo, "mean": None, "std": None, "frames": None}

    def forward(self, lr_audio: torch.Tensor, channels: int = 1):
        return self.to_spectro(lr_audio, mask=self.mask, channels=channels)

    def hr_forward(self, hr_audio: torch.Tensor, channels: int = 1):
        return self.to_spectro(hr_audio, mask=self.mask_hr, channels=channels,
                               mask_size=int(self.n_fft * (1 - self.sr_sampling_rate / self.hr_sampling_rate) // 2))

    # ------------------------------------------------------------------ inverse transform
    @torch.no_grad()
    def to_audio(self, log_spectro: torch.Tensor, norm_param=None, pha=None, out_length: Optional[int] = None):
        """normalised spectrogram [B, 1, F, N] (or [B, F, N]) fp32 -> audio [B, 1, 1, (F-1)*hop]."""
        _require_cuda(log_spectro, "Audio2MDCT.to_audio")
        if not self._fused:
            return self._to_audio_generic(log_spectro, norm_param, pha, out_length)
        if norm_param is not None:
            self._check_norm_param(norm_param)
        s = log_spectro
        if s.dim() == 4:
            assert s.shape[1] == 1, "to_audio expects a single-channel spectrogram"
            s = s[:, 0]
        assert s.dim() == 3 and s.shape[-1] == self.n_fft // 2
        s = s.to(torch.float32)
        if s.stride(-1) != 1 or s.stride(-2) != s.shape[-1]:
            s = s.contiguous()
        B, F, nb = s.shape
        full = max(F - 1, 0) * self.hop_length
        out_len = full if out_length is None else min(full, int(out_length))
        dt = torch.float64 if self.precision == _lib.F64 else torch.float32
        audio = torch.empty((B, 1, 1, out_len), dtype=dt, device=s.device)
        if B and out_len:
            with torch.cuda.device(s.device):
                _lib.check(_lib.lib().mdctgan_mdct2audio_inverse(
                    self._mdct._plan(s.device).handle, s.data_ptr(), B, F, s.stride(0) if B > 1 else F * nb, self._cnorm,
                    audio.data_ptr(), out_len, out_len, self.precision, _stream_ptr(s.device)))
        return audio

    # ------------------------------------------------------------------ secondary encodings (dB / explicit / per-sample min-max)
    def _apply_mask(self, out, B, F, nb, mask_size, channels):
        """pix2pixHD_model.py:57-80 on the 1-channel lr_spectro, then the second network channel |s|*2+lo from it (:400-402)."""
        if mask_size == -1:
            mask_size = int(nb * (1 - 1 / self.up_ratio))
        if mask_size <= 0:      # reference: mask_size == 0 would make an empty slice -> no-op
            return
        if self.fit_residual:
            out[:, 0, :, nb - mask_size:] = 0
        else:
            noise = torch.randn(B, 1, F, mask_size, device=out.device)
            out[:, 0:1, :, nb - mask_size:] = noise / (noise.max() - noise.min())
        if channels == 2:
            out[:, 1, :, nb - mask_size:] = out[:, 0, :, nb - mask_size:].abs() * 2 + float(self.norm_range[0])

    def _to_spectro_generic(self, audio, B, F, nb, mask, mask_size, channels, out):
        """dB / --explicit_encoding / no --abs_norm (pix2pixHD_model.py:32-125): raw MDCT4 launch -> spectro_encode (+ per-plane min /
        max) -> spectro_affine.  Returns (log_spectro fp32 [B, C, F, N], pha = sign(spectro) [* noise], norm_param) like the reference."""
        from ctypes import c_double, c_int, c_int64, c_void_p

        spec, _ = self._mdct(audio)                             # raw coefficients [B, F, N] (fp32; fp64 in the fp64 flavour)
        spec = spec.reshape(B, F, nb).contiguous()
        dev = spec.device
        C = 2 if self._mode == _lib.MODE_EXPLICIT else 1
        want2 = channels == 2 and C == 1
        enc = torch.empty((B, C, F, nb), dtype=torch.float64, device=dev)
        sign = torch.empty((B, 1, F, nb), dtype=torch.float32, device=dev)
        minmax = None if self.abs_norm else torch.empty((B, C, 2), dtype=torch.float32, device=dev)
        res = torch.empty((B, C, F, nb), dtype=torch.float32, device=dev)
        L = _lib.lib()
        L.mdctgan_spectro_encode.argtypes = [c_void_p, c_int, c_int64, c_int64, c_int, c_double, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p]
        L.mdctgan_spectro_affine.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_double, c_double, c_double, c_double, c_void_p, c_void_p]
        if B and F:
            with torch.cuda.device(dev):
                st = _stream_ptr(dev)
                _lib.check(L.mdctgan_spectro_encode(spec.data_ptr(), _lib.F64 if spec.dtype == torch.float64 else _lib.F32, B, F * nb, self._mode,
                                                    float(self.arcsinh_gain), float(self.alpha), float(self.min_value), enc.data_ptr(),
                                                    sign.data_ptr(), minmax.data_ptr() if minmax is not None else None, st))
                _lib.check(L.mdctgan_spectro_affine(enc.data_ptr(), B * C, F * nb, minmax.data_ptr() if minmax is not None else None,
                                                    float(self.src_range[0]), float(self.src_range[1]), float(self.norm_range[0]),
                                                    float(self.norm_range[1]), res.data_ptr(), st))
        if want2:                                               # the generator's second input channel |s|*2+lo (:400-402)
            two = torch.empty((B, 2, F, nb), dtype=torch.float32, device=dev) if out is None else out
            two[:, 0:1] = res
            two[:, 1:2] = res.abs() * 2 + float(self.norm_range[0])
            res = two
        elif out is not None:
            out.copy_(res)
            res = out
        pha = sign
        if not self.explicit_encoding:                          # :49-54: sign * min-max-scaled noise
            noise = torch.randn(sign.shape, device=dev)
            pha = sign * ((noise - noise.min()) / (noise.max() - noise.min()))
        if mask:
            self._apply_mask(res, B, F, nb, mask_size, 2 if want2 else 1)
        if minmax is not None:
            lo, hi = minmax[:, :, 0, None, None].contiguous(), minmax[:, :, 1, None, None].contiguous()
        else:
            lo, hi = self._src_minmax(dev)
        return res, pha, {"max": hi, "min": lo, "mean": None, "std": None, "frames": None}

    def _to_audio_generic(self, log_spectro, norm_param, pha, out_length):
        """pix2pixHD_model.py:127-163 for the secondary encodings: spectro_decode (denormalise, dB -> amplitude, channel recombination /
        phase product with the random pseudo phase of the frames beyond F / up_ratio) -> raw IMDCT4 launch."""
        from ctypes import c_double, c_int, c_int64, c_void_p

        s = log_spectro.to(torch.float32)
        C = 2 if self._mode == _lib.MODE_EXPLICIT else 1
        if s.dim() == 3:
            s = s[:, None]
        assert s.dim() == 4 and s.shape[1] == C and s.shape[-1] == self.n_fft // 2, f"to_audio expects [B, {C}, F, {self.n_fft // 2}]"
        s = s.contiguous()
        B, _, F, nb = s.shape
        dev = s.device
        minmax = None
        lo_v, hi_v = float(self.src_range[0]), float(self.src_range[1])
        if norm_param is not None:
            lo, hi = torch.as_tensor(norm_param["min"]), torch.as_tensor(norm_param["max"])
            if lo.numel() == 1:
                lo_v, hi_v = float(lo.reshape(-1)[0]), float(hi.reshape(-1)[0])
            else:
                minmax = torch.stack((lo.to(dev, torch.float32).reshape(B, C), hi.to(dev, torch.float32).reshape(B, C)), dim=-1).contiguous()
        mult = None
        if self._mode == _lib.MODE_DB and self.up_ratio > 1:    # :150-157 (the reference concatenates along the FRAME axis)
            if pha is None:
                raise ValueError("to_audio: the dB encoding needs the `pha` returned by to_spectro")
            ph = pha.to(dev, torch.float32).reshape(B, F, nb)
            keep = int(F * (1 / self.up_ratio))
            pseudo = (2 * torch.randint(low=0, high=2, size=(B, F - keep, nb), device=dev) - 1).to(torch.float32)
            mult = torch.cat((ph[:, :keep], pseudo), dim=1).contiguous()
        raw = torch.empty((B, F, nb), dtype=torch.float64, device=dev)
        L = _lib.lib()
        L.mdctgan_spectro_decode.argtypes = [c_void_p, c_int64, c_int64, c_int, c_double, c_double, c_double, c_void_p, c_double, c_double,
                                             c_double, c_double, c_void_p, c_void_p, c_void_p]
        if B and F:
            with torch.cuda.device(dev):
                _lib.check(L.mdctgan_spectro_decode(s.data_ptr(), B, F * nb, self._mode, float(self.arcsinh_gain), float(self.alpha),
                                                    float(self.min_value), minmax.data_ptr() if minmax is not None else None, lo_v, hi_v,
                                                    float(self.norm_range[0]), float(self.norm_range[1]),
                                                    mult.data_ptr() if mult is not None else None, raw.data_ptr(), _stream_ptr(dev)))
        if self.precision != _lib.F64:
            raw = raw.to(torch.float32)
        audio, _ = self._imdct(raw)
        return audio if out_length is None else audio[..., :int(out_length)]

    # ------------------------------------------------------------------ un-fused pieces (API parity)
    @torch.no_grad()
    def normalize(self, spectro: torch.Tensor):
        """pix2pixHD_model.py:83-125 on an existing spectrogram (arcsinh / raw branch, abs_norm): returns
        (log_spectro, audio_max, audio_min, mean, std) with the spectrogram's dtype (the reference's is fp64); `mean` / `std`
        are None (computed and never consumed in the reference)."""
        from ctypes import c_int, c_int64, c_void_p, POINTER

        _require_cuda(spectro, "Audio2MDCT.normalize")
        x = spectro.contiguous()
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float32)
        y = torch.empty_like(x)
        L = _lib.lib()
        L.mdctgan_spectro_normalize.argtypes = [c_void_p, c_void_p, c_int64, POINTER(_lib._Norm), c_int, c_void_p]
        if x.numel():
            with torch.cuda.device(x.device):
                _lib.check(L.mdctgan_spectro_normalize(x.data_ptr(), y.data_ptr(), x.numel(), self._cnorm,
                                                       _lib.F64 if x.dtype == torch.float64 else _lib.F32, _stream_ptr(x.device)))
        lo, hi = self._src_minmax(x.device)
        return y, hi, lo, None, None

    @torch.no_grad()
    def denormalize(self, log_spectro: torch.Tensor, min: torch.Tensor, max: torch.Tensor):
        """pix2pixHD_model.py:127-137: normalised spectrogram -> fp64 MDCT coefficients (arcsinh / raw branch)."""
        from ctypes import c_int, c_int64, c_void_p, POINTER

        _require_cuda(log_spectro, "Audio2MDCT.denormalize")
        self._check_norm_param({"min": min, "max": max})
        lo_v, hi_v = float(torch.as_tensor(min).reshape(-1)[0]), float(torch.as_tensor(max).reshape(-1)[0])
        cn = _lib.NormSpec(self._norm.mode, self._norm.gain, (lo_v, hi_v), self._norm.norm_range).c()
        x = log_spectro.contiguous()
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float32)
        y = torch.empty(x.shape, dtype=torch.float64, device=x.device)
        L = _lib.lib()
        L.mdctgan_spectro_denormalize.argtypes = [c_void_p, c_void_p, c_int64, POINTER(_lib._Norm), c_int, c_void_p]
        if x.numel():
            with torch.cuda.device(x.device):
                _lib.check(L.mdctgan_spectro_denormalize(x.data_ptr(), y.data_ptr(), x.numel(), cn,
                                                         _lib.F64 if x.dtype == torch.float64 else _lib.F32, _stream_ptr(x.device)))
        return y


# =====================================================================================================
# Model facade (reference: models/pix2pixHD_model.py:203-714, models/base_model.py)
# =====================================================================================================
import os  # noqa: E402

from . import networks  # noqa: E402
from .. import nn_ops as _ops  # noqa: E402


# opt-in: split the gradient exchange into an early bucket (overlapped with the generator sweep) + the rest (3 collectives per step)
BUCKETED_ALLREDUCE = os.environ.get("MDCTGAN_BUCKETED_ALLREDUCE", "0") == "1"
# Pipelined update (default): per-bucket [all-reduce ->] Adam -> weight images on an update stream as soon as the bucket's gradients
# are complete, overlapping the rest of the backward sweep.  "0": one all-reduce + whole-network Adam + packing after the sweeps.
PIPELINED_UPDATE = os.environ.get("MDCTGAN_PIPELINED_UPDATE", "1") != "0"
COMM_STREAM = os.environ.get("MDCTGAN_COMM_STREAM", "1") != "0"      # per-bucket collectives on their own stream (A/B switch)


class BaseModel(torch.nn.Module):
    def name(self):
        return "BaseModel"

    def initialize(self, opt):
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        self.isTrain = opt.isTrain
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        if not len(self.gpu_ids):
            raise RuntimeError("mdctgan_b200 needs --gpu_ids >= 0: there is no CPU path")
        self.device = torch.device("cuda", self.gpu_ids[0])

    def save_network(self, network, network_label, epoch_label, gpu_ids=None):
        """`<checkpoints_dir>/<name>/<epoch>_net_<label>.pth` = plain state_dict (base_model.py:43-46)."""
        os.makedirs(self.save_dir, exist_ok=True)
        # parameters are views into one flat buffer (optim.FlatBucket): clone so the file holds plain per-tensor storages
        torch.save({k: v.detach().contiguous().clone() for k, v in network.state_dict().items()},
                   os.path.join(self.save_dir, "%s_net_%s.pth" % (epoch_label, network_label)))

    def load_network(self, network, network_label, epoch_label, save_dir=""):
        """Three-level fallback of the reference (base_model.py:49-111): strict load -> keys present in the
        model -> `--param_key_map` remap of the second key component."""
        save_path = os.path.join(save_dir or self.save_dir, "%s_net_%s.pth" % (epoch_label, network_label))
        if not os.path.isfile(save_path):
            if network_label == "G":
                raise FileNotFoundError("Generator must exist! (%s)" % save_path)
            print("%s not exists yet!" % save_path)
            return
        pretrained = torch.load(save_path, map_location="cpu")
        try:
            network.load_state_dict(pretrained)
        except RuntimeError:
            model_dict = network.state_dict()
            try:
                network.load_state_dict({k: v for k, v in pretrained.items() if k in model_dict})
            except RuntimeError:
                module_map = getattr(self.opt, "param_key_map", None) or {}
                for name, param in pretrained.items():
                    if name not in model_dict or param.size() != model_dict[name].size():
                        parts = name.split(".")
                        key = parts[0] + "." + parts[1] if len(parts) > 1 else name
                        if key not in module_map:
                            continue
                        parts[1] = module_map[key]
                        name = ".".join(parts)
                    model_dict[name] = param
                network.load_state_dict(model_dict)
        network.to(self.device)


class Pix2PixHDModel(BaseModel):
    """Drop-in for the reference facade (models/pix2pixHD_model.py:203-714): `inference`, `forward`, `_forward` (the
    four losses as differentiable 0-dim tensors), `optimizer_G` / `optimizer_D`, `save`, `update_learning_rate`;
    plus `train_step` = the whole train.py:160-202 iteration as one call (what bench.py times)."""

    loss_names = ["G_GAN", "G_GAN_Feat", "D_real", "D_fake"]

    def name(self):
        return "Pix2PixHDModel"

    def initialize(self, opt):
        BaseModel.initialize(self, opt)
        for k, v in vars(opt).items():
            setattr(self, k, v)
        self.isTrain = opt.isTrain
        input_nc = opt.label_nc if getattr(opt, "label_nc", 0) != 0 else opt.input_nc
        self.preprocess = Audio2MDCT(opt, device=self.device, precision=getattr(opt, "mdct_precision", "mixed"))
        self.freeze = opt.freeze_g_d or opt.freeze_g_u or opt.freeze_l_d or opt.freeze_l_u
        self.netG = networks.define_G(input_nc, opt.output_nc, opt.ngf, opt.netG, opt.n_downsample_global, opt.n_blocks_global,
                                      opt.n_local_enhancers, opt.n_blocks_local, opt.norm, gpu_ids=self.gpu_ids,
                                      upsample_type=opt.upsample_type, downsample_type=opt.downsample_type,
                                      input_size=(opt.bins, opt.n_fft // 2), n_attn_g=opt.n_blocks_attn_g, n_attn_l=opt.n_blocks_attn_l,
                                      proj_factor_g=opt.proj_factor_g, heads_g=opt.heads_g, dim_head_g=opt.dim_head_g,
                                      proj_factor_l=opt.proj_factor_l, heads_l=opt.heads_l, dim_head_l=opt.dim_head_l)
        self.netG.set_freeze(opt.freeze_g_d, opt.freeze_g_u, opt.freeze_l_d, opt.freeze_l_u)
        if self.isTrain:
            self.netD = networks.define_D(input_nc + opt.output_nc, opt.ndf, opt.n_layers_D, opt.norm, opt.no_lsgan, opt.num_D,
                                          not opt.no_ganFeat_loss, gpu_ids=self.gpu_ids)
        if not self.isTrain or opt.continue_train or opt.load_pretrain:
            pretrained_path = "" if not self.isTrain else opt.load_pretrain
            self.load_network(self.netG, "G", opt.which_epoch, pretrained_path)
            if self.isTrain:
                self.load_network(self.netD, "D", opt.which_epoch, pretrained_path)
        self._two_channel = bool(self.abs_spectro and self.arcsinh_transform)
        if self.isTrain:
            if opt.pool_size > 0 and len(self.gpu_ids) > 1:
                raise NotImplementedError("Fake Pool Not Implemented for MultiGPU")
            if opt.pool_size > 0:
                raise NotImplementedError("--pool_size > 0 (image history buffer) is out of scope: identity at the reference default 0")
            from ..optim import FlatBucket, FusedAdam

            self.old_lr = opt.lr
            self.limit_aux_loss = False
            self.loss_names = [n for n in ["G_GAN", "G_GAN_Feat", "D_real", "D_fake"] if not (n == "G_GAN_Feat" and opt.no_ganFeat_loss)]
            self.netG.to(self.device)
            self.netD.to(self.device)
            # every parameter becomes a view into one flat buffer per network; gradients likewise (the all-reduce bucket)
            nG, nD = FlatBucket.padded_numel(self.netG), FlatBucket.padded_numel(self.netD)
            self.grad_all = torch.zeros(nG + nD, dtype=torch.float32, device=self.device)     # [grad_G | grad_D]: THE all-reduce bucket
            self.bucket_G = FlatBucket(self.netG, self.grad_all[:nG])
            self.bucket_D = FlatBucket(self.netD, self.grad_all[nG:])
            graph_safe = bool(getattr(opt, "graph_safe_adam", True))
            params_G = None
            if opt.niter_fix_global > 0:      # only the local enhancer trains at first (pix2pixHD_model.py:333-347)
                prefix = "model" + str(opt.n_local_enhancers)
                params_G = [p for k, p in self.netG.named_parameters() if k.startswith(prefix)]
                print("------------- Only training the local enhancer network (for %d epochs) ------------" % opt.niter_fix_global)
            self.optimizer_G = FusedAdam(self.bucket_G, lr=opt.lr, betas=(opt.beta1, 0.999), graph_safe=graph_safe, params=params_G)
            self.optimizer_D = FusedAdam(self.bucket_D, lr=opt.lr, betas=(opt.beta1, 0.999), graph_safe=graph_safe)
            self._graph = None
            # the reference-shaped API (_forward / loss.backward()) replays captured segments after a short eager warm-up (runtime.GraphedAPI)
            from ..runtime import GraphedAPI

            self._graph_api = GraphedAPI(self) if os.environ.get("MDCTGAN_GRAPH_API", "1") != "0" else None
            # update buckets: contiguous flat ranges of parameters whose Adam step and kernel-side weight images are issued as soon as
            # their gradients are complete (while the rest of the backward sweep still runs); one weight packer per bucket
            self._plan_update_buckets()
            self._packed_at = None

    # ---- generator graph ---------------------------------------------------------------------------------
    def _lr_input(self, lr_audio):
        """(lr_spectro [B,1,F,N], generator input [B,1|2,F,N]) from ONE fused MDCT launch (second channel
        |s|*2+lo written by the same kernel, pix2pixHD_model.py:400-402,624-628)."""
        ch = 2 if self._two_channel else 1
        lr_input, lr_pha, lr_norm_param = self.preprocess.forward(lr_audio, channels=ch)
        return lr_input[:, :1], lr_input, lr_pha, lr_norm_param

    def forward(self, lr_audio, hr_audio):
        """Generator half of the training graph (pix2pixHD_model.py:394-414), no autograd in this round."""
        if getattr(self, "packer", None) is not None:
            self._refresh_weight_images()
        lr_spectro, lr_input, lr_pha, lr_norm_param = self._lr_input(lr_audio)
        hr_spectro, hr_pha, hr_norm_param = self.preprocess.hr_forward(hr_audio)
        sr_spectro = self.netG.forward(lr_input)
        if self.fit_residual:
            sr_spectro = _ops.residual_scale_add(sr_spectro, lr_spectro, 0, 1.0)
        return sr_spectro, None, hr_spectro, hr_pha, hr_norm_param, lr_spectro, lr_pha, lr_norm_param

    def _pack_key(self):
        """What the kernel-side weight images were derived from: the optimisers (identity and step counts) and the version counter of
        EVERY parameter (load_state_dict and in-place edits bump them; FusedAdam.step writes through raw pointers, hence the step counts)."""
        return (id(self.optimizer_G), id(self.optimizer_D), self.optimizer_G.step_count, self.optimizer_D.step_count,
                sum(p._version for p in self.bucket_G.params), sum(p._version for p in self.bucket_D.params))

    def _refresh_weight_images(self):
        """Re-pack after optimiser steps / load_state_dict / in-place parameter edits."""
        key = self._pack_key()
        if key != self._packed_at:
            self.packer.refresh()
            self._packed_at = key

    def _forward(self, lr_audio, hr_audio, infer=False):
        """pix2pixHD_model.py:416-616: returns [[G_GAN, G_GAN_Feat, D_real, D_fake] (loss_filter order), sr_spectro or
        None].  The losses are 0-dim CUDA tensors; `loss_G.backward()` / `loss_D.backward()` (train.py:185-201) run the
        generator / discriminator backward sweeps of train_ops.GanGraph, which write straight into the flat gradient
        buckets behind `optimizer_G` / `optimizer_D`."""
        from .. import train_ops as T

        if not self.isTrain:
            raise RuntimeError("_forward needs a training model (opt.isTrain)")
        self._refresh_weight_images()
        graph = self._graph_api.forward(lr_audio, hr_audio) if self._graph_api is not None else None
        if graph is None:                                 # eager (warm-up iterations of a batch shape, or MDCTGAN_GRAPH_API=0)
            graph = T.GanGraph(self)
            with _ops.stats_pass(self.device):
                graph.forward(lr_audio, hr_audio)
        self._graph = graph
        # autograd anchors: any parameter that requires grad (the first one may be frozen: --freeze_g_d, niter_fix_global)
        anchor_g = next((p for p in self.bucket_G.params if p.requires_grad), None)
        anchor_d = next((p for p in self.bucket_D.params if p.requires_grad), None)
        if anchor_g is None or anchor_d is None:
            raise RuntimeError("_forward: every parameter of netG (or netD) is frozen; nothing to differentiate")
        g_gan, g_feat, d_real, d_fake = T.loss_tensors(graph, anchor_g, anchor_d)
        losses = [g_gan] + ([] if self.no_ganFeat_loss else [g_feat]) + [d_real, d_fake]
        return [losses, None if not infer else graph.sr_spectro]

    def train_step(self, lr_audio, hr_audio, world_size: int = 1, all_reduce=None):
        if not _ops.STREAM_PRIORITY:
            return self._train_step(lr_audio, hr_audio, world_size, all_reduce)
        cur = torch.cuda.current_stream(self.device)      # the dependent chain on a high-priority stream (nn_ops.STREAM_PRIORITY)
        hi = _ops.aux_stream(self.device, "main_hi")
        hi.wait_stream(cur)
        with torch.cuda.stream(hi):
            out = self._train_step(lr_audio, hr_audio, world_size, all_reduce)
        cur.wait_stream(hi)
        return out

    def _train_step(self, lr_audio, hr_audio, world_size: int = 1, all_reduce=None):
        """One whole iteration of train.py:160-202 without autograd in the loop: forward, generator sweep,
        discriminator sweep, [ONE all-reduce over the two flat gradient buckets], both Adam steps.  Mathematically the
        reference order (G step before D backward) because loss_D never depends on the updated G.
        Returns the 4-element fp32 loss vector (device)."""
        from .. import train_ops as T

        graph = T.GanGraph(self)
        self._refresh_weight_images()
        pipelined = PIPELINED_UPDATE and not BUCKETED_ALLREDUCE and _ops.SIDE_STREAM_WGRAD
        with _ops.stats_pass(self.device):
            self.grad_all.zero_()                          # one memset for both buckets
            self.optimizer_G.grad_scale = self.optimizer_D.grad_scale = 1.0 / world_size
            self.optimizer_G.begin_step()                  # step counters first: the bucket updates below run on their own stream
            self.optimizer_D.begin_step()
            if pipelined:
                upd = _ops.aux_stream(self.device, "update")
                upd.wait_stream(torch.cuda.current_stream(self.device))
                for bk in self._buckets:
                    bk.done = False
                    if bk.mark is not None and bk.net == "G":
                        _ops._marks[id(bk.mark)] = (lambda _bk=bk: self._flush_bucket(_bk, all_reduce, None))
            losses = graph.forward(lr_audio, hr_audio)
            # the two sweeps only read the tapes and write disjoint gradients: the discriminator sweep runs on its own stream,
            # concurrently with the generator sweep; weight-gradient kernels of both go to the side stream
            half = self._half_scalar()
            main = torch.cuda.current_stream(self.device)
            # bucketed exchange (opt-in): all-reduce the early-complete half of the generator gradients on a communication stream
            # while the rest of the generator sweep still runs
            eb = self._early_bucket() if (all_reduce is not None and BUCKETED_ALLREDUCE) else None
            comm = None
            if eb is not None:
                lo_b, hi_b, trig = eb
                comm = _ops.aux_stream(self.device, "comm")

                def _hook(stream, _lo=lo_b, _hi=hi_b):
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    comm.wait_event(ev)
                    with torch.cuda.stream(comm):
                        all_reduce(self.grad_all[_lo:_hi])

                _ops._wgrad_hooks[id(trig)] = _hook
            if _ops.PARALLEL_BRANCHES:
                sD = _ops.aux_stream(self.device, "sweep_D")
                sD.wait_stream(main)
                with torch.cuda.stream(sD):
                    graph.backward_D(half, half, join=False)
                    ev_sD = torch.cuda.Event()
                    ev_sD.record(sD)
                after_D = None
                if pipelined:
                    # the discriminator's update overlaps the generator sweep -- but only once that sweep has passed through the
                    # discriminator itself (it back-propagates the generator loss through the OLD discriminator weights)
                    def after_D():
                        evs = [ev_sD]
                        for st_ in [main] + _ops.side_streams(self.device):
                            e_ = torch.cuda.Event()
                            e_.record(st_)
                            evs.append(e_)
                        for bk in self._buckets:
                            if bk.net == "D":
                                self._flush_bucket(bk, all_reduce, evs)
                graph.backward_G(join=False, after_D=after_D)
                main.wait_stream(sD)
            else:
                graph.backward_G(join=False)
                graph.backward_D(half, half, join=False)
            _ops.join_side_work(self.device)              # the exchange / optimiser read the gradients from here on
            if pipelined:
                _ops._marks.clear()
                for bk in self._buckets:                   # whatever no tape mark flushed (frozen trunks, unmarked leftovers)
                    if not bk.done:
                        self._flush_bucket(bk, all_reduce, None)
                main.wait_stream(upd)
            elif eb is not None:
                _ops._wgrad_hooks.pop(id(trig), None)
                main.wait_stream(comm)
                all_reduce(self.grad_all[:lo_b])
                all_reduce(self.grad_all[hi_b:])
            elif all_reduce is not None:
                all_reduce(self.grad_all)                  # ONE collective per step
            if not pipelined:
                for o in (self.optimizer_G, self.optimizer_D):
                    o.step_range(0, o.bucket.numel)
                self.packer.refresh()                      # the next forward (and a CUDA-graph replay) sees the updated weights
            self.optimizer_G.end_step()
            self.optimizer_D.end_step()
            self._packed_at = self._pack_key()
        graph.release()
        return losses

    def _plan_update_buckets(self):
        """Update buckets of the pipelined train step.  Generator: the top-level children of its sequential trunks (`model`, and
        `model1_1` / `model1_2` of a LocalEnhancer), grouped in backward order into contiguous flat ranges of >= 1/8 of the
        parameters; a tape mark in front of a bucket's first child fires when the backward sweep has passed all of it
        (nn_ops._MarkOp).  Discriminator: one bucket, flushed when its sweep has been issued.  Each bucket owns the weight packer
        of its layers."""
        from types import SimpleNamespace

        from ..packing import WeightPacker

        nG = self.bucket_G.numel
        offs = {id(p): (o, o + (p.numel() + 3) // 4 * 4) for p, o in zip(self.bucket_G.params, self.bucket_G.offsets)}
        buckets, covered = [], 0
        target = max(nG // 8, 1 << 20)
        trunks = [m for m in (getattr(self.netG, n, None) for n in ("model", "model1_1", "model1_2")) if isinstance(m, torch.nn.Sequential)]
        for trunk in trunks:
            kids = [c for c in trunk if next(c.parameters(), None) is not None]
            group, size = [], 0
            for c in reversed(kids):
                group.insert(0, c)
                size += sum(p.numel() for p in c.parameters())
                if size >= target or c is kids[0]:
                    ps = [p for m in group for p in m.parameters()]
                    lo, hi = min(offs[id(p)][0] for p in ps), max(offs[id(p)][1] for p in ps)
                    if hi - lo == sum(offs[id(p)][1] - offs[id(p)][0] for p in ps):      # contiguous in the flat buffer
                        buckets.append(SimpleNamespace(net="G", lo=lo, hi=hi, glo=lo, ghi=hi, mark=group[0], modules=list(group), done=False))
                        covered += hi - lo
                    group, size = [], 0
        if covered != nG or not buckets:          # unknown generator structure: one bucket, flushed after the sweeps
            buckets = [SimpleNamespace(net="G", lo=0, hi=nG, glo=0, ghi=nG, mark=None, modules=[self.netG], done=False)]
        nD = self.bucket_D.numel
        buckets.append(SimpleNamespace(net="D", lo=0, hi=nD, glo=nG, ghi=nG + nD, mark=None, modules=[self.netD], done=False))
        for bk in buckets:
            has_conv = any(isinstance(m, (networks.Conv2d, networks.ConvTranspose2d)) for mod in bk.modules for m in mod.modules())
            bk.packer = WeightPacker(*bk.modules) if has_conv else None
        self._buckets = buckets
        self.packer = SimpleNamespace(refresh=lambda: [bk.packer.refresh() for bk in self._buckets if bk.packer is not None],
                                      parts=[bk.packer for bk in buckets if bk.packer is not None])

    def _flush_bucket(self, bk, all_reduce, events):
        """[all-reduce ->] Adam -> weight images of one update bucket on the update stream, after everything enqueued so far on the
        current stream and on the weight-gradient side stream (`events` = None), or after `events`."""
        if bk.done:
            return
        bk.done = True
        upd = _ops.aux_stream(self.device, "update")
        cur = torch.cuda.current_stream(self.device)
        if events is None:
            events = []
            for st in [cur] + _ops.side_streams(self.device):
                ev = torch.cuda.Event()
                ev.record(st)
                events.append(ev)
        if all_reduce is not None and COMM_STREAM:
            # the exchange of bucket i+1 overlaps the Adam / weight-image launches of bucket i: collectives on their own stream (same order
            # on every rank: the bucket order), the update stream only waits for its bucket's collective
            comm = _ops.aux_stream(self.device, "comm")
            for ev in events:
                comm.wait_event(ev)
            with torch.cuda.stream(comm):
                all_reduce(self.grad_all[bk.glo:bk.ghi])
                ev_c = torch.cuda.Event()
                ev_c.record(comm)
            events, all_reduce = [ev_c], None
        for ev in events:
            upd.wait_event(ev)
        opt_ = self.optimizer_G if bk.net == "G" else self.optimizer_D
        with torch.cuda.stream(upd):
            if all_reduce is not None:
                all_reduce(self.grad_all[bk.glo:bk.ghi])
            opt_.step_range(bk.lo, bk.hi)
            if bk.packer is not None:
                bk.packer.refresh()

    def _early_bucket(self):
        """(lo, hi, trigger module) of the flat gradient range that is complete long before the end of the generator sweep: the
        second half of the global trunk's residual blocks and everything after them in the trunk (back-propagated first; ~45 % of
        all gradient bytes in cfg4).  The trigger is the convolution whose weight gradient is enqueued last in that range."""
        if hasattr(self, "_eb"):
            return self._eb
        self._eb = None
        trunk = getattr(self.netG, "model", None)
        if trunk is not None:
            rbs = [m for m in trunk if isinstance(m, networks.ResnetBlock)]
            tparams = list(trunk.parameters())
            if len(rbs) >= 2 and tparams:
                first = rbs[len(rbs) // 2]
                b = self.bucket_G
                idx = {id(p): i for i, p in enumerate(b.params)}
                i0, i1 = idx[id(first.conv_block[1].weight)], idx[id(tparams[-1])]
                if all(id(p) in {id(q) for q in tparams} for p in b.params[i0:i1 + 1]):       # contiguous run of trunk parameters
                    lo, hi = b.offsets[i0], b.offsets[i1] + (b.params[i1].numel() + 3) // 4 * 4
                    self._eb = (lo, hi, first.conv_block[1])
        return self._eb

    def _half_scalar(self):
        if not hasattr(self, "_half"):
            self._half = torch.full((), 0.5, dtype=torch.float32, device=self.device)
        return self._half

    def update_fixed_params(self):
        """After --niter_fix_global epochs also finetune the global generator: a fresh Adam over all of G, like the reference
        (pix2pixHD_model.py:654-662)."""
        from ..optim import FusedAdam

        self.optimizer_G = FusedAdam(self.bucket_G, lr=self.lr, betas=(self.beta1, 0.999), graph_safe=self.optimizer_G.step_dev is not None)
        if getattr(self, "verbose", False):
            print("------------ Now also finetuning global generator -----------")

    def update_learning_rate(self):
        """Linear decay, pix2pixHD_model.py:664-673."""
        lrd = self.lr / self.niter_decay
        lr = self.old_lr - lrd
        for opt_ in (self.optimizer_D, self.optimizer_G):
            for g in opt_.param_groups:
                g["lr"] = lr
        if getattr(self, "verbose", False):
            print("update learning rate: %f -> %f" % (self.old_lr, lr))
        self.old_lr = lr

    def get_current_visuals(self):
        """Device references of the last step (the reference copies three tensors to the host EVERY step,
        pix2pixHD_model.py:569-613; here they are materialised only when asked for)."""
        g = getattr(self, "_graph", None)
        if g is None:
            return {}
        return {"lable_spectro": g.lr_spectro[0, 0].detach().cpu().numpy(), "generated_spectro": g.sr_spectro[0, 0].detach().cpu().numpy(),
                "real_spectro": g.hr_spectro[0, 0].detach().cpu().numpy()}

    @torch.no_grad()
    def inference(self, lr_audio):
        """pix2pixHD_model.py:618-638 -> (sr_spectro, sr_audio, lr_pha, lr_norm_param, lr_spectro)."""
        if getattr(self, "packer", None) is not None:
            self._refresh_weight_images()
        lr_spectro, lr_input, lr_pha, lr_norm_param = self._lr_input(lr_audio)
        sr_spectro = self.netG.forward(lr_input)
        if self.fit_residual:
            lr_part = int(sr_spectro.size(-1) / self.preprocess.up_ratio)
            sr_spectro = _ops.residual_scale_add(sr_spectro, lr_spectro, lr_part, 1e-3)
        sr_audio = self.preprocess.to_audio(sr_spectro, lr_norm_param, lr_pha)
        return sr_spectro, sr_audio, lr_pha, lr_norm_param, lr_spectro

    def save(self, which_epoch):
        self.save_network(self.netG, "G", which_epoch, self.gpu_ids)
        if hasattr(self, "netD"):
            self.save_network(self.netD, "D", which_epoch, self.gpu_ids)


class InferenceModel(Pix2PixHDModel):
    def forward(self, lr_audio):
        return self.inference(lr_audio)
