"""MDCT4 / IMDCT4 with the reference's constructor and forward signatures (models/mdct.py:359-489),
executed by the fused sm_100a kernels in libmdctgan_b200.so.  CUDA tensors only; no fallback."""
from __future__ import annotations

import torch

from .. import _lib

_PREC = {"fp64": _lib.F64, "float64": _lib.F64, torch.float64: _lib.F64,
         "fp32": _lib.F32, "float32": _lib.F32, torch.float32: _lib.F32,
         "mixed": _lib.MIXED}


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor, got device={t.device}; mdctgan_b200 has no CPU path")


class _TransformBase(torch.nn.Module):
    def __init__(self, n_fft, hop_length, win_length, window, center, pad_mode, device, precision):
        super().__init__()
        self.n_fft = n_fft
        self.pad_mode = pad_mode
        self.device = device
        self.hop_length = hop_length
        self.center = center
        if window is None:
            window = torch.ones
        if callable(window):
            self.win_length = int(win_length)
            window = window(self.win_length)
        else:
            self.win_length = len(window)
        assert self.win_length <= self.n_fft, "Window lenth %d should be no more than fft length %d" % (self.win_length, self.n_fft)
        assert self.hop_length <= self.win_length, "You hopped more than one frame"
        if not center or pad_mode != "constant":
            raise NotImplementedError("mdctgan_b200 implements the configuration the reference model uses: "
                                      "center=True, pad_mode='constant' (pix2pixHD_model.py:26-30)")
        self._window_host = window.detach().to("cpu", torch.float32).contiguous()
        self.register_buffer("window", self._window_host.clone(), persistent=False)
        if precision not in _PREC:
            raise ValueError(f"precision must be 'fp64', 'mixed' or 'fp32', got {precision!r}")
        self.precision = _PREC[precision]
        self._plans = {}

    def _plan(self, device: torch.device) -> _lib.Plan:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        p = self._plans.get(idx)
        if p is None:
            with torch.cuda.device(idx):
                p = _lib.Plan(self.n_fft, self.hop_length, self.win_length, self._window_host.numpy())
            self._plans[idx] = p
        return p

    @property
    def _dtype(self):
        return torch.float64 if self.precision == _lib.F64 else torch.float32


class MDCT4(_TransformBase):
    """signal [T] or [..., T] fp32 -> (coefficients [..., F, n_fft/2], frames).

    ``precision='fp64'`` (default) returns float64 like the reference (mdct.py:387-390,421-423);
    ``'mixed'`` runs the same fp64 butterflies on fp32 tensors (round trip 0.3 eps*peak, inside the 2-ulp bar);
    ``'fp32'`` is the all-fp32 fast flavour (round trip 2-2.8 eps*peak).  ``frames`` (windowed frames, mdct.py:410-412) is only
    materialised when ``return_frames=True``; otherwise an ``empty(1)`` placeholder like the reference.
    """

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window=None, center=True, pad_mode="constant",
                 device="cuda", precision="fp64") -> None:
        super().__init__(n_fft, hop_length, win_length, window, center, pad_mode, device, precision)

    def frame_count(self, signal: torch.Tensor) -> int:
        # len(signal) is dim 0: the sample count for 1-D input, the batch size for N-D input (mdct.py:394)
        return _lib.frame_count(signal.shape[-1], signal.shape[0], self.hop_length, self.win_length, self.center)

    @torch.no_grad()
    def forward(self, signal: torch.Tensor, return_frames: bool = False):
        _require_cuda(signal, "MDCT4.forward")
        lead = signal.shape[:-1]
        T = signal.shape[-1]
        F = self.frame_count(signal)
        x = signal.to(torch.float32).reshape(-1, T)
        if x.stride(-1) != 1:
            x = x.contiguous()
        B = x.shape[0]
        spec = torch.empty((B, F, self.n_fft // 2), dtype=self._dtype, device=x.device)
        if B and F:
            with torch.cuda.device(x.device):
                _lib.check(_lib.lib().mdctgan_mdct4_forward(self._plan(x.device).handle, x.data_ptr(), B, T, x.stride(0) if B > 1 else T,
                                                            F, spec.data_ptr(), F * (self.n_fft // 2), self.precision,
                                                            _stream_ptr(x.device)))
        spec = spec.reshape(*lead, F, self.n_fft // 2)
        if return_frames:
            frames = self.frames(signal, F)
        else:
            frames = torch.empty(1)
        return spec, frames

    def frames(self, signal: torch.Tensor, F: int | None = None) -> torch.Tensor:
        """Windowed fp32 frames [..., F, win] (mdct.py:403-412); diagnostic, not on the hot path."""
        F = self.frame_count(signal) if F is None else F
        need = (F - 1) * self.hop_length + self.win_length if F else 0
        x = torch.nn.functional.pad(signal, (self.hop_length, max(need - self.hop_length - signal.shape[-1], 0)))
        return x[..., :need].unfold(-1, self.win_length, self.hop_length) * self.window.to(signal.device)


class IMDCT4(_TransformBase):
    """coefficients [B, F, n_fft/2] -> (audio [B, 1, 1, (F-1)*hop (or out_length)], frames) (mdct.py:457-489)."""

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window=None, center=True, pad_mode="constant",
                 out_length=None, device="cuda", precision="fp64") -> None:
        super().__init__(n_fft, hop_length, win_length, window, center, pad_mode, device, precision)
        self.out_length = out_length

    @torch.no_grad()
    def forward(self, signal: torch.Tensor, return_frames: bool = False):
        assert signal.dim() == 3, "Only tensors shaped in BHW are supported, got tensor of shape %s" % (str(signal.size()))
        assert signal.size()[-1] == self.n_fft // 2, \
            "The last dim of input tensor should match the n_fft. Expected %d ,got %d" % (self.n_fft, signal.size()[-1])
        if return_frames:
            raise NotImplementedError("IMDCT4(return_frames=True): the fused kernel never materialises the windowed "
                                      "time frames (only dead code in the reference consumes them)")
        _require_cuda(signal, "IMDCT4.forward")
        B, F, N = signal.shape
        x = signal.to(self._dtype).contiguous()
        full = max(F - 1, 0) * self.hop_length
        out_len = full if self.out_length is None else min(full, max(int(self.out_length), 0))
        audio = torch.empty((B, 1, 1, out_len), dtype=self._dtype, device=x.device)
        if B and out_len:
            with torch.cuda.device(x.device):
                _lib.check(_lib.lib().mdctgan_imdct4_inverse(self._plan(x.device).handle, x.data_ptr(), B, F, F * N, audio.data_ptr(),
                                                             out_len, out_len, self.precision, _stream_ptr(x.device)))
        return audio, torch.zeros(1)


class FastMDCT4(MDCT4):
    """Import-compatible alias promised by the reference README (README.md:100): fp32 flavour of MDCT4.
    Like the reference's FastMDCT4 (mdct.py:596-628) it returns a 4-D [B, C, F, n_fft/2] fp32 tensor for
    [B, C, T] input; 1-D / 2-D inputs keep MDCT4's shapes."""

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window=None, center=True, pad_mode="constant",
                 device="cuda") -> None:
        super().__init__(n_fft, hop_length, win_length, window, center, pad_mode, device, precision="fp32")


class FastIMDCT4(IMDCT4):
    """fp32 flavour of IMDCT4 (reference: mdct.py:631-747, which is not runnable as shipped)."""

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window=None, center=True, pad_mode="constant",
                 out_length=None, device="cuda") -> None:
        super().__init__(n_fft, hop_length, win_length, window, center, pad_mode, out_length, device, precision="fp32")

    def forward(self, signal: torch.Tensor, return_frames: bool = False):
        if signal.dim() == 4:   # [B, C, F, N] as FastMDCT4 emits
            b, c, f, n = signal.shape
            audio, fr = super().forward(signal.reshape(b * c, f, n), return_frames)
            return audio.reshape(b, c, 1, -1), fr
        return super().forward(signal, return_frames)
