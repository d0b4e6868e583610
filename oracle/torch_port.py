"""Oracle (CPU, torch) -- the reference's OWN formulation of the transform path, op for op, on torch-CPU.
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py for who may import it).

`oracle/mdct_oracle.py` is the numpy restatement used as the parity checker.  This module restates the
same functions with the torch ops the reference itself issues (pad -> unfold -> window multiply ->
complex128 pre-twiddle -> 512-point torch.fft.fft -> slice -> post-twiddle -> real; and the inverse with
fold), so that timing it on the host cores measures what the reference's torch-CPU path costs
(bench.py `cpu_baseline` and `--impl reference`; kind = "port", because /root/reference itself cannot
travel to the GPU box).  Citations are into /root/reference:

  MDCT4Port        models/mdct.py:365-425
  IMDCT4Port       models/mdct.py:429-489
  Audio2MDCTPort   models/pix2pixHD_model.py:32-47,81 (to_spectro), :96-123 (normalize, arcsinh+abs_norm),
                   :127-133 (denormalize), :139-163 (to_audio)

Pinned against tests/golden/mdct_golden.npz (reference outputs) by tests/test_oracle_golden.py.
"""
import math

import torch
import torch.nn.functional as F


def kbdwin(n: int, beta: float = 12.0) -> torch.Tensor:   # util/util.py:179-186
    k = torch.kaiser_window(window_length=n // 2 + 1, beta=beta * math.pi, periodic=False)
    half = torch.sqrt(torch.cumsum(k, dim=0) / k.sum())[:-1]
    return torch.cat((half, half.flip(0)))


class MDCT4Port:
    def __init__(self, n_fft=512, hop=256, window=None, device="cpu"):
        self.n, self.hop = n_fft, hop
        self.w = (kbdwin(n_fft) if window is None else window).to(device)
        self.win = len(self.w)
        m = torch.arange(0, n_fft, dtype=torch.float64)
        self.pre = torch.exp(-1j * torch.pi / n_fft * m).to(device)
        k = torch.arange(1, n_fft, 2, dtype=torch.float64)
        self.post = torch.exp(-1j * (torch.pi / (2 * n_fft) + torch.pi / 4) * k).to(device)

    def __call__(self, x: torch.Tensor, return_frames: bool = False):
        start = self.hop
        extra = int(len(x)) % self.hop           # len() is dim 0 (reference quirk)
        end = start + (self.hop - extra if extra else 0)
        fr = F.pad(x, (start, end)).unfold(-1, self.win, self.hop) * self.w
        frames = fr.clone() if return_frames else torch.empty(1)
        if self.n > self.win:
            fr = F.pad(fr, (0, self.n - self.win))
        z = torch.fft.fft(fr * self.pre)[..., : self.n // 2]
        return torch.real(self.post * z), frames


class IMDCT4Port:
    def __init__(self, n_fft=512, hop=256, window=None, out_length=None, device="cpu"):
        self.n, self.hop, self.out_length = n_fft, hop, out_length
        self.w = (kbdwin(n_fft) if window is None else window).to(device)
        self.win = len(self.w)
        k = torch.arange(1, n_fft, 2, dtype=torch.float64)
        self.pre = torch.exp(-1j * (torch.pi / (2 * n_fft) + torch.pi / 4) * k).to(device)
        m = torch.arange(0, 2 * n_fft, 2, dtype=torch.float64)
        self.post = torch.exp(-1j * torch.pi / (2 * n_fft) * m).to(device)

    def __call__(self, spec: torch.Tensor):
        assert spec.dim() == 3 and spec.shape[-1] == self.n // 2
        y = torch.real(torch.fft.fft(self.pre * spec, n=self.n) * self.post)[..., : self.win] * self.w
        total = (y.shape[-2] - 1) * self.hop + self.win
        out = 4 / self.n * F.fold(y.transpose(-1, -2), kernel_size=(1, self.win), stride=(1, self.hop), output_size=(1, total))
        out = out[..., self.win // 2: -self.win // 2]
        return out if self.out_length is None else out[..., : self.out_length]


class Audio2MDCTPort:
    """arcsinh + abs_norm branch (the configs' branch)."""

    def __init__(self, gain=1000.0, src_range=(-5.0, 5.0), norm_range=(-1.0, 1.0), n_fft=512, hop=256, device="cpu"):
        """`device="cuda"`: the same torch ops on the GPU (cuFFT) -- the torch-eager "kernel to beat" leg of bench.py."""
        self.gain, self.src, self.rng, self.device = gain, src_range, norm_range, device
        self.w = kbdwin(n_fft)
        self.fwd = MDCT4Port(n_fft, hop, self.w, device=device)
        self.inv = IMDCT4Port(n_fft, hop, self.w, device=device)
        self.ln10 = torch.log(torch.tensor(10.0)).to(device)   # fp32 constant, pix2pixHD_model.py:100,133
        self.lo = torch.tensor([src_range[0]])[None, None, None, :].to(device)
        self.hi = torch.tensor([src_range[1]])[None, None, None, :].to(device)

    def to_spectro(self, audio: torch.Tensor):
        spec, frames = self.fwd(audio, True)         # the reference always asks for the frames clone (:34)
        spec = spec.unsqueeze(1)
        pha = torch.sign(spec)
        s = torch.arcsinh(self.gain * spec) / self.ln10
        mean, std = s.mean().float(), s.var().sqrt().float()      # computed and never consumed (:108-109)
        noise = torch.randn(pha.size(), device=spec.device)       # the throw-away draw of :49-54
        noise = (noise - noise.min()) / (noise.max() - noise.min())
        pha = pha * noise
        lo, hi = self.lo, self.hi
        s = (s - lo) / (hi - lo)
        s = s * (self.rng[1] - self.rng[0]) + self.rng[0]
        return s.float(), pha, {"max": hi, "min": lo, "mean": mean, "std": std, "frames": frames}

    def to_audio(self, s: torch.Tensor, prm, pha=None):
        x = (s.to(torch.float64) - self.rng[0]) / (self.rng[1] - self.rng[0])
        x = x * (prm["max"] - prm["min"]) + prm["min"]
        x = torch.sinh(x * self.ln10) / self.gain
        return self.inv(x.squeeze(1))
