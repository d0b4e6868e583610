"""Oracle (CPU, numpy) for the transform half of the hot path.  TEST INFRASTRUCTURE ONLY.

Restates, function by function, what the reference computes (citations are into
/root/reference):

  kbdwin            util/util.py:179-186
  frame_count       models/mdct.py:394-407   (incl. the len(signal)==batch quirk)
  mdct4             models/mdct.py:392-425   (512-pt complex128 FFT formulation)
  mdct4_closed      the cosine sum the FFT formulation evaluates
  imdct4            models/mdct.py:457-489
  imdct4_closed     SURVEY.md a-3
  compress/expand   models/pix2pixHD_model.py:83-137 (normalize / denormalize)
  to_spectro        models/pix2pixHD_model.py:32-81  (deterministic part)
  to_audio          models/pix2pixHD_model.py:139-163

Pinned by tests/test_oracle_golden.py against tests/golden/*.npz, which were
produced by running the reference itself (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

LN10_F32 = float(np.log(np.float32(10.0), dtype=np.float32))  # pix2pixHD_model.py:100,133


# --------------------------------------------------------------------------- window
def kbdwin(n: int, beta: float = 12.0) -> np.ndarray:
    """fp32 Kaiser-Bessel-derived window exactly as util/util.py:179-186 builds it.

    The reference calls torch.kaiser_window (fp32); we call the same library
    routine so the fp32 bits agree, then do the cumsum/sqrt in fp32 like it does.
    """
    import torch

    assert n % 2 == 0, "N must be even"
    k = torch.kaiser_window(window_length=n // 2 + 1, beta=beta * math.pi, periodic=False)
    half = torch.sqrt(torch.cumsum(k, dim=0) / k.sum())[:-1]
    return torch.cat((half, half.flip(0))).numpy().astype(np.float32)


def kbdwin_f64(n: int, beta: float = 12.0) -> np.ndarray:
    """Same window evaluated in fp64 (Princen-Bradley residual ~1e-16)."""
    k = np.kaiser(n // 2 + 1, beta * math.pi)
    half = np.sqrt(np.cumsum(k) / k.sum())[:-1]
    return np.concatenate((half, half[::-1]))


# --------------------------------------------------------------------------- framing
def frame_count(t: int, hop: int, win: int, dim0: int, center: bool = True) -> tuple[int, int, int]:
    """(start_pad, end_pad, frames) per models/mdct.py:394-407.

    `dim0` is len(signal): the number of samples for 1-D input but the *batch
    size* for 2-D input (reference quirk).
    """
    start = hop if center else 0
    extra = dim0 % hop
    end = start + (hop - extra if extra else 0)
    total = start + t + end
    frames = (total - win) // hop + 1 if total >= win else 0
    return start, end, frames


def _frames(signal: np.ndarray, hop: int, win: int, center: bool) -> np.ndarray:
    x = np.asarray(signal, dtype=np.float32)
    dim0 = x.shape[0]
    t = x.shape[-1]
    start, end, f = frame_count(t, hop, win, dim0, center)
    pad = [(0, 0)] * (x.ndim - 1) + [(start, end)]
    xp = np.pad(x, pad)
    idx = np.arange(f)[:, None] * hop + np.arange(win)[None, :]
    return xp[..., idx]  # [..., F, win] fp32


# --------------------------------------------------------------------------- MDCT4
def mdct4(signal, window, n_fft: int = 512, hop: int = 256, center: bool = True):
    """fp64 spectrogram [..., F, n_fft/2] and the fp32 windowed frames (mdct.py:392-425)."""
    w = np.asarray(window, dtype=np.float32)
    win = w.shape[0]
    fr = _frames(signal, hop, win, center) * w  # fp32 * fp32 -> fp32 (mdct.py:410)
    z = np.zeros(fr.shape[:-1] + (n_fft,), dtype=np.float64)
    z[..., :win] = fr
    m = np.arange(n_fft, dtype=np.float64)
    pre = np.exp(-1j * np.pi / n_fft * m)
    k = np.arange(1, n_fft, 2, dtype=np.float64)
    post = np.exp(-1j * (np.pi / (2 * n_fft) + np.pi / 4) * k)
    spec = np.fft.fft(z * pre, axis=-1)[..., : n_fft // 2]
    return np.real(post * spec), fr


def _cos_table(n_fft: int) -> np.ndarray:
    nb = n_fft // 2
    m = np.arange(n_fft, dtype=np.float64)[:, None]
    k = np.arange(nb, dtype=np.float64)[None, :]
    return np.cos(np.pi / nb * (m + 0.5 + nb / 2) * (k + 0.5))


def mdct4_closed(signal, window, n_fft: int = 512, hop: int = 256, center: bool = True):
    """X[t,k] = sum_m w[m] x_pad[t*hop+m] cos(pi/N (m+1/2+N/2)(k+1/2)), N = n_fft/2."""
    w = np.asarray(window)
    win = w.shape[0]
    assert win == n_fft, "closed form written for win_length == n_fft"
    fr = _frames(signal, hop, win, center)
    if w.dtype == np.float32:
        z = (fr * w).astype(np.float64)
    else:
        z = fr.astype(np.float64) * w
    return z @ _cos_table(n_fft)


# --------------------------------------------------------------------------- IMDCT4
def imdct4(spec, window, n_fft: int = 512, hop: int = 256, center: bool = True, out_length=None):
    """fp64 audio [B,1,1,T] from spec [B,F,n_fft/2] (mdct.py:457-489)."""
    x = np.asarray(spec, dtype=np.float64)
    assert x.ndim == 3 and x.shape[-1] == n_fft // 2
    w = np.asarray(window)
    win = w.shape[0]
    k = np.arange(1, n_fft, 2, dtype=np.float64)
    pre = np.exp(-1j * (np.pi / (2 * n_fft) + np.pi / 4) * k)
    m = np.arange(0, 2 * n_fft, 2, dtype=np.float64)
    post = np.exp(-1j * np.pi / (2 * n_fft) * m)
    y = np.real(np.fft.fft(pre * x, n=n_fft, axis=-1) * post)[..., :win]
    y = y * w.astype(np.float64) if w.dtype != np.float64 else y * w
    return _overlap_add(y, n_fft, hop, win, center, out_length)


def imdct4_closed(spec, window, n_fft: int = 512, hop: int = 256, center: bool = True, out_length=None):
    x = np.asarray(spec, dtype=np.float64)
    w = np.asarray(window).astype(np.float64)
    y = (x @ _cos_table(n_fft).T) * w
    return _overlap_add(y, n_fft, hop, w.shape[0], center, out_length)


def _overlap_add(y, n_fft, hop, win, center, out_length):
    b, f, _ = y.shape
    total = (f - 1) * hop + win
    out = np.zeros((b, total), dtype=np.float64)
    for t in range(f):  # fold == overlap-add (mdct.py:480-482)
        out[:, t * hop : t * hop + win] += y[:, t]
    out *= 4.0 / n_fft
    if center:
        out = out[:, win // 2 : total - win // 2]
    if out_length is not None:
        out = out[:, :out_length]
    return out[:, None, None, :]


# --------------------------------------------------------------------------- compress / normalise
def _amp_to_db(x, amin):
    """torchaudio.functional.amplitude_to_DB(x, multiplier=20.0, amin, db_multiplier=1.0) as the reference calls it."""
    return 20.0 * np.log10(np.clip(x, amin, None)) - 20.0


def compress(spec, *, arcsinh_transform=True, arcsinh_gain=500.0, raw_mdct=False, explicit_encoding=False, alpha=0.6, min_value=1e-7,
             abs_norm=True, src_range=(-5.0, 5.0), norm_range=(0.0, 1.0)):
    """Audio2MDCT.normalize (pix2pixHD_model.py:83-125): explicit_encoding (2 channels) / arcsinh / raw / dB, abs-norm or per-plane min-max.

    Returns (log_spectro fp64, max, min, mean fp32, std fp32).  With abs_norm the
    max/min are the fp32 constants the reference builds (shape [1,1,1,1]).
    """
    x = np.asarray(spec, dtype=np.float64)
    if explicit_encoding:      # :84-95
        neg = 0.5 * (np.abs(x) - x)
        pos = x + neg
        s = np.concatenate((_amp_to_db(alpha * pos + (1 - alpha) * neg, min_value), _amp_to_db((1 - alpha) * pos + alpha * neg, min_value)), axis=1)
    elif arcsinh_transform:
        s = np.arcsinh(arcsinh_gain * x) / LN10_F32
    elif raw_mdct:
        s = x
    else:  # :104-106
        s = _amp_to_db(np.abs(x) + min_value, min_value)
    mean = np.float32(s.mean())
    std = np.float32(np.sqrt(s.var(ddof=1))) if s.size > 1 else np.float32(np.nan)
    if abs_norm:
        lo_src = np.full((1, 1, 1, 1), src_range[0], dtype=np.float32)
        hi_src = np.full((1, 1, 1, 1), src_range[1], dtype=np.float32)
    else:
        flat = s.reshape(s.shape[0], s.shape[1], -1)
        hi_src = flat.max(-1)[:, :, None, None].astype(np.float32)
        lo_src = flat.min(-1)[:, :, None, None].astype(np.float32)
    s = (s - lo_src) / (hi_src - lo_src)
    s = s * (norm_range[1] - norm_range[0]) + norm_range[0]
    return s, hi_src, lo_src, mean, std


def expand(log_spectro, lo_src, hi_src, *, arcsinh_transform=True, arcsinh_gain=500.0, raw_mdct=False, explicit_encoding=False, alpha=0.6,
           min_value=1e-7, norm_range=(0.0, 1.0)):
    """Audio2MDCT.denormalize (pix2pixHD_model.py:127-137); the explicit_encoding flag is only accepted (it acts in to_audio)."""
    s = (np.asarray(log_spectro).astype(np.float64) - norm_range[0]) / (norm_range[1] - norm_range[0])
    s = s * (np.asarray(hi_src) - np.asarray(lo_src)) + np.asarray(lo_src)
    if arcsinh_transform:
        return np.sinh(s * LN10_F32) / arcsinh_gain
    if raw_mdct:
        return s
    return 10.0 * np.power(np.power(10.0, 0.1 * s), 0.5) - min_value  # aF.DB_to_amplitude(x, 10.0, 0.5) - min_value


def to_spectro(audio, window, *, n_fft=512, hop=256, **kw):
    """Deterministic part of Audio2MDCT.to_spectro (pix2pixHD_model.py:32-81).

    Returns (log_spectro fp32 [B,1,F,N], sign fp64 [B,1,F,N], hi, lo).  The
    reference multiplies `sign` by min-max-scaled randn noise (:49-54); that
    factor is random and unused in arcsinh mode, so it is not reproduced.
    """
    spec, _ = mdct4(audio, window, n_fft, hop)
    spec = spec[:, None]
    s, hi_src, lo_src, _, _ = compress(spec, **kw)
    return s.astype(np.float32), np.sign(spec), hi_src, lo_src


def to_audio(log_spectro, lo_src, hi_src, window, *, n_fft=512, hop=256, pha=None, **kw):
    """Audio2MDCT.to_audio (pix2pixHD_model.py:139-163).  dB mode: `pha` is the phase multiplier actually applied (the reference
    builds it from the sign, noise and a random pseudo phase for the upper frames, :150-157 -- the caller passes the result)."""
    kw.pop("abs_norm", None)
    kw.pop("src_range", None)
    x = expand(log_spectro, lo_src, hi_src, **kw)
    if kw.get("explicit_encoding", False):
        x = ((x[:, 0] - x[:, 1]) / (2 * kw.get("alpha", 0.6) - 1))[:, None]
    elif not kw.get("arcsinh_transform", True) and not kw.get("raw_mdct", False) and pha is not None:
        x = x * np.asarray(pha, dtype=np.float64)
    return imdct4(x[:, 0], window, n_fft, hop)
