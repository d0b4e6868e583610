"""The tensor half of AudioDataset.__getitem__ on the device (reference: data/audio_dataset.py:54-82,104-110): the three resamples
(data/resample.py), the optional SNR-controlled noise injection and the crop / zero-pad to `segment_length`.  File decoding and the
random crop offset stay with the caller (the reference reads them through torchaudio.load on DataLoader workers)."""
from __future__ import annotations

from ctypes import c_double, c_int64, c_void_p

import torch

from .. import _lib
from .. import nn_ops as ops
from ..longform import _L as _LF
from .resample import make_lr_hr


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def add_noise(lr_waveform: torch.Tensor, snr: float, segment_length: int, noise: torch.Tensor = None) -> torch.Tensor:
    """data/audio_dataset.py:72-78.  `noise`: the N(0, 1) draw with lr_waveform's shape (default: torch.randn on the device, the
    call the reference makes on the CPU)."""
    if not lr_waveform.is_cuda:
        raise RuntimeError("add_noise: expected a CUDA tensor; mdctgan_b200 has no CPU path")
    x = lr_waveform.to(torch.float32).contiguous()
    if noise is None:
        noise = torch.randn(x.shape, device=x.device)
    nz = noise.to(device=x.device, dtype=torch.float32).contiguous()
    if nz.shape != x.shape:
        raise ValueError(f"add_noise: noise shape {tuple(nz.shape)} != waveform shape {tuple(x.shape)}")
    out = torch.empty_like(x)
    scratch = torch.empty(3, dtype=torch.float64, device=x.device)
    L = ops._L()
    L.mdctgan_add_noise.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_void_p, c_void_p]
    with torch.cuda.device(x.device):
        _lib.check(L.mdctgan_add_noise(x.data_ptr(), nz.data_ptr(), out.data_ptr(), x.numel(), float(segment_length), float(snr),
                                       scratch.data_ptr(), _stream(x)))
    return out


def fit_segment(waveform: torch.Tensor, segment_length: int) -> torch.Tensor:
    """AudioDataset.seg_pad_audio (data/audio_dataset.py:104-110): [1, L] -> [segment_length], cropped or zero-padded at the end."""
    if not waveform.is_cuda:
        raise RuntimeError("fit_segment: expected a CUDA tensor; mdctgan_b200 has no CPU path")
    a = waveform.reshape(-1).to(torch.float32).contiguous()
    out = torch.empty((1, segment_length), dtype=torch.float32, device=a.device)
    if a.numel() == 0:
        return out.zero_()[0]
    with torch.cuda.device(a.device):
        _lib.check(_LF().mdctgan_segment_gather(a.data_ptr(), a.numel(), out.data_ptr(), 1, segment_length, 0, _stream(a)))
    return out[0]


def training_pair(waveform: torch.Tensor, orig_sample_rate: int, lr_sampling_rate: int, hr_sampling_rate: int, segment_length: int,
                  add_noise_snr: float = None, noise: torch.Tensor = None):
    """AudioDataset.__getitem__ after the file read: {'HR_audio': [segment_length], 'LR_audio': [segment_length]} on the device."""
    hr, lr = make_lr_hr(waveform, orig_sample_rate, lr_sampling_rate, hr_sampling_rate)
    if add_noise_snr is not None:
        lr = add_noise(lr, add_noise_snr, segment_length, noise)
    return {"HR_audio": fit_segment(hr, segment_length), "LR_audio": fit_segment(lr, segment_length)}
