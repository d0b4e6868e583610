"""GPU resampling for the data path in front of the model (reference: data/audio_dataset.py:66-71 and :169-177 call
torchaudio.functional.resample on CPU DataLoader workers: HR clip -> LR rate -> back to the HR rate).

`resample()` follows torchaudio's published algorithm (sinc interpolation with a Hann window, lowpass_filter_width 6, rolloff 0.99):
a polyphase FIR table [new/gcd][2*width + orig/gcd] built once per rate pair on the host, applied by `mdctgan_resample_fir`."""
from __future__ import annotations

import math
from ctypes import c_int, c_int64, c_void_p

import torch

from .. import _lib
from .. import nn_ops as ops

_tables = {}


def sinc_resample_table(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """(table fp32 [new][K], width, orig, new) with orig / new reduced by their gcd -- the Hann-windowed sinc filters of
    torchaudio.functional.resample, evaluated in fp32 in the same order of operations."""
    if not (int(orig_freq) == orig_freq and int(new_freq) == new_freq) or orig_freq <= 0 or new_freq <= 0:
        raise ValueError("resample: frequencies must be positive integers")
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float32)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float32)[:, None, None] / new + idx
    t *= base
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    kern = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kern *= window * (base / orig)
    return kern.reshape(new, 2 * width + orig).contiguous(), width, orig, new


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """torchaudio.functional.resample(waveform, orig_freq, new_freq) on the device: [..., time] fp32 CUDA -> [..., ceil(time*new/orig)]."""
    if not waveform.is_cuda:
        raise RuntimeError("resample: expected a CUDA tensor; mdctgan_b200 has no CPU path")
    if orig_freq == new_freq:
        return waveform
    key = (int(orig_freq), int(new_freq), waveform.device.index)
    if key not in _tables:
        tab, width, orig, new = sinc_resample_table(orig_freq, new_freq)
        _tables[key] = (tab.to(waveform.device), width, orig, new)
    tab, width, orig, new = _tables[key]
    shape = waveform.shape
    x = waveform.to(torch.float32).reshape(-1, shape[-1]).contiguous()
    rows, L = x.shape
    target = int(math.ceil(new * L / orig))
    y = torch.empty((rows, target), dtype=torch.float32, device=x.device)
    Lb = ops._L()
    Lb.mdctgan_resample_fir.argtypes = [c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p]
    if rows and target:
        with torch.cuda.device(x.device):
            _lib.check(Lb.mdctgan_resample_fir(x.data_ptr(), rows, L, tab.data_ptr(), tab.shape[1], orig, new, width, y.data_ptr(), target,
                                               torch.cuda.current_stream(x.device).cuda_stream))
    return y.reshape(shape[:-1] + (target,))


def make_lr_hr(waveform: torch.Tensor, orig_sample_rate: int, lr_sampling_rate: int, hr_sampling_rate: int):
    """The three resamples of AudioDataset.__getitem__ (data/audio_dataset.py:66-71): (hr_waveform, lr_waveform at the HR rate)."""
    hr = resample(waveform, orig_sample_rate, hr_sampling_rate)
    lr = resample(resample(waveform, orig_sample_rate, lr_sampling_rate), lr_sampling_rate, hr_sampling_rate)
    return hr, lr
