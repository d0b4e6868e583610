"""Host-side helpers with the reference's names (util/util.py)."""
import math

import torch


def kbdwin(N: int, beta: float = 12.0, device="cpu") -> torch.Tensor:
    """MATLAB-style Kaiser-Bessel-derived window, fp32 (reference: util/util.py:179-186).

    Built with the same torch library routine on the same device as the reference builds it
    (CPU by default, then moved), so the fp32 bits -- which set the MDCT round-trip error floor --
    are identical.
    """
    assert N % 2 == 0, "N must be even"
    k = torch.kaiser_window(window_length=N // 2 + 1, beta=beta * math.pi, periodic=False, device=device)
    half = torch.sqrt(torch.cumsum(k, dim=0) / k.sum())[:-1]
    return torch.cat((half, half.flip(dims=(0,))), dim=0)


_lsd_windows = {}


def compute_matrics(hr_audio: torch.Tensor, lr_audio: torch.Tensor, sr_audio: torch.Tensor, opt):
    """MSE, SNR(sr), SNR(lr), 0, 0, 0, LSD -- the reference's evaluation metrics (util/util.py:132-177; callers train.py:116-117,
    generate_audio.py:59-60) computed on the device by two kernels of libmdctgan_b200.so: per-row error / energy sums
    (`mdctgan_metrics_rows`) and the log-spectral distance over the 2*n_fft STFT with the kbdwin(2*win_length) window
    (`mdctgan_lsd_frames`: the hr and sr frames share one complex FFT).  Audio: [..., T]; fp32 on the device (fp64 inputs are
    rounded once).  Returns Python floats like the reference's `.item()` calls."""
    from ctypes import c_double, c_int, c_int64, c_void_p

    from .. import _lib
    from .. import nn_ops as ops

    dev = sr_audio.device
    if dev.type != "cuda":
        raise RuntimeError("compute_matrics: expected CUDA tensors; mdctgan_b200 has no CPU path")
    T = sr_audio.shape[-1]
    hr = hr_audio.to(dev, torch.float32).reshape(-1, T).contiguous()
    lr = lr_audio.to(dev, torch.float32).reshape(-1, T).contiguous()
    sr = sr_audio.to(torch.float32).reshape(-1, T).contiguous()
    if not (hr.shape == lr.shape == sr.shape):
        raise ValueError(f"compute_matrics: shapes differ: {tuple(hr_audio.shape)}, {tuple(lr_audio.shape)}, {tuple(sr_audio.shape)}")
    rows = hr.shape[0]
    L = ops._L()
    L.mdctgan_metrics_rows.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]
    L.mdctgan_lsd_frame_count.restype = c_int64
    L.mdctgan_lsd_frame_count.argtypes = [c_int64, c_int, c_int, c_int]
    L.mdctgan_lsd_frames.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]
    n_fft, hop, win = 2 * opt.n_fft, 2 * opt.hop_length, 2 * opt.win_length
    if win != n_fft:
        raise NotImplementedError("compute_matrics: win_length != n_fft is not a reference configuration (options/audio_config.py)")
    key = (win, dev.index)
    if key not in _lsd_windows:
        _lsd_windows[key] = kbdwin(win).to(dev)
    acc = torch.zeros(rows * 3 + 1, dtype=torch.float64, device=dev)
    center = 1 if getattr(opt, "center", False) else 0
    frames = int(L.mdctgan_lsd_frame_count(T, n_fft, hop, center))
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.mdctgan_metrics_rows(hr.data_ptr(), lr.data_ptr(), sr.data_ptr(), rows, T, acc.data_ptr(), st))
        _lib.check(L.mdctgan_lsd_frames(hr.data_ptr(), sr.data_ptr(), rows, T, n_fft, hop, _lsd_windows[key].data_ptr(), center,
                                        acc[rows * 3:].data_ptr(), st))
    a = acc.cpu()                                        # one small D2H copy; the rest is host arithmetic on rows*3 + 1 numbers
    r = a[:rows * 3].view(rows, 3)
    mse = (r[:, 0].sum() / (rows * T)).item()
    snr_sr = (10 * torch.log10(r[:, 1] / r[:, 0])).mean().item()
    snr_lr = (10 * torch.log10(r[:, 1] / r[:, 2])).mean().item()
    lsd = (a[rows * 3] / max(rows * frames, 1)).item()
    return mse, snr_sr, snr_lr, 0, 0, 0, lsd
