"""Oracle (CPU, torch) for the evaluation metrics -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/util/util.py:132-177 (compute_matrics): MSE, SNR of sr / lr against hr, log-spectral distance over
the power spectrogram torchaudio.functional.spectrogram(n_fft = 2*opt.n_fft, hop = 2*opt.hop_length, win = 2*opt.win_length,
window = kbdwin(2*win), center = opt.center, power = 2) -- written with torch.stft, which is what aF.spectrogram calls.
Pinned by tests/golden/metrics_golden.npz (the reference's own compute_matrics on seeded signals)."""
import torch

from .torch_port import kbdwin


def compute_matrics(hr_audio, lr_audio, sr_audio, n_fft=512, hop_length=256, win_length=512, center=True):
    mse = ((sr_audio - hr_audio) ** 2).mean().item()
    snr_sr = (10 * torch.log10(torch.sum(hr_audio ** 2, dim=-1) / torch.sum((sr_audio - hr_audio) ** 2, dim=-1))).mean().item()
    snr_lr = (10 * torch.log10(torch.sum(hr_audio ** 2, dim=-1) / torch.sum((lr_audio - hr_audio) ** 2, dim=-1))).mean().item()
    w = kbdwin(2 * win_length).to(hr_audio.dtype)

    def power_spec(x):
        shape = x.shape
        z = torch.stft(x.reshape(-1, shape[-1]), n_fft=2 * n_fft, hop_length=2 * hop_length, win_length=2 * win_length, window=w,
                       center=center, pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
        return (z.abs() ** 2).reshape(shape[:-1] + z.shape[-2:])

    hl, sl = torch.log10(power_spec(hr_audio) + 1e-6), torch.log10(power_spec(sr_audio) + 1e-6)
    lsd = torch.sqrt(torch.mean((hl - sl) ** 2, dim=-2)).mean().item()
    return mse, snr_sr, snr_lr, 0, 0, 0, lsd
