"""Oracle (CPU, torch) for the tensor half of AudioDataset.__getitem__ -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/data/audio_dataset.py:66-82 (three torchaudio resamples, noise injection :72-78, seg_pad_audio :104-110).
Pinned against tests/golden/augment_golden.npz = outputs of the reference's own AudioDataset.__getitem__ with `readaudio` stubbed
to return a seeded waveform (tests/golden/make_golden.py augment) by tests/test_augment.py."""
import torch
import torch.nn.functional as F


def add_noise(lr_waveform, snr, segment_length, noise):
    noise = noise - noise.mean()
    signal_power = torch.sum(lr_waveform ** 2) / segment_length
    noise_var = signal_power / 10 ** (snr / 10)
    noise = torch.sqrt(noise_var) / noise.std() * noise
    return lr_waveform + noise


def seg_pad_audio(waveform, segment_length):
    if waveform.size(1) >= segment_length:
        return waveform[0][:segment_length]
    return F.pad(waveform, (0, segment_length - waveform.size(1)), "constant")


def training_pair(waveform, orig_sr, lr_sr, hr_sr, segment_length, snr=None, noise=None):
    import torchaudio.functional as aF

    hr = aF.resample(waveform, orig_sr, hr_sr)
    lr = aF.resample(aF.resample(waveform, orig_sr, lr_sr), lr_sr, hr_sr)
    if snr is not None:
        lr = add_noise(lr, snr, segment_length, noise)
    return {"HR_audio": seg_pad_audio(hr, segment_length).squeeze(0), "LR_audio": seg_pad_audio(lr, segment_length).squeeze(0)}
