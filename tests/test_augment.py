"""The tensor half of AudioDataset.__getitem__ (mdctgan_b200/data/augment.py; reference data/audio_dataset.py:54-82,104-110).
CPU: the oracle restatement against goldens produced by the reference's own __getitem__ (readaudio stubbed, seeded torch.randn).
GPU: resample -> add_noise -> crop / pad kernels through the C ABI against the same goldens."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
from make_golden import AUGMENT_CASES  # noqa: E402
from oracle import augment_oracle as AO  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "augment_golden.npz")))


def _inputs(case):
    L, osr, lsr, hsr, seg, noise_on, snr, seed = case
    g = torch.Generator().manual_seed(seed)
    wave = 0.1 * torch.randn(1, L, generator=g)
    torch.manual_seed(seed + 1)
    n_lr = int(np.ceil(int(np.ceil(L * lsr / osr)) * hsr / lsr))           # length after orig -> lr -> hr
    noise = torch.randn(1, n_lr) if noise_on else None
    return wave, noise, f"{L}_{osr}_{lsr}_{seg}_{int(noise_on)}"


@pytest.mark.parametrize("case", AUGMENT_CASES)
def test_oracle_matches_reference_getitem(gold, case):
    L, osr, lsr, hsr, seg, noise_on, snr, seed = case
    wave, noise, key = _inputs(case)
    item = AO.training_pair(wave, osr, lsr, hsr, seg, snr if noise_on else None, noise)
    assert item["HR_audio"].shape == (seg,) and item["LR_audio"].shape == (seg,)
    assert np.array_equal(item["HR_audio"].numpy(), gold[f"hr_{key}"])
    assert np.array_equal(item["LR_audio"].numpy(), gold[f"lr_{key}"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", AUGMENT_CASES)
def test_device_pair_matches_reference_getitem(gold, case):
    """Resampling is within 2e-6 of peak of torchaudio (tests/test_resample.py); the injected noise is scaled from fp64 sums instead
    of torch's fp32 reductions: 1e-6 relative on the noise amplitude.  The crop / pad is exact."""
    from mdctgan_b200.data.augment import add_noise, fit_segment, training_pair

    L, osr, lsr, hsr, seg, noise_on, snr, seed = case
    dev = torch.device("cuda:0")
    wave, noise, key = _inputs(case)
    item = training_pair(wave.to(dev), osr, lsr, hsr, seg, snr if noise_on else None, noise)
    for k, gk in (("HR_audio", f"hr_{key}"), ("LR_audio", f"lr_{key}")):
        got, ref = item[k].cpu().numpy(), gold[gk]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 4e-6 * np.abs(ref).max(), k
        assert np.array_equal(got == 0, ref == 0) or L >= seg * osr / hsr          # the zero-padded tail is exactly zero
    # noise injection alone on identical inputs: within 2e-6 relative of the oracle
    x = 0.1 * torch.randn(1, 5000)
    nz = torch.randn(1, 5000)
    ref = AO.add_noise(x, 20.0, 4000, nz)
    got = add_noise(x.to(dev), 20.0, 4000, nz).cpu()
    assert float((got - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
    # crop / pad: exact
    assert torch.equal(fit_segment(x.to(dev), 4000).cpu(), x[0, :4000])
    assert torch.equal(fit_segment(x.to(dev), 6000).cpu(), torch.cat((x[0], torch.zeros(1000))))
