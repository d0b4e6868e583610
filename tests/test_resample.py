"""Sinc resampling of the data path in front of the model (mdctgan_b200/data/resample.py; the reference calls
torchaudio.functional.resample, data/audio_dataset.py:66-71).  CPU: the fp64 oracle restatement and our fp32 filter table against
goldens made by torchaudio itself.  GPU: the FIR kernel through the C ABI against the same goldens (2e-6 of the signal peak: fp32
filter taps, different summation order)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
from make_golden import RESAMPLE_CASES, resample_wave  # noqa: E402
from oracle import resample_oracle as RO  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "resample_golden.npz")))


@pytest.mark.parametrize("i", range(len(RESAMPLE_CASES)))
def test_oracle_and_table_match_torchaudio(gold, i):
    o, n, L = RESAMPLE_CASES[i]
    x = resample_wave(L, 300 + i)
    want = gold[f"r_{o}_{n}_{L}"]
    got = RO.resample(x.numpy(), o, n)
    assert got.shape == want.shape
    # torchaudio evaluates the filter taps in fp32 (arange / orig_freq etc.): against the fp64 restatement that costs ~1e-7 for small
    # rate ratios and ~2e-5 for 147:160 (44.1 -> 48 kHz)
    assert np.abs(got - want).max() < (5e-5 if o == 44100 else 2e-6) * np.abs(want).max()
    from mdctgan_b200.data.resample import sinc_resample_table      # host-side table builder (no GPU needed)

    tab, width, orig, new = sinc_resample_table(o, n)
    assert tab.shape == (new, 2 * width + orig) and tab.dtype == torch.float32
    import torchaudio.functional.functional as Fn
    import math

    ref_tab, ref_w = Fn._get_sinc_resample_kernel(o, n, math.gcd(o, n), dtype=torch.float32)
    assert ref_w == width and torch.equal(ref_tab.reshape(new, -1), tab)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(RESAMPLE_CASES)))
def test_kernel_matches_torchaudio(gold, i):
    from mdctgan_b200.data.resample import make_lr_hr, resample

    o, n, L = RESAMPLE_CASES[i]
    dev = torch.device("cuda:0")
    x = resample_wave(L, 300 + i)
    want = gold[f"r_{o}_{n}_{L}"]
    got = resample(x.to(dev), o, n).cpu().numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-6 * np.abs(want).max()
    if i == 0:      # the reference's HR -> LR -> HR chain keeps the length of a clip whose length divides the ratio
        hr, lr = make_lr_hr(resample_wave(4000, 1).to(dev), 48000, 12000, 48000)
        assert hr.shape == lr.shape == (2, 4000)
