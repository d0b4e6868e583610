"""Evaluation metrics (mdctgan_b200/util/util.py:compute_matrics; reference util/util.py:132-177).
CPU: the oracle restatement against goldens made by the reference's own compute_matrics.  GPU: the kernels through the C ABI
against those goldens: 1e-5 relative on MSE / SNR (double accumulation of fp32 data), 2e-4 relative on the LSD (fp32 1024-point FFT
and log10 against the reference's fp32 torch.stft)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
from make_golden import METRIC_CASES, metric_signals  # noqa: E402
from oracle import metrics_oracle as MO  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "metrics_golden.npz")))


@pytest.mark.parametrize("case", METRIC_CASES)
def test_oracle_matches_reference(gold, case):
    rows, T, center, seed = case
    hr, lr, sr = metric_signals(rows, T, seed)
    got = np.array(MO.compute_matrics(hr, lr, sr, center=center), dtype=np.float64)
    np.testing.assert_allclose(got, gold[f"m_{rows}_{T}_{int(center)}"], rtol=1e-5, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("case", METRIC_CASES)
def test_kernels_match_reference(gold, case):
    from mdctgan_b200.util.util import compute_matrics

    rows, T, center, seed = case
    hr, lr, sr = metric_signals(rows, T, seed)
    opt = SimpleNamespace(n_fft=512, hop_length=256, win_length=512, center=center)
    dev = torch.device("cuda:0")
    got = np.array(compute_matrics(hr.to(dev), lr.to(dev), sr.to(dev), opt), dtype=np.float64)
    want = gold[f"m_{rows}_{T}_{int(center)}"]
    np.testing.assert_allclose(got[:3], want[:3], rtol=1e-5)
    assert got[3] == got[4] == got[5] == 0
    np.testing.assert_allclose(got[6], want[6], rtol=2e-4)
