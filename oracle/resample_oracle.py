"""Oracle (CPU, numpy fp64) for the sinc resampler -- TEST INFRASTRUCTURE ONLY.

The reference resamples with the third-party torchaudio.functional.resample (data/audio_dataset.py:66-71, :169-177; torchaudio is a
pip dependency, present in this image: 2.11).  Restated here from its published algorithm (sinc_interp_hann, lowpass_filter_width 6,
rolloff 0.99) in fp64; pinned by tests/golden/resample_golden.npz = outputs of torchaudio's own function on seeded waveforms
(tests/golden/make_golden.py resample)."""
import math

import numpy as np


def resample(x, orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    x = np.asarray(x, dtype=np.float64)
    if orig_freq == new_freq:
        return x
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = (np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    kern = np.where(t == 0, 1.0, np.sin(t) / np.where(t == 0, 1.0, t)) * window * (base / orig)      # [new][K]
    rows = x.reshape(-1, x.shape[-1])
    L = rows.shape[-1]
    target = int(math.ceil(new * L / orig))
    xp = np.pad(rows, ((0, 0), (width, width + orig)))
    K = kern.shape[1]
    nq = (xp.shape[-1] - K) // orig + 1
    out = np.empty((rows.shape[0], nq, new))
    for q in range(nq):
        out[:, q, :] = xp[:, q * orig:q * orig + K] @ kern.T
    return out.reshape(rows.shape[0], -1)[:, :target].reshape(x.shape[:-1] + (target,))
