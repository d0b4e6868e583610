"""Transform constants (reference: options/audio_config.py:1-12)."""
N_FFT, HOP_LENGTH, WIN_LENGTH = 512, 256, 512
LR_SAMPLE_RATE, HR_SAMPLE_RATE, SR_SAMPLE_RATE = 8000, 48000, 48000
BINS = 128
assert BINS % 16 == 0
CENTER = True
FRAME_LENGTH = (BINS - 1) * HOP_LENGTH if CENTER else (BINS - 1) * HOP_LENGTH + WIN_LENGTH
