"""Long-form segmentation / overlap-add (mdctgan_b200/longform.py; reference: data/audio_dataset.py:153-167,
generate_audio.py:40-53).  CPU: the oracle restatement against goldens made with the reference's own
AudioTestDataset.seg_pad_audio.  GPU: the kernels through the C ABI against the oracle, bit-exact (pure data movement, one
multiply by 0.5 and one add per sample, in the reference's order), incl. a 60 s clip (BASELINE configs[4]) and a clip shorter
than one segment; plus the end-to-end generator against per-segment inference."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
from make_golden import LONGFORM_CASES, LONGFORM_FULL, longform_clip  # noqa: E402
from oracle import longform_oracle as LO  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "longform_golden.npz")))


def _ck(t):
    return np.array([float(t.double().sum()), float((t.double() ** 2).sum())])


@pytest.mark.parametrize("case", LONGFORM_CASES)
def test_oracle_matches_reference(gold, case):
    L, seg, ov = case
    key = f"{L}_{seg}_{ov}"
    segs = LO.seg_pad_audio(longform_clip(L, seg, ov), seg, ov)
    assert list(segs.shape) == list(gold[f"segshape_{key}"])
    np.testing.assert_allclose(_ck(segs), gold[f"segck_{key}"], rtol=1e-13)
    ola = LO.overlap_add(segs.double().reshape(-1, 1, 1, seg), seg, ov)
    assert list(ola.shape) == list(gold[f"olashape_{key}"])
    np.testing.assert_allclose(_ck(ola), gold[f"olack_{key}"], rtol=1e-13)
    if case in LONGFORM_FULL:
        assert np.array_equal(segs.numpy(), gold[f"seg_{key}"]) and np.array_equal(ola.numpy(), gold[f"ola_{key}"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", LONGFORM_CASES)
def test_kernels_bit_exact(gold, case):
    from mdctgan_b200 import longform as LF

    L, seg, ov = case
    key = f"{L}_{seg}_{ov}"
    dev = torch.device("cuda:0")
    clip = longform_clip(L, seg, ov)
    segs = LF.seg_pad_audio(clip.to(dev), seg, ov)
    ref = LO.seg_pad_audio(clip, seg, ov)
    assert torch.equal(segs.cpu(), ref)
    for dt in (torch.float64, torch.float32):
        got = LF.overlap_add(segs.to(dt).reshape(-1, 1, 1, seg), ov)
        want = LO.overlap_add(ref.to(dt).reshape(-1, 1, 1, seg), seg, ov)
        assert got.shape == want.shape and torch.equal(got.cpu(), want)
    idx = np.linspace(0, want.shape[-1] - 1, 257).astype(np.int64)
    got64 = LF.overlap_add(segs.double(), ov).cpu()
    assert np.array_equal(got64[0, idx].numpy(), gold[f"olaprobe_{key}"])


@pytest.mark.gpu
def test_longform_generator_matches_per_segment_inference():
    """LongFormGenerator (graphs, batches) == the reference loop: segment, inference per batch, overlap-add."""
    from test_model_gpu import our_opt

    from mdctgan_b200 import longform as LF
    from mdctgan_b200.models.models import create_model

    opt = our_opt("inf_small", gpu="0")
    opt.isTrain = False
    opt.checkpoints_dir, opt.name = "/tmp/mdctgan_lf", "lf"
    torch.manual_seed(11)
    trainer_opt = our_opt("inf_small", gpu="0")
    trainer_opt.checkpoints_dir, trainer_opt.name = "/tmp/mdctgan_lf", "lf"
    m0 = create_model(trainer_opt)
    m0.save("latest")
    model = create_model(opt)
    model.eval()
    dev = model.device
    clip = longform_clip(30000, 3840, 128).to(dev)
    for ov in (0, 128):
        gen = LF.LongFormGenerator(model, batch_size=4)
        got = gen(clip, 3840, ov)
        segs = LO.seg_pad_audio(clip.cpu(), 3840, ov)
        outs = [model.inference(segs[i:i + 4].to(dev))[1].cpu() for i in range(0, segs.shape[0], 4)]
        want = LO.overlap_add(torch.cat(outs, dim=0), 3840, ov)
        assert got.shape == want.shape
        assert torch.allclose(got.cpu(), want.to(got.dtype), rtol=0, atol=1e-6)


@pytest.mark.parametrize("case", [(100000, 7936, 256), (2880000, 32512, 256), (9000, 3840, 128), (63488, 7936, 0)])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_shard_layout_covers_the_clip(case, world):
    """Host logic of the segment-sharded generation: contiguous runs, every segment owned once, the shard outputs
    tile the assembled clip with exactly `overlap` shared samples at interior edges."""
    from mdctgan_b200 import longform as LF

    L, seg, ov = case
    n = LO.seg_pad_audio(longform_clip(L, seg, ov), seg, ov).shape[0]
    runs = [LF.shard_segments(n, r, world) for r in range(world)]
    assert runs[0][0] == 0 and runs[-1][1] == n and all(runs[i][1] == runs[i + 1][0] for i in range(world - 1))
    step = seg - ov
    total = (n - 1) * step + seg - 2 * ov
    pos = 0
    for lo, hi in runs:
        if hi <= lo:
            continue
        length = (hi - lo - 1) * step + seg - (ov if lo == 0 else 0) - (ov if hi == n else 0)
        off = LF.shard_offset(lo, seg, ov)
        assert off == (pos - ov if lo > 0 else 0)
        pos = off + length
    assert pos == total


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(100000, 7936, 256), (9000, 3840, 128), (63488, 7936, 0), (2880000, 32512, 256)])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_overlap_add_bit_exact(case, world):
    """Every rank folds its own contiguous run of segments; summing the `overlap` shared samples at the shard edges reproduces
    the single-device fold bit for bit (the same two-term sums)."""
    from mdctgan_b200 import longform as LF

    L, seg, ov = case
    dev = torch.device("cuda:0")
    segs = LF.seg_pad_audio(longform_clip(L, seg, ov).to(dev), seg, ov)
    n = segs.shape[0]
    for dt in (torch.float32, torch.float64):
        y = (segs * 1.25 + 0.01).to(dt)          # stands for the generated segments
        whole = LF.overlap_add(y, ov)
        parts = []
        for r in range(world):
            lo, hi = LF.shard_segments(n, r, world)
            if hi > lo:
                parts.append((LF.overlap_add(y[lo:hi], ov, (ov if lo == 0 else 0, ov if hi == n else 0)), lo))
        got = LF.stitch_shards(parts, seg, ov)
        assert got.shape == whole.shape and torch.equal(got, whole)
