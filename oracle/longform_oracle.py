"""Oracle (CPU, torch) for long-form segmentation / overlap-add -- TEST INFRASTRUCTURE ONLY.

  seg_pad_audio   /root/reference/data/audio_dataset.py:153-167 (AudioTestDataset.seg_pad_audio), restated
  overlap_add     /root/reference/generate_audio.py:40-53 (script-level code: restated op for op)
Pinned by tests/golden/longform_golden.npz: `seg_*` entries come from the reference's own AudioTestDataset.seg_pad_audio
(tests/golden/make_golden.py longform), `ola_*` entries from the restated script block run on those segments.
"""
from math import ceil

import torch
import torch.nn.functional as F


def seg_pad_audio(audio: torch.Tensor, segment_length: int, overlap: int) -> torch.Tensor:
    audio = audio.squeeze(0)
    length = len(audio)
    if length >= segment_length:
        num_segments = int(ceil(length / segment_length))
        audio = F.pad(audio, (overlap, segment_length * num_segments - length + overlap), "constant")
        return audio.unfold(dimension=0, size=segment_length, step=segment_length - overlap)
    return F.pad(audio, (0, segment_length - length), "constant").unsqueeze(0)


def overlap_add(audio: torch.Tensor, segment_length: int, gen_overlap: int) -> torch.Tensor:
    """audio: [n_seg, 1, 1, seg] (what torch.cat of the sr_audio batches gives)."""
    n = audio.shape[0]
    stride = segment_length - gen_overlap
    if gen_overlap > 0:
        out_len = (n - 1) * stride + segment_length
        audio = audio.clone()
        audio[..., :gen_overlap] *= 0.5
        audio[..., -gen_overlap:] *= 0.5
        audio = audio.reshape(n, segment_length).transpose(-1, -2)
        audio = F.fold(audio, kernel_size=(1, segment_length), stride=(1, stride), output_size=(1, out_len)).squeeze(0)
        return audio[..., gen_overlap:-gen_overlap]
    return audio.reshape(1, -1)
