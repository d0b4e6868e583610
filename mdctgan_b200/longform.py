"""Long-form super-resolution of a whole clip on the device (reference: generate_audio.py:29-53 +
AudioTestDataset.seg_pad_audio, data/audio_dataset.py:153-167).

The reference cuts the clip into fixed segments on the CPU, runs `model.inference` per DataLoader batch, copies every
result to the host and overlap-adds there with `fold`.  Here the clip is segmented by one kernel, the batches run
back to back on the device (a CUDA graph per batch shape), and the generated segments are overlap-added by one kernel;
one D2H copy at the end.  Segment boundaries are kept exactly (InstanceNorm makes segmentation part of the numerics)."""
from __future__ import annotations

from ctypes import c_int, c_int64, c_void_p

import torch

from . import _lib
from . import nn_ops as ops

_bound = False


def _L():
    global _bound
    L = ops._L()
    if not _bound:
        L.mdctgan_segment_count.restype = c_int64
        L.mdctgan_segment_count.argtypes = [c_int64, c_int, c_int]
        L.mdctgan_segment_gather.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p]
        L.mdctgan_segment_ola.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]
        L.mdctgan_segment_ola_part.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]
        _bound = True
    return L


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def seg_pad_audio(audio: torch.Tensor, segment_length: int, overlap: int = 0) -> torch.Tensor:
    """[L] or [1, L] fp32 CUDA -> [n_seg, segment_length] (AudioTestDataset.seg_pad_audio)."""
    if not audio.is_cuda:
        raise RuntimeError("seg_pad_audio: expected a CUDA tensor; mdctgan_b200 has no CPU path")
    a = audio.reshape(-1).to(torch.float32).contiguous()
    n = int(_L().mdctgan_segment_count(a.numel(), segment_length, overlap))
    out = torch.empty((n, segment_length), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_L().mdctgan_segment_gather(a.data_ptr(), a.numel(), out.data_ptr(), n, segment_length, overlap, _stream(a)))
    return out


def overlap_add(segments: torch.Tensor, overlap: int = 0, crop=None) -> torch.Tensor:
    """generated segments [n_seg, (1, 1,) seg] fp32 / fp64 CUDA -> [1, L'] (generate_audio.py:40-53).
    `crop` = (begin, end) samples cropped from the fold (default (overlap, overlap) = the reference's whole-clip form); a run of
    segments that is one shard of a clip keeps its interior edges: crop 0 there (see `stitch_shards`)."""
    if not segments.is_cuda:
        raise RuntimeError("overlap_add: expected a CUDA tensor; mdctgan_b200 has no CPU path")
    x = segments.reshape(segments.shape[0], segments.shape[-1]).contiguous()
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float32)
    n, seg = x.shape
    cb, ce = (overlap, overlap) if crop is None else (int(crop[0]), int(crop[1]))
    out = torch.empty((1, (n - 1) * (seg - overlap) + seg - cb - ce), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_L().mdctgan_segment_ola_part(x.data_ptr(), out.data_ptr(), n, seg, overlap, cb, ce,
                                                 _lib.F64 if x.dtype == torch.float64 else _lib.F32, _stream(x)))
    return out


def shard_segments(n_seg: int, rank: int, world: int):
    """Contiguous run [lo, hi) of the clip's segments owned by `rank` (parallel.shard_range)."""
    from .parallel import shard_range

    r = shard_range(n_seg, world, rank)
    return r.start, r.stop


def shard_offset(lo: int, segment_length: int, overlap: int) -> int:
    """Index in the assembled clip of the first sample of the shard whose run starts at segment `lo`."""
    return lo * (segment_length - overlap) - (overlap if lo > 0 else 0)


def stitch_shards(parts, segment_length: int, overlap: int) -> torch.Tensor:
    """Assemble the per-rank outputs of `LongFormGenerator.generate_shard`: parts = [(audio [1, L_r], lo_r), ...] in any order,
    all on one device / host.  Neighbouring shards overlap by `overlap` half-weighted samples, which are summed (the same two-term
    sum the single-device fold performs, so the result is bit-identical to it)."""
    parts = [(a, lo) for a, lo in parts if a is not None and a.numel()]
    end = max(shard_offset(lo, segment_length, overlap) + a.shape[-1] for a, lo in parts)
    out = torch.zeros((1, end), dtype=parts[0][0].dtype, device=parts[0][0].device)
    for a, lo in sorted(parts, key=lambda t: t[1]):
        o = shard_offset(lo, segment_length, overlap)
        out[:, o:o + a.shape[-1]] += a.to(out.device)
    return out


class LongFormGenerator:
    """`generate_audio.py` as one call: clip -> segments -> model.inference in batches -> overlap-add."""

    def __init__(self, model, batch_size: int = 16, use_graph: bool = True):
        self.model, self.batch_size, self.use_graph = model, batch_size, use_graph
        self._graphs = {}

    def _infer(self, batch: torch.Tensor) -> torch.Tensor:
        if not self.use_graph:
            return self.model.inference(batch)[1]
        from .runtime import GraphedInference

        key = tuple(batch.shape)
        if key not in self._graphs:
            self._graphs[key] = GraphedInference(self.model, batch.shape[0], batch.shape[1], warmup=2)
        return self._graphs[key](batch)[1]

    @torch.no_grad()
    def __call__(self, lr_audio: torch.Tensor, segment_length: int = None, gen_overlap: int = None) -> torch.Tensor:
        m = self.model
        seg = int(segment_length if segment_length is not None else m.opt.segment_length)
        ov = int(gen_overlap if gen_overlap is not None else getattr(m.opt, "gen_overlap", 0))
        return self.generate_shard(lr_audio, seg, ov, 0, 1)[0]

    @torch.no_grad()
    def generate_shard(self, lr_audio: torch.Tensor, segment_length: int, gen_overlap: int, rank: int, world: int):
        """This rank's share of the clip: the contiguous run [lo, hi) of its segments is generated and folded on this device, no
        collective.  Returns (audio [1, L_r], lo); `stitch_shards` assembles the ranks' outputs.  world = 1 is the whole clip."""
        m = self.model
        seg, ov = int(segment_length), int(gen_overlap)
        segs = seg_pad_audio(lr_audio.to(m.device), seg, ov)
        n = segs.shape[0]
        if lr_audio.numel() < seg:
            ov = 0                                   # a clip shorter than one segment is end-padded only (audio_dataset.py:163-166)
        lo, hi = shard_segments(n, rank, world)
        self.last_segments = (hi - lo, n)
        if hi <= lo:
            return torch.zeros((1, 0), dtype=torch.float32, device=m.device), lo
        outs = []
        for i in range(lo, hi, self.batch_size):
            outs.append(self._infer(segs[i:min(i + self.batch_size, hi)]).reshape(-1, seg).clone())
        crop = (ov if lo == 0 else 0, ov if hi == n else 0)
        return overlap_add(torch.cat(outs, dim=0), ov, crop), lo
