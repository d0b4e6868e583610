"""Flat parameter / gradient buckets and the fused Adam step.

Reference: two `torch.optim.Adam(params, lr=opt.lr, betas=(opt.beta1, 0.999))` (models/pix2pixHD_model.py:350-364)
stepped from train.py:185-202.  Here every parameter of a network is a view into ONE flat fp32 buffer (so are the
gradients: that buffer is also the NCCL all-reduce bucket) and `step()` is one kernel launch
(`mdctgan_adam_flat`) per contiguous run of trainable parameters.  The optimizer keeps the torch.optim.Optimizer
surface train.py touches: `zero_grad()`, `step()`, `param_groups[i]['lr']`, `state_dict()`.
"""
from __future__ import annotations

from typing import List

import torch

import os

from . import _lib
from . import nn_ops as ops

KN_LAYOUT = os.environ.get("MDCTGAN_KN_LAYOUT", "1") != "0"      # convolution weights stored [kh][kw][Cin][Cout] in the flat buckets


def _kn_views(module: torch.nn.Module):
    """id(weight) -> permutation that maps the kernel-side storage order [kh][kw][Cin][Cout] back to the parameter's logical shape,
    for every convolution weight of `module`.  In that order the master copy IS the [kh*kw*Cin][Cout] image the direct kernels
    read, the tcgen05 weight gradient writes whole 128-byte rows (co contiguous) instead of scattered 4-byte reductions, and
    both GEMM geometries (forward: MN-major, input gradient: K-major) see 128-byte runs.  The logical tensors (state_dict,
    load_state_dict, checkpoints: reference layout [Cout][Cin][kh][kw] / [Cin][Cout][kh][kw]) are unchanged -- only strides."""
    from .models.networks import Conv2d, ConvTranspose2d

    out = {}
    for m in module.modules():
        if isinstance(m, ConvTranspose2d):
            out[id(m.weight)] = ((2, 3, 0, 1), (2, 3, 0, 1))      # logical [ci][co][kh][kw] -> storage (kh, kw, ci, co); and back
        elif isinstance(m, Conv2d):
            out[id(m.weight)] = ((2, 3, 1, 0), (3, 2, 0, 1))      # logical [co][ci][kh][kw] -> storage (kh, kw, ci, co); and back
    return out


class FlatBucket:
    """All parameters of `module` re-pointed into one flat buffer (16-byte aligned segments), with a matching flat
    gradient buffer whose views are installed as `p.grad`.  Convolution weights are stored in the kernel-side order
    [kh][kw][Cin][Cout] (`_kn_views`); their `.data` / `.grad` are permuted views with the reference's logical shape."""

    @staticmethod
    def padded_numel(module: torch.nn.Module) -> int:
        return sum((p.numel() + 3) // 4 * 4 for p in module.parameters())

    def __init__(self, module: torch.nn.Module, grad_storage: torch.Tensor = None):
        """`grad_storage`: an existing flat fp32 tensor of padded_numel(module) elements to hold the gradients (lets
        several buckets share ONE buffer = one all-reduce per step)."""
        params = [p for p in module.parameters()]
        if not params:
            raise ValueError("FlatBucket: module has no parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatBucket: parameters must live on a CUDA device (no CPU path)")
        self.params: List[torch.nn.Parameter] = params
        self.offsets, off = [], 0
        for p in params:
            if p.dtype != torch.float32:
                raise RuntimeError("FlatBucket: fp32 master parameters only")
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        if grad_storage is None:
            grad_storage = torch.zeros(off, dtype=torch.float32, device=dev)
        assert grad_storage.numel() == off and grad_storage.dtype == torch.float32 and grad_storage.is_contiguous()
        self.grad = grad_storage
        self._kn = _kn_views(module) if KN_LAYOUT else {}
        with torch.no_grad():
            for p, o in zip(params, self.offsets):
                self._view(self.flat, p, o).copy_(p)
                p.data = self._view(self.flat, p, o)
                p.grad = self._view(self.grad, p, o)

    def _view(self, flat: torch.Tensor, p: torch.Tensor, o: int) -> torch.Tensor:
        """The logical-shape view of parameter `p`'s segment of a flat buffer (permuted for convolution weights)."""
        seg = flat[o:o + p.numel()]
        perm = self._kn.get(id(p))
        if perm is None:
            return seg.view_as(p)
        to_storage, to_logical = perm
        return seg.view([p.shape[i] for i in to_storage]).permute(*to_logical)

    def trainable_runs(self, subset=None):
        """Maximal contiguous [begin, end) runs of the flat buffer whose parameters require grad (and are in `subset`, a set of
        parameter ids, when given)."""
        runs, cur = [], None
        for p, o in zip(self.params, self.offsets):
            end = o + (p.numel() + 3) // 4 * 4
            if p.requires_grad and (subset is None or id(p) in subset):
                cur = [o, end] if cur is None else [cur[0], end]
            elif cur is not None:
                runs.append(tuple(cur))
                cur = None
        if cur is not None:
            runs.append(tuple(cur))
        return runs

    def reattach_grads(self):
        """Re-install the gradient views (after something set p.grad = None)."""
        for p, o in zip(self.params, self.offsets):
            p.grad = self._view(self.grad, p, o)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay, no amsgrad) over a FlatBucket.  `grad_scale` multiplies the
    gradient inside the kernel (1/world_size after a sum all-reduce)."""

    def __init__(self, bucket: FlatBucket, lr=2e-4, betas=(0.5, 0.999), eps=1e-8, graph_safe: bool = False, params=None):
        """`params`: optional subset of the bucket's parameters to optimise (the reference's --niter_fix_global phase trains
        only the local enhancer, pix2pixHD_model.py:333-347); the others keep receiving gradients but are not stepped."""
        self._subset = None if params is None else {id(p) for p in params}
        trainable = [p for p in bucket.params if p.requires_grad and (self._subset is None or id(p) in self._subset)]
        super().__init__(trainable, dict(lr=lr, betas=betas, eps=eps))
        self.bucket = bucket
        self.exp_avg = torch.zeros_like(bucket.flat)
        self.exp_avg_sq = torch.zeros_like(bucket.flat)
        self.step_count = 0
        self.grad_scale = 1.0
        # a device-side step counter lets a captured CUDA graph of the step replay with the right bias corrections
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=bucket.flat.device) if graph_safe else None

    def zero_grad(self, set_to_none: bool = True):
        """One memset of the flat gradient bucket; the `p.grad` views stay installed (the backward kernels
        accumulate straight into them)."""
        self.bucket.grad.zero_()
        for p in self.bucket.params:
            if p.grad is None:
                self.bucket.reattach_grads()
                break

    @torch.no_grad()
    def begin_step(self):
        """Advance the step counters (host and device) once per optimiser step, before the first `step_range`."""
        self.step_count += 1
        if self.step_dev is not None:
            b = self.bucket
            with torch.cuda.device(b.flat.device):
                _lib.check(ops._L().mdctgan_counter_inc(self.step_dev.data_ptr(), torch.cuda.current_stream(b.flat.device).cuda_stream))

    @torch.no_grad()
    def step_range(self, lo: int, hi: int):
        """Adam on the trainable parameters inside the flat range [lo, hi) -- one launch per contiguous run, on the current stream.
        The pipelined train step updates a range as soon as its gradients are complete (models/pix2pixHD_model.py)."""
        g = self.param_groups[0]
        b = self.bucket
        L = ops._L()
        with torch.cuda.device(b.flat.device):
            st = torch.cuda.current_stream(b.flat.device).cuda_stream
            for rlo, rhi in b.trainable_runs(self._subset):
                rlo, rhi = max(rlo, lo), min(rhi, hi)
                if rhi <= rlo:
                    continue
                _lib.check(L.mdctgan_adam_flat(b.flat[rlo:rhi].data_ptr(), b.grad[rlo:rhi].data_ptr(), self.exp_avg[rlo:rhi].data_ptr(),
                                               self.exp_avg_sq[rlo:rhi].data_ptr(), rhi - rlo, float(g["lr"]), float(g["betas"][0]),
                                               float(g["betas"][1]), float(g["eps"]), float(self.grad_scale), self.step_count,
                                               self.step_dev.data_ptr() if self.step_dev is not None else None, st))

    def end_step(self):
        for p in self.bucket.params:          # kernel-side weight images (packed / tcgen05) are keyed on the version counter
            p._version_bump = getattr(p, "_version_bump", 0) + 1

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedAdam.step(closure)")
        self.begin_step()
        self.step_range(0, self.bucket.numel)
        self.end_step()
        return None

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        if self.step_dev is not None:
            self.step_dev.fill_(self.step_count)
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
