"""The GAN train step of the reference as one explicit graph over the kernels of libmdctgan_b200.so.

Reference: Pix2PixHDModel._forward (models/pix2pixHD_model.py:416-451) builds, with torch autograd,

    sr = netG(cat(lr, |lr|*2+lo)) [+ lr]                                   (:394-414)
    D_fake     = sum_scales MSE(netD(cat(lr, sr_in.detach()))[-1], 0)       (:429-431, networks.py:127-137)
    D_real     = sum_scales MSE(netD(cat(lr, hr_in))[-1], 1)                (:434-435)
    G_GAN      = sum_scales MSE(netD(cat(lr, sr_in))[-1], 1)                (:439-441)
    G_GAN_Feat = sum_i sum_j<last 1/num_D * 4/(n_layers_D+1) * lambda_feat * L1(fake_ij, real_ij.detach())   (:443-451)

and train.py:175-202 backpropagates loss_G = G_GAN + G_GAN_Feat into G and loss_D = 0.5 (D_fake + D_real) into D.

Here: the detached and the attached fake passes of D are the same numbers, so D runs ONCE on the batch
[fake ; real] (InstanceNorm is per sample, so batching the two passes is exact); the forward records two tapes
(nn_ops.Tape); `backward_G` sweeps the D tape on the fake half for input gradients only (the reference also
produces D weight gradients there and throws them away, train.py:194) and then the G tape; `backward_D` sweeps the
D tape on both halves with weight gradients.  Weight gradients land in the flat buckets (optim.FlatBucket).
"""
from __future__ import annotations

from typing import List, Optional

import ctypes

import torch

from . import _lib
from . import nn_ops as ops
from .nn_ops import Feat

LOSS_NAMES = ["G_GAN", "G_GAN_Feat", "D_real", "D_fake"]


def _st(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _LossItem(ctypes.Structure):
    _fields_ = [("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("g", ctypes.c_void_p), ("n", ctypes.c_int64), ("coef", ctypes.c_double),
                ("gscale", ctypes.c_void_p), ("target", ctypes.c_float), ("kind", ctypes.c_int32), ("slot", ctypes.c_int32), ("pad_", ctypes.c_int32)]


MAX_LOSS_ITEMS = 24


def _multi_loss(L, items, acc_ptr, stream, backward=False):
    """items: list of dicts (a, b, g, n, coef, gscale, target, kind, slot).  One launch for up to MAX_LOSS_ITEMS of them."""
    for i in range(0, len(items), MAX_LOSS_ITEMS):
        chunk = items[i:i + MAX_LOSS_ITEMS]
        arr = (_LossItem * len(chunk))()
        for r, it in zip(arr, chunk):
            r.a, r.b, r.g, r.n, r.coef = it["a"], it.get("b"), it.get("g"), it["n"], it["coef"]
            r.gscale, r.target, r.kind, r.slot = it.get("gscale"), it.get("target", 0.0), it["kind"], it.get("slot", 0)
        if backward:
            _lib.check(L.mdctgan_multi_loss_bwd(arr, len(chunk), stream))
        else:
            _lib.check(L.mdctgan_multi_loss_fwd(arr, len(chunk), acc_ptr, 4, stream))


class GanGraph:
    """One evaluation of the training graph for a batch; holds the tapes until both backward sweeps ran."""

    def __init__(self, model):
        self.m = model
        self.tapeG: Optional[ops.Tape] = None
        self.tapeD: Optional[ops.Tape] = None

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, lr_audio: torch.Tensor, hr_audio: torch.Tensor):
        m = self.m
        if not m._two_channel:
            raise NotImplementedError("train step: only the --abs_spectro --arcsinh_transform input encoding (every shipped script) is built")
        L = ops._L()
        if m.no_lsgan and not m.no_ganFeat_loss:
            # reference: with the feature-matching outputs on (getIntermFeat) the sigmoid stage is never applied (networks.py:686), and
            # nn.BCELoss rejects the raw logits -- fail like it does instead of training on something else
            raise RuntimeError("--no_lsgan needs --no_ganFeat_loss: with intermediate features the reference's discriminator returns logits "
                               "(networks.py:671-686) and nn.BCELoss raises on values outside [0, 1]")
        dev = m.device
        lr_spectro, lr_input, _, _ = m._lr_input(lr_audio)              # [B,1,F,N] view of [B,2,F,N]
        hr_spectro, _, _ = m.preprocess.hr_forward(hr_audio)             # [B,1,F,N]
        B, _, Fr, N = lr_input.shape
        self.B, self.Fr, self.N = B, Fr, N
        lo = float(m.norm_range[0])
        self.tapeG, self.tapeD = ops.Tape(), ops.Tape()
        with torch.cuda.device(dev):
            with ops.recording(self.tapeG):
                self.g_out = m.netG.run(ops.to_nhwc(lr_input))           # [B,F,N,1], tanh applied
            sr = self.g_out.x.view(B, 1, Fr, N)
            if m.fit_residual:
                sr = ops.residual_scale_add(sr, lr_spectro, 0, 1.0)       # training: plain sum (:407-408)
            self.sr_spectro, self.lr_spectro, self.hr_spectro = sr, lr_spectro, hr_spectro
            # discriminator input, NHWC [2B,F,N,3]: rows 0..B-1 fake, B..2B-1 real
            din = torch.empty((2 * B, Fr, N, 3), dtype=torch.float32, device=dev)
            lr_stride = lr_spectro.stride(0)
            _lib.check(L.mdctgan_disc_input_fwd(lr_spectro.data_ptr(), lr_stride, sr.data_ptr(), din.data_ptr(), B, Fr * N, lo, _st(din)))
            _lib.check(L.mdctgan_disc_input_fwd(lr_spectro.data_ptr(), lr_stride, hr_spectro.data_ptr(), din[B:].data_ptr(), B, Fr * N, lo,
                                                _st(din)))
            self.din = Feat(din)
            with ops.recording(self.tapeD):
                self.feats: List[List[Feat]] = m.netD.run_features(self.din)
            # losses -> acc[4] (float64) -> fp32 0-dim tensors
            acc = torch.zeros(4, dtype=torch.float64, device=dev)
            st = _st(acc)
            a = acc.data_ptr()
            self.feat_coef = (1.0 / m.num_D) * (4.0 / (m.n_layers_D + 1)) * float(m.lambda_feat)
            gan_kind = 2 if m.no_lsgan else 0                                 # GANLoss: nn.BCELoss / nn.MSELoss (networks.py:105-108)
            items = []
            for scale in self.feats:
                pred = scale[-1].x                                        # [2B,h,w,1]
                n = pred[:B].numel()
                fake_p, real_p = pred.data_ptr(), pred[B:].data_ptr()
                items.append(dict(a=fake_p, n=n, target=1.0, coef=1.0 / n, kind=gan_kind, slot=0))      # G_GAN
                items.append(dict(a=real_p, n=n, target=1.0, coef=1.0 / n, kind=gan_kind, slot=2))      # D_real
                items.append(dict(a=fake_p, n=n, target=0.0, coef=1.0 / n, kind=gan_kind, slot=3))      # D_fake
                if not m.no_ganFeat_loss:
                    for f in scale[:-1]:
                        nf = f.x[:B].numel()
                        items.append(dict(a=f.x.data_ptr(), b=f.x[B:].data_ptr(), n=nf, coef=self.feat_coef / nf, kind=1, slot=1))
            _multi_loss(L, items, a, st)                                      # every loss term of the step: one launch
            self.losses = torch.empty(4, dtype=torch.float32, device=dev)
            _lib.check(L.mdctgan_f64_to_f32(a, self.losses.data_ptr(), 4, st))
        return self.losses

    # ------------------------------------------------------------------ backward sweeps
    @torch.no_grad()
    def backward_G(self, g_gan: Optional[torch.Tensor] = None, g_feat: Optional[torch.Tensor] = None, use_gan=True, use_feat=True, join=True,
                   after_D=None):
        """d(g_gan * G_GAN + g_feat * G_GAN_Feat) / d(G parameters) accumulated into their .grad (the flat bucket).
        g_*: 0-dim CUDA tensors (upstream gradients of the loss tensors) or None = 1."""
        m, B = self.m, self.B
        L = ops._L()
        G = ops.GradMap()
        with torch.cuda.device(m.device):
            items = []
            for scale in self.feats:
                pred = scale[-1]
                if use_gan:
                    g = torch.empty_like(pred.x[:B])
                    n = g.numel()
                    items.append(dict(a=pred.x.data_ptr(), g=g.data_ptr(), n=n, target=1.0, coef=1.0 / n, gscale=ops._ptr(g_gan),
                                      kind=2 if m.no_lsgan else 0))
                    G.add(pred, g)
                if use_feat and not m.no_ganFeat_loss:
                    for f in scale[:-1]:
                        g = torch.empty_like(f.x[:B])
                        n = g.numel()
                        items.append(dict(a=f.x.data_ptr(), b=f.x[B:].data_ptr(), g=g.data_ptr(), n=n, coef=self.feat_coef / n,
                                          gscale=ops._ptr(g_feat), kind=1))
                        G.add(f, g)
            if items:
                _multi_loss(L, items, None, _st(self.din.x), backward=True)     # every gradient seed of the sweep: one launch
            self.tapeD.backward(G, wgrad=False, nb=B, join=False)
            if after_D is not None:                                       # from here on nothing reads the discriminator's weights
                after_D()
            gin = G.pop(self.din)                                         # [B,F,N,3]
            dsr = torch.empty((B, self.Fr, self.N, 1), dtype=torch.float32, device=m.device)
            _lib.check(L.mdctgan_disc_input_bwd(gin.data_ptr(), self.sr_spectro.data_ptr(), dsr.data_ptr(), dsr.numel(), _st(dsr)))
            GG = ops.GradMap()
            GG.add(self.g_out, dsr)                                       # fit_residual's "+ lr" passes the gradient through
            self.tapeG.backward(GG, wgrad=True, nb=None, join=False)
            if join:
                ops.join_side_work(m.device)

    @torch.no_grad()
    def backward_D(self, g_real: Optional[torch.Tensor] = None, g_fake: Optional[torch.Tensor] = None, join=True):
        """d(g_real * D_real + g_fake * D_fake) / d(D parameters) accumulated into their .grad."""
        m, B = self.m, self.B
        L = ops._L()
        G = ops.GradMap()
        self.din.needs_grad = False                                       # the inputs are data / detached here
        try:
            with torch.cuda.device(m.device):
                items = []
                kind = 2 if m.no_lsgan else 0
                for scale in self.feats:
                    pred = scale[-1]
                    g = torch.empty_like(pred.x)
                    n = pred.x[:B].numel()
                    items.append(dict(a=pred.x.data_ptr(), g=g.data_ptr(), n=n, target=0.0, coef=1.0 / n, gscale=ops._ptr(g_fake), kind=kind))
                    items.append(dict(a=pred.x[B:].data_ptr(), g=g[B:].data_ptr(), n=n, target=1.0, coef=1.0 / n, gscale=ops._ptr(g_real), kind=kind))
                    G.add(pred, g)
                _multi_loss(L, items, None, _st(self.din.x), backward=True)
                self.tapeD.backward(G, wgrad=True, nb=None, join=False)
                if join:
                    ops.join_side_work(m.device)
        finally:
            self.din.needs_grad = True

    def release(self):
        self.tapeG = self.tapeD = None
        self.feats = None


# ---------------------------------------------------------------------------------------- autograd glue
class _GLosses(torch.autograd.Function):
    """(G_GAN, G_GAN_Feat) as differentiable 0-dim tensors: `.backward()` runs GanGraph.backward_G."""

    @staticmethod
    def forward(ctx, anchor, graph: GanGraph):
        ctx.graph = graph
        return graph.losses[0].clone(), graph.losses[1].clone()

    @staticmethod
    def backward(ctx, g0, g1):
        gr = ctx.graph
        seg = getattr(gr, "_api_segments", None)
        if seg is not None:          # captured sweep (runtime.GraphedAPI)
            from .runtime import GraphedAPI

            GraphedAPI.backward(seg, "G", (g0, g1))
            return None, None
        gr.backward_G(g0.contiguous() if g0 is not None else None, g1.contiguous() if g1 is not None else None,
                      use_gan=g0 is not None, use_feat=g1 is not None)
        return None, None


class _DLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, graph: GanGraph):
        ctx.graph = graph
        return graph.losses[2].clone(), graph.losses[3].clone()

    @staticmethod
    def backward(ctx, g_real, g_fake):
        gr = ctx.graph
        seg = getattr(gr, "_api_segments", None)
        if seg is not None:          # captured sweep (runtime.GraphedAPI)
            from .runtime import GraphedAPI

            GraphedAPI.backward(seg, "D", (g_real, g_fake))
            return None, None
        z = None
        if g_real is None or g_fake is None:
            z = torch.zeros((), dtype=torch.float32, device=gr.losses.device)
        gr.backward_D((g_real if g_real is not None else z).contiguous(), (g_fake if g_fake is not None else z).contiguous())
        return None, None


def loss_tensors(graph: GanGraph, anchor_g: torch.Tensor, anchor_d: torch.Tensor):
    """The four reference losses [G_GAN, G_GAN_Feat, D_real, D_fake] wired to the two backward sweeps.  The anchors
    are parameters (requires_grad) that tie the nodes into autograd; their .grad is written by our kernels, the
    autograd engine itself accumulates nothing."""
    with torch.enable_grad():
        g_gan, g_feat = _GLosses.apply(anchor_g, graph)
        d_real, d_fake = _DLosses.apply(anchor_d, graph)
    return [g_gan, g_feat, d_real, d_fake]
