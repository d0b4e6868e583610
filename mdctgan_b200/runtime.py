"""Launch-bound inner loops captured in CUDA graphs.

`model.inference` at the reference's batch sizes is ~70 short kernels; issued eagerly from Python the host
is the bottleneck.  `GraphedInference` captures one inference pass for a fixed (batch, segment length) into a
CUDA graph and replays it: the kernels are exactly the ones the eager path launches (same C-ABI calls on the
capturing stream), the parameters are read in place, so weight updates / `load_state_dict` need no
re-capture (only re-packing, which happens outside the graph when a weight version changes -> re-capture
is triggered automatically)."""
from __future__ import annotations

import torch


class GraphedInference:
    def __init__(self, model, batch: int, samples: int, warmup: int = 3):
        self.model = model
        dev = model.device
        self.static_in = torch.zeros(batch, samples, dtype=torch.float32, device=dev)
        self._versions = None
        self._capture(warmup)

    def _weight_versions(self):
        return tuple(p._version for p in self.model.netG.parameters())

    def _capture(self, warmup: int):
        side = torch.cuda.Stream(self.static_in.device)
        side.wait_stream(torch.cuda.current_stream(self.static_in.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):          # first calls create plans, set attributes, pack weights
                self.model.inference(self.static_in)
        torch.cuda.current_stream(self.static_in.device).wait_stream(side)
        torch.cuda.synchronize(self.static_in.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self.model.inference(self.static_in)
        self._versions = self._weight_versions()

    def replay(self):
        """Run the captured pass on whatever `static_in` holds; returns the static outputs
        (sr_spectro, sr_audio, lr_pha, lr_norm_param, lr_spectro)."""
        if self._weight_versions() != self._versions:
            self._capture(1)
        self.graph.replay()
        return self.static_out

    def __call__(self, lr_audio: torch.Tensor):
        self.static_in.copy_(lr_audio, non_blocking=True)
        return self.replay()


class GraphedTrainStep:
    """One whole training iteration (train.py:160-202) captured in a CUDA graph and replayed: ~600 kernel launches of
    libmdctgan_b200.so (+ the NCCL all-reduce of the flat gradient bucket when world_size > 1) per replay, no Python in
    between.  The Adam step counter lives on the device (optim.FusedAdam graph_safe), the re-packing of the updated
    weights into the kernel-side images is part of the captured sequence.  A learning-rate change needs a re-capture
    (`recapture()`; the reference changes it once per epoch, pix2pixHD_model.py:664-673)."""

    def __init__(self, model, batch: int, samples: int, world_size: int = 1, all_reduce=None, warmup: int = 3):
        self.model, self.world_size, self.all_reduce = model, world_size, all_reduce
        dev = model.device
        self.lr_in = torch.zeros(batch, samples, dtype=torch.float32, device=dev)
        self.hr_in = torch.zeros(batch, samples, dtype=torch.float32, device=dev)
        self._warmup = warmup
        self.graph = None

    # ---- warm-up and capture must not train: everything a step mutates is snapshotted before and put back after
    def _state(self):
        m = self.model
        opts = (m.optimizer_G, m.optimizer_D)
        return {"flat": [m.bucket_G.flat.clone(), m.bucket_D.flat.clone()],
                "adam": [(o.exp_avg.clone(), o.exp_avg_sq.clone(), None if o.step_dev is None else o.step_dev.clone(), o.step_count) for o in opts],
                "buffers": [b.clone() for net in (m.netG, m.netD) for b in net.buffers()]}

    def _restore(self, st):
        m = self.model
        with torch.no_grad():
            m.bucket_G.flat.copy_(st["flat"][0])
            m.bucket_D.flat.copy_(st["flat"][1])
            for o, (ea, eas, sd, sc) in zip((m.optimizer_G, m.optimizer_D), st["adam"]):
                o.exp_avg.copy_(ea)
                o.exp_avg_sq.copy_(eas)
                if sd is not None:
                    o.step_dev.copy_(sd)
                o.step_count = sc
            bufs = [b for net in (m.netG, m.netD) for b in net.buffers()]
            for b, v in zip(bufs, st["buffers"]):
                b.copy_(v)
        m.packer.refresh()          # kernel-side weight images of the restored weights (the captured step re-packs at its end)
        m._packed_at = m._pack_key()

    def recapture(self):
        dev = self.lr_in.device
        if any(o.step_dev is None for o in (self.model.optimizer_G, self.model.optimizer_D)):
            raise RuntimeError("GraphedTrainStep needs the device-side Adam step counter (opt.graph_safe_adam)")
        state = self._state()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(self._warmup, 2)):     # plans, attributes, weight packs, allocator pools, NCCL warm
                self.model.train_step(self.lr_in, self.hr_in, self.world_size, self.all_reduce)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = self.model.train_step(self.lr_in, self.hr_in, self.world_size, self.all_reduce)
        self._restore(state)        # also undoes the host-side step_count bump of the (not executed) captured step
        torch.cuda.synchronize(dev)
        self._key = self._capture_key()

    def _capture_key(self):
        m = self.model
        return (id(m.optimizer_G), id(m.optimizer_D), m.optimizer_G.param_groups[0]["lr"], m.optimizer_D.param_groups[0]["lr"])

    def replay(self):
        """Re-captures when the learning rate changed (update_learning_rate) or an optimiser was replaced (update_fixed_params)."""
        if self.graph is None or self._capture_key() != self._key:
            self.recapture()
        self.graph.replay()
        for o in (self.model.optimizer_G, self.model.optimizer_D):
            o.step_count += 1
        self.model._packed_at = self.model._pack_key()
        return self.losses

    def __call__(self, lr_audio: torch.Tensor, hr_audio: torch.Tensor):
        """lr_audio / hr_audio: [batch, samples] fp32, device or (pinned) host.  Returns the 4 losses
        [G_GAN, G_GAN_Feat, D_real, D_fake] as a device tensor (static buffer of the graph)."""
        self.lr_in.copy_(lr_audio, non_blocking=True)
        self.hr_in.copy_(hr_audio, non_blocking=True)
        return self.replay()


class _ApiSegments:
    """Captured segments of one (batch shape, trainable-set) key of GraphedAPI."""

    def __init__(self):
        self.calls = 0
        self.g_fwd = None
        self.graph = None            # train_ops.GanGraph whose tapes live in the graphs' memory pool
        self.bwd = {}                # ("G" | "D", which upstream gradients exist) -> (CUDAGraph, static upstream scalars)


class GraphedAPI:
    """The reference's own call sequence -- train.py:160-202 verbatim: `model._forward(lr, hr)` -> `loss_G.backward()` ->
    `optimizer_G.step()` -> `loss_D.backward()` -> `optimizer_D.step()` -- at CUDA-graph speed, with no change to the caller.
    Issued eagerly, one iteration is ~460 launches through ctypes and the host is the bottleneck (12.5 ms per cfg4 step on a B200).
    After `WARMUP` eager iterations on a given batch shape the three launch-heavy pieces are captured, each in the call that needs it,
    and replayed from then on:
        forward segment    2 MDCT launches, generator, discriminator on [fake ; real], the four loss reductions   (in `_forward`)
        generator sweep    what `loss_G.backward()` runs (train_ops.GanGraph.backward_G)
        discriminator sweep what `loss_D.backward()` runs
    The segments share one memory pool (the tapes recorded by the forward segment are read by the sweeps) and are replayed in capture
    order.  `zero_grad()`, both Adam steps and the weight-image refresh stay eager (a handful of launches).  The upstream gradients
    (1, 1, 0.5, 0.5 in train.py; the GradScaler's scale under --fp16) are copied into static device scalars before a sweep replays.
    Opt out with MDCTGAN_GRAPH_API=0."""

    WARMUP = 3

    def __init__(self, model):
        self.model = model
        self.segments = {}

    def _key(self, lr_audio, hr_audio):
        m = self.model
        trainable = hash(tuple(p.requires_grad for p in m.bucket_G.params) + tuple(p.requires_grad for p in m.bucket_D.params))
        return (tuple(lr_audio.shape), tuple(hr_audio.shape), trainable, m.netG.training, m.netD.training)

    def forward(self, lr_audio, hr_audio):
        """Returns the GanGraph of this iteration (losses computed), or None while the shape is still in its eager warm-up."""
        from . import nn_ops as ops
        from . import train_ops as T

        m = self.model
        if lr_audio.dim() != 2 or hr_audio.dim() != 2 or not lr_audio.is_cuda:
            return None
        seg = self.segments.setdefault(self._key(lr_audio, hr_audio), _ApiSegments())
        if seg.calls < self.WARMUP:
            seg.calls += 1
            return None
        dev = m.device
        if seg.g_fwd is None:
            seg.lr_in = torch.empty_like(lr_audio, dtype=torch.float32)
            seg.hr_in = torch.empty_like(hr_audio, dtype=torch.float32)
            seg.lr_in.copy_(lr_audio)
            seg.hr_in.copy_(hr_audio)
            torch.cuda.synchronize(dev)
            seg.g_fwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(seg.g_fwd):
                graph = T.GanGraph(m)
                with ops.stats_pass(dev):
                    graph.forward(seg.lr_in, seg.hr_in)
            graph._api_segments = seg
            seg.graph, seg.api = graph, self
        else:
            seg.lr_in.copy_(lr_audio, non_blocking=True)
            seg.hr_in.copy_(hr_audio, non_blocking=True)
        seg.g_fwd.replay()
        return seg.graph

    @staticmethod
    def backward(seg, which, grads):
        """`which` = "G" (grads = upstream of G_GAN, G_GAN_Feat) or "D" (upstream of D_real, D_fake); entries may be None."""
        dev = seg.graph.m.device
        key = (which, tuple(g is not None for g in grads))
        if key not in seg.bwd:
            statics = [None if g is None else torch.ones((), dtype=torch.float32, device=dev) for g in grads]
            for s_, g in zip(statics, grads):
                if s_ is not None:
                    s_.copy_(g.reshape(()))
            torch.cuda.synchronize(dev)
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg, pool=seg.g_fwd.pool()):
                if which == "G":
                    seg.graph.backward_G(statics[0], statics[1], use_gan=statics[0] is not None, use_feat=statics[1] is not None)
                else:
                    z = torch.zeros((), dtype=torch.float32, device=dev)
                    seg.graph.backward_D(statics[0] if statics[0] is not None else z, statics[1] if statics[1] is not None else z)
            seg.bwd[key] = (cg, statics)
        cg, statics = seg.bwd[key]
        for s_, g in zip(statics, grads):
            if s_ is not None:
                s_.copy_(g.reshape(()), non_blocking=True)
        cg.replay()
