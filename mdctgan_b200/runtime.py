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

    def recapture(self):
        dev = self.lr_in.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(self._warmup, 2)):     # plans, attributes, weight packs, Adam state, NCCL warm
                self.model.train_step(self.lr_in, self.hr_in, self.world_size, self.all_reduce)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = self.model.train_step(self.lr_in, self.hr_in, self.world_size, self.all_reduce)
        self._lr = self.model.optimizer_G.param_groups[0]["lr"]

    def replay(self):
        if self.graph is None or self.model.optimizer_G.param_groups[0]["lr"] != self._lr:
            self.recapture()
        self.graph.replay()
        for o in (self.model.optimizer_G, self.model.optimizer_D):
            o.step_count += 1
        return self.losses

    def __call__(self, lr_audio: torch.Tensor, hr_audio: torch.Tensor):
        """lr_audio / hr_audio: [batch, samples] fp32, device or (pinned) host.  Returns the 4 losses
        [G_GAN, G_GAN_Feat, D_real, D_fake] as a device tensor (static buffer of the graph)."""
        self.lr_in.copy_(lr_audio, non_blocking=True)
        self.hr_in.copy_(hr_audio, non_blocking=True)
        return self.replay()
