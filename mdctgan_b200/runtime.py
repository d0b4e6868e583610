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
