"""Phase timeline of the tcgen05 convolution (debug aid): python tools/umma_trace.py"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from mdctgan_b200 import _lib, nn_ops as ops  # noqa: E402
from tools.conv_bench import LAYERS  # noqa: E402

NAMES = ["gt_ns", "start", "prologue", "1st chunk", "gather done", "acc ready", "staged", "cluster1", "stored", "stats", "cluster2", "mma issued"]


def trace(dev, spec, engine="umma"):
    name, B, Cin, H, W, Cout, k, stride, pad, reflect, transposed = spec
    ops.CONV_ENGINE = engine
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    w = (torch.randn((Cin, Cout, k, k) if transposed else (Cout, Cin, k, k), generator=g) * 0.05).to(dev)
    kn = ops.pack_conv_weight(w, transposed)
    um = ops.pack_conv_weight_umma(kn)
    f = ops.Feat(x, scale=torch.ones(B * Cin, device=dev), shift=torch.zeros(B * Cin, device=dev), per_sample=True, act=ops.ACT_RELU)
    kw = dict(kh=k, kw=k, stride=stride, pad=pad, pad_mode=ops.PAD_REFLECT if reflect else ops.PAD_ZERO, transposed=transposed,
              output_padding=1 if transposed else 0, want_stats=True)
    L = ops._L()
    L.mdctgan_conv2d_umma_set_trace.argtypes = [ctypes.c_void_p]
    buf = torch.zeros(4096 * 16, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.conv2d(f, kn, None, w_umma=um, **kw)
    torch.cuda.synchronize()
    L.mdctgan_conv2d_umma_set_trace(buf.data_ptr())
    ops.conv2d(f, kn, None, w_umma=um, **kw)
    torch.cuda.synchronize()
    L.mdctgan_conv2d_umma_set_trace(None)
    t = buf.cpu().view(-1, 16)
    t = t[t[:, 1] != 0]
    n = t.shape[0]
    gt = t[:, 0] - t[:, 0].min()
    rel = (t[:, 1:12] - t[:, 1:2]).double()
    print(f"== {name}: {n} CTAs; CTA start spread (globaltimer) min/median/max = {gt.min().item()}/{gt.median().item()}/{gt.max().item()} ns")
    med = rel.median(dim=0).values
    mx = rel.max(dim=0).values
    for i, nm in enumerate(NAMES[1:]):
        print(f"   {nm:12s} median {med[i].item():9.0f} cyc   max {mx[i].item():9.0f} cyc")


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    for idx in (int(a) for a in sys.argv[1:]) if len(sys.argv) > 1 else (0, 1, 6):
        trace(dev, LAYERS[idx])
        trace(dev, LAYERS[idx], "tf32")
