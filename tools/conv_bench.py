"""Per-launch device time of single convolution layers, measured as a CUDA graph of back-to-back launches
(no host launch gaps), for the tcgen05 and the direct engines.  python tools/conv_bench.py"""
import sys

import torch

sys.path.insert(0, ".")
from mdctgan_b200 import nn_ops as ops  # noqa: E402

LAYERS = [
    # name, B, Cin, H, W, Cout, k, stride, pad, reflect, transposed
    ("res 256->256 @4x32", 4, 256, 4, 32, 256, 3, 1, 1, True, False),
    ("down 32->64 s2 @32x256", 4, 32, 32, 256, 64, 3, 2, 1, False, False),
    ("down 64->128 s2", 4, 64, 16, 128, 128, 3, 2, 1, False, False),
    ("down 128->256 s2", 4, 128, 8, 64, 256, 3, 2, 1, False, False),
    ("up 256->128 T", 4, 256, 4, 32, 128, 3, 2, 1, False, True),
    ("up 128->64 T", 4, 128, 8, 64, 64, 3, 2, 1, False, True),
    ("up 64->32 T", 4, 64, 16, 128, 32, 3, 2, 1, False, True),
    ("res 512->512 @2x16 b8", 8, 512, 2, 16, 512, 3, 1, 1, True, False),
    ("res 64->64 @16x128 b8", 8, 64, 16, 128, 64, 3, 1, 1, True, False),
    ("D 64->128 k4 s2 @17x129", 4, 64, 17, 129, 128, 4, 2, 2, False, False),
    ("res 512->512 @2x16 b4 (cfg4 bottleneck)", 4, 512, 2, 16, 512, 3, 1, 1, True, False),
]


def time_layer(dev, spec, engine, nlayers=8, reps=20):
    name, B, Cin, H, W, Cout, k, stride, pad, reflect, transposed = spec
    ops.CONV_ENGINE = engine
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    ws = []
    for _ in range(nlayers):          # distinct weights per launch, like consecutive layers of the network
        w = (torch.randn((Cin, Cout, k, k) if transposed else (Cout, Cin, k, k), generator=g) * 0.05).to(dev)
        kn = ops.pack_conv_weight(w, transposed)
        ws.append((kn, ops.pack_conv_weight_umma(kn) if engine != "direct" else None))
    bias = torch.zeros(Cout, device=dev)
    scale = torch.ones(B * Cin, device=dev)
    shift = torch.zeros(B * Cin, device=dev)
    f = ops.Feat(x, scale=scale, shift=shift, per_sample=True, act=ops.ACT_RELU)
    kw = dict(kh=k, kw=k, stride=stride, pad=pad, pad_mode=ops.PAD_REFLECT if reflect else ops.PAD_ZERO, transposed=transposed,
              output_padding=1 if transposed else 0, want_stats=True)

    def burst():
        for kn, um in ws:
            with ops.stats_pass(dev):
                ops.conv2d(f, kn, bias, w_umma=um, **kw)

    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        burst()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        burst()
    for _ in range(3):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    # each launch is preceded by the arena memset of its stats_pass: subtract nothing, report it as is
    us = e0.elapsed_time(e1) * 1e3 / (reps * nlayers)
    Ho = (H - 1) * stride - 2 * pad + k + 1 if transposed else (H + 2 * pad - k) // stride + 1
    Wo = (W - 1) * stride - 2 * pad + k + 1 if transposed else (W + 2 * pad - k) // stride + 1
    flops = 2.0 * B * Ho * Wo * Cout * k * k * Cin / (stride * stride if transposed else 1)
    return us, flops


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    for spec in LAYERS:
        row = [f"{spec[0]:26s}"]
        for engine in ("umma", "tf32", "direct"):
            us, flops = time_layer(dev, spec, engine)
            row.append(f"{engine} {us:7.1f} us ({flops / us / 1e6:6.1f} TF/s)")
        print("  ".join(row), flush=True)
