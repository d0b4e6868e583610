"""GPU parity of the tcgen05 implicit-GEMM convolution (mdctgan_b200/csrc/conv_umma.cuh) through the C ABI.

Checked against (a) torch-CPU fp32 convolutions of the same layer (reference layers: models/networks.py
:327-350 stride-2 / transposed, :440-457 residual 3x3 reflect, :649-670 PatchGAN 4x4) and (b) the direct
fp32 FFMA kernel of the same library on the same device.

Tolerance: the 3xTF32 split keeps every product to ~2^-21 relative and accumulates in fp32, so the bar is the
fp32 one (2e-5 rel-L2 per layer vs a CPU fp32 reference with a different summation order).  The single-pass
TF32 engine is checked at 2e-3 (10-bit mantissa inputs)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


UMMA_CASES = [
    # B, Cin, H, W, Cout, k, stride, pad, reflect, transposed
    (4, 256, 4, 32, 256, 3, 1, 1, True, False),    # cfg2 residual conv: 4 M-tiles x 4 N-tiles, cluster of 8 over K
    (1, 256, 4, 32, 256, 3, 1, 1, True, False),    # one M-tile
    (3, 64, 4, 32, 64, 3, 1, 1, True, False),      # M = 384
    (8, 512, 2, 16, 512, 3, 1, 1, True, False),    # cfg3 global bottleneck: 32-pixel planes, a tile spans 4 samples
    (2, 32, 16, 64, 64, 3, 2, 1, False, False),    # stride-2 downsample, zero padding
    (4, 128, 8, 64, 256, 3, 2, 1, False, False),
    (2, 64, 8, 16, 32, 3, 2, 1, False, True),      # ConvTranspose2d k3 s2 p1 op1, Cout = 32 (BN = 32)
    (4, 256, 4, 32, 128, 3, 2, 1, False, True),
    (2, 64, 5, 9, 32, 4, 2, 2, False, True),       # transposed k4 s2 p2 op1 -> 9x17: odd output, ragged parity classes (PatchGAN dgrad)
    (3, 128, 3, 17, 64, 4, 2, 2, False, True),     # 5x33 output
    (2, 64, 9, 17, 128, 4, 1, 2, False, False),    # PatchGAN 4x4 s1 p2, odd plane (M = 2*12*20 = 480: partial tile)
    (2, 64, 17, 33, 128, 4, 2, 2, False, False),   # PatchGAN 4x4 s2 p2
    (2, 128, 5, 9, 96, 1, 1, 0, False, False),     # 1x1 (BottleStack), Cout % 64 != 0, K = 128 (4 chunks)
    (1, 40, 7, 5, 32, 5, 1, 2, False, False),      # Cin % 32 != 0: K chunks straddle taps, K = 1000 padded to 1024
    (5, 32, 32, 256, 32, 3, 1, 1, True, False),    # many tiles (320), no split
    (4, 512, 2, 16, 512, 3, 1, 1, True, False),    # cfg4 bottleneck: BN = 128 tiles, 3-stage ring, cluster of 8
    (2, 128, 16, 64, 128, 3, 1, 1, True, False),   # BN = 128, 8 M-tiles per sample
    (2, 256, 5, 33, 512, 4, 1, 2, False, False),   # PatchGAN 4x4 s1 -> 512 channels (BN = 128), ragged last tile
    # planes smaller than a tile: tiles span samples (conv_umma.cuh `span`)
    (5, 64, 3, 10, 64, 3, 1, 1, True, False),      # 30-pixel planes: a tile straddles 5 samples, none aligned
    (6, 32, 2, 6, 96, 3, 1, 1, False, False),      # 12-pixel planes, 72 rows in all: one partial tile
    (4, 64, 2, 8, 32, 3, 2, 1, False, True),       # ConvTranspose2d, 16-pixel parity classes spanning the 4 samples
    (3, 64, 3, 5, 32, 4, 2, 2, False, True),       # ragged parity classes (5x9 output), spanning
    (7, 128, 3, 11, 128, 3, 2, 1, False, False),   # stride 2 -> 2x6 planes, 84 rows
]


def _torch_ref(x, w, bias, k, stride, pad, reflect, transposed):
    if transposed:
        return F.conv_transpose2d(x, w, bias, stride=stride, padding=pad, output_padding=1)
    if reflect:
        return F.conv2d(F.pad(x, (pad,) * 4, mode="reflect"), w, bias, stride=stride)
    return F.conv2d(x, w, bias, stride=stride, padding=pad)


@pytest.mark.parametrize("engine,tol", [("umma", 2e-5), ("tf32", 2e-3)])
@pytest.mark.parametrize("case", UMMA_CASES)
def test_conv2d_umma_matches_torch_and_direct(dev, case, engine, tol, monkeypatch):
    from mdctgan_b200 import nn_ops as ops

    monkeypatch.setattr(ops, "CONV_ENGINE", engine)
    B, Cin, H, W, Cout, k, stride, pad, reflect, transposed = case
    g = torch.Generator().manual_seed(abs(hash(case)) % 2**31)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn((Cin, Cout, k, k) if transposed else (Cout, Cin, k, k), generator=g) * 0.1
    bias = torch.randn(Cout, generator=g)
    ref = _torch_ref(x, w, bias, k, stride, pad, reflect, transposed)
    assert ops.umma_supported(Cin, Cout)
    w_kn = ops.pack_conv_weight(w.to(dev), transposed)
    w_um = ops.pack_conv_weight_umma(w_kn)
    kw = dict(kh=k, kw=k, stride=stride, pad=pad, pad_mode=ops.PAD_REFLECT if reflect else ops.PAD_ZERO, transposed=transposed,
              output_padding=1 if transposed else 0, want_stats=True)
    f = ops.Feat(_nhwc(x).to(dev))
    y = ops.conv2d(f, w_kn, bias.to(dev), w_umma=w_um, **kw)
    y_direct = ops.conv2d(f, w_kn, bias.to(dev), w_umma=None, **kw)
    torch.cuda.synchronize()
    got = _nchw(y.x.cpu())
    assert got.shape == ref.shape
    assert rel_l2(got.numpy(), ref.numpy()) < tol
    assert rel_l2(y.x.cpu().numpy(), y_direct.x.cpu().numpy()) < tol
    st = y.stats.cpu()
    rt = 1e-4 if engine == "umma" else 5e-3
    # plane sums cancel (zero-mean data): the absolute bar is the per-element bar (tol x rms) summed over the plane in quadrature, x4
    atol = 4.0 * tol * float(ref.pow(2).mean().sqrt()) * float(ref.shape[2] * ref.shape[3]) ** 0.5
    np.testing.assert_allclose(st[..., 0].numpy(), ref.double().sum(dim=(2, 3)).numpy(), rtol=rt, atol=max(atol, 1e-3))
    np.testing.assert_allclose(st[..., 1].numpy(), (ref.double() ** 2).sum(dim=(2, 3)).numpy(), rtol=rt)


@pytest.mark.parametrize("eager_norm", [False, True])
def test_conv2d_umma_fused_norm_act_chain_and_determinism(dev, monkeypatch, eager_norm):
    """ResnetBlock through the tensor-core path: x + IN(conv(relu(IN(conv(x))))), deferred norm applied in the gather."""
    from mdctgan_b200 import nn_ops as ops

    monkeypatch.setattr(ops, "CONV_ENGINE", "umma")
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 4, 128, 4, 32
    x = torch.randn(B, C, H, W, generator=g)
    w1, w2 = torch.randn(C, C, 3, 3, generator=g) * 0.05, torch.randn(C, C, 3, 3, generator=g) * 0.05
    b1, b2 = torch.randn(C, generator=g), torch.randn(C, generator=g)
    h = F.relu(F.instance_norm(F.conv2d(F.pad(x, (1,) * 4, mode="reflect"), w1, b1)))
    ref = x + F.instance_norm(F.conv2d(F.pad(h, (1,) * 4, mode="reflect"), w2, b2))

    def block():
        f = ops.Feat(_nhwc(x).to(dev))
        k1, k2 = ops.pack_conv_weight(w1.to(dev)), ops.pack_conv_weight(w2.to(dev))
        a = ops.conv2d(f, k1, b1.to(dev), kh=3, kw=3, pad=1, pad_mode=ops.PAD_REFLECT, want_stats=True, w_umma=ops.pack_conv_weight_umma(k1))
        a = ops.with_act(ops.finalize_norm(a, eager=eager_norm), ops.ACT_RELU)   # raw statistics, or explicit scale / shift
        c = ops.conv2d(a, k2, b2.to(dev), kh=3, kw=3, pad=1, pad_mode=ops.PAD_REFLECT, want_stats=True, w_umma=ops.pack_conv_weight_umma(k2))
        return ops.combine(f, ops.finalize_norm(c, eager=eager_norm)).x.cpu()

    out1, out2 = block(), block()
    assert rel_l2(_nchw(out1).numpy(), ref.numpy()) < 2e-5
    # the cluster reduction adds the K slices in rank order: the tensor itself is bit-reproducible; the
    # statistics go through double atomics (order-dependent in the last bits), hence a tolerance here
    assert rel_l2(out1.numpy(), out2.numpy()) < 1e-6


def test_conv2d_umma_epilogue_activation_and_batchnorm_input(dev, monkeypatch):
    from mdctgan_b200 import nn_ops as ops

    monkeypatch.setattr(ops, "CONV_ENGINE", "umma")
    g = torch.Generator().manual_seed(9)
    B, C, H, W, Co = 2, 64, 9, 17, 128
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Co, C, 4, 4, generator=g) * 0.05
    b = torch.randn(Co, generator=g)
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    xin = F.leaky_relu(x * scale[None, :, None, None] + shift[None, :, None, None], 0.2)
    ref = F.leaky_relu(F.conv2d(xin, w, b, stride=2, padding=2), 0.2)
    f = ops.Feat(_nhwc(x).to(dev), scale=scale.to(dev), shift=shift.to(dev), per_sample=False, act=ops.ACT_LEAKY)
    k = ops.pack_conv_weight(w.to(dev))
    y = ops.conv2d(f, k, b.to(dev), kh=4, kw=4, stride=2, pad=2, act=ops.ACT_LEAKY, w_umma=ops.pack_conv_weight_umma(k))
    assert rel_l2(_nchw(y.x.cpu()).numpy(), ref.numpy()) < 2e-5


def test_conv2d_umma_rejects_unsupported_shapes(dev):
    from mdctgan_b200 import nn_ops as ops

    assert not ops.umma_supported(2, 32)      # stem: Cin = 2
    assert not ops.umma_supported(32, 1)      # head: Cout = 1
    x = ops.Feat(torch.zeros(1, 4, 4, 6, device=dev))
    with pytest.raises(RuntimeError, match="conv2d_umma"):
        ops.conv2d(x, torch.zeros(6, 8, device=dev), None, kh=1, kw=1, w_umma=torch.zeros(64, device=dev))


def test_stats_arena_and_small_plane_fallback(dev, monkeypatch):
    """32-pixel planes (cfg3 global bottleneck): a 128-row tile spans 4 samples, so the deferred InstanceNorm is
    resolved by the finalize kernel; inside a stats_pass the statistics are slices of the zeroed arena."""
    from mdctgan_b200 import nn_ops as ops

    monkeypatch.setattr(ops, "CONV_ENGINE", "umma")
    g = torch.Generator().manual_seed(11)
    B, C, H, W = 8, 64, 2, 16
    x = torch.randn(B, C, H, W, generator=g)
    w1, w2 = torch.randn(C, C, 3, 3, generator=g) * 0.05, torch.randn(C, C, 3, 3, generator=g) * 0.05
    h = F.relu(F.instance_norm(F.conv2d(F.pad(x, (1,) * 4, mode="reflect"), w1)))
    ref = F.instance_norm(F.conv2d(F.pad(h, (1,) * 4, mode="reflect"), w2))
    k1, k2 = ops.pack_conv_weight(w1.to(dev)), ops.pack_conv_weight(w2.to(dev))
    u1, u2 = ops.pack_conv_weight_umma(k1), ops.pack_conv_weight_umma(k2)
    for _ in range(2):      # second pass re-zeroes the arena
        with ops.stats_pass(dev):
            f = ops.Feat(_nhwc(x).to(dev))
            a = ops.conv2d(f, k1, None, kh=3, kw=3, pad=1, pad_mode=ops.PAD_REFLECT, want_stats=True, w_umma=u1)
            assert a.stats.data_ptr() != 0 and a.stats._base is not None      # a view of the arena
            a = ops.with_act(ops.finalize_norm(a), ops.ACT_RELU)
            c = ops.conv2d(a, k2, None, kh=3, kw=3, pad=1, pad_mode=ops.PAD_REFLECT, want_stats=True, w_umma=u2)
            out = ops.materialize(ops.finalize_norm(c)).x.cpu()
        assert rel_l2(_nchw(out).numpy(), ref.numpy()) < 2e-5
