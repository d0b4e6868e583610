"""GPU parity of the network kernels (convolutions with fused pad / norm / activation, attention, pooling)
through the C ABI, against torch-CPU fp32 references of the same ops and against the golden outputs of the
reference's own define_G / define_D (tests/golden/nets_golden.npz).

Tolerance: fp32 kernels with a different summation order than the CPU reference -> 2e-5 rel-L2 per layer,
1e-4 rel-L2 on whole-network outputs (the north_star waveform bar is 1e-3)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_l2

sys.path.insert(0, GOLDEN)
from make_golden_nets import NET_CASES, make_input  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def nets_golden():
    return dict(np.load(os.path.join(GOLDEN, "nets_golden.npz")))


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


CONV_CASES = [
    # B, Cin, H, W, Cout, k, stride, pad, reflect, transposed
    (2, 2, 16, 64, 32, 7, 1, 3, True, False),      # stem: ReflPad3 + 7x7, Cin = 2
    (2, 32, 16, 64, 64, 3, 2, 1, False, False),    # downsample
    (3, 64, 4, 32, 64, 3, 1, 1, True, False),      # residual-block conv
    (1, 256, 4, 32, 256, 3, 1, 1, True, False),
    (2, 64, 8, 16, 32, 3, 2, 1, False, True),      # ConvTranspose2d k3 s2 p1 op1
    (2, 3, 17, 33, 64, 4, 2, 2, False, False),     # discriminator first layer, odd sizes, Cin = 3
    (2, 64, 9, 17, 128, 4, 1, 2, False, False),
    (2, 128, 5, 9, 96, 1, 1, 0, False, False),     # 1x1 (BottleStack projections), Cout not a multiple of 64
    (2, 32, 16, 64, 1, 7, 1, 3, True, False),      # generator head, Cout = 1
    (2, 128, 6, 10, 1, 4, 1, 2, False, False),     # discriminator head, Cout = 1
    (1, 40, 7, 5, 24, 5, 1, 2, False, False),      # 5x5 (ConvResBlock / InterpolateUpsample shapes)
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_matches_torch(dev, case):
    from mdctgan_b200 import nn_ops as ops

    B, Cin, H, W, Cout, k, stride, pad, reflect, transposed = case
    g = torch.Generator().manual_seed(hash(case) % 2**31)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn((Cin, Cout, k, k) if transposed else (Cout, Cin, k, k), generator=g) * 0.1
    bias = torch.randn(Cout, generator=g)
    if transposed:
        ref = F.conv_transpose2d(x, w, bias, stride=stride, padding=pad, output_padding=1)
    elif reflect:
        ref = F.conv2d(F.pad(x, (pad,) * 4, mode="reflect"), w, bias, stride=stride)
    else:
        ref = F.conv2d(x, w, bias, stride=stride, padding=pad)
    f = ops.Feat(_nhwc(x).to(dev))
    y = ops.conv2d(f, ops.pack_conv_weight(w.to(dev), transposed), bias.to(dev), kh=k, kw=k, stride=stride, pad=pad,
                   pad_mode=ops.PAD_REFLECT if reflect else ops.PAD_ZERO, transposed=transposed, output_padding=1 if transposed else 0,
                   want_stats=Cout > 1)
    got = _nchw(y.x.cpu())
    assert got.shape == ref.shape
    assert rel_l2(got.numpy(), ref.numpy()) < 2e-5
    if Cout > 1:   # epilogue statistics == sums of the output
        st = y.stats.cpu()
        np.testing.assert_allclose(st[..., 0].numpy(), ref.double().sum(dim=(2, 3)).numpy(), rtol=1e-4, atol=1e-3)
        np.testing.assert_allclose(st[..., 1].numpy(), (ref.double() ** 2).sum(dim=(2, 3)).numpy(), rtol=1e-4)


def test_conv_norm_act_conv_chain(dev):
    """conv -> InstanceNorm -> ReLU -> conv, the norm + activation applied in the consumer's gather; and the
    residual combine."""
    from mdctgan_b200 import nn_ops as ops

    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 16, 8, 32, generator=g)
    w1, b1 = torch.randn(32, 16, 3, 3, generator=g) * 0.1, torch.randn(32, generator=g)
    w2, b2 = torch.randn(16, 32, 3, 3, generator=g) * 0.1, torch.randn(16, generator=g)
    h = F.relu(F.instance_norm(F.conv2d(F.pad(x, (1,) * 4, mode="reflect"), w1, b1)))
    ref = x + F.instance_norm(F.conv2d(F.pad(h, (1,) * 4, mode="reflect"), w2, b2))
    f = ops.Feat(_nhwc(x).to(dev))
    a = ops.conv2d(f, ops.pack_conv_weight(w1.to(dev)), b1.to(dev), kh=3, kw=3, pad=1, pad_mode=ops.PAD_REFLECT, want_stats=True)
    a = ops.with_act(ops.finalize_norm(a), ops.ACT_RELU)
    c = ops.conv2d(a, ops.pack_conv_weight(w2.to(dev)), b2.to(dev), kh=3, kw=3, pad=1, pad_mode=ops.PAD_REFLECT, want_stats=True)
    out = ops.combine(f, ops.finalize_norm(c))
    assert rel_l2(_nchw(out.x.cpu()).numpy(), ref.numpy()) < 2e-5
    # LeakyReLU epilogue + BatchNorm (train / eval) on the consumer side
    gamma, beta = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g)
    rm, rv = torch.zeros(32), torch.ones(32)
    conv = F.conv2d(x, w1, b1, padding=1)
    ref_bn = F.leaky_relu(F.batch_norm(conv, rm.clone(), rv.clone(), gamma, beta, True, 0.1, 1e-5), 0.2)
    a = ops.conv2d(f, ops.pack_conv_weight(w1.to(dev)), b1.to(dev), kh=3, kw=3, pad=1, want_stats=True)
    rmd, rvd = rm.to(dev), rv.to(dev)
    a = ops.with_act(ops.finalize_norm(a, mode=1, gamma=gamma.to(dev), beta=beta.to(dev), running_mean=rmd, running_var=rvd), ops.ACT_LEAKY)
    assert rel_l2(_nchw(ops.materialize(a).x.cpu()).numpy(), ref_bn.numpy()) < 2e-5
    rm2, rv2 = rm.clone(), rv.clone()
    F.batch_norm(conv, rm2, rv2, gamma, beta, True, 0.1, 1e-5)
    np.testing.assert_allclose(rmd.cpu().numpy(), rm2.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rvd.cpu().numpy(), rv2.numpy(), rtol=1e-5, atol=1e-6)


def test_avgpool_and_layout(dev):
    from mdctgan_b200 import nn_ops as ops

    g = torch.Generator().manual_seed(6)
    for shape in ((2, 3, 32, 256), (1, 2, 17, 33), (2, 5, 4, 2)):
        x = torch.randn(shape, generator=g)
        f = ops.to_nhwc(x.to(dev))
        assert torch.equal(f.x.cpu(), _nhwc(x))
        assert torch.equal(ops.to_nchw(f).cpu(), x)
        ref = F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False)
        got = ops.to_nchw(ops.avgpool3s2(f)).cpu()
        assert got.shape == ref.shape and rel_l2(got.numpy(), ref.numpy()) < 1e-6


@pytest.mark.parametrize("L_hw,heads,d", [((2, 16), 4, 64), ((8, 16), 6, 128), ((3, 5), 2, 32)])
def test_attention_matches_oracle(dev, L_hw, heads, d):
    from mdctgan_b200 import nn_ops as ops
    from oracle import networks_oracle as NO

    g = torch.Generator().manual_seed(7)
    hh, ww = L_hw
    cin = 48
    x = torch.randn(2, cin, hh, ww, generator=g)
    sd = {"a.to_qkv.weight": torch.randn(3 * heads * d, cin, 1, 1, generator=g) * 0.2,
          "a.pos_emb.height": torch.randn(hh, d, generator=g) * d ** -0.5, "a.pos_emb.width": torch.randn(ww, d, generator=g) * d ** -0.5}
    ref = NO.attention(sd, "a", x, heads, d)
    qkv = ops.conv2d(ops.Feat(_nhwc(x).to(dev)), ops.pack_conv_weight(sd["a.to_qkv.weight"].to(dev)), None, kh=1, kw=1)
    out = ops.attention(qkv, sd["a.pos_emb.height"].to(dev), sd["a.pos_emb.width"].to(dev), heads, d, d ** -0.5)
    assert rel_l2(_nchw(out.x.cpu()).numpy(), ref.numpy()) < 2e-5
    np.testing.assert_allclose(out.stats.cpu()[..., 0].numpy(), ref.double().sum(dim=(2, 3)).numpy(), rtol=1e-4, atol=1e-3)


def _build(kind, kw, seed, dev):
    from mdctgan_b200.models import networks

    torch.manual_seed(seed)
    net = networks.define_G(**kw) if kind == "G" else networks.define_D(**kw)
    return net.to(dev)


@pytest.mark.parametrize("name", ["g_small", "l_small", "cfg2", "local_noattn", "cfg3", "g_small_rc", "trainsh"])
def test_generator_matches_reference_output(dev, nets_golden, name):
    import mdctgan_b200

    kind, kw, shape, seed = NET_CASES[name]
    net = _build(kind, kw, seed, dev).eval()
    x = make_input(shape, seed).to(dev)
    n0 = mdctgan_b200.launch_count()
    y = net(x)
    assert mdctgan_b200.launch_count() > n0
    assert y.shape == nets_golden[f"{name}_y"].shape
    err = rel_l2(y.cpu().numpy(), nets_golden[f"{name}_y"])
    assert err < 1e-4, err
    if name == "cfg3":      # BatchNorm with batch statistics + running-buffer update
        net.train()
        yt = net(x)
        # BatchNorm over 2 x 32 tokens: batch statistics of 64 values amplify fp32 summation-order differences
        assert rel_l2(yt.cpu().numpy(), nets_golden["cfg3_y_train"]) < 5e-4
        bn = net.model[17].net[0].net[1]
        assert int(bn.num_batches_tracked) == 1 and not torch.equal(bn.running_mean, torch.zeros_like(bn.running_mean))


@pytest.mark.parametrize("name", ["d_small", "d3"])
def test_discriminator_matches_reference_output(dev, nets_golden, name):
    kind, kw, shape, seed = NET_CASES[name]
    net = _build(kind, kw, seed, dev).eval()
    res = net(make_input(shape, seed).to(dev))
    assert len(res) == kw["num_D"] and all(len(r) == kw["n_layers_D"] + 2 for r in res)
    for i, feats in enumerate(res):
        assert rel_l2(feats[-1].cpu().numpy(), nets_golden[f"{name}_pred{i}"]) < 1e-4
        for j, f in enumerate(feats):
            st = nets_golden[f"{name}_f{i}{j}_stats"]
            assert tuple(f.shape) == tuple(int(v) for v in st[2:])
            assert abs(float((f.double() ** 2).sum()) - st[1]) <= 2e-4 * st[1]
        if name == "d_small":
            assert rel_l2(feats[1].cpu().numpy(), nets_golden[f"{name}_feat{i}1"]) < 1e-4


def test_generator_against_oracle_other_batch_and_frames(dev):
    """Seeded oracle comparison at a shape the goldens do not cover (B = 5, 64 frames)."""
    from oracle import networks_oracle as NO

    kw = dict(input_nc=2, output_nc=1, ngf=16, netG="global", n_downsample_global=3, n_blocks_global=3, norm="instance", input_size=(64, 256))
    net = _build("G", kw, 21, dev).eval()
    x = make_input((5, 2, 64, 256), 21)
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    with torch.no_grad():
        ref = NO.global_generator(sd, x, 3, 3)
    assert rel_l2(net(x.to(dev)).cpu().numpy(), ref.numpy()) < 1e-4


def test_leaf_layers_do_not_compute_and_cpu_input_raises(dev):
    from mdctgan_b200.models import networks

    conv = networks.Conv2d(4, 4, 3).to(dev)
    with pytest.raises(RuntimeError, match="parameters only"):
        conv(torch.zeros(1, 4, 8, 8, device=dev))
    net = networks.define_G(2, 1, 8, "global", 2, 1, input_size=(16, 64)).to(dev)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        net(torch.zeros(1, 2, 16, 64))
