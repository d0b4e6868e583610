"""CPU: pin oracle/networks_oracle.py and the drop-in module tree against tests/golden/nets_golden.npz
(outputs / state_dict checksums of the reference's own define_G / define_D on seeded weights)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import networks_oracle as NO

sys.path.insert(0, GOLDEN)
from make_golden_nets import NET_CASES, make_input, state_checksum  # noqa: E402


@pytest.fixture(scope="module")
def nets_golden():
    return dict(np.load(os.path.join(GOLDEN, "nets_golden.npz")))


def build_ours(kind, kw, seed):
    from mdctgan_b200.models import networks

    torch.manual_seed(seed)
    return networks.define_G(**kw) if kind == "G" else networks.define_D(**kw)


@pytest.mark.parametrize("name", list(NET_CASES))
def test_state_dict_layout_and_seeded_init_match_reference(nets_golden, name):
    """Same keys, shapes AND the same seeded initial values as the reference (construction order, default
    init of ConvTranspose2d / biases, weights_init rule, BottleStack override: SURVEY 3d)."""
    kind, kw, shape, seed = NET_CASES[name]
    sd = build_ours(kind, kw, seed).state_dict()
    assert list(sd.keys()) == list(nets_golden[f"{name}_keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(nets_golden[f"{name}_shapes"])
    np.testing.assert_allclose(state_checksum(sd), nets_golden[f"{name}_cksum"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["g_small", "l_small", "cfg2", "local_noattn", "cfg3", "g_small_rc", "trainsh"])
def test_generator_oracle_matches_reference(nets_golden, name):
    kind, kw, shape, seed = NET_CASES[name]
    sd = build_ours(kind, kw, seed).state_dict()
    x = make_input(shape, seed)
    common = dict(n_attn=kw.get("n_attn_g", 0), heads=kw.get("heads_g", 4), dim_head=kw.get("dim_head_g", 128),
                  down=kw.get("downsample_type", "conv"), up=kw.get("upsample_type", "transconv"))
    with torch.no_grad():
        if kw["netG"] == "global":
            y = NO.global_generator(sd, x, kw["n_downsample_global"], kw["n_blocks_global"], **common)
        else:
            y = NO.local_enhancer(sd, x, kw["n_downsample_global"], kw["n_blocks_global"], kw["n_blocks_local"], **common)
    assert rel_l2(y.numpy(), nets_golden[f"{name}_y"]) < 2e-6
    if name == "cfg3":
        with torch.no_grad():
            yt = NO.local_enhancer(sd, x, 3, 9, 3, training=True, **common)
        assert rel_l2(yt.numpy(), nets_golden["cfg3_y_train"]) < 2e-6


@pytest.mark.parametrize("name", ["d_small", "d3"])
def test_discriminator_oracle_matches_reference(nets_golden, name):
    kind, kw, shape, seed = NET_CASES[name]
    sd = build_ours(kind, kw, seed).state_dict()
    x = make_input(shape, seed)
    with torch.no_grad():
        res = NO.multiscale_d(sd, x, kw["num_D"], kw["n_layers_D"])
    for i, feats in enumerate(res):
        assert rel_l2(feats[-1].numpy(), nets_golden[f"{name}_pred{i}"]) < 2e-6
        for j, f in enumerate(feats):
            st = nets_golden[f"{name}_f{i}{j}_stats"]
            assert tuple(f.shape) == tuple(int(v) for v in st[2:])
            assert abs(float((f.double() ** 2).sum()) - st[1]) <= 1e-5 * st[1]


INFER_CFG = {
    "inf_cfg3": dict(n_down=3, n_blocks_global=9, n_blocks_local=3, n_attn=2, heads=4, dim_head=64, fit_residual=True),
    "inf_small": dict(n_down=2, n_blocks_global=2, n_blocks_local=1, n_attn=0, fit_residual=False),
}


def our_opt(name, gpu="-1"):
    from make_golden_nets import INFER_FLAGS
    from mdctgan_b200.options.train_options import TrainOptions

    flags = INFER_FLAGS[name][0]
    base = ["--name", "g", "--gpu_ids", gpu, "--lr_sampling_rate", "12000", "--sr_sampling_rate", "48000", "--arcsinh_transform",
            "--abs_spectro", "--arcsinh_gain", "1000", "--center", "--norm_range", "-1", "1", "--abs_norm", "--src_range", "-5", "5"]
    return TrainOptions().parse(save=False, args=base + flags)


@pytest.mark.parametrize("name", ["inf_small", "inf_cfg3"])
def test_inference_oracle_matches_reference(nets_golden, name):
    from make_golden_nets import INFER_FLAGS
    from mdctgan_b200.models import networks
    from oracle import model_oracle as MOD

    g = nets_golden
    opt = our_opt(name)
    torch.manual_seed(INFER_FLAGS[name][3])
    netG = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.netG, opt.n_downsample_global, opt.n_blocks_global,
                             opt.n_local_enhancers, opt.n_blocks_local, opt.norm, input_size=(opt.bins, opt.n_fft // 2),
                             n_attn_g=opt.n_blocks_attn_g, heads_g=opt.heads_g, dim_head_g=opt.dim_head_g)
    np.testing.assert_allclose(state_checksum(netG.state_dict()), g[f"{name}_G_cksum"], rtol=1e-12, atol=1e-12)
    sr, audio, lr = MOD.inference(netG.state_dict(), g[f"{name}_lr_audio"], **INFER_CFG[name])
    assert np.abs(lr - g[f"{name}_lr_spectro"]).max() <= 6e-8     # 1 ulp flips where the fp64 FFT orders differ
    assert rel_l2(sr, g[f"{name}_sr_spectro"]) < 2e-6
    assert rel_l2(audio, g[f"{name}_sr_audio"]) < 2e-5    # sinh(ln10*5*s) amplifies the 1e-6 spectrogram difference ~10x
