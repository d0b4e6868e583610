"""CPU: pin oracle/train_oracle.py (losses, both backward passes, both Adam steps) against tests/golden/train_golden.npz
= the reference's own create_model(opt)._forward / backward / optimizer steps on seeded weights and audio
(tests/golden/make_golden_nets.py train; reference: models/pix2pixHD_model.py:416-451, train.py:175-202)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import train_oracle as TO

sys.path.insert(0, GOLDEN)
from make_golden_nets import TRAIN_FLAGS, TRAIN_STEPS, state_checksum  # noqa: E402


def flags_to_cfg(flags):
    g = lambda k, d=None: flags[flags.index(k) + 1] if k in flags else d   # noqa: E731
    return dict(netG=g("--netG"), ngf=int(g("--ngf")), n_down=int(g("--n_downsample_global")), n_blocks_global=int(g("--n_blocks_global")),
                n_blocks_local=int(g("--n_blocks_local")), n_attn=int(g("--n_blocks_attn_g")), heads=int(g("--heads_g", 4)),
                dim_head=int(g("--dim_head_g", 128)), num_D=int(g("--num_D")), n_layers_D=int(g("--n_layers_D", 3)), ndf=int(g("--ndf", 64)),
                bins=int(g("--bins")), fit_residual="--fit_residual" in flags, down=g("--downsample_type", "conv"), up=g("--upsample_type", "transconv"))


def build_nets(cfg, seed, device="cpu"):
    """Our parameter-holding module trees, seeded like the reference's create_model (G first, then D)."""
    from mdctgan_b200.models import networks

    torch.manual_seed(seed)
    G = networks.define_G(2, 1, cfg["ngf"], cfg["netG"], cfg["n_down"], cfg["n_blocks_global"], 1, cfg["n_blocks_local"], "instance",
                          input_size=(cfg["bins"], 256), n_attn_g=cfg["n_attn"], heads_g=cfg["heads"], dim_head_g=cfg["dim_head"],
                          upsample_type=cfg["up"], downsample_type=cfg["down"])
    D = networks.define_D(3, cfg["ndf"], cfg["n_layers_D"], "instance", False, cfg["num_D"], True)
    return G, D


@pytest.fixture(scope="module")
def train_golden():
    return dict(np.load(os.path.join(GOLDEN, "train_golden.npz")))


@pytest.mark.parametrize("name", list(TRAIN_FLAGS))
def test_train_oracle_matches_reference(train_golden, name):
    flags, batch, T, seed = TRAIN_FLAGS[name]
    cfg = flags_to_cfg(flags)
    G, D = build_nets(cfg, seed)
    np.testing.assert_allclose(state_checksum(G.state_dict()), train_golden[f"{name}_G_cksum0"], rtol=1e-12)
    np.testing.assert_allclose(state_checksum(D.state_dict()), train_golden[f"{name}_D_cksum0"], rtol=1e-12)
    kw = {k: cfg[k] for k in ("netG", "n_down", "n_blocks_global", "n_blocks_local", "n_attn", "heads", "dim_head", "num_D", "n_layers_D",
                              "fit_residual", "down", "up")}
    out = TO.train_step(G.state_dict(), D.state_dict(), train_golden[f"{name}_lr_audio"], train_golden[f"{name}_hr_audio"], steps=TRAIN_STEPS,
                        **kw)
    np.testing.assert_allclose(np.array(out["losses"]), train_golden[f"{name}_losses"], rtol=2e-4)
    keysG = list(train_golden[f"{name}_gradG_keys"])
    ck = state_checksum({k: out["gradG"][k] for k in keysG})
    np.testing.assert_allclose(ck[:, 1], train_golden[f"{name}_gradG_cksum"][:, 1], rtol=2e-3)
    keysD = list(train_golden[f"{name}_gradD_keys"])
    ck = state_checksum({k: out["gradD"][k] for k in keysD})
    np.testing.assert_allclose(ck[:, 1], train_golden[f"{name}_gradD_cksum"][:, 1], rtol=2e-3)
    if name.startswith("tr_small"):
        for k in keysG:
            if f"{name}_gradG::{k}" in train_golden and float(np.abs(train_golden[f"{name}_gradG::{k}"]).max()) > 1e-4:
                assert rel_l2(out["gradG"][k].numpy(), train_golden[f"{name}_gradG::{k}"]) < 1e-3, k
    if name == "tr_small":
        for k in keysD:
            assert rel_l2(out["gradD"][k].numpy(), train_golden[f"{name}_gradD::{k}"]) < 1e-3, k
        for k, v in out["paramsG"].items():
            if v.dtype.is_floating_point and "running_" not in k and v.dim() >= 2:
                assert rel_l2(v.numpy(), train_golden[f"{name}_G_after::{k}"]) < 1e-3, k
    # post-step parameter checksums (sum of squares).  Adam turns a gradient into a +-lr step whatever its size, so entries whose true
    # gradient is zero (every bias in front of an InstanceNorm) move by rounding-noise signs: those tensors are held to the
    # Cauchy-Schwarz bound |d sum p^2| <= 2 |p| |dp|, |dp| <= 2 lr steps sqrt(n); weight tensors to 2e-4 relative.
    def check_after(params, ref, lr=2e-4):
        items = [(k, v) for k, v in params.items()]
        ck = state_checksum(dict(items))
        for i, (k, v) in enumerate(items):
            if not v.dtype.is_floating_point or "running_" in k or "num_batches" in k:
                continue
            if v.dim() >= 2:
                np.testing.assert_allclose(ck[i, 1], ref[i, 1], rtol=2e-4, err_msg=k)
            else:
                bound = 2.0 * np.sqrt(ref[i, 1]) * (2 * lr * TRAIN_STEPS) * np.sqrt(v.numel()) + 1e-12
                assert abs(ck[i, 1] - ref[i, 1]) <= bound, (k, ck[i, 1], ref[i, 1], bound)

    check_after(out["paramsG"], train_golden[f"{name}_G_cksum_after"])
    check_after(out["paramsD"], train_golden[f"{name}_D_cksum_after"])


@pytest.mark.parametrize("name", ["tr_small_bce", "tr_small_attnl"])
def test_train_oracle_option_branches_match_reference(name):
    """The oracle's --no_lsgan (sigmoid head + BCE, no intermediate features) and --n_blocks_attn_l (local attention sandwich) branches
    against the reference's own train iteration (tests/golden/train_{bce,attnl}_golden.npz): losses and every gradient tensor of the first
    iteration.  The attention-sandwich configuration is badly conditioned (see tests/test_train_gpu.py): on the same CPU, in fp32, this
    functional restatement and the reference's module graph already differ by percent-level amounts on the far end of the generator."""
    from make_golden_nets import TRAIN_ATTNL_FLAGS, TRAIN_BCE_FLAGS
    from mdctgan_b200.models import networks

    bce = name == "tr_small_bce"
    gold = dict(np.load(os.path.join(GOLDEN, "train_bce_golden.npz" if bce else "train_attnl_golden.npz")))
    flags, batch, T, seed = (TRAIN_BCE_FLAGS if bce else TRAIN_ATTNL_FLAGS)[name]
    cfg = flags_to_cfg(flags)
    g = lambda k, d: int(flags[flags.index(k) + 1]) if k in flags else d   # noqa: E731
    n_attn_l, heads_l, dim_head_l = g("--n_blocks_attn_l", 0), g("--heads_l", 4), g("--dim_head_l", 128)
    torch.manual_seed(seed)
    G = networks.define_G(2, 1, cfg["ngf"], cfg["netG"], cfg["n_down"], cfg["n_blocks_global"], 1, cfg["n_blocks_local"], "instance",
                          input_size=(cfg["bins"], 256), n_attn_g=cfg["n_attn"], heads_g=cfg["heads"], dim_head_g=cfg["dim_head"],
                          n_attn_l=n_attn_l, heads_l=heads_l, dim_head_l=dim_head_l)
    D = networks.define_D(3, cfg["ndf"], cfg["n_layers_D"], "instance", bce, cfg["num_D"], not bce)
    np.testing.assert_allclose(state_checksum(G.state_dict()), gold[f"{name}_G_cksum0"], rtol=1e-12)
    np.testing.assert_allclose(state_checksum(D.state_dict()), gold[f"{name}_D_cksum0"], rtol=1e-12)
    kw = {k: cfg[k] for k in ("netG", "n_down", "n_blocks_global", "n_blocks_local", "n_attn", "heads", "dim_head", "num_D", "n_layers_D",
                              "fit_residual", "down", "up")}
    kw.update(lsgan=not bce, use_feat=not bce, n_attn_l=n_attn_l, heads_l=heads_l, dim_head_l=dim_head_l)
    out = TO.train_step(G.state_dict(), D.state_dict(), gold[f"{name}_lr_audio"], gold[f"{name}_hr_audio"], steps=1, **kw)
    names = list(gold[f"{name}_loss_names"])
    got = dict(zip(["G_GAN", "G_GAN_Feat", "D_real", "D_fake"], out["losses"][0]))
    np.testing.assert_allclose([got[k] for k in names], gold[f"{name}_losses"][0], rtol=2e-4)
    for tag in ("gradG", "gradD"):
        ref = {k[len(name) + len(tag) + 3:]: v for k, v in gold.items() if k.startswith(f"{name}_{tag}::")}
        gmax = max(float(np.abs(v).max()) for v in ref.values())
        worst = 0.0
        for k, v in ref.items():
            if float(np.abs(v).max()) < 1e-4 * gmax:
                continue
            assert k in out[tag], (tag, k)
            worst = max(worst, rel_l2(out[tag][k].numpy(), v))
        assert worst < (1e-3 if bce else 0.25), (tag, worst)
