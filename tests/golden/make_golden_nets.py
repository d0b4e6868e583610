#!/usr/bin/env python
"""Generate tests/golden/nets_golden.npz by running the REFERENCE's define_G / define_D (read-only
/root/reference) on CPU with seeded weights and inputs.  Only numeric outputs and per-tensor checksums of
the seeded state_dict are stored (the weights are re-created from the seed by whoever checks).

    python tests/golden/make_golden.py nets
"""
import contextlib
import io
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (kind, kwargs, input shape, seed)
NET_CASES = {
    # BASELINE configs[1]: GlobalGenerator only, ngf 32, 32 frames
    "cfg2": ("G", dict(input_nc=2, output_nc=1, ngf=32, netG="global", n_downsample_global=3, n_blocks_global=9, norm="instance",
                       input_size=(32, 256)), (2, 2, 32, 256), 1234),
    "g_small": ("G", dict(input_nc=2, output_nc=1, ngf=8, netG="global", n_downsample_global=2, n_blocks_global=2, norm="instance",
                          input_size=(16, 64)), (3, 2, 16, 64), 11),
    # BASELINE configs[2] without / with the two BottleStack attention layers
    "local_noattn": ("G", dict(input_nc=2, output_nc=1, ngf=32, netG="local", n_downsample_global=3, n_blocks_global=9,
                               n_local_enhancers=1, n_blocks_local=3, norm="instance", input_size=(32, 256)), (2, 2, 32, 256), 77),
    "cfg3": ("G", dict(input_nc=2, output_nc=1, ngf=32, netG="local", n_downsample_global=3, n_blocks_global=9, n_local_enhancers=1,
                       n_blocks_local=3, norm="instance", input_size=(32, 256), n_attn_g=2, heads_g=4, dim_head_g=64, proj_factor_g=4),
             (2, 2, 32, 256), 78),
    "l_small": ("G", dict(input_nc=2, output_nc=1, ngf=8, netG="local", n_downsample_global=2, n_blocks_global=2, n_local_enhancers=1,
                          n_blocks_local=1, norm="instance", input_size=(16, 64)), (2, 2, 16, 64), 12),
    # the reference's shipped training recipe (train.sh): resconv down / interpolate up, ngf 56, 3 attention layers (6 heads x 128)
    "trainsh": ("G", dict(input_nc=2, output_nc=1, ngf=56, netG="local", n_downsample_global=3, n_blocks_global=4, n_local_enhancers=1,
                          n_blocks_local=3, norm="instance", input_size=(128, 256), n_attn_g=3, heads_g=6, dim_head_g=128, proj_factor_g=4,
                          upsample_type="interpolate", downsample_type="resconv"), (1, 2, 128, 256), 79),
    "g_small_rc": ("G", dict(input_nc=2, output_nc=1, ngf=8, netG="global", n_downsample_global=2, n_blocks_global=2, norm="instance",
                             input_size=(16, 64), upsample_type="interpolate", downsample_type="resconv"), (3, 2, 16, 64), 14),
    "d3": ("D", dict(input_nc=3, ndf=64, n_layers_D=3, norm="instance", use_sigmoid=False, num_D=3, getIntermFeat=True),
           (2, 3, 32, 256), 99),
    "d_small": ("D", dict(input_nc=3, ndf=8, n_layers_D=2, norm="instance", use_sigmoid=False, num_D=2, getIntermFeat=True),
                (2, 3, 16, 64), 13),
}


def make_input(shape, seed):
    g = torch.Generator().manual_seed(seed + 1000)
    return (0.5 * torch.randn(shape, generator=g)).clamp(-1, 1)


def state_checksum(sd):
    """[n_tensors, 2]: (sum, sum of squares) in float64, in state_dict order."""
    return np.array([[float(v.double().sum()), float((v.double() ** 2).sum())] for v in sd.values()])


def build_reference(kind, kw, seed):
    from models import networks

    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = networks.define_G(**kw) if kind == "G" else networks.define_D(**kw)
    return net


def gen_nets():
    out = {}
    for name, (kind, kw, shape, seed) in NET_CASES.items():
        net = build_reference(kind, kw, seed)
        net.eval()
        x = make_input(shape, seed)
        with torch.no_grad():
            y = net(x)
        sd = net.state_dict()
        out[f"{name}_keys"] = np.array(list(sd.keys()))
        out[f"{name}_shapes"] = np.array([str(tuple(v.shape)) for v in sd.values()])
        out[f"{name}_cksum"] = state_checksum(sd)
        if kind == "G":
            out[f"{name}_y"] = y.numpy()
        else:
            for i, feats in enumerate(y):
                for j, f in enumerate(feats):
                    out[f"{name}_f{i}{j}_stats"] = np.array([float(f.double().sum()), float((f.double() ** 2).sum())] + list(f.shape))
                out[f"{name}_pred{i}"] = feats[-1].numpy()
                out[f"{name}_feat{i}1"] = feats[1].numpy() if name == "d_small" else np.zeros(0, np.float32)
        if name == "cfg3":   # BatchNorm in training mode as well (batch statistics)
            net.train()
            with torch.no_grad():
                out["cfg3_y_train"] = net(x).numpy()
        print(name, "done")
    np.savez_compressed(os.path.join(HERE, "nets_golden.npz"), **out)
    print("wrote nets_golden.npz")


INFER_FLAGS = {
    # BASELINE configs[2]: LocalEnhancer + 2 BottleStack attention layers, 12 -> 48 kHz, 32-frame segments, --fit_residual
    "inf_cfg3": (["--netG", "local", "--ngf", "32", "--n_downsample_global", "3", "--n_blocks_global", "9", "--n_blocks_attn_g", "2",
                  "--heads_g", "4", "--dim_head_g", "64", "--n_blocks_local", "3", "--num_D", "3", "--segment_length", "7936",
                  "--bins", "32", "--fit_residual"], 2, 7936, 4242),
    "inf_small": (["--netG", "local", "--ngf", "8", "--n_downsample_global", "2", "--n_blocks_global", "2", "--n_blocks_attn_g", "0",
                   "--n_blocks_local", "1", "--num_D", "2", "--segment_length", "3840", "--bins", "16"], 3, 3840, 4243),
}


def make_lr_audio(batch, T, seed, cutoff_hz=6000.0, sr=48000.0):
    """Seeded band-limited 'LR' audio at the HR rate: 0.1*randn low-passed by an FFT brick wall at 6 kHz."""
    rng = np.random.default_rng(seed)
    x = 0.1 * rng.standard_normal((batch, T))
    X = np.fft.rfft(x, axis=-1)
    X[:, np.fft.rfftfreq(T, 1.0 / sr) > cutoff_hz] = 0
    return torch.from_numpy(np.fft.irfft(X, n=T, axis=-1).astype(np.float32))


def gen_infer():
    """Pix2PixHDModel.inference of the reference (create_model(opt)) on seeded weights / audio."""
    from make_golden import ref_opt
    from models.models import create_model

    path = os.path.join(HERE, "nets_golden.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    for name, (flags, batch, T, seed) in INFER_FLAGS.items():
        opt = ref_opt(flags)
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            model = create_model(opt)
        model.eval()
        lr = make_lr_audio(batch, T, seed)
        sr_spectro, sr_audio, lr_pha, prm, lr_spectro = model.inference(lr)
        out[f"{name}_lr_audio"] = lr.numpy()
        out[f"{name}_sr_spectro"] = sr_spectro.numpy()
        out[f"{name}_sr_audio"] = sr_audio.numpy()
        out[f"{name}_lr_spectro"] = lr_spectro.numpy()
        out[f"{name}_G_cksum"] = state_checksum(model.netG.state_dict())
        print(name, "done", sr_audio.shape, sr_audio.dtype)
    np.savez_compressed(path, **out)


TRAIN_FLAGS = {
    # BASELINE configs[3] per-GPU slice: LocalEnhancer + 2 attention layers, num_D 3, feature matching, --fit_residual (batch 2 here)
    "tr_cfg4": (["--netG", "local", "--ngf", "32", "--n_downsample_global", "3", "--n_blocks_global", "9", "--n_blocks_attn_g", "2",
                 "--heads_g", "4", "--dim_head_g", "64", "--n_blocks_local", "3", "--num_D", "3", "--segment_length", "7936",
                 "--bins", "32", "--fit_residual"], 2, 7936, 5151),
    "tr_small": (["--netG", "local", "--ngf", "8", "--n_downsample_global", "2", "--n_blocks_global", "2", "--n_blocks_attn_g", "1",
                  "--heads_g", "2", "--dim_head_g", "32", "--n_blocks_local", "1", "--num_D", "2", "--n_layers_D", "2", "--ndf", "8",
                  "--segment_length", "3840", "--bins", "16", "--fit_residual"], 3, 3840, 5152),
}
TRAIN_FLAGS["tr_small_rc"] = (TRAIN_FLAGS["tr_small"][0] + ["--upsample_type", "interpolate", "--downsample_type", "resconv"], 2, 3840, 5153)
TRAIN_STEPS = 2


def make_hr_audio(batch, T, seed):
    rng = np.random.default_rng(seed + 17)
    return torch.from_numpy((0.1 * rng.standard_normal((batch, T))).astype(np.float32))


def gen_train():
    """create_model(opt) of the reference in training mode: _forward, loss_G / loss_D backward, both Adam steps
    (train.py:160-202), TRAIN_STEPS iterations on one seeded batch."""
    from make_golden import ref_opt
    from models.models import create_model

    out = {}
    for name, (flags, batch, T, seed) in TRAIN_FLAGS.items():
        opt = ref_opt(flags)
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            model = create_model(opt)
        model.train()
        lr, hr = make_lr_audio(batch, T, seed), make_hr_audio(batch, T, seed)
        out[f"{name}_lr_audio"], out[f"{name}_hr_audio"] = lr.numpy(), hr.numpy()
        out[f"{name}_G_cksum0"] = state_checksum(model.netG.state_dict())
        out[f"{name}_D_cksum0"] = state_checksum(model.netD.state_dict())
        losses_all = []
        for it in range(TRAIN_STEPS):
            losses, _ = model._forward(lr, hr)
            d = dict(zip(model.loss_names, losses))
            losses_all.append([float(d[k]) for k in ("G_GAN", "G_GAN_Feat", "D_real", "D_fake")])
            loss_D = (d["D_fake"] + d["D_real"]) * 0.5
            loss_G = d["G_GAN"] + d["G_GAN_Feat"]
            model.optimizer_G.zero_grad()
            loss_G.backward()
            if it == 0:
                gG = {k: p.grad for k, p in model.netG.named_parameters()}
                out[f"{name}_gradG_keys"] = np.array(list(gG.keys()))
                out[f"{name}_gradG_cksum"] = state_checksum(gG)
                if name.startswith("tr_small"):
                    for k, v in gG.items():
                        if name == "tr_small" or v.numel() <= 2500:
                            out[f"{name}_gradG::{k}"] = v.numpy().copy()
            model.optimizer_G.step()
            model.optimizer_D.zero_grad()
            loss_D.backward()
            if it == 0:
                gD = {k: p.grad for k, p in model.netD.named_parameters()}
                out[f"{name}_gradD_keys"] = np.array(list(gD.keys()))
                out[f"{name}_gradD_cksum"] = state_checksum(gD)
                if name == "tr_small":
                    for k, v in gD.items():
                        out[f"{name}_gradD::{k}"] = v.numpy().copy()
            model.optimizer_D.step()
        out[f"{name}_losses"] = np.array(losses_all)
        out[f"{name}_G_cksum_after"] = state_checksum(model.netG.state_dict())
        out[f"{name}_D_cksum_after"] = state_checksum(model.netD.state_dict())
        if name == "tr_small":
            for k, v in model.netG.state_dict().items():
                out[f"{name}_G_after::{k}"] = v.numpy().copy()
        print(name, "done", losses_all)
    np.savez_compressed(os.path.join(HERE, "train_golden.npz"), **out)
    print("wrote train_golden.npz")


# Option branches outside the shipped recipes, each pinned by the reference's own train iteration (train.py:160-202):
#   tr_small_bce    --no_lsgan (nn.BCELoss on a sigmoid PatchGAN head, networks.py:105-108,671-672).  The reference's discriminator applies
#                   the sigmoid only without intermediate features (networks.py:686 walks models 0 .. n_layers + 1), so the flag works
#                   with --no_ganFeat_loss only.
#   tr_small_attnl  --n_blocks_attn_l 1 (attention sandwich of the local branch, networks.py:218-237: weight-shared down / up layers,
#                   BottleStack with a projection shortcut).  bins 32: the 8x-down map must be at least 1 x 1 after // 16.
TRAIN_BCE_FLAGS = {"tr_small_bce": (TRAIN_FLAGS["tr_small"][0] + ["--no_lsgan", "--no_ganFeat_loss"], 3, 3840, 5154)}
_attnl = [a for a in TRAIN_FLAGS["tr_small"][0]]
_attnl[_attnl.index("--bins") + 1] = "32"
_attnl[_attnl.index("--segment_length") + 1] = "7936"
_attnl[_attnl.index("--n_blocks_local") + 1] = "3"
TRAIN_ATTNL_FLAGS = {"tr_small_attnl": (_attnl + ["--n_blocks_attn_l", "1", "--heads_l", "2", "--dim_head_l", "32"], 2, 7936, 5155)}
TRAIN_EXTRA = {"train_bce_golden.npz": TRAIN_BCE_FLAGS, "train_attnl_golden.npz": TRAIN_ATTNL_FLAGS}


def gen_train_extra(which=None):
    """The reference's own train iteration (train.py:160-202) under the option branches above: losses of TRAIN_STEPS iterations, every
    gradient tensor of the first one, parameter checksums after the last -> tests/golden/train_{bce,attnl}_golden.npz."""
    from make_golden import ref_opt
    from models.models import create_model

    for fname, table in TRAIN_EXTRA.items():
        if which is not None and fname != which:
            continue
        out = {}
        for name, (flags, batch, T, seed) in table.items():
            opt = ref_opt(flags)
            torch.manual_seed(seed)
            with contextlib.redirect_stdout(io.StringIO()):
                model = create_model(opt)
            model.train()
            lr, hr = make_lr_audio(batch, T, seed), make_hr_audio(batch, T, seed)
            out[f"{name}_lr_audio"], out[f"{name}_hr_audio"] = lr.numpy(), hr.numpy()
            out[f"{name}_G_cksum0"] = state_checksum(model.netG.state_dict())
            out[f"{name}_D_cksum0"] = state_checksum(model.netD.state_dict())
            out[f"{name}_G_keys"] = np.array(list(model.netG.state_dict().keys()))
            out[f"{name}_D_keys"] = np.array(list(model.netD.state_dict().keys()))
            out[f"{name}_loss_names"] = np.array(list(model.loss_names))
            losses_all = []
            for it in range(TRAIN_STEPS):
                losses, _ = model._forward(lr, hr)
                d = dict(zip(model.loss_names, losses))
                losses_all.append([float(d[k].detach()) for k in model.loss_names])
                loss_D = (d["D_fake"] + d["D_real"]) * 0.5
                loss_G = d["G_GAN"] + d.get("G_GAN_Feat", 0)
                model.optimizer_G.zero_grad()
                loss_G.backward()
                if it == 0:
                    for k, p in model.netG.named_parameters():
                        out[f"{name}_gradG::{k}"] = p.grad.numpy().copy()
                model.optimizer_G.step()
                model.optimizer_D.zero_grad()
                loss_D.backward()
                if it == 0:
                    for k, p in model.netD.named_parameters():
                        out[f"{name}_gradD::{k}"] = p.grad.numpy().copy()
                model.optimizer_D.step()
            out[f"{name}_losses"] = np.array(losses_all)
            out[f"{name}_G_cksum_after"] = state_checksum(model.netG.state_dict())
            out[f"{name}_D_cksum_after"] = state_checksum(model.netD.state_dict())
            print(name, "done", list(model.loss_names), losses_all)
        np.savez_compressed(os.path.join(HERE, fname), **out)
        print("wrote", fname)
