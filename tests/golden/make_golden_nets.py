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


def gen_train():
    raise SystemExit("train-step goldens: not generated in this round")
