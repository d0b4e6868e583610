"""Diagnostic (GPU box): forward accuracy of the generator at BASELINE cfg4 -- ours vs the fp64 oracle vs the fp32 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
from make_golden_nets import TRAIN_FLAGS
from oracle import train_oracle as TO
from test_oracle_train import build_nets, flags_to_cfg
from test_train_gpu import _build_model
from mdctgan_b200 import nn_ops as ops

dev = torch.device("cuda:0")
gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "train_golden.npz")))
name = "tr_cfg4"
flags, batch, T, seed = TRAIN_FLAGS[name]
cfg = flags_to_cfg(flags)
G0, D0 = build_nets(cfg, seed)
kw = {k: cfg[k] for k in ("netG", "n_down", "n_blocks_global", "n_blocks_local", "n_attn", "heads", "dim_head", "num_D", "n_layers_D", "fit_residual", "down", "up")}
lr_a, hr_a = gold[f"{name}_lr_audio"], gold[f"{name}_hr_audio"]
sdG, sdD = G0.state_dict(), D0.state_dict()
r32 = TO.train_step(sdG, sdD, lr_a, hr_a, steps=1, **kw)
r64 = TO.train_step(sdG, sdD, lr_a, hr_a, steps=1, dtype=torch.float64, **kw)
truth = r64["sr_spectro"].double()
print("oracle fp32 vs fp64: sr rel-L2 %.3e" % float((r32["sr_spectro"].double() - truth).norm() / truth.norm()))
for eng in ("umma", "direct", "tf32"):
    ops.CONV_ENGINE = eng
    model = _build_model(flags, seed, dev)
    model.netG.load_state_dict(sdG); model.netD.load_state_dict(sdD)
    model._refresh_weight_images()
    s1 = TO.spectro(lr_a).to(dev)
    x = torch.cat((s1, s1.abs() * 2 - 1.0), dim=1).contiguous()
    sr = model.netG.forward(x) + s1
    print("%-6s vs fp64: sr rel-L2 %.3e   vs oracle fp32 %.3e" % (eng, float((sr.cpu().double() - truth).norm() / truth.norm()),
                                                               float((sr.cpu().double() - r32["sr_spectro"].double()).norm() / truth.norm())))
