"""Oracle (CPU) for Pix2PixHDModel.inference -- TEST INFRASTRUCTURE ONLY.

Composition of the pinned pieces, following /root/reference/models/pix2pixHD_model.py:618-638:
  to_spectro (oracle/mdct_oracle.py) -> cat(s, |s|*2+lo) -> generator (oracle/networks_oracle.py)
  -> [fit_residual: sr[..., :N/up_ratio] *= 1e-3; sr += lr] -> to_audio (oracle/mdct_oracle.py)
Pinned against tests/golden/nets_golden.npz (`inf_*` entries = the reference's own create_model(opt).inference).
"""
import numpy as np
import torch

from . import mdct_oracle as MO
from . import networks_oracle as NO


def inference(sd, lr_audio, *, netG="local", n_down=3, n_blocks_global=9, n_blocks_local=3, n_attn=0, heads=4, dim_head=128,
              fit_residual=False, up_ratio=4.0, gain=1000.0, src_range=(-5.0, 5.0), norm_range=(-1.0, 1.0)):
    w = MO.kbdwin(512)
    kw = dict(arcsinh_transform=True, arcsinh_gain=gain, abs_norm=True, src_range=src_range, norm_range=norm_range)
    lr_spectro, _, hi, lo = MO.to_spectro(np.asarray(lr_audio, dtype=np.float32), w, **kw)
    s = torch.from_numpy(lr_spectro)
    x = torch.cat((s, s.abs() * 2 + norm_range[0]), dim=1)
    with torch.no_grad():
        if netG == "global":
            sr = NO.global_generator(sd, x, n_down, n_blocks_global, n_attn, heads, dim_head)
        else:
            sr = NO.local_enhancer(sd, x, n_down, n_blocks_global, n_blocks_local, n_attn, heads, dim_head)
    if fit_residual:
        lr_part = int(sr.shape[-1] / up_ratio)
        sr = sr.clone()
        sr[..., :lr_part] *= 1e-3
        sr = sr + s
    audio = MO.to_audio(sr.numpy(), lo, hi, w, arcsinh_transform=True, arcsinh_gain=gain, norm_range=norm_range)
    return sr.numpy(), audio, lr_spectro
