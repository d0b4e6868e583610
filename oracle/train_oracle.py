"""Oracle (CPU, torch autograd) for the GAN train step -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/models/pix2pixHD_model.py:394-451 (forward, _forward: three discriminator passes, LSGAN
and feature-matching losses, networks.py:97-137) and /root/reference/train.py:175-202 (loss_G / loss_D, two
backward passes, two Adam steps, pix2pixHD_model.py:350-364) over reference-layout state_dicts, with
oracle/mdct_oracle.py for the spectrograms and oracle/networks_oracle.py for G and D.
Pinned against tests/golden/train_golden.npz (the reference's own create_model(opt)._forward / backward / Adam on
seeded weights and audio, tests/golden/make_golden_nets.py train) by tests/test_oracle_train.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import mdct_oracle as MO
from . import networks_oracle as NO


def spectro(audio, gain=1000.0, src_range=(-5.0, 5.0), norm_range=(-1.0, 1.0)):
    w = MO.kbdwin(512)
    s, _, _, _ = MO.to_spectro(np.asarray(audio, dtype=np.float32), w, arcsinh_transform=True, arcsinh_gain=gain, abs_norm=True,
                               src_range=src_range, norm_range=norm_range)
    return torch.from_numpy(s)


def losses(pG, pD, lr_spectro, hr_spectro, *, netG="local", n_down=3, n_blocks_global=9, n_blocks_local=3, n_attn=0, heads=4, dim_head=128,
           num_D=3, n_layers_D=3, lambda_feat=10.0, fit_residual=True, lo=-1.0, use_feat=True, down="conv", up="transconv", lsgan=True,
           n_attn_l=0, heads_l=4, dim_head_l=128):
    """The four loss tensors [G_GAN, G_GAN_Feat, D_real, D_fake] as autograd functions of the parameter dicts."""
    x = torch.cat((lr_spectro, lr_spectro.abs() * 2 + lo), dim=1)
    if netG == "global":
        sr = NO.global_generator(pG, x, n_down, n_blocks_global, n_attn, heads, dim_head, training=True, down=down, up=up)
    else:
        sr = NO.local_enhancer(pG, x, n_down, n_blocks_global, n_blocks_local, n_attn, heads, dim_head, training=True, down=down, up=up,
                               n_attn_l=n_attn_l, heads_l=heads_l, dim_head_l=dim_head_l)
    if fit_residual:
        sr = sr + lr_spectro
    sr_in = torch.cat((sr, sr.abs() * 2 + lo), dim=1)
    hr_in = torch.cat((hr_spectro, hr_spectro.abs() * 2 + lo), dim=1)
    dkw = dict(interm=use_feat, sigmoid=not lsgan)          # define_D(..., use_sigmoid = no_lsgan, getIntermFeat = not no_ganFeat_loss)
    pred_fake_pool = NO.multiscale_d(pD, torch.cat((lr_spectro, sr_in.detach()), dim=1), num_D, n_layers_D, **dkw)
    pred_real = NO.multiscale_d(pD, torch.cat((lr_spectro, hr_in), dim=1), num_D, n_layers_D, **dkw)
    pred_fake = NO.multiscale_d(pD, torch.cat((lr_spectro, sr_in), dim=1), num_D, n_layers_D, **dkw)
    gan = F.mse_loss if lsgan else F.binary_cross_entropy       # GANLoss: nn.MSELoss / nn.BCELoss (networks.py:105-108)
    d_fake = sum(gan(p[-1], torch.zeros_like(p[-1])) for p in pred_fake_pool)
    d_real = sum(gan(p[-1], torch.ones_like(p[-1])) for p in pred_real)
    g_gan = sum(gan(p[-1], torch.ones_like(p[-1])) for p in pred_fake)
    g_feat = torch.zeros((), device=lr_spectro.device)
    if use_feat:
        fw, dw = 4.0 / (n_layers_D + 1), 1.0 / num_D
        for i in range(num_D):
            for j in range(len(pred_fake[i]) - 1):
                g_feat = g_feat + dw * fw * F.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) * lambda_feat
    return [g_gan, g_feat, d_real, d_fake], sr


def train_step(sdG, sdD, lr_audio, hr_audio, *, lr=2e-4, beta1=0.5, steps=1, dtype=None, **cfg):
    """`steps` iterations of train.py:160-202 on the same batch.  Returns per-step losses, the gradients of the
    first step and the parameters after the last one.  `dtype=torch.float64` runs the same graph in double precision on the
    same fp32 inputs and weights: the ground truth the fp32 implementations (this oracle in fp32, the CUDA path) are measured
    against in the conditioning-aware gradient test (tests/test_train_gpu.py)."""
    if dtype is not None:
        sdG = {k: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in sdG.items()}
        sdD = {k: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in sdD.items()}
    floatsG = {k for k, v in sdG.items() if v.dtype.is_floating_point and "running_" not in k}
    pG = {k: (v.clone().requires_grad_(True) if k in floatsG else v.clone()) for k, v in sdG.items()}
    pD = {k: v.clone().requires_grad_(True) for k, v in sdD.items()}
    optG = torch.optim.Adam([pG[k] for k in pG if k in floatsG], lr=lr, betas=(beta1, 0.999))
    optD = torch.optim.Adam(list(pD.values()), lr=lr, betas=(beta1, 0.999))
    ls, hs = spectro(lr_audio), spectro(hr_audio)
    if dtype is not None:
        ls, hs = ls.to(dtype), hs.to(dtype)
    out = {"losses": []}
    for it in range(steps):
        (g_gan, g_feat, d_real, d_fake), sr = losses(pG, pD, ls, hs, **cfg)
        out["losses"].append([float(g_gan), float(g_feat), float(d_real), float(d_fake)])
        loss_D = (d_fake + d_real) * 0.5
        loss_G = g_gan + g_feat
        optG.zero_grad()
        loss_G.backward()
        if it == 0:
            out["gradG"] = {k: pG[k].grad.detach().clone() for k in pG if k in floatsG and pG[k].grad is not None}
            out["sr_spectro"] = sr.detach().clone()
        optG.step()
        optD.zero_grad()
        loss_D.backward()
        if it == 0:
            out["gradD"] = {k: v.grad.detach().clone() for k, v in pD.items()}
        optD.step()
    out["paramsG"] = {k: v.detach().clone() for k, v in pG.items()}
    out["paramsD"] = {k: v.detach().clone() for k, v in pD.items()}
    return out


def make_stepper(sdG, sdD, lr_audio, hr_audio, *, lr=2e-4, beta1=0.5, device="cpu", cuda_graph=False, **cfg):
    """A closure running one full iteration of train.py:160-202 per call (spectrograms included), for the timed baselines of
    bench.py; returns the four loss values of the iteration.  `device="cuda"` runs the very same torch ops on the GPU (cuDNN /
    cuFFT, torch autograd, torch.optim.Adam): the torch-eager "kernel to beat" (SURVEY.md 8d); `cuda_graph=True` additionally
    captures the whole iteration in one torch.cuda.CUDAGraph (capturable Adam) and the closure replays it (returns None)."""
    dev = torch.device(device)
    floatsG = {k for k, v in sdG.items() if v.dtype.is_floating_point and "running_" not in k}
    pG = {k: (v.detach().to(dev).clone().requires_grad_(True) if k in floatsG else v.detach().to(dev).clone()) for k, v in sdG.items()}
    pD = {k: v.detach().to(dev).clone().requires_grad_(True) for k, v in sdD.items()}
    kw = dict(capturable=True) if cuda_graph else {}
    optG = torch.optim.Adam([pG[k] for k in pG if k in floatsG], lr=lr, betas=(beta1, 0.999), **kw)
    optD = torch.optim.Adam(list(pD.values()), lr=lr, betas=(beta1, 0.999), **kw)
    from . import torch_port as P

    a2m = P.Audio2MDCTPort(1000.0, (-5.0, 5.0), (-1.0, 1.0), 512, 256, device=dev)
    lr_t, hr_t = torch.as_tensor(lr_audio).to(dev), torch.as_tensor(hr_audio).to(dev)

    def iteration():
        with torch.no_grad():
            ls, hs = a2m.to_spectro(lr_t)[0], a2m.to_spectro(hr_t)[0]      # the reference's complex128 512-pt FFT formulation
        (g_gan, g_feat, d_real, d_fake), _ = losses(pG, pD, ls, hs, **cfg)
        loss_D = (d_fake + d_real) * 0.5
        loss_G = g_gan + g_feat
        optG.zero_grad()
        loss_G.backward()
        optG.step()
        optD.zero_grad()
        loss_D.backward()
        optD.step()
        return g_gan, g_feat, d_real, d_fake

    if not cuda_graph:
        def step():
            return [float(v.detach()) for v in iteration()]

        return step

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            iteration()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    optG.zero_grad(set_to_none=True)
    optD.zero_grad(set_to_none=True)
    with torch.cuda.graph(graph):
        static_losses = torch.stack([v.detach() for v in iteration()])

    def replay():
        graph.replay()
        return static_losses

    return replay
