"""GPU parity of the model facade: create_model(opt).inference(lr_audio) against the reference's own
create_model(opt).inference on the same seeded weights and audio (tests/golden/nets_golden.npz `inf_*`).
Bar (north_star): reconstructed waveform within 1e-3 relative L2 of the reference."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2

sys.path.insert(0, GOLDEN)
from make_golden_nets import INFER_FLAGS  # noqa: E402
from test_oracle_nets import INFER_CFG, our_opt  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nets_golden():
    return dict(np.load(os.path.join(GOLDEN, "nets_golden.npz")))


@pytest.mark.parametrize("name", ["inf_small", "inf_cfg3"])
@pytest.mark.parametrize("prec", ["fp32", "fp64"])
def test_inference_matches_reference(nets_golden, name, prec):
    import mdctgan_b200
    from mdctgan_b200.models.models import create_model

    g = nets_golden
    opt = our_opt(name, gpu="0")
    opt.mdct_precision = prec
    opt.num_D = 1
    from mdctgan_b200.models import networks
    from mdctgan_b200.models.pix2pixHD_model import InferenceModel

    model = InferenceModel()
    model.initialize(opt)            # isTrain=True: fresh weights instead of a checkpoint on disk
    # the golden weights were drawn from the CPU generator (the reference ran with --gpu_ids -1); with gpu_ids=[0]
    # define_G moves the net to the GPU before weights_init, i.e. draws from the CUDA generator -> load the CPU-seeded ones
    torch.manual_seed(INFER_FLAGS[name][3])
    cpu_G = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.netG, opt.n_downsample_global, opt.n_blocks_global,
                              opt.n_local_enhancers, opt.n_blocks_local, opt.norm, input_size=(opt.bins, opt.n_fft // 2),
                              n_attn_g=opt.n_blocks_attn_g, heads_g=opt.heads_g, dim_head_g=opt.dim_head_g)
    model.netG.load_state_dict(cpu_G.state_dict())
    model.eval()
    n0 = mdctgan_b200.launch_count()
    sr_spectro, sr_audio, lr_pha, prm, lr_spectro = model.inference(torch.from_numpy(g[f"{name}_lr_audio"]).cuda())
    assert mdctgan_b200.launch_count() - n0 > 10
    assert sr_audio.shape == g[f"{name}_sr_audio"].shape and sr_spectro.shape == g[f"{name}_sr_spectro"].shape
    assert sr_audio.dtype == (torch.float64 if prec == "fp64" else torch.float32)
    assert np.abs(lr_spectro.cpu().numpy() - g[f"{name}_lr_spectro"]).max() <= (6e-8 if prec == "fp64" else 5e-5)
    e_s = rel_l2(sr_spectro.cpu().numpy(), g[f"{name}_sr_spectro"])
    e_a = rel_l2(sr_audio.cpu().numpy(), g[f"{name}_sr_audio"])
    assert e_s < 1e-4, e_s
    assert e_a < 1e-3, e_a            # north_star bar; see DESIGN.md for the measured value


def test_inference_against_oracle_other_shape():
    """Seeded oracle comparison at a batch / segment length the goldens do not cover (B = 5, 64 frames)."""
    from mdctgan_b200.models.pix2pixHD_model import InferenceModel
    from make_golden_nets import make_lr_audio
    from oracle import model_oracle as MOD

    opt = our_opt("inf_small", gpu="0")
    opt.segment_length, opt.bins, opt.fit_residual, opt.num_D = 16128, 64, True, 1
    torch.manual_seed(5)
    model = InferenceModel()
    model.initialize(opt)
    model.eval()
    lr = make_lr_audio(5, 16128, 99)
    sd = {k: v.cpu() for k, v in model.netG.state_dict().items()}
    ref_sr, ref_audio, _ = MOD.inference(sd, lr.numpy(), n_down=2, n_blocks_global=2, n_blocks_local=1, fit_residual=True, up_ratio=4.0)
    sr_spectro, sr_audio, *_ = model.inference(lr.cuda())
    assert rel_l2(sr_spectro.cpu().numpy(), ref_sr) < 1e-4
    assert rel_l2(sr_audio.cpu().numpy(), ref_audio) < 1e-3


def test_checkpoint_round_trip(tmp_path):
    """save() writes <name>/<epoch>_net_G.pth with the reference key names; an InferenceModel loads it."""
    from mdctgan_b200.models.models import create_model

    opt = our_opt("inf_small", gpu="0")
    opt.checkpoints_dir, opt.name = str(tmp_path), "ck"
    torch.manual_seed(3)
    m = create_model(opt)
    m.save("latest")
    sd = torch.load(os.path.join(str(tmp_path), "ck", "latest_net_G.pth"))
    assert "model1_1.1.weight" in sd and "model.1.weight" in sd
    assert os.path.exists(os.path.join(str(tmp_path), "ck", "latest_net_D.pth"))
    opt.isTrain = False
    m2 = create_model(opt)
    for k, v in m.netG.state_dict().items():
        assert torch.equal(v, m2.netG.state_dict()[k])
    m.train()
    losses, _ = m._forward(0.1 * torch.randn(2, 3840).cuda(), 0.1 * torch.randn(2, 3840).cuda())
    assert len(losses) == len(m.loss_names) == 4 and all(torch.isfinite(v).item() and v.dim() == 0 for v in losses)


def test_niter_fix_global_trains_local_enhancer_only(tmp_path):
    """--niter_fix_global > 0 (pix2pixHD_model.py:333-347): optimizer_G steps only the `model1_*` parameters until
    update_fixed_params() (:654-662) replaces it by an Adam over the whole generator."""
    from mdctgan_b200.models.models import create_model

    opt = our_opt("inf_small", gpu="0")
    opt.checkpoints_dir, opt.name, opt.niter_fix_global = str(tmp_path), "fix", 2
    torch.manual_seed(4)
    m = create_model(opt)
    m.train()
    before = {k: v.detach().clone() for k, v in m.netG.named_parameters()}
    x, y = 0.05 * torch.randn(2, 3840).cuda(), 0.1 * torch.randn(2, 3840).cuda()
    m.train_step(x, y)
    moved = {k: not torch.equal(v, before[k]) for k, v in m.netG.named_parameters()}
    assert all(moved[k] for k in moved if k.startswith("model1") and k.endswith("weight"))
    assert not any(moved[k] for k in moved if not k.startswith("model1"))
    m.update_fixed_params()
    m.train_step(x, y)
    assert any(not torch.equal(v, before[k]) for k, v in m.netG.named_parameters() if not k.startswith("model1"))


def test_train_py_fp16_sequence_with_gradscaler(tmp_path):
    """train.py:160-202 with --fp16 verbatim (autocast around _forward, ONE GradScaler, scaler.scale(loss).backward(),
    scaler.step(optimizer), one scaler.update()): the loss scale reaches our backward kernels as the upstream gradient of the loss
    tensors, GradScaler unscales the flat-bucket views in place and steps FusedAdam.  The kernels compute in fp32 either way, and the
    scale is a power of two, so the unscaled gradients equal the plain run's."""
    from torch.cuda.amp import GradScaler, autocast

    from mdctgan_b200.models.models import create_model

    def build():
        opt = our_opt("inf_small", gpu="0")
        opt.checkpoints_dir, opt.name, opt.fp16 = str(tmp_path), "amp", True
        torch.manual_seed(8)
        torch.cuda.manual_seed(8)
        m = create_model(opt)
        m.train()
        return m

    g = torch.Generator().manual_seed(1)
    lr_a, hr_a = (0.05 * torch.randn(2, 3840, generator=g)).cuda(), (0.1 * torch.randn(2, 3840, generator=g)).cuda()
    m_amp, m_ref = build(), build()
    m_ref.bucket_G.flat.copy_(m_amp.bucket_G.flat)
    m_ref.bucket_D.flat.copy_(m_amp.bucket_D.flat)
    m_ref._refresh_weight_images()
    # ---- train.py --fp16
    scaler = GradScaler()
    optimizer_G, optimizer_D = m_amp.optimizer_G, m_amp.optimizer_D
    with autocast():
        losses, _ = m_amp._forward(lr_a, hr_a, infer=False)
    losses = [torch.mean(x) if not isinstance(x, int) else x for x in losses]
    loss_dict = dict(zip(m_amp.loss_names, losses))
    loss_D = (loss_dict["D_fake"] + loss_dict["D_real"]) * 0.5
    loss_G = loss_dict["G_GAN"] + loss_dict.get("G_GAN_Feat", 0)
    optimizer_G.zero_grad()
    scaler.scale(loss_G).backward()
    scaler.step(optimizer_G)
    gG_amp = m_amp.bucket_G.grad.clone()
    optimizer_D.zero_grad()
    scaler.scale(loss_D).backward()
    scaler.step(optimizer_D)
    scaler.update()
    gD_amp = m_amp.bucket_D.grad.clone()
    # ---- plain sequence on the same weights
    ls, _ = m_ref._forward(lr_a, hr_a)
    d = dict(zip(m_ref.loss_names, ls))
    m_ref.optimizer_G.zero_grad()
    (d["G_GAN"] + d["G_GAN_Feat"]).backward()
    gG = m_ref.bucket_G.grad.clone()
    m_ref.optimizer_G.step()
    m_ref.optimizer_D.zero_grad()
    ((d["D_fake"] + d["D_real"]) * 0.5).backward()
    gD = m_ref.bucket_D.grad.clone()
    m_ref.optimizer_D.step()
    assert torch.allclose(torch.stack([x.detach() for x in losses]), torch.stack([x.detach() for x in ls]), rtol=1e-6)
    assert float((gG_amp - gG).norm() / gG.norm()) < 1e-4 and float((gD_amp - gD).norm() / gD.norm()) < 1e-4
    assert m_amp.optimizer_G.step_count == 1 and m_amp.optimizer_D.step_count == 1
    moved = float((m_ref.bucket_G.flat - m_amp.bucket_G.flat).norm()) / float(2e-4 * m_ref.bucket_G.flat.numel() ** 0.5)
    assert moved < 0.2, moved


def test_load_network_fallbacks(tmp_path):
    """The reference's three-level load (models/base_model.py:49-111): strict -> "pretrained has excessive layers" (keep the keys
    the model has) -> "--param_key_map" remap of the second key component for renumbered layers."""
    from mdctgan_b200.models.models import create_model

    opt = our_opt("inf_small", gpu="0")
    opt.checkpoints_dir, opt.name = str(tmp_path), "fb"
    torch.manual_seed(5)
    src = create_model(opt)
    sd = {k: v.detach().cpu().clone() for k, v in src.netG.state_dict().items()}
    d = os.path.join(str(tmp_path), "fb")
    os.makedirs(d, exist_ok=True)
    # (2) excessive layers: extra keys in the file are dropped
    extra = dict(sd)
    extra["model.99.weight"] = torch.zeros(3, 3)
    extra["not_a_layer.bias"] = torch.ones(7)
    torch.save(extra, os.path.join(d, "a_net_G.pth"))
    torch.manual_seed(6)
    dst = create_model(opt)
    assert not torch.equal(dst.netG.state_dict()["model.1.weight"].cpu(), sd["model.1.weight"])
    dst.load_network(dst.netG, "G", "a")
    for k, v in sd.items():
        assert torch.equal(dst.netG.state_dict()[k].cpu(), v), k
    # (3) renumbered layers: the file calls the model's `model.4` layer `model.3` -> --param_key_map model.3:4
    #     (and lacks model.4 itself, so neither the strict nor the filtered load can succeed)
    ren = {}
    for k, v in sd.items():
        parts = k.split(".")
        if parts[0] == "model" and parts[1] == "4":
            parts[1] = "3"
        ren[".".join(parts)] = v
    assert "model.3.weight" in ren and "model.4.weight" not in ren
    torch.save(ren, os.path.join(d, "b_net_G.pth"))
    torch.manual_seed(7)
    dst2 = create_model(opt)
    dst2.opt.param_key_map = {"model.3": "4"}
    dst2.load_network(dst2.netG, "G", "b")
    for k, v in sd.items():
        assert torch.equal(dst2.netG.state_dict()[k].cpu(), v), k
    # the loaded weights are what the kernels then use (weight images are re-derived)
    dst2.eval()
    src.eval()
    x = 0.05 * torch.randn(2, 3840).cuda()
    assert torch.equal(dst2.inference(x)[1], src.inference(x)[1])
    # a missing generator file is an error, a missing discriminator file is not (base_model.py:54-57)
    with pytest.raises(FileNotFoundError):
        dst2.load_network(dst2.netG, "G", "nope")
    dst2.load_network(dst2.netD, "D", "nope")


def test_graphed_train_step_equals_eager_steps(tmp_path):
    """runtime.GraphedTrainStep: capture / warm-up must not train: N replays == N eager train_step calls from the same
    state (float atomics in the weight-gradient reductions make the comparison a tolerance, not bit equality), the host step counter
    follows the device one, and a new optimiser (update_fixed_params) or learning rate triggers a re-capture."""
    from mdctgan_b200.models.models import create_model
    from mdctgan_b200.runtime import GraphedTrainStep

    opt = our_opt("inf_small", gpu="0")
    opt.checkpoints_dir, opt.name = str(tmp_path), "g"
    x, y = 0.05 * torch.randn(2, 3840).cuda(), 0.1 * torch.randn(2, 3840).cuda()
    torch.manual_seed(8)
    a = create_model(opt)
    a.train()
    torch.manual_seed(8)
    b = create_model(opt)
    b.train()
    b.netG.load_state_dict(a.netG.state_dict())
    b.netD.load_state_dict(a.netD.state_dict())
    w0 = a.bucket_G.flat.clone()
    gts = GraphedTrainStep(a, 2, 3840)
    gts.lr_in.copy_(x)
    gts.hr_in.copy_(y)
    gts.recapture()
    assert torch.equal(a.bucket_G.flat, w0) and a.optimizer_G.step_count == 0 and int(a.optimizer_G.step_dev) == 0      # no training yet
    la, lb = [], []
    for _ in range(3):
        la.append(gts(x, y).clone())
        lb.append(b.train_step(x, y).clone())
    assert a.optimizer_G.step_count == 3 == int(a.optimizer_G.step_dev) and a.optimizer_D.step_count == 3
    for u, v in zip(la, lb):
        assert torch.allclose(u, v, rtol=2e-3, atol=0), (u, v)
    assert torch.allclose(la[0], lb[0], rtol=1e-5, atol=0)                # first step: identical weights
    moved = float((b.bucket_G.flat - w0).norm())
    assert float((a.bucket_G.flat - b.bucket_G.flat).norm()) < 0.05 * moved
    # learning-rate change -> re-capture, still no extra training
    a.update_learning_rate()
    w1, sc = a.bucket_G.flat.clone(), a.optimizer_G.step_count
    gts.lr_in.copy_(x)
    gts.recapture()
    assert torch.equal(a.bucket_G.flat, w1) and a.optimizer_G.step_count == sc
    gts(x, y)
    assert a.optimizer_G.step_count == sc + 1 and not torch.equal(a.bucket_G.flat, w1)
