#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (read-only /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [mdct] [nets] [train]

The reference is imported unmodified with the three stubs SURVEY.md section 8c lists
(torch_scatter, matplotlib -> dummies; bottleneck_transformer_pytorch -> our
restatement oracle/bottlestack_ref.py, which is therefore "parity unpinned").
Nothing from the reference is copied into the repo: only its numeric outputs.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MDCTGAN_REFERENCE", "/root/reference")


def import_reference():
    sys.path.insert(0, ROOT)
    ts = types.ModuleType("torch_scatter")
    ts.scatter = None
    sys.modules["torch_scatter"] = ts
    mp = types.ModuleType("matplotlib")
    pp = types.ModuleType("matplotlib.pyplot")
    pp.switch_backend = lambda *a, **k: None
    mp.pyplot = pp
    sys.modules["matplotlib"] = mp
    sys.modules["matplotlib.pyplot"] = pp
    import oracle.bottlestack_ref as bs

    sys.modules["bottleneck_transformer_pytorch"] = bs
    sys.path.insert(0, REF)


def ref_opt(extra=()):
    """TrainOptions().parse() with the hot-path flags."""
    import tempfile

    from options.train_options import TrainOptions

    tmp = tempfile.mkdtemp()
    argv = ["x", "--name", "g", "--checkpoints_dir", tmp, "--gpu_ids", "-1",
            "--lr_sampling_rate", "12000", "--sr_sampling_rate", "48000",
            "--arcsinh_transform", "--abs_spectro", "--arcsinh_gain", "1000", "--center",
            "--norm_range", "-1", "1", "--abs_norm", "--src_range", "-5", "5"] + list(extra)
    old = sys.argv
    sys.argv = argv
    try:
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            opt = TrainOptions().parse()
    finally:
        sys.argv = old
    return opt


LONGFORM_CASES = [(100000, 7936, 0), (100000, 7936, 256), (50000, 32512, 4096), (5000, 7936, 0), (5000, 7936, 256), (63488, 7936, 0),
                  (63488, 7936, 512), (9000, 3840, 128), (2880000, 32512, 256)]
LONGFORM_FULL = {(5000, 7936, 256), (9000, 3840, 128)}      # cases whose full outputs are stored (the others: checksums + probes)


def longform_clip(L, seg, ov):
    torch.manual_seed(L + seg + ov)
    return 0.1 * torch.randn(1, L)


def gen_longform():
    """AudioTestDataset.seg_pad_audio of the reference itself (data/audio_dataset.py:153-167) on seeded clips; the overlap-add
    block of generate_audio.py:40-53 is script-level code, so it is run through its restatement (oracle/longform_oracle.py)."""
    import torchaudio

    if not hasattr(torchaudio, "set_audio_backend"):
        torchaudio.set_audio_backend = lambda *a, **k: None      # removed API, called at import time (audio_dataset.py:9)
    from data.audio_dataset import AudioTestDataset

    from oracle import longform_oracle as LO

    out = {}
    for (L, seg, ov) in LONGFORM_CASES:
        clip = longform_clip(L, seg, ov)
        ds = object.__new__(AudioTestDataset)
        ds.segment_length, ds.overlap = seg, ov
        segs = ds.seg_pad_audio(clip.clone())
        assert torch.equal(segs, LO.seg_pad_audio(clip.clone(), seg, ov))
        key = f"{L}_{seg}_{ov}"
        ola = LO.overlap_add(segs.double().reshape(-1, 1, 1, seg), seg, ov)
        out[f"segshape_{key}"] = np.array(segs.shape)
        out[f"segck_{key}"] = np.array([float(segs.double().sum()), float((segs.double() ** 2).sum())])
        out[f"olashape_{key}"] = np.array(ola.shape)
        out[f"olack_{key}"] = np.array([float(ola.sum()), float((ola ** 2).sum())])
        idx = np.linspace(0, ola.shape[-1] - 1, 257).astype(np.int64)
        out[f"olaprobe_{key}"] = ola[0, idx].numpy()
        if (L, seg, ov) in LONGFORM_FULL:
            out[f"seg_{key}"] = segs.numpy()
            out[f"ola_{key}"] = ola.numpy()
        print(key, tuple(segs.shape), tuple(ola.shape))
    np.savez_compressed(os.path.join(HERE, "longform_golden.npz"), **out)


METRIC_CASES = [(1, 20000, True, 3), (4, 7936, True, 4), (2, 32512, False, 5), (1, 2870528 // 8, True, 6)]


def metric_signals(rows, T, seed):
    g = torch.Generator().manual_seed(seed)
    hr = 0.1 * torch.randn(rows, T, generator=g)
    lr = hr + 0.05 * torch.randn(rows, T, generator=g)
    sr = hr + 0.01 * torch.randn(rows, T, generator=g)
    return hr, lr, sr


def gen_metrics():
    """compute_matrics of the reference itself (util/util.py:132-177)."""
    from types import SimpleNamespace

    from util.util import compute_matrics

    out = {}
    for (rows, T, center, seed) in METRIC_CASES:
        hr, lr, sr = metric_signals(rows, T, seed)
        opt = SimpleNamespace(n_fft=512, hop_length=256, win_length=512, center=center, hr_sampling_rate=48000)
        out[f"m_{rows}_{T}_{int(center)}"] = np.array(compute_matrics(hr, lr, sr, opt), dtype=np.float64)
        print(rows, T, center, out[f"m_{rows}_{T}_{int(center)}"])
    np.savez_compressed(os.path.join(HERE, "metrics_golden.npz"), **out)


RESAMPLE_CASES = [(48000, 12000, 4001), (12000, 48000, 1000), (48000, 16000, 4000), (16000, 48000, 1333), (44100, 48000, 2205),
                  (48000, 8000, 3000)]


def resample_wave(L, seed):
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(2, L, generator=g)


def gen_resample():
    """torchaudio.functional.resample -- the function the reference calls (data/audio_dataset.py:66-71) -- on seeded waveforms."""
    import torchaudio.functional as aF

    out = {}
    for i, (o, n, L) in enumerate(RESAMPLE_CASES):
        out[f"r_{o}_{n}_{L}"] = aF.resample(resample_wave(L, 300 + i), o, n).numpy()
        print(o, n, L, out[f"r_{o}_{n}_{L}"].shape)
    np.savez_compressed(os.path.join(HERE, "resample_golden.npz"), **out)


AUGMENT_CASES = [  # (L, orig_sr, lr_sr, hr_sr, segment_length, add_noise, snr, seed)
    (40000, 48000, 12000, 48000, 32512, True, 55.0, 801), (20000, 48000, 16000, 48000, 32512, True, 30.0, 802),
    (30000, 44100, 8000, 48000, 7936, False, 55.0, 803), (9000, 48000, 12000, 48000, 7936, True, 10.0, 804)]


def gen_augment():
    """AudioDataset.__getitem__ of the reference itself (data/audio_dataset.py:54-82) with `readaudio` stubbed to return a seeded
    waveform: the three resamples, the noise injection (torch.randn under a known seed) and seg_pad_audio."""
    import torchaudio

    if not hasattr(torchaudio, "set_audio_backend"):
        torchaudio.set_audio_backend = lambda *a, **k: None
    from data.audio_dataset import AudioDataset

    out = {}
    for (L, osr, lsr, hsr, seg, noise_on, snr, seed) in AUGMENT_CASES:
        ds = object.__new__(AudioDataset)
        ds.hr_sampling_rate, ds.lr_sampling_rate, ds.segment_length, ds.add_noise, ds.snr = hsr, lsr, seg, noise_on, snr
        g = torch.Generator().manual_seed(seed)
        wave = 0.1 * torch.randn(1, L, generator=g)
        ds.readaudio = lambda idx, _w=wave, _sr=osr: (_w.clone(), _sr)
        torch.manual_seed(seed + 1)                     # the draw __getitem__ makes: torch.randn(lr_waveform.size())
        item = ds[0]
        key = f"{L}_{osr}_{lsr}_{seg}_{int(noise_on)}"
        out[f"hr_{key}"], out[f"lr_{key}"] = item["HR_audio"].numpy(), item["LR_audio"].numpy()
        print(key, item["HR_audio"].shape, item["LR_audio"].shape)
    np.savez_compressed(os.path.join(HERE, "augment_golden.npz"), **out)


OPTION_ARGV = [[], ["--netG", "local", "--ngf", "56", "--fp16", "--batchSize", "20", "--lr", "1.5e-4", "--upsample_type", "interpolate",
                     "--downsample_type", "resconv", "--n_blocks_attn_g", "3", "--heads_g", "6", "--niter", "60", "--niter_decay", "60",
                     "--num_D", "3", "--fit_residual", "--param_key_map", "model.1:2,model.4:5", "--gen_overlap", "256", "--phase", "test"]]


def gen_options():
    """The Namespace TrainOptions().parse() of the reference returns (options/base_options.py, train_options.py) for the default
    command line of ref_opt() and for the train.sh one: the flag surface create_model(opt) consumes."""
    import json

    out = []
    for extra in OPTION_ARGV:
        opt = ref_opt(extra)
        d = {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in vars(opt).items() if k != "checkpoints_dir"}
        d = {k: (v if isinstance(v, (int, float, str, bool, list, dict, type(None))) else str(v)) for k, v in d.items()}
        out.append(d)
    with open(os.path.join(HERE, "options_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("options:", len(out[0]), "fields")


ENCODING_CASES = [("db_abs", dict(arcsinh_transform=False, raw_mdct=False, explicit_encoding=False, abs_norm=True, src_range=[-180.0, 20.0])),
                  ("db_minmax", dict(arcsinh_transform=False, raw_mdct=False, explicit_encoding=False, abs_norm=False)),
                  ("explicit_minmax", dict(arcsinh_transform=False, raw_mdct=False, explicit_encoding=True, abs_norm=False)),
                  ("explicit_abs", dict(arcsinh_transform=False, raw_mdct=False, explicit_encoding=True, abs_norm=True, src_range=[-180.0, 20.0])),
                  ("arcsinh_minmax", dict(arcsinh_transform=True, raw_mdct=False, explicit_encoding=False, abs_norm=False)),
                  ("raw_minmax", dict(arcsinh_transform=False, raw_mdct=True, explicit_encoding=False, abs_norm=False))]


def gen_encodings():
    """Audio2MDCT.to_spectro / to_audio of the reference itself (pix2pixHD_model.py:32-163) on the secondary encodings: dB magnitude
    (default options), --explicit_encoding, and the per-sample min / max normalisation (no --abs_norm) of every encoding.  The random
    factors are kept out of the golden: `pha` is passed as the plain sign (the reference multiplies it by noise, :49-54), and
    up_ratio is set just above 1 so that only the LAST frame receives the random pseudo phase of :150-157 (tests compare the samples
    no later frame touches: [: (F - 2) * hop])."""
    from models.pix2pixHD_model import Audio2MDCT

    out = {}
    g = torch.Generator().manual_seed(33)
    audio = 0.1 * torch.randn(2, 7936, generator=g)
    audio[1] *= 0.01
    logs = (torch.rand(2, 2, 32, 256, generator=g) * 2 - 1).float()
    out["audio"], out["log_spectro"] = audio.numpy(), logs.numpy()
    for tag, over in ENCODING_CASES:
        opt = ref_opt(["--segment_length", "7936", "--bins", "32"])
        for k, v in over.items():
            setattr(opt, k, v)
        a2m = Audio2MDCT(opt)
        torch.manual_seed(5)
        s, pha, npar = a2m.to_spectro(audio.clone())
        out[f"{tag}_spectro"] = s.numpy()
        out[f"{tag}_max"], out[f"{tag}_min"] = np.asarray(npar["max"].numpy(), dtype=np.float32), np.asarray(npar["min"].numpy(), dtype=np.float32)
        raw, _ = a2m._mdct(audio.clone(), True)
        sign = torch.sign(raw.unsqueeze(1))
        out["sign"] = sign.numpy().astype(np.float32)
        a2m.up_ratio = 1.0001
        rt = a2m.to_audio(s.clone(), npar, sign.clone())
        C = s.shape[1]
        out[f"{tag}_decode"] = a2m.to_audio(logs[:, :C].clone(), npar, sign.clone()).numpy()
        keep = 30 * 256
        print(tag, tuple(s.shape), s.dtype, tuple(rt.shape), rt.dtype, "round trip max err", float((rt.reshape(2, -1)[:, :keep].float() - audio[:, :keep]).abs().max()))
    np.savez_compressed(os.path.join(HERE, "encodings_golden.npz"), **out)


def gen_normalize():
    """Audio2MDCT.normalize / denormalize of the reference itself (pix2pixHD_model.py:83-137), arcsinh and raw branches, abs_norm."""
    from models.pix2pixHD_model import Audio2MDCT

    out = {}
    g = torch.Generator().manual_seed(21)
    spectro = 0.02 * torch.randn(1, 1, 4, 256, generator=g, dtype=torch.float64)
    logs = (torch.rand(1, 1, 4, 256, generator=g) * 2 - 1).float()
    out["spectro"], out["log_spectro"] = spectro.numpy(), logs.numpy()
    for tag, extra in (("arcsinh", []), ("raw", ["--raw_mdct"])):
        flags = ["--segment_length", "7936", "--bins", "32"] + extra
        opt = ref_opt(flags)
        if tag == "raw":
            opt.arcsinh_transform = False
        a2m = Audio2MDCT(opt)
        y, mx, mn, mean, std = a2m.normalize(spectro.clone())
        out[f"{tag}_norm"] = y.numpy()
        out[f"{tag}_denorm"] = a2m.denormalize(logs.clone(), mn, mx).numpy()
        print(tag, y.dtype, out[f"{tag}_denorm"].dtype)
    np.savez_compressed(os.path.join(HERE, "normalize_golden.npz"), **out)


def gen_mdct():
    from models.mdct import IMDCT4, MDCT4
    from models.pix2pixHD_model import Audio2MDCT
    from util.util import kbdwin

    out = {}
    w = kbdwin(512)
    out["kbdwin512"] = w.numpy()
    out["kbdwin1024"] = kbdwin(1024).numpy()
    out["kbdwin64"] = kbdwin(64).numpy()
    fwd = MDCT4(n_fft=512, hop_length=256, win_length=512, window=w, device="cpu")
    inv = IMDCT4(n_fft=512, hop_length=256, win_length=512, window=w, device="cpu")

    # cfg1: one 8192-sample clip, 1-D input (BASELINE.json configs[0])
    torch.manual_seed(0)
    x = 0.1 * torch.randn(8192)
    spec, frames = fwd(x, True)
    audio, _ = inv(spec.unsqueeze(0).clone())
    out["c1_x"] = x.numpy()
    out["c1_spec"] = spec.numpy()
    out["c1_frames"] = frames.numpy()
    out["c1_audio"] = audio.numpy()

    # batched 2-D input, 32 frames (network segment length 7936)
    torch.manual_seed(1)
    xb = 0.1 * torch.randn(4, 7936)
    specb, _ = fwd(xb)
    audiob, _ = inv(specb.clone())
    out["b4_x"] = xb.numpy()
    out["b4_spec"] = specb.numpy()
    out["b4_audio"] = audiob.numpy()

    # ragged lengths: the len(signal)==batch quirk (mdct.py:394-402)
    torch.manual_seed(2)
    xr = 0.1 * torch.randn(3, 1000)
    specr, _ = fwd(xr)
    out["r3_x"] = xr.numpy()
    out["r3_spec"] = specr.numpy()
    spec1, _ = fwd(xr[0])
    out["r1_spec"] = spec1.numpy()
    xq = 0.1 * torch.randn(4, 8193)
    out["q4_x"] = xq.numpy()
    out["q4_spec"] = fwd(xq)[0].numpy()
    out["q1_spec"] = fwd(xq[0])[0].numpy()
    # IMDCT with out_length crop
    inv_c = IMDCT4(n_fft=512, hop_length=256, win_length=512, window=w, out_length=1000, device="cpu")
    out["r3_audio_crop"] = inv_c(specr.clone())[0].numpy()

    # a second transform size (n_fft 1024 / hop 512) to pin the generic formulas
    w2 = kbdwin(1024)
    f2 = MDCT4(n_fft=1024, hop_length=512, win_length=1024, window=w2, device="cpu")
    i2 = IMDCT4(n_fft=1024, hop_length=512, win_length=1024, window=w2, device="cpu")
    torch.manual_seed(3)
    x2 = 0.1 * torch.randn(2, 4096)
    s2 = f2(x2)[0]
    out["n1024_x"] = x2.numpy()
    out["n1024_spec"] = s2.numpy()
    out["n1024_audio"] = i2(s2.clone())[0].numpy()

    # Audio2MDCT: arcsinh + abs_norm (the configs' branch), gain 1000, range (-1,1)
    opt = ref_opt(["--segment_length", "7936", "--bins", "32"])
    pre = Audio2MDCT(opt)
    torch.manual_seed(4)
    ls, pha, prm = pre.to_spectro(xb)
    out["a2m_log_spectro"] = ls.numpy()
    out["a2m_max"] = prm["max"].numpy()
    out["a2m_min"] = prm["min"].numpy()
    out["a2m_mean"] = prm["mean"].numpy()
    out["a2m_std"] = prm["std"].numpy()
    out["a2m_audio"] = pre.to_audio(ls, prm, pha).numpy()
    # per-sample min/max branch (no --abs_norm) and raw_mdct branch
    opt2 = ref_opt(["--segment_length", "7936", "--bins", "32"])
    opt2.abs_norm = False
    pre2 = Audio2MDCT(opt2)
    ls2, pha2, prm2 = pre2.to_spectro(xb)
    out["mm_log_spectro"] = ls2.numpy()
    out["mm_max"] = prm2["max"].numpy()
    out["mm_min"] = prm2["min"].numpy()
    out["mm_audio"] = pre2.to_audio(ls2, prm2, pha2).numpy()
    opt3 = ref_opt(["--segment_length", "7936", "--bins", "32"])
    opt3.arcsinh_transform = False
    opt3.raw_mdct = True
    pre3 = Audio2MDCT(opt3)
    ls3, pha3, prm3 = pre3.to_spectro(xb)
    out["raw_log_spectro"] = ls3.numpy()
    out["raw_audio"] = pre3.to_audio(ls3, prm3, pha3).numpy()

    np.savez_compressed(os.path.join(HERE, "mdct_golden.npz"), **out)
    print("wrote mdct_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    import_reference()
    what = sys.argv[1:] or ["mdct"]
    if "mdct" in what:
        gen_mdct()
    if "longform" in what:
        gen_longform()
    if "metrics" in what:
        gen_metrics()
    if "normalize" in what:
        gen_normalize()
    if "encodings" in what:
        gen_encodings()
    if "resample" in what:
        gen_resample()
    if "augment" in what:
        gen_augment()
    if "options" in what:
        gen_options()
    if "nets" in what or "train" in what or "infer" in what:   # ("train_bce" has its own entry below)
        from make_golden_nets import gen_nets, gen_train  # noqa: E402

        if "nets" in what:
            gen_nets()
        if "infer" in what:
            from make_golden_nets import gen_infer

            gen_infer()
        if "train" in what:
            gen_train()
    if "train_bce" in what or "train_attnl" in what:
        from make_golden_nets import gen_train_extra  # noqa: E402

        for w in what:
            if w in ("train_bce", "train_attnl"):
                gen_train_extra(w + "_golden.npz")
