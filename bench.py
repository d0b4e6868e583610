#!/usr/bin/env python
"""bench.py -- the GAN train step of the MDCT -> generator -> IMDCT hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): train-step audio-seconds per second.  Workload = BASELINE configs[3] at its per-GPU slice:
LocalEnhancer (ngf 32, 3 downsamplings, 9 global + 3 local residual blocks, 2 BottleStack attention layers with
4 heads x 64) + MultiscaleDiscriminator (num_D 3, 3 layers, ndf 64) + LSGAN + feature matching + two Adams,
--fit_residual, 12 -> 48 kHz, fp32, 4 segments of 7936 samples (32 frames; 8192 samples give 33 frames, which the
reference generator cannot process) per GPU: global batch 4 N (32 at N = 8, the configuration the
metric is quoted on), weak scaling, ONE NCCL all-reduce of the flat [grad_G | grad_D] bucket per step.
A "step" = one iteration of train.py:160-202: 2 fused MDCT launches, G forward, D forward on [fake ; real], the
four losses, generator sweep, discriminator sweep, all-reduce, both Adam steps.  Rank 0 prints ONE JSON line:

  value        whole-job audio-s/s, batch resident in HBM, the step captured as a CUDA graph (runtime.GraphedTrainStep)
               and replayed K times, CUDA events, max over ranks
  e2e          same metric through the public API from pinned HOST audio: H2D of lr / hr audio, the step, D2H of the
               four losses, every step, stream-synchronised
  roofline     the kernel with the largest share of the step, timed with CUDA events around every C-ABI launch of an
               eager pass inside bench.py
  strong       the same step at the FIXED global batch 32 (SURVEY.md 8d cfg4 "report both"): 32 / N segments per GPU
  mdct         the transform half at HBM-roofline scale (8192 clips x 8192 samples per GPU, and the README shape
               [64, 32512]): GSamp/s and achieved GB/s of the raw MDCT4 / IMDCT4 kernels and of the fused
               (compress + abs-norm) ones vs MEASURED_PEAKS.json, for the flavour that passes the 2-ulp round-trip bar
               ("mixed": fp64 butterflies on fp32 tensors; checked in-line) and for the all-fp32 one
  longform     BASELINE configs[4]: generate_audio.py on a 60 s clip, its segments sharded over the N ranks
  configs      cfg2 / cfg3 generator inference steps and the train.sh recipe train step (ngf 56, 128 frames, batch 20)
  gpu_baseline the reference's own code (baseline/_ref, torch eager, cuDNN / cuFFT) on the SAME GPU: the kernel to beat
  cpu_baseline the reference's own code on this box's host cores (kind "reference"; "port" = oracle/train_oracle.py
               when baseline/_ref is absent)
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 48000
SEG = 7936             # samples per segment (32 frames)
BATCH = 4
N_FFT, HOP, NBINS = 512, 256, 256
GAIN, SRC, RNG = 1000.0, (-5.0, 5.0), (-1.0, 1.0)
NET_KW = dict(netG="local", n_down=3, n_blocks_global=9, n_blocks_local=3, n_attn=2, heads=4, dim_head=64, num_D=3, n_layers_D=3,
              fit_residual=True)
METRIC, UNIT = "train-step audio-sec/sec (cfg4 per-GPU slice: LocalEnhancer+2 attn, num_D 3, feat-match, 2x Adam, batch 4 x 7936 samples per GPU, fp32)", "audio-s/s"
OPT_ARGS = ["--name", "bench", "--lr_sampling_rate", "12000", "--sr_sampling_rate", "48000", "--arcsinh_transform", "--abs_spectro",
            "--arcsinh_gain", "1000", "--center", "--norm_range", "-1", "1", "--abs_norm", "--src_range", "-5", "5", "--netG", "local",
            "--ngf", "32", "--n_downsample_global", "3", "--n_blocks_global", "9", "--n_blocks_attn_g", "2", "--heads_g", "4",
            "--dim_head_g", "64", "--n_blocks_local", "3", "--num_D", "3", "--n_layers_D", "3", "--ndf", "64", "--lambda_feat", "10",
            "--segment_length", str(SEG), "--bins", "32", "--fit_residual", "--lr", "0.0002", "--beta1", "0.5"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s)"


def workload_config(world=1):
    return {"workload": "cfg4 (BASELINE configs[3]) per-GPU slice: full GAN train step, LocalEnhancer ngf 32 / 3 down / 9+3 blocks / 2 attention "
                        "layers (4 heads x 64) + MultiscaleDiscriminator num_D 3 + LSGAN + feature matching (lambda 10) + 2x Adam(2e-4, 0.5), "
                        "--fit_residual, 12->48 kHz, fp32, 4 segments x 7936 samples (32 frames x 256 bins) per GPU",
            "batch_per_gpu": BATCH, "global_batch": BATCH * world, "samples_per_segment": SEG, "n_fft": N_FFT, "hop": HOP,
            "parallelism": f"dp{world}: batch-sharded, one all-reduce of the flat 54.7 M-element gradient bucket per step",
            "l2": "per step the kernels stream 219 MB of weights, 219 MB of gradients and 438 MB of Adam state (larger than the 126 MB L2) "
                  "plus ~60 MB of activations; the `mdct` sub-benchmark uses 268 MB of audio"}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed regions run."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok, self.err = False, ""
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "note": "no NVML samples " + self.err}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------- CPU port
def make_lr_audio(batch, T, seed):
    import numpy as np
    import torch

    rng = np.random.default_rng(seed)
    x = 0.1 * rng.standard_normal((batch, T))
    X = np.fft.rfft(x, axis=-1)
    X[:, np.fft.rfftfreq(T, 1.0 / SR) > 6000.0] = 0
    return torch.from_numpy(np.fft.irfft(X, n=T, axis=-1).astype(np.float32))


def make_hr_audio(batch, T, seed):
    import numpy as np
    import torch

    rng = np.random.default_rng(seed + 17)
    return torch.from_numpy((0.1 * rng.standard_normal((batch, T))).astype(np.float32))


def cpu_port_run(seconds_budget=None, steps=None, warmup=1, init=None):
    """The reference's torch-CPU formulation of the train step: oracle/train_oracle.py (make_stepper).  `init` = (state_dict G,
    state_dict D) to start from (the GPU arm passes its own initial weights, so the first CPU step doubles as a cross-check of the
    four losses); the losses of the first step come back as "first_losses".  Used only where baseline/_ref is absent."""
    import torch

    from mdctgan_b200.models import networks
    from oracle import train_oracle as TO

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    if init is None:
        G = networks.define_G(2, 1, 32, "local", 3, 9, 1, 3, "instance", input_size=(32, 256), n_attn_g=2, heads_g=4, dim_head_g=64)  # parameter holders
        D = networks.define_D(3, 64, 3, "instance", False, 3, True)
        init = (G.state_dict(), D.state_dict())
    step = TO.make_stepper(init[0], init[1], make_lr_audio(BATCH, SEG, 42), make_hr_audio(BATCH, SEG, 42), **NET_KW)
    return _time_cpu_steps(step, seconds_budget, steps, warmup, "port",
                           "torch-CPU {t} threads, fp32 networks + complex128 transform, torch autograd + torch.optim.Adam "
                           "(oracle/train_oracle.py: the reference's formulation restated; baseline/_ref absent)")


def cpu_reference_run(seconds_budget=None, steps=None, warmup=1, init=None):
    """The UNMODIFIED reference (baseline/_ref: create_model(TrainOptions().parse()) + the train.py:160-202 loop) on the host
    cores, all threads.  `init`: the GPU arm's initial state_dicts, loaded into the reference's own modules (same keys), so the
    first step doubles as a cross-check of the four losses against the real reference."""
    import torch

    from baseline import ref_runner as R

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, model = R.make_stepper(OPT_ARGS, make_lr_audio(BATCH, SEG, 42), make_hr_audio(BATCH, SEG, 42), device="cpu", seed=1234)
    if init is not None:
        model.netG.load_state_dict(init[0])
        model.netD.load_state_dict(init[1])
    return _time_cpu_steps(step, seconds_budget, steps, warmup, "reference",
                           "torch-CPU {t} threads: the reference's own create_model / Pix2PixHDModel._forward / torch autograd / "
                           "torch.optim.Adam (baseline/_ref, unmodified), fp32 networks + complex128 transform")


def _time_cpu_steps(step, seconds_budget, steps, warmup, kind, how):
    import torch

    first_losses = None
    for _ in range(max(warmup, 1)):
        out = step()
        first_losses = first_losses if first_losses is not None else out
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if steps is not None and len(times) >= steps:
            break
        if seconds_budget is not None and (time.perf_counter() - t_start) >= seconds_budget:
            break
    total = sum(times)
    return {"value": BATCH * SEG / SR * len(times) / total, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
            "sample": f"the same train step (batch {BATCH} x {SEG} samples), {len(times)} steps, " + how.format(t=torch.get_num_threads()),
            "ms_per_step": 1e3 * total / len(times), "first_losses": first_losses, "steps": len(times)}


def cpu_baseline_run(**kw):
    from baseline import ref_runner as R

    return cpu_reference_run(**kw) if R.available() else cpu_port_run(**kw)


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the step on this box's host cores (rank 0 only), same
    warm-up count as the GPU arm, every step a full train-step on the arm's own batch; bounded to ~150 s of timed work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    warm = max(args.warmup, 3)
    r = cpu_baseline_run(steps=args.steps, seconds_budget=150.0, warmup=warm)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 networks / f64 transform (reference dtypes)", "data": "synthetic", "config": workload_config(1),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- transform sub-benchmark
def bench_mdct(dev, steps, hbm_peak):
    """The transform half at roofline scale, per GPU: 8192 clips x 8192 samples (268 MB of audio, > L2) and the README shape
    [64, 32512] (README.md:102-109).  For each arithmetic flavour: the raw MDCT4 / IMDCT4 kernels (fp32 coefficients) and the
    fused ones (compress + abs-norm, 2-channel forward).  The 2-ulp round-trip bar of north_star is checked here on the raw pair
    of the quoted flavour, over every clip."""
    import torch

    import mdctgan_b200
    from mdctgan_b200.models.mdct import IMDCT4, MDCT4
    from mdctgan_b200.models.pix2pixHD_model import Audio2MDCT, default_audio_opt
    from mdctgan_b200.util.util import kbdwin

    EPS = 2.0 ** -23
    st = torch.cuda.current_stream(dev)
    w = kbdwin(N_FFT)

    def timed(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            fn()
        e1.record(st)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps

    def entry(kernel, ms, nbytes, n):
        return {"kernel": kernel, "avg_launch_ms": ms, "algorithmic_bytes_per_launch": nbytes, "achieved_gbs": nbytes / (ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / hbm_peak, "gsamp_per_s": n / (ms * 1e-3) / 1e9}

    out = {"quoted_flavour": "mixed", "hbm_peak_gbs": hbm_peak, "shapes": {}}
    n0 = mdctgan_b200.launch_count()
    for (B, T) in ((8192, 8192), (64, 32512)):
        F = T // HOP + 1
        torch.manual_seed(7)
        x = 0.1 * torch.randn(B, T, device=dev)
        spec2 = torch.empty(B, 2, F, NBINS, device=dev)
        n = B * T
        shape = {}
        for prec in ("mixed", "fp32"):
            fwd, inv = MDCT4(N_FFT, HOP, N_FFT, w, device=dev, precision=prec), IMDCT4(N_FFT, HOP, N_FFT, w, device=dev, precision=prec)
            a2m = Audio2MDCT(default_audio_opt(arcsinh_gain=GAIN, src_range=SRC, norm_range=RNG, gpu_ids=[dev.index]), device=dev, precision=prec)
            spec = fwd(x)[0]
            y = inv(spec)[0].reshape(B, -1)
            err = ((y.double() - x[:, :y.shape[1]].double()).abs().amax(dim=1) / x.abs().amax(dim=1).double()).max().item() / EPS
            if prec == "mixed":
                assert err <= 2.0, f"MDCT4 -> IMDCT4 round trip {err} eps*peak breaks the 2-ulp bar"
            core = "double" if prec == "mixed" else "float"
            t_f, t_i = timed(lambda: fwd(x)), timed(lambda: inv(spec))
            t_ff = timed(lambda: a2m.to_spectro(x, channels=2, out=spec2))
            s1 = spec2[:, 0].contiguous()
            t_fi = timed(lambda: a2m.to_audio(s1))
            ya = a2m.to_audio(s1).reshape(B, -1)
            rt = ((ya[:64].double() - x[:64, :ya.shape[1]].double()).norm() / x[:64].double().norm()).item()
            assert rt < 1e-3, rt
            shape[prec] = {
                "round_trip_max_abs_err_in_eps_peak": err, "passes_2ulp": bool(err <= 2.0),
                "raw_forward": entry(f"mdct4_fwd_kernel<{core},0,float>", t_f, 4 * n + 4 * B * F * NBINS, n),
                "raw_inverse": entry(f"imdct4_inv_kernel<{core},float,float,0>", t_i, 4 * B * F * NBINS + 4 * n, n),
                "raw_round_trip_gsamp_per_s": n / ((t_f + t_i) * 1e-3) / 1e9,
                "fused_forward_2ch": entry(f"mdct4_fwd_kernel<{core},1,float>", t_ff, 4 * n + 8 * B * F * NBINS, n),
                "fused_inverse": entry(f"imdct4_inv_kernel<{core},float,float,1>", t_fi, 4 * B * F * NBINS + 4 * n, n),
                "fused_round_trip_rel_l2": rt}
            del spec, y, s1, ya
        out["shapes"][f"{B}x{T}"] = shape
        del x, spec2
    out["launches"] = mdctgan_b200.launch_count() - n0
    q = out["shapes"]["8192x8192"]["mixed"]
    out.update({"workload": "8192 clips x 8192 samples per GPU; raw = MDCT4 / IMDCT4 on fp32 tensors, fused = + arcsinh/abs-norm (2-channel forward)",
                "gsamp_per_s_round_trip": q["raw_round_trip_gsamp_per_s"], "forward": q["raw_forward"], "inverse": q["raw_inverse"]})
    return out


# ---------------------------------------------------------------------------------------------- long-form sub-benchmark
def bench_longform(dev, rank=0, world=1, reps=3):
    """BASELINE configs[4]: generate_audio.py on a 60 s clip, 16 -> 48 kHz, 32512-sample segments (128 frames), gen_overlap 256,
    LocalEnhancer + 2 attention layers: pinned host clip -> H2D -> on-device segmentation -> batched inference (CUDA graphs) ->
    on-device overlap-add -> D2H.  With N ranks every rank takes a contiguous run of the segments (parallel.shard_range, no
    collective; the shard outputs overlap by gen_overlap samples and are summed by whoever assembles them).  Returns this rank's
    seconds per clip; the caller takes the max over ranks."""
    import torch

    from mdctgan_b200.longform import LongFormGenerator
    from mdctgan_b200.models.models import create_model
    from mdctgan_b200.options.train_options import TrainOptions

    seg, ov, secs = 32512, 256, 60
    args = [a for a in OPT_ARGS]
    for k, v in (("--segment_length", str(seg)), ("--bins", "128"), ("--lr_sampling_rate", "16000")):
        args[args.index(k) + 1] = v
    opt = TrainOptions().parse(save=False, args=args + ["--gpu_ids", str(dev.index), "--gen_overlap", str(ov)])
    opt.checkpoints_dir = "/tmp/mdctgan_bench"
    torch.manual_seed(99)
    model = create_model(opt)
    model.eval()
    clip = make_lr_audio(1, secs * SR, 5).pin_memory()
    gen = LongFormGenerator(model, batch_size=16)
    out, _ = gen.generate_shard(clip.to(dev, non_blocking=True), seg, ov, rank, world)          # captures the batch shapes
    out_h = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        out, _ = gen.generate_shard(clip.to(dev, non_blocking=True), seg, ov, rank, world)
        out_h.copy_(out, non_blocking=True)
        torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / reps
    n_seg = gen.last_segments
    dt0 = None
    try:        # the reference script's own value, --gen_overlap 0 (generate_audio.sh:4-15)
        gen.generate_shard(clip.to(dev, non_blocking=True), seg, 0, rank, world)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            o0, _ = gen.generate_shard(clip.to(dev, non_blocking=True), seg, 0, rank, world)
            o0.cpu()
        dt0 = (time.perf_counter() - t0) / reps
    except Exception as e:  # noqa: BLE001
        dt0 = repr(e)[:200]
    del gen, model
    return {"workload": f"{secs} s clip @48 kHz -> {n_seg[1]} x {seg}-sample segments, gen_overlap {ov}, LocalEnhancer+2 attn at 128 frames, "
                        f"batches of 16, fp32; {world} rank(s), contiguous segment runs, no collective", "seconds_per_clip": dt,
            "segments_this_rank": n_seg[0], "clip_seconds": secs, "h2d_bytes": clip.numel() * 4, "d2h_bytes": out.numel() * out.element_size(),
            "seconds_per_clip_gen_overlap_0": dt0}


# ---------------------------------------------------------------------------------------------- other BASELINE configs
def _build_model(extra, dev, seed=1234):
    import torch

    from mdctgan_b200.models.models import create_model
    from mdctgan_b200.options.train_options import TrainOptions

    args = [a for a in OPT_ARGS]
    flags = []
    for k, v in extra:
        if v is None:
            flags.append(k)
        elif k in args:
            args[args.index(k) + 1] = v
        else:
            flags += [k, v]
    opt = TrainOptions().parse(save=False, args=args + flags + ["--gpu_ids", str(dev.index)])
    opt.checkpoints_dir = "/tmp/mdctgan_bench"
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    return create_model(opt)


def _graph_time(replay, steps, dev):
    import torch

    st = torch.cuda.current_stream(dev)
    for _ in range(3):
        replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        replay()
    e1.record(st)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def bench_configs(dev, steps, bf16_peak):
    """The BASELINE configs that are not the headline: cfg2 (GlobalGenerator inference, batch 4), cfg3 (LocalEnhancer + 2 attention
    layers inference, batch 8; the bf16 engine is not built: run by the fp32-class 3xTF32 engine and by the single-pass TF32 one)
    and the train step of the reference's shipped recipe train.sh (ngf 56, resconv / interpolate, 3 attention layers of 6 x 128,
    128 frames, batch 20).  Each: a CUDA-graph replay of the step, CUDA events; tensor-roofline fraction = algorithmic conv+matmul
    FLOPs of the step (probed with torch's flop counter) / time / measured dense bf16 peak."""
    import torch

    from mdctgan_b200 import nn_ops
    from mdctgan_b200.runtime import GraphedInference, GraphedTrainStep

    res = {}

    def infer_line(name, extra, batch, T, gflop_per_sample, engines):
        model = _build_model(extra, dev)
        model.eval()
        lr = make_lr_audio(batch, T, 3).to(dev)
        for eng in engines:
            prev = nn_ops.CONV_ENGINE
            nn_ops.CONV_ENGINE = eng
            try:
                gi = GraphedInference(model, batch, T, warmup=2)
                gi.static_in.copy_(lr)
                ms = _graph_time(gi.replay, steps, dev)
            finally:
                nn_ops.CONV_ENGINE = prev
            fl = gflop_per_sample * 1e9 * batch
            res[f"{name}[{eng}]"] = {"step": "model.inference: fused MDCT -> generator -> fused IMDCT, CUDA-graph replay", "batch": batch, "samples": T,
                                     "engine": "3xTF32 (fp32-class)" if eng == "umma" else "single-pass TF32", "ms_per_step": ms,
                                     "audio_sec_per_sec": batch * T / SR / (ms * 1e-3), "algorithmic_gflop_per_step": fl / 1e9,
                                     "tflops": fl / (ms * 1e-3) / 1e12, "frac_of_bf16_peak": fl / (ms * 1e-3) / 1e12 / bf16_peak}
            del gi
        del model

    infer_line("cfg2_global_generator_fwd_b4", [("--netG", "global"), ("--n_blocks_attn_g", "0")], 4, SEG, 3.25, ("umma",))
    infer_line("cfg3_local_enhancer_attn_fwd_b8", [], 8, SEG, 4.37, ("umma", "tf32"))
    # train.sh: the reference's shipped recipe (train.sh:3-17)
    try:
        extra = [("--ngf", "56"), ("--n_blocks_global", "4"), ("--n_blocks_attn_g", "3"), ("--heads_g", "6"), ("--dim_head_g", "128"),
                 ("--segment_length", "32512"), ("--bins", "128"), ("--lr_sampling_rate", "16000"), ("--upsample_type", "interpolate"),
                 ("--downsample_type", "resconv"), ("--lr", "0.00015")]
        model = _build_model(extra, dev)
        model.train()
        b, T = 20, 32512
        gts = GraphedTrainStep(model, b, T, 1, None, warmup=2)
        gts.lr_in.copy_(make_lr_audio(b, T, 11).to(dev))
        gts.hr_in.copy_(make_hr_audio(b, T, 11).to(dev))
        ms = _graph_time(gts.replay, max(3, min(steps, 10)), dev)
        fl = (3 * 103.5 + 9 * 5.20) * 1e9 * b          # 3 G_fwd + 9 D_fwd per sample, F = 128
        res["train_sh_recipe_step_b20"] = {"step": "train.sh recipe: ngf 56, resconv/interpolate, 3 attn x (6 heads x 128), 128 frames x 256 bins, batch 20, "
                                                   "num_D 3, full GAN train step, CUDA-graph replay", "batch": b, "samples": T, "ms_per_step": ms,
                                           "audio_sec_per_sec": b * T / SR / (ms * 1e-3), "algorithmic_gflop_per_step": fl / 1e9,
                                           "tflops": fl / (ms * 1e-3) / 1e12, "frac_of_bf16_peak": fl / (ms * 1e-3) / 1e12 / bf16_peak,
                                           "losses": [float(v) for v in gts.losses.cpu()]}
        del gts, model
    except Exception as e:  # noqa: BLE001
        res["train_sh_recipe_step_b20"] = {"error": repr(e)[:300]}
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------- the GPU kernel to beat
def bench_gpu_baseline(dev, sd0, lr, hr, steps):
    """The same train step by the reference's OWN code on this GPU ("the bar on B200 is torch-eager"):
      reference_eager_*   baseline/_ref create_model(opt) with --gpu_ids 0 + the verbatim train.py:160-202 loop, torch eager,
                          cudnn.benchmark = True like train.py:26, weights = the GPU arm's initial weights; with torch's default
                          TF32 policy (cuDNN convolutions TF32, matmul fp32) and with TF32 off (true fp32)
      port_cuda_graph_*   oracle/train_oracle.py (the same torch ops, functional) captured whole in ONE torch CUDA graph with
                          capturable Adam: the reference's kernels without its Python / launch overhead and without its per-step
                          .cpu() snapshots -- the strongest torch baseline available here
    CUDA events around `steps` iterations after 5 warm-up iterations (cuDNN autotuning included in the warm-up)."""
    import torch

    from baseline import ref_runner as R
    from oracle import train_oracle as TO

    out = {}
    st = torch.cuda.current_stream(dev)
    audio_s = BATCH * SEG / SR

    def time_steps(step, n, warm=5):
        for _ in range(warm):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(st)
        for _ in range(n):
            step()
        e1.record(st)
        torch.cuda.synchronize(dev)
        wall = (time.perf_counter() - t0) / n * 1e3
        return max(e0.elapsed_time(e1) / n, wall)

    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.benchmark = True
        for name, tf32 in (("tf32_default", True), ("fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            if R.available():
                try:
                    step, model = R.make_stepper(OPT_ARGS, lr.cpu(), hr.cpu(), device=dev, seed=1234)
                    model.netG.load_state_dict(sd0[0])
                    model.netD.load_state_dict(sd0[1])
                    first = step()
                    ms = time_steps(lambda: step(False), steps)
                    out[f"reference_eager_{name}"] = {"ms_per_step": ms, "audio_sec_per_sec": audio_s / (ms * 1e-3), "first_losses": first,
                                                      "what": "baseline/_ref (unmodified reference) on cuda, torch eager, train.py:160-202 verbatim"}
                    del step, model
                except Exception as e:  # noqa: BLE001
                    out[f"reference_eager_{name}"] = {"error": repr(e)[:300]}
            try:
                replay = TO.make_stepper(sd0[0], sd0[1], lr, hr, device=dev, cuda_graph=True, **NET_KW)
                ms = time_steps(replay, steps, warm=3)
                out[f"port_cuda_graph_{name}"] = {"ms_per_step": ms, "audio_sec_per_sec": audio_s / (ms * 1e-3),
                                                  "what": "oracle/train_oracle.py on cuda, whole step in one torch CUDA graph (capturable Adam)"}
                del replay
            except Exception as e:  # noqa: BLE001
                out[f"port_cuda_graph_{name}"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return out


def bench_reference_transform_and_longform(dev):
    """Baselines for BASELINE configs[0] / [4] by the reference's OWN code (baseline/_ref, unmodified):
      mdct      models.mdct.MDCT4 / IMDCT4 (complex128 512-point FFT formulation) on `cuda` -- torch / cuFFT, the kernel to beat for
                the transform half -- at the README shape [64, 32512] and at [1024, 8192], and on the host cores at the README shape
                (README.md:102-109 publishes 9.61 ms for it on an RTX 3070 laptop GPU)
      longform  generate_audio.py's loop (seg_pad_audio -> model.inference per batch of 16 segments -> .cpu() -> fold overlap-add,
                generate_audio.py:24-53) on a 60 s clip, on `cuda`, and on the host cores on a bounded sample (one batch of 16 segments,
                scaled to the 89 segments of the clip)"""
    import torch

    from baseline import ref_runner as R
    from oracle import longform_oracle as LO

    if not R.available():
        return {"unavailable": "baseline/_ref is missing"}
    R.import_reference()
    from models.mdct import IMDCT4 as RefIMDCT4, MDCT4 as RefMDCT4
    from util.util import kbdwin as ref_kbdwin

    out = {"mdct": {}, "longform": {}}

    def wall(fn, reps, sync):
        fn()
        if sync:
            torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        if sync:
            torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / reps * 1e3

    for where, shapes in (("cuda", ((64, 32512), (1024, 8192))), ("cpu", ((64, 32512),))):
        d = dev if where == "cuda" else torch.device("cpu")
        w = ref_kbdwin(N_FFT).to(d)
        fwd = RefMDCT4(n_fft=N_FFT, hop_length=HOP, win_length=N_FFT, window=w, device=d)
        inv = RefIMDCT4(n_fft=N_FFT, hop_length=HOP, win_length=N_FFT, window=w, device=d)
        for (B, T) in shapes:
            torch.manual_seed(7)
            x = (0.1 * torch.randn(B, T)).to(d)
            with torch.no_grad():
                spec = fwd(x)[0]
                t_f = wall(lambda: fwd(x), 5 if where == "cuda" else 3, where == "cuda")
                t_i = wall(lambda: inv(spec), 5 if where == "cuda" else 3, where == "cuda")
            n = B * T
            out["mdct"][f"reference_{where}_{B}x{T}"] = {"forward_ms": t_f, "inverse_ms": t_i, "forward_gsamp_per_s": n / t_f / 1e6,
                                                         "inverse_gsamp_per_s": n / t_i / 1e6,
                                                         "what": f"models.mdct.MDCT4 / IMDCT4 of the unmodified reference on {where}"
                                                                 + (f" ({os.cpu_count()} threads)" if where == "cpu" else " (torch eager, cuFFT)")}
            del x, spec
    # ---- long-form
    seg, ov, secs = 32512, 256, 60
    args = [a for a in OPT_ARGS]
    for k, v in (("--segment_length", str(seg)), ("--bins", "128"), ("--lr_sampling_rate", "16000")):
        args[args.index(k) + 1] = v
    clip = make_lr_audio(1, secs * SR, 5)
    segs = LO.seg_pad_audio(clip.clone(), seg, ov)                       # AudioTestDataset.seg_pad_audio (pinned to the reference by tests)
    n_seg = segs.shape[0]
    for where in ("cuda", "cpu"):
        d = dev if where == "cuda" else torch.device("cpu")
        _, model = R.make_stepper(args, segs[:4], segs[:4], device=d, seed=99)
        model.eval()

        def generate(n_batches):
            outs = []
            with torch.no_grad():
                for i in range(0, min(n_seg, 16 * n_batches), 16):
                    sr_audio = model.inference(segs[i:i + 16].to(d))[1]
                    outs.append(sr_audio.cpu())                        # generate_audio.py:36 copies every batch to the host
            return torch.cat(outs, dim=0)

        if where == "cuda":
            generate(1)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            au = generate(10 ** 6)
            LO.overlap_add(au.double().reshape(-1, 1, 1, au.shape[-1]), seg, ov)
            dt = time.perf_counter() - t0
            out["longform"]["reference_cuda"] = {"seconds_per_clip": dt, "real_time_factor": secs / dt, "segments": n_seg,
                                                 "what": "generate_audio.py:24-53 with the unmodified reference model on cuda (torch eager), 60 s clip"}
        else:
            t0 = time.perf_counter()
            generate(1)
            dt16 = time.perf_counter() - t0
            dt = dt16 * n_seg / 16.0
            out["longform"]["reference_cpu"] = {"seconds_per_clip": dt, "real_time_factor": secs / dt, "cores": os.cpu_count(),
                                                "sample": f"one batch of 16 of the {n_seg} segments ({dt16:.2f} s), scaled to the clip",
                                                "what": "the same loop on the host cores"}
        del model
        torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------- per-launch profile
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of the same workload (profiles/, gpurun s5f / s5m)
NCU_TRAFFIC = {
    "wgrad[fp32] k3 s1 512->512 B4 @2x16": (10.08e6, "profiles/r01_train_ncu_full_wgrad.txt: 10.06-10.10 MB read, 0 B written back within the "
                                                     "kernel (the 9.4 MB of gradient atomics resolve in L2); algorithmic minimum 9.4 MB (dW) + 0.5 MB"),
    "conv[tcgen05] k3 s1 512->512 B4 @2x16": (19.2e6, "profiles/r01_train_ncu_full_wgrad_umma.txt: 19.2 MB read = the TF32 hi|lo weight image (2 x 9.4 MB)"),
    "conv[tcgen05] k3 s1 512->512 B4 @4x18": (19.2e6, "profiles/r01_train_ncu_full_wgrad_umma.txt: 19.2 MB read = the TF32 hi|lo weight image (2 x 9.4 MB)"),
}

KERNEL_ENTRIES = ["mdctgan_conv2d_nhwc", "mdctgan_conv2d_umma", "mdctgan_conv2d_wgrad", "mdctgan_norm_finalize", "mdctgan_norm_apply",
                  "mdctgan_norm_act_bwd", "mdctgan_norm_act_bwd_folded", "mdctgan_act_bwd", "mdctgan_add", "mdctgan_reflect_pad_bwd", "mdctgan_reflect_pad_bwd_add", "mdctgan_avgpool3s2_nhwc",
                  "mdctgan_avgpool3s2_bwd", "mdctgan_attention_abs_pos", "mdctgan_attention_abs_pos_bwd", "mdctgan_mse_const_fwd",
                  "mdctgan_mse_const_bwd", "mdctgan_l1_pair_fwd", "mdctgan_l1_pair_bwd", "mdctgan_multi_loss_fwd", "mdctgan_multi_loss_bwd", "mdctgan_f64_to_f32", "mdctgan_disc_input_fwd",
                  "mdctgan_disc_input_bwd", "mdctgan_adam_flat", "mdctgan_counter_inc", "mdctgan_conv2d_umma_pack_weight",
                  "mdctgan_pack_weights_multi", "mdctgan_pack_weights_tiled",
                  "mdctgan_nchw_to_nhwc", "mdctgan_nhwc_to_nchw", "mdctgan_residual_scale_add", "mdctgan_audio2mdct_forward",
                  "mdctgan_mdct2audio_inverse"]


class LaunchProfiler:
    """CUDA events around every C-ABI launch of an eager pass (bench-only instrumentation)."""

    def __init__(self, dev):
        import torch

        from mdctgan_b200 import nn_ops

        self.torch, self.dev = torch, dev
        self.L = nn_ops._L()
        self.orig, self.records = {}, []

    def __enter__(self):
        torch = self.torch
        from mdctgan_b200 import nn_ops

        self._side, self._par = nn_ops.SIDE_STREAM_WGRAD, nn_ops.PARALLEL_BRANCHES
        nn_ops.SIDE_STREAM_WGRAD = nn_ops.PARALLEL_BRANCHES = False        # per-kernel times: everything serialised on one stream
        for n in KERNEL_ENTRIES:
            if not hasattr(self.L, n):
                continue
            f = getattr(self.L, n)
            self.orig[n] = f

            def wrap(*a, _f=f, _n=n):
                st = torch.cuda.current_stream(self.dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                rc = _f(*a)
                e1.record(st)
                tag, flops, nbytes = _n.replace("mdctgan_", ""), 0.0, 0.0
                if _n in ("mdctgan_conv2d_nhwc", "mdctgan_conv2d_umma"):
                    B, H, W, Cin, Ho, Wo, Cout, kh, stride, transposed = a[1], a[2], a[3], a[4], a[8], a[9], a[10], a[11], a[13], a[16]
                    eng = "tcgen05" if _n.endswith("umma") else "fp32"
                    tag = f"conv[{eng}] k{kh} s{stride}{'T' if transposed else ''} {Cin}->{Cout} B{B} @{Ho}x{Wo}"
                    flops = 2.0 * B * Ho * Wo * Cout * kh * kh * Cin / (stride * stride if transposed else 1)
                elif _n == "mdctgan_conv2d_wgrad":
                    B, Cin, Ho, Wo, Cout, kh, stride, transposed = a[1], a[4], a[6], a[7], a[8], a[9], a[11], a[14]
                    tag = f"wgrad[{'tcgen05' if a[-2] else 'fp32'}] k{kh} s{stride}{'T' if transposed else ''} {Cin}->{Cout} B{B} @{Ho}x{Wo}"
                    flops = 2.0 * B * Ho * Wo * Cout * kh * kh * Cin / (stride * stride if transposed else 1)
                elif _n == "mdctgan_adam_flat":
                    nbytes = 28.0 * a[4]          # p, g, m, v read (16 B) + p, m, v written (12 B) per parameter
                elif _n == "mdctgan_norm_act_bwd":
                    nbytes = 4.0 * a[13] * a[14] * a[15] * 5   # x, dv read twice + dx written
                elif _n == "mdctgan_norm_act_bwd_folded":
                    nbytes = 4.0 * a[13] * a[14] * a[15] * a[16] * 5
                elif _n == "mdctgan_norm_apply":
                    nbytes = 4.0 * a[15] * a[16] * a[17] * (3 if a[6] else 2)
                self.records.append((tag, _n, e0, e1, flops, nbytes))
                return rc

            setattr(self.L, n, wrap)
        return self

    def __exit__(self, *exc):
        from mdctgan_b200 import nn_ops

        nn_ops.SIDE_STREAM_WGRAD, nn_ops.PARALLEL_BRANCHES = self._side, self._par
        for n, f in self.orig.items():
            setattr(self.L, n, f)

    def table(self):
        """tag -> [launches, total ms, flops per launch, bytes per launch, entry point]"""
        self.torch.cuda.synchronize(self.dev)
        agg = {}
        for tag, entry, e0, e1, flops, nbytes in self.records:
            a = agg.setdefault(tag, [0, 0.0, flops, nbytes, entry])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
        return agg


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import mdctgan_b200
    from mdctgan_b200.models.models import create_model
    from mdctgan_b200.options.train_options import TrainOptions
    from mdctgan_b200.runtime import GraphedTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, bf16_peak, peak_src = peaks()

    opt = TrainOptions().parse(save=False, args=OPT_ARGS + ["--gpu_ids", str(local)])
    opt.checkpoints_dir = "/tmp/mdctgan_bench"
    torch.manual_seed(1234)                      # identical initial weights on every rank
    torch.cuda.manual_seed(1234)
    model = create_model(opt)
    model.train()
    lr = make_lr_audio(BATCH, SEG, 42 + rank).to(dev)      # a different batch per rank
    hr = make_hr_audio(BATCH, SEG, 42 + rank).to(dev)
    warm = max(args.warmup, 3)

    from mdctgan_b200.parallel import GradExchange, broadcast_flat

    all_reduce = GradExchange() if world > 1 else None      # ONE sum all-reduce of the flat [grad_G | grad_D] bucket per step
    broadcast_flat([model.bucket_G.flat, model.bucket_D.flat])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the initial weights, kept for the cpu_baseline leg (its first step on the same batch doubles as a cross-check of the losses)
    sd0 = None
    if rank == 0:
        sd0 = ({k: v.detach().cpu().clone() for k, v in model.netG.state_dict().items()},
               {k: v.detach().cpu().clone() for k, v in model.netD.state_dict().items()})
    # ---- eager steps: launches per step, per-kernel table (CUDA events around every C-ABI launch)
    first = model.train_step(lr, hr, world, all_reduce).cpu().tolist()
    assert all(v == v and abs(v) < 1e6 for v in first), f"train-step losses {first}"
    for _ in range(2):
        model.train_step(lr, hr, world, all_reduce)
    n0 = mdctgan_b200.launch_count()
    model.train_step(lr, hr, world, all_reduce)
    launches_per_step = mdctgan_b200.launch_count() - n0
    with LaunchProfiler(dev) as prof:
        for _ in range(3):
            model.train_step(lr, hr, world, all_reduce)
    table = prof.table()
    tot_ms = sum(v[1] for v in table.values())

    # ---- timed region: the step as a CUDA graph, K replays
    gts = GraphedTrainStep(model, BATCH, SEG, world, all_reduce, warmup=2)
    gts.lr_in.copy_(lr)
    gts.hr_in.copy_(hr)
    gts.recapture()
    st = torch.cuda.current_stream(dev)
    for _ in range(warm):
        gts.replay()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        gts.replay()
    e1.record(st)
    barrier()
    step_ms = e0.elapsed_time(e1) / args.steps
    last_losses = gts.losses.cpu().tolist()

    # ---- e2e: pinned host audio -> H2D -> graph replay -> D2H of the four losses, every step
    lr_h, hr_h = lr.cpu().pin_memory(), hr.cpu().pin_memory()
    loss_h = torch.empty(4, dtype=torch.float32).pin_memory()
    for _ in range(3):
        loss_h.copy_(gts(lr_h, hr_h), non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss_h.copy_(gts(lr_h, hr_h), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()     # the training loop reads the losses of a step before the next one
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps

    # ---- the reference-shaped API (train.py:160-202 verbatim: _forward, two backward(), two optimizer.step()), eager, for the record
    def api_step():
        ls_, _ = model._forward(lr, hr)
        d = dict(zip(model.loss_names, ls_))
        loss_D = (d["D_fake"] + d["D_real"]) * 0.5
        loss_G = d["G_GAN"] + d["G_GAN_Feat"]
        model.optimizer_G.zero_grad()
        loss_G.backward()
        model.optimizer_G.step()
        model.optimizer_D.zero_grad()
        loss_D.backward()
        model.optimizer_D.step()

    for _ in range(7):            # runtime.GraphedAPI: 3 eager iterations, then the forward / sweep segments are captured and replayed
        api_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    n_api = max(5, min(args.steps, 30))
    for _ in range(n_api):
        api_step()
    torch.cuda.synchronize(dev)
    api_ms = 1e3 * (time.perf_counter() - t0) / n_api
    model._graph_api, api_saved = None, model._graph_api      # the same sequence with every launch issued eagerly, for the record
    for _ in range(2):
        api_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(5):
        api_step()
    torch.cuda.synchronize(dev)
    api_eager_ms = 1e3 * (time.perf_counter() - t0) / 5
    model._graph_api = api_saved
    clocks = sampler.stop()

    # ---- strong scaling (SURVEY.md 8d cfg4 "report both"): the FIXED global batch 32, 32 / N segments per GPU
    strong = None
    G_BATCH = 32
    if G_BATCH % world == 0 and not args.no_strong:
        bs = G_BATCH // world
        if bs == BATCH:
            strong = {"global_batch": G_BATCH, "batch_per_gpu": bs, "ms_per_step": step_ms, "note": "same shape as the weak leg at this N"}
        else:
            del gts
            torch.cuda.empty_cache()
            lr_s, hr_s = make_lr_audio(bs, SEG, 142 + rank).to(dev), make_hr_audio(bs, SEG, 142 + rank).to(dev)
            gs = GraphedTrainStep(model, bs, SEG, world, all_reduce, warmup=2)
            gs.lr_in.copy_(lr_s)
            gs.hr_in.copy_(hr_s)
            for _ in range(3):
                gs.replay()
            n_s = max(5, min(args.steps, 30))
            barrier()
            e0.record(st)
            for _ in range(n_s):
                gs.replay()
            e1.record(st)
            barrier()
            strong = {"global_batch": G_BATCH, "batch_per_gpu": bs, "ms_per_step": e0.elapsed_time(e1) / n_s, "steps": n_s}
            del gs
            torch.cuda.empty_cache()

    # ---- transform and long-form sub-benchmarks: every rank runs its shard (no collective), max over ranks below
    mdct = bench_mdct(dev, 30, hbm_peak)
    longform = None
    if not args.no_longform:
        try:
            longform = bench_longform(dev, rank, world)
        except Exception as e:  # noqa: BLE001  (a sub-benchmark must not take the headline line down)
            longform = {"error": repr(e)[:300]}
    lf_s = longform["seconds_per_clip"] if (longform and "seconds_per_clip" in longform) else 0.0
    sh = mdct["shapes"]["8192x8192"]["mixed"]
    vals = torch.tensor([step_ms, e2e_ms, strong["ms_per_step"] if strong else 0.0, lf_s, sh["raw_forward"]["avg_launch_ms"],
                         sh["raw_inverse"]["avg_launch_ms"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms, strong_ms, lf_s, mf_ms, mi_ms = vals.tolist()
    if world > 1:
        # No collective follows.  Leave without the NCCL / CUDA-graph teardown: destroying a communicator that a live captured
        # graph still references can block forever (seen on the 2-GPU run), and the driver waits for every rank to exit.
        torch.cuda.synchronize(dev)
        dist.barrier()
        if rank != 0:
            sys.stdout.flush()
            os._exit(0)

    if rank == 0:
        audio_s = world * BATCH * SEG / SR
        if strong:
            strong.update({"ms_per_step": strong_ms, "value": G_BATCH * SEG / SR / (strong_ms * 1e-3), "unit": UNIT, "scaling": "strong",
                           "n_gpus": world})
        if longform and "seconds_per_clip" in longform:
            longform.update({"seconds_per_clip": lf_s, "real_time_factor": longform["clip_seconds"] / lf_s,
                             "audio_sec_per_sec": longform["clip_seconds"] / lf_s, "n_gpus": world})
        nsamp = 8192 * 8192
        mdct["aggregate"] = {"n_gpus": world, "scaling": "weak (8192 x 8192 samples per GPU, no collective)",
                             "raw_forward_gsamp_per_s": world * nsamp / (mf_ms * 1e-3) / 1e9, "raw_inverse_gsamp_per_s": world * nsamp / (mi_ms * 1e-3) / 1e9,
                             "raw_round_trip_gsamp_per_s": world * nsamp / ((mf_ms + mi_ms) * 1e-3) / 1e9}
        configs = gpu_base = ref_tl = None
        if world == 1 and not args.no_extras:
            try:
                configs = bench_configs(dev, max(5, min(args.steps, 30)), bf16_peak)
            except Exception as e:  # noqa: BLE001
                configs = {"error": repr(e)[:300]}
            try:
                gpu_base = bench_gpu_baseline(dev, sd0, lr, hr, max(5, min(args.steps, 20)))
            except Exception as e:  # noqa: BLE001
                gpu_base = {"error": repr(e)[:300]}
            try:
                ref_tl = bench_reference_transform_and_longform(dev)
            except Exception as e:  # noqa: BLE001
                ref_tl = {"error": repr(e)[:300]}
        cpu = cpu_baseline_run(seconds_budget=args.cpu_seconds, warmup=1, init=sd0)
        # world 1: the GPU's first step and the CPU baseline's first step saw the same weights and the same batch
        err = max(abs(a - b) / abs(b) for a, b in zip(first, cpu["first_losses"])) if world == 1 else None
        if err is not None:
            assert err < 2e-3, f"train-step losses {first} vs the CPU baseline's {cpu['first_losses']}"
        # the dominant KERNEL = the entry point (one __global__ template) with the largest share of the step; its launches differ in
        # shape, so achieved = sum of algorithmic flops (bytes) over its launches / sum of their durations
        ent = {}
        for tag, v in table.items():
            e = ent.setdefault(v[4], [0, 0.0, 0.0, 0.0, None, 0.0])
            e[0] += v[0]; e[1] += v[1]; e[2] += v[2] * v[0]; e[3] += v[3] * v[0]
            if v[1] > e[5]:
                e[4], e[5] = tag, v[1]
        dom_entry, dom = max(ent.items(), key=lambda kv: kv[1][1])
        dom_tag, dom_ms = dom[4], dom[1] / dom[0]
        kname = {"mdctgan_conv2d_umma": "conv2d_umma_kernel (tcgen05 implicit-GEMM convolution: forward + input-gradient launches)",
                 "mdctgan_conv2d_wgrad": "conv_wgrad_umma_kernel / conv_wgrad_kernel (weight gradient)"}.get(dom_entry, dom_entry.replace("mdctgan_", ""))
        if dom[2] > 0:
            ach = dom[2] / (dom[1] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak, "traffic": None,
                    "algorithmic_flops_per_launch": dom[2] / dom[0], "largest_shape": dom_tag,
                    "note": "algorithmic flops = 2*M*N*K of every launch of the kernel in a step / their summed durations; peak = measured dense bf16 "
                            "cuBLAS burst (MEASURED_PEAKS.json); the engine computes 3xTF32 (3 MMAs per product at half the bf16 rate), so 1/6 of "
                            "this peak is its own ceiling"}
        else:
            ach = dom[3] / (dom[1] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                    "algorithmic_bytes_per_launch": dom[3] / dom[0], "largest_shape": dom_tag}
        if dom_tag in NCU_TRAFFIC:
            roof["traffic"], roof["traffic_source"] = NCU_TRAFFIC[dom_tag]
            roof["traffic_is_for"] = dom_tag
        roof.update({"peak_source": peak_src, "avg_launch_ms": dom_ms, "launches_per_step": dom[0] // 3, "share_of_step": dom[1] / tot_ms})
        by_entry = {}
        for tag, v in table.items():
            e = by_entry.setdefault(v[4].replace("mdctgan_", ""), [0, 0.0])
            e[0] += v[0] // 3
            e[1] += v[1] / 3
        kernel_table = {k: {"launches_per_step": v[0] // 3, "ms_per_step": round(v[1] / 3, 4), "share": round(v[1] / tot_ms, 4),
                            **({"tflops": round(v[2] / (v[1] / v[0] * 1e-3) / 1e12, 2)} if v[2] else {})}
                        for k, v in sorted(table.items(), key=lambda kv: -kv[1][1])[:60]}
        step_flops = 28.9e9 * BATCH       # 3 G_fwd + 9 D_fwd per sample at F = 32
        line = {
            "metric": METRIC, "value": audio_s / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "roofline": roof,
            "step_roofline": {"algorithmic_gflop_per_gpu_step": step_flops / 1e9, "tflops": step_flops / (step_ms * 1e-3) / 1e12,
                              "frac_of_bf16_peak": step_flops / (step_ms * 1e-3) / 1e12 / bf16_peak,
                              "hbm_floor_ms": 3.8e9 / (hbm_peak * 1e9) * 1e3, "note": "whole step: conv+matmul flops / step time; HBM floor = "
                              "weights + gradients + Adam state streamed once (DESIGN.md 5)"},
            "kernel_table": kernel_table,
            "entry_point_table": {k: {"launches_per_step": v[0], "ms_per_step": round(v[1], 4)} for k, v in
                                  sorted(by_entry.items(), key=lambda kv: -kv[1][1])},
            "eager_sum_of_kernels_ms": tot_ms / 3,
            "strong": strong,
            "mdct": mdct,
            "longform": longform,
            "configs": configs,
            "gpu_baseline": gpu_base,
            "reference_transform_longform": ref_tl,
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": audio_s / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * BATCH * SEG * 4, "d2h_bytes_per_step": 16,
                    "ms_per_step": e2e_ms, "api": "runtime.GraphedTrainStep(model)(pinned lr_audio, pinned hr_audio) -> 4 losses copied to "
                                                  "pinned host memory, stream-synchronised every step"},
            "reference_api": {"value": audio_s / world / (api_ms * 1e-3), "ms_per_step": api_ms, "n_gpus": 1, "ms_per_step_all_eager": api_eager_ms,
                              "api": "train.py:160-202 verbatim on rank 0: model._forward -> loss_G.backward() -> optimizer_G.step() -> "
                                     "loss_D.backward() -> optimizer_D.step(); after 3 eager iterations the forward and the two sweeps replay "
                                     "captured segments (runtime.GraphedAPI), zero_grad / Adam / weight images stay eager; no all-reduce; "
                                     "ms_per_step_all_eager = the same with MDCTGAN_GRAPH_API=0 (host-launch bound)"},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step, "clocks": clocks,
            "losses_first_step_rel_err_vs_cpu_baseline": err, "losses_last_step": last_losses,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-longform", dest="no_longform", action="store_true")
    ap.add_argument("--no-strong", dest="no_strong", action="store_true")
    ap.add_argument("--no-extras", dest="no_extras", action="store_true", help="skip the cfg2 / cfg3 / train.sh lines and the gpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
