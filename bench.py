#!/usr/bin/env python
"""bench.py -- the MDCT -> generator -> IMDCT hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload = BASELINE.json configs[1]: GlobalGenerator only (n_blocks_attn_g = 0), 12 -> 48 kHz, ngf = 32,
n_downsample_global = 3, 9 residual blocks, batch 4, fp32, segments of 7936 samples (= 32 frames: 8192
samples give 33 frames, which the reference generator cannot process).  A "step" is one
`model.inference(lr_audio)` pass over one batch: fused MDCT+arcsinh/abs-norm kernel -> generator kernels ->
--fit_residual -> fused denorm+IMDCT kernel.  (The train step of the BASELINE metric -- discriminator,
losses, backward, Adam -- is not built in this round; DESIGN.md.)  Rank 0 prints ONE JSON line:

  value        whole-job audio-seconds per second, batch resident in HBM, the captured CUDA graph of the step
               replayed K times, CUDA events, max over ranks; N > 1 = N independent replicas (weak scaling,
               no collective on this path)
  e2e          same metric through the public API from pinned HOST audio: H2D copy, the step, D2H copy of
               the reconstructed audio, every step
  roofline     the dominant kernel of the step (the residual-block convolution), timed with CUDA events
               around every launch of an eager pass inside bench.py
  mdct         the transform half on its own at HBM-roofline scale (8192 clips x 8192 samples): GSamp/s and
               achieved GB/s of the fused forward / inverse kernels vs MEASURED_PEAKS.json
  cpu_baseline the reference's torch-CPU formulation of the same step (oracle/: torch_port + networks_oracle,
               kind "port") on this box's cores
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
G_KW = dict(ngf=32, n_down=3, n_blocks=9)
METRIC, UNIT = "inference-step audio-sec/sec (cfg2: MDCT4 -> GlobalGenerator ngf32 -> IMDCT4, batch 4 x 7936 samples, fp32)", "audio-s/s"
OPT_ARGS = ["--name", "bench", "--lr_sampling_rate", "12000", "--sr_sampling_rate", "48000", "--arcsinh_transform", "--abs_spectro",
            "--arcsinh_gain", "1000", "--center", "--norm_range", "-1", "1", "--abs_norm", "--src_range", "-5", "5", "--netG", "global",
            "--ngf", "32", "--n_downsample_global", "3", "--n_blocks_global", "9", "--n_blocks_attn_g", "0", "--segment_length", str(SEG),
            "--bins", "32", "--fit_residual", "--num_D", "1"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s)"


def workload_config():
    return {"workload": "cfg2 (BASELINE configs[1]): GlobalGenerator only, ngf 32, 3 downsamplings, 9 residual blocks, no attention, "
                        "12->48 kHz, --fit_residual, batch 4 segments x 7936 samples (32 frames x 256 bins), fp32; inference step",
            "batch_per_gpu": BATCH, "samples_per_segment": SEG, "n_fft": N_FFT, "hop": HOP,
            "l2": "working set (46 MB of weights + 20 MB of activations) is smaller than L2 by construction of the reference config; "
                  "the `mdct` sub-benchmark uses inputs larger than L2 (268 MB audio)"}


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


def cpu_port_run(seconds_budget=None, steps=None, warmup=1):
    """The reference's torch-CPU formulation of the step: oracle/torch_port.py (transform, complex128 FFT) +
    oracle/networks_oracle.py (the same F.conv2d / instance_norm calls the reference's modules make)."""
    import torch

    from mdctgan_b200.models import networks
    from oracle import networks_oracle as NO
    from oracle import torch_port as P

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    net = networks.define_G(2, 1, 32, "global", 3, 9, norm="instance", input_size=(32, 256))   # parameter holder only
    sd = net.state_dict()
    x = make_lr_audio(BATCH, SEG, 7)
    a2m = P.Audio2MDCTPort(GAIN, SRC, RNG, N_FFT, HOP)

    def step():
        with torch.no_grad():
            s, pha, prm = a2m.to_spectro(x)
            inp = torch.cat((s, s.abs() * 2 + RNG[0]), dim=1)
            sr = NO.global_generator(sd, inp, 3, 9)
            sr[..., :64] *= 1e-3
            sr = sr + s
            return a2m.to_audio(sr, prm, pha)

    for _ in range(warmup):
        step()
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if steps is not None and len(times) >= steps:
            break
        if steps is None and (time.perf_counter() - t_start) >= seconds_budget:
            break
    total = sum(times)
    return {"value": BATCH * SEG / SR * len(times) / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"the same step (batch {BATCH} x {SEG} samples), {len(times)} steps, torch-CPU {torch.get_num_threads()} threads, fp32 "
                      f"network + complex128 transform (oracle/torch_port.py + oracle/networks_oracle.py)",
            "ms_per_step": 1e3 * total / len(times)}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps = min(args.steps, 40)
    r = cpu_port_run(steps=steps, warmup=min(args.warmup, 3))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 network / f64 transform (reference dtypes)", "data": "synthetic", "config": workload_config(),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- transform sub-benchmark
def bench_mdct(dev, steps, hbm_peak):
    """The transform half at roofline scale: 8192 clips x 8192 samples, fused forward (2-channel spectrogram) + fused inverse."""
    import torch

    import mdctgan_b200
    from mdctgan_b200.models.pix2pixHD_model import Audio2MDCT, default_audio_opt

    B, T = 8192, 8192
    F = T // HOP + 1
    a2m = Audio2MDCT(default_audio_opt(arcsinh_gain=GAIN, src_range=SRC, norm_range=RNG, gpu_ids=[dev.index]), device=dev)
    x = 0.1 * torch.randn(B, T, device=dev)
    spec = torch.empty(B, 2, F, NBINS, device=dev)
    st = torch.cuda.current_stream(dev)
    for _ in range(5):
        a2m.to_spectro(x, channels=2, out=spec)
        y = a2m.to_audio(spec[:, 0])
    torch.cuda.synchronize(dev)
    n0 = mdctgan_b200.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for k in range(steps):
        evs[k][0].record(st)
        a2m.to_spectro(x, channels=2, out=spec)
        evs[k][1].record(st)
        y = a2m.to_audio(spec[:, 0])
        evs[k][2].record(st)
    torch.cuda.synchronize(dev)
    total = evs[0][0].elapsed_time(evs[-1][2]) / steps
    fwd = statistics.fmean(e[0].elapsed_time(e[1]) for e in evs)
    inv = statistics.fmean(e[1].elapsed_time(e[2]) for e in evs)
    rt = ((y.reshape(B, T)[:64].double() - x[:64].double()).norm() / x[:64].double().norm()).item()
    assert rt < 1e-3, rt
    fb, ib = B * (4 * T + 8 * F * NBINS), B * (4 * F * NBINS + 4 * T)
    return {"workload": f"{B} clips x {T} samples, fused MDCT4+arcsinh/abs-norm (2-channel fp32 spectrogram) -> fused denorm+IMDCT4",
            "gsamp_per_s_round_trip": B * T / (total * 1e-3) / 1e9, "ms_per_round_trip": total, "launches": mdctgan_b200.launch_count() - n0,
            "forward": {"kernel": "mdct4_fwd_kernel<float,1>", "avg_launch_ms": fwd, "algorithmic_bytes_per_launch": fb,
                        "achieved_gbs": fb / (fwd * 1e-3) / 1e9, "frac_of_hbm_peak": fb / (fwd * 1e-3) / 1e9 / hbm_peak,
                        "gsamp_per_s": B * T / (fwd * 1e-3) / 1e9},
            "inverse": {"kernel": "imdct4_inv_kernel<float,float,float,1>", "avg_launch_ms": inv, "algorithmic_bytes_per_launch": ib,
                        "achieved_gbs": ib / (inv * 1e-3) / 1e9, "frac_of_hbm_peak": ib / (inv * 1e-3) / 1e9 / hbm_peak,
                        "gsamp_per_s": B * T / (inv * 1e-3) / 1e9},
            "round_trip_rel_l2": rt, "hbm_peak_gbs": hbm_peak}


# ---------------------------------------------------------------------------------------------- per-launch profile
class LaunchProfiler:
    """CUDA events around every C-ABI launch of an eager pass (bench-only instrumentation)."""

    def __init__(self, dev):
        import torch

        from mdctgan_b200 import _lib, nn_ops

        self.torch, self.dev = torch, dev
        self.L = nn_ops._L()
        self.names = [n for n in dir(self.L) if n.startswith("mdctgan_")] or []
        self.names = ["mdctgan_conv2d_nhwc", "mdctgan_conv2d_umma", "mdctgan_norm_finalize", "mdctgan_norm_apply", "mdctgan_attention_abs_pos",
                      "mdctgan_avgpool3s2_nhwc", "mdctgan_nchw_to_nhwc", "mdctgan_nhwc_to_nchw", "mdctgan_residual_scale_add",
                      "mdctgan_audio2mdct_forward", "mdctgan_mdct2audio_inverse"]
        self.orig, self.records = {}, []

    def __enter__(self):
        torch = self.torch
        for n in self.names:
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
                tag = _n
                if _n in ("mdctgan_conv2d_nhwc", "mdctgan_conv2d_umma"):
                    B, H, W, Cin, Cout, kh, stride, transposed = a[1], a[2], a[3], a[4], a[10], a[11], a[13], a[16]
                    Ho, Wo = a[8], a[9]
                    eng = "umma" if _n.endswith("umma") else "direct"
                    tag = f"conv[{eng}] k{kh} s{stride}{' T' if transposed else ''} {Cin}->{Cout} @{Ho}x{Wo}"
                    self.records.append((tag, e0, e1, 2.0 * B * Ho * Wo * Cout * kh * kh * Cin / (stride * stride if transposed else 1)))
                else:
                    self.records.append((tag, e0, e1, 0.0))
                return rc

            setattr(self.L, n, wrap)
        return self

    def __exit__(self, *exc):
        for n, f in self.orig.items():
            setattr(self.L, n, f)

    def table(self):
        self.torch.cuda.synchronize(self.dev)
        agg = {}
        for tag, e0, e1, flops in self.records:
            t = e0.elapsed_time(e1)
            a = agg.setdefault(tag, [0, 0.0, flops])
            a[0] += 1
            a[1] += t
        return agg


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import mdctgan_b200
    from mdctgan_b200.models.models import create_model
    from mdctgan_b200.options.train_options import TrainOptions
    from mdctgan_b200.runtime import GraphedInference

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
    torch.manual_seed(1234)                      # identical weights on every rank
    model = create_model(opt)
    model.eval()
    lr = make_lr_audio(BATCH, SEG, 42 + rank).to(dev)
    warm = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- eager warm-up + correctness guard against the CPU oracle (outside every timed region)
    for _ in range(warm):
        out = model.inference(lr)
    if rank == 0:
        from oracle import model_oracle as MOD

        sd = {k: v.cpu() for k, v in model.netG.state_dict().items()}
        _, ref_audio, _ = MOD.inference(sd, lr.cpu().numpy(), netG="global", n_down=3, n_blocks_global=9, fit_residual=True, up_ratio=4.0)
        err = float(((out[1].cpu().double().numpy() - ref_audio) ** 2).sum() ** 0.5 / (ref_audio ** 2).sum() ** 0.5)
        assert err < 1e-3, f"waveform rel-L2 vs oracle {err} breaks the 1e-3 bar"
    else:
        err = None

    # ---- timed region: the step as a CUDA graph, K replays
    gi = GraphedInference(model, BATCH, SEG, warmup=2)
    gi.static_in.copy_(lr)
    st = torch.cuda.current_stream(dev)
    for _ in range(warm):
        gi.replay()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        gi.replay()
    e1.record(st)
    barrier()
    step_ms = e0.elapsed_time(e1) / args.steps

    # ---- launches per step + per-kernel table from an eager, event-instrumented pass
    n0 = mdctgan_b200.launch_count()
    model.inference(lr)
    launches_per_step = mdctgan_b200.launch_count() - n0
    with LaunchProfiler(dev) as prof:
        for _ in range(5):
            model.inference(lr)
    table = prof.table()
    tot_ms = sum(v[1] for v in table.values())
    dom_tag, dom = max(((k, v) for k, v in table.items() if k.startswith("conv")), key=lambda kv: kv[1][1])
    dom_ms = dom[1] / dom[0]
    kernel_table = {k: {"launches_per_step": v[0] // 5, "ms_per_step": v[1] / 5, "share": v[1] / tot_ms} for k, v in
                    sorted(table.items(), key=lambda kv: -kv[1][1])[:8]}

    # ---- e2e: pinned host audio -> H2D -> graph replay -> D2H of the reconstructed audio, every step
    xh = lr.cpu().pin_memory()
    yh = torch.empty(BATCH, 1, 1, SEG, dtype=torch.float32).pin_memory()
    for _ in range(3):
        yh.copy_(gi(xh)[1], non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        yh.copy_(gi(xh)[1], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()     # the caller consumes each result before submitting the next batch
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()

    mdct = bench_mdct(dev, 50, hbm_peak) if rank == 0 else None

    vals = torch.tensor([step_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms = vals.tolist()

    if rank == 0:
        audio_s = world * BATCH * SEG / SR
        cpu = cpu_port_run(seconds_budget=args.cpu_seconds, warmup=1)
        flops = dom[2]
        line = {
            "metric": METRIC, "value": audio_s / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(),
            "roofline": {"bound": "tensor", "kernel": f"{'conv2d_umma_kernel<64,true> (tcgen05 kind::tf32, 3xTF32 split, cluster split-K)' if 'umma' in dom_tag else 'conv2d_nhwc_kernel (fp32 FFMA)'}: {dom_tag}",
                         "achieved": flops / (dom_ms * 1e-3) / 1e12, "peak": bf16_peak, "unit": "TFLOP/s",
                         "frac": flops / (dom_ms * 1e-3) / 1e12 / bf16_peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_flops_per_launch": flops, "avg_launch_ms": dom_ms, "share_of_step": dom[1] / tot_ms,
                         "note": "algorithmic flops = 2*M*N*K of the convolution (the 3xTF32 split issues 3x that on the tensor pipe); "
                                 "peak = measured dense bf16 cuBLAS burst; at batch 4 x 4x32 pixels the layer is 0.6 GFLOP over 2.4 MB of "
                                 "weights: latency / weight-bandwidth bound, not tensor bound (DESIGN.md)"},
            "kernel_table": kernel_table,
            "mdct": mdct,
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": audio_s / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": BATCH * SEG * 4, "d2h_bytes_per_step": BATCH * SEG * 4,
                    "ms_per_step": e2e_ms, "api": "GraphedInference(model)(pinned lr_audio) -> sr_audio copied to pinned host memory, "
                                                  "stream-synchronised every step"},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step, "clocks": clocks,
            "waveform_rel_l2_vs_oracle": err,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
