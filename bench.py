#!/usr/bin/env python
"""bench.py -- the hot path (audio -> fused MDCT+normalise -> [generator, when built] -> fused denormalise+IMDCT
-> audio) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clips B] [--precision fp32|fp64]

A "step" is one pass of the path over one batch of `--clips` synthetic 48 kHz / 8192-sample clips per GPU
(BASELINE.json north_star: "synthetic 48 kHz / 8192-sample segments").  Rank 0 prints ONE JSON line.

  value      whole-job GSamp/s, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e        same metric through the public API from pinned HOST buffers: H2D of the audio, both kernels,
             D2H of the reconstructed audio, all inside the timed region
  roofline   the dominant kernel (fused forward: 4 B/sample in + 4*C*F*256/T B/sample out) against the measured
             HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the reference's torch-CPU formulation (oracle/torch_port.py, kind "port") on this box's cores,
             bounded sample

`--impl reference` times that CPU port alone (rank 0 only) with the same metric / unit / config keys.
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

T_CLIP = 8192          # samples per clip (BASELINE.json)
SR = 48000
N_FFT, HOP, NBINS = 512, 256, 256
F_CLIP = T_CLIP // HOP + 1
CHANNELS = 2           # the generator's input: s and |s|*2+lo (pix2pixHD_model.py:400-402)
GAIN, SRC, RNG = 1000.0, (-5.0, 5.0), (-1.0, 1.0)
METRIC, UNIT = "MDCT GSamp/s (audio->fused MDCT4+arcsinh/abs-norm->fused denorm+IMDCT4->audio round trip)", "GSamp/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
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
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if bit and (r & bit):
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
                    "note": "no NVML samples" + (": " + self.err if not self.ok else "")}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------- CPU port
def cpu_port_run(clips, seconds_budget=None, steps=None, warmup=1):
    """Time the reference's torch-CPU formulation (oracle/torch_port.py) on `clips` clips per step."""
    import torch

    from oracle import torch_port as P

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    x = 0.1 * torch.randn(clips, T_CLIP)
    a2m = P.Audio2MDCTPort(GAIN, SRC, RNG, N_FFT, HOP)

    def step():
        s, pha, prm = a2m.to_spectro(x)
        s2 = torch.cat((s, s.abs() * 2 + RNG[0]), dim=1)      # pix2pixHD_model.py:400-402
        return a2m.to_audio(s2[:, :1], prm, pha)

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
    gsamp = clips * T_CLIP * len(times) / total / 1e9
    return {"value": gsamp, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{clips} clips x {T_CLIP} samples per step, {len(times)} steps, torch-CPU {torch.get_num_threads()} threads, "
                      f"oracle/torch_port.py (reference formulation: complex128 512-pt FFT)",
            "ms_per_step": 1e3 * total / len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clips = args.ref_clips
    # keep the whole run within a few minutes: one step of 256 clips takes ~0.1-0.3 s on a server CPU
    r = cpu_port_run(clips, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex128 FFT, reference dtypes)", "data": "synthetic",
        "config": workload_config(clips, "fp64"),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(clips, precision):
    return {"workload": f"cfg1-batched: MDCT4->IMDCT4 round trip with fused arcsinh/abs-norm, {clips} clips x {T_CLIP} samples "
                        f"@48 kHz per GPU per step ({F_CLIP} frames x {NBINS} bins, {CHANNELS}-channel spectrogram)",
            "clips_per_gpu": clips, "samples_per_clip": T_CLIP, "n_fft": N_FFT, "hop": HOP, "precision": precision,
            "generator": "not in the timed path yet (round 1: transform half of the hot path)",
            "l2": "inputs larger than L2: no flush needed"}


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import mdctgan_b200
    from mdctgan_b200.models.pix2pixHD_model import Audio2MDCT, default_audio_opt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, F = args.clips, T_CLIP, F_CLIP
    torch.manual_seed(42 + rank)
    a2m = Audio2MDCT(default_audio_opt(arcsinh_gain=GAIN, src_range=SRC, norm_range=RNG, gpu_ids=[local]), device=dev,
                     precision=args.precision)
    out_dt = torch.float64 if args.precision == "fp64" else torch.float32
    x = 0.1 * torch.randn(B, T, device=dev)
    spec = torch.empty(B, CHANNELS, F, NBINS, device=dev, dtype=torch.float32)
    stream = torch.cuda.current_stream(dev)

    def step():
        a2m.to_spectro(x, channels=CHANNELS, out=spec)
        return a2m.to_audio(spec[:, 0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        y = step()
    barrier()

    # ---- timed region: K steps, per-kernel events on the launching stream
    sampler = ClockSampler(local)
    sampler.start()
    n0 = mdctgan_b200.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        evs[k][0].record(stream)
        a2m.to_spectro(x, channels=CHANNELS, out=spec)
        evs[k][1].record(stream)
        y = a2m.to_audio(spec[:, 0])
        evs[k][2].record(stream)
    barrier()
    launches = mdctgan_b200.launch_count() - n0
    total_ms = evs[0][0].elapsed_time(evs[-1][2])
    fwd_ms = statistics.fmean(e[0].elapsed_time(e[1]) for e in evs)
    inv_ms = statistics.fmean(e[1].elapsed_time(e[2]) for e in evs)
    clocks = sampler.stop()

    # ---- parity guard on the data just timed (cheap, outside the timed region)
    rt = ((y.reshape(B, T)[:64].double() - x[:64].double()).norm() / x[:64].double().norm()).item()
    assert rt < 1e-3, f"round trip rel-L2 {rt} breaks the 1e-3 bar"

    # ---- e2e: pinned host audio -> H2D -> both kernels -> D2H audio, through the public API
    Be = min(B, args.e2e_clips)
    xh = torch.empty(Be, T, dtype=torch.float32).pin_memory()
    xh.copy_(x[:Be])
    yh = torch.empty(Be, 1, 1, T, dtype=out_dt).pin_memory()
    nchunk = 4
    streams = [torch.cuda.Stream(dev) for _ in range(nchunk)]
    bounds = [(i * Be // nchunk, (i + 1) * Be // nchunk) for i in range(nchunk)]
    xd = [torch.empty(b1 - b0, T, device=dev) for b0, b1 in bounds]
    sd = [torch.empty(b1 - b0, CHANNELS, F, NBINS, device=dev) for b0, b1 in bounds]

    def e2e_step():
        for i, (b0, b1) in enumerate(bounds):
            with torch.cuda.stream(streams[i]):
                xd[i].copy_(xh[b0:b1], non_blocking=True)
                a2m.to_spectro(xd[i], channels=CHANNELS, out=sd[i])
                yh[b0:b1].copy_(a2m.to_audio(sd[i][:, 0]), non_blocking=True)
        for s in streams:
            s.synchronize()

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    assert ((yh.reshape(Be, T)[:8].double() - xh[:8].double()).norm() / xh[:8].double().norm()).item() < 1e-3

    # ---- max over ranks
    vals = torch.tensor([total_ms, fwd_ms, inv_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    total_ms, fwd_ms, inv_ms, e2e_ms = vals.tolist()

    if rank == 0:
        ms_per_step = total_ms / args.steps
        samples_step = world * B * T
        value = samples_step / (ms_per_step * 1e-3) / 1e9
        # algorithmic bytes of the dominant (forward) kernel per launch: audio in + C-channel spectrogram out
        fwd_bytes = B * (4 * T + 4 * CHANNELS * F * NBINS)
        inv_bytes = B * (4 * F * NBINS + (8 if args.precision == "fp64" else 4) * T)
        peak, peak_src = peaks()
        ach = fwd_bytes / (fwd_ms * 1e-3) / 1e9
        ach_inv = inv_bytes / (inv_ms * 1e-3) / 1e9
        cpu = cpu_port_run(args.ref_clips, seconds_budget=args.cpu_seconds, warmup=1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
            "config": workload_config(B, args.precision),
            "audio_sec_per_sec": samples_step / SR / (ms_per_step * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "mdct4_fwd_kernel<float,1> (fused MDCT4+arcsinh+abs-norm, 2 channels)"
                         if args.precision == "fp32" else "mdct4_fwd_kernel<double,1>",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": fwd_bytes, "avg_launch_ms": fwd_ms,
                         "inverse_kernel": {"achieved": ach_inv, "frac": ach_inv / peak, "avg_launch_ms": inv_ms,
                                            "algorithmic_bytes_per_launch": inv_bytes}},
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": world * Be * T / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": Be * T * 4,
                    "d2h_bytes_per_step": Be * T * yh.element_size(), "ms_per_step": e2e_ms, "clips_per_gpu": Be,
                    "api": "Audio2MDCT.to_spectro / to_audio on pinned host tensors, 4 chunks on 4 streams"},
            "gpu_launches": launches, "clocks": clocks, "round_trip_rel_l2": rt,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=8192, help="clips per GPU per step (8192 x 8192 samples = 268 MB of audio)")
    ap.add_argument("--e2e-clips", type=int, default=8192)
    ap.add_argument("--ref-clips", type=int, default=256, help="clips per step of the CPU port (bounded sample)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 50:       # bounded: the CPU port needs ~0.2 s per 256-clip step
            args.steps = 50
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
