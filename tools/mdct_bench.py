#!/usr/bin/env python
"""Times the MDCT4 / IMDCT4 kernels of every arithmetic flavour on one GPU (CUDA events, L2-exceeding buffers) and reports
GSamp/s, achieved algorithmic GB/s, the fraction of the measured HBM peak and the round-trip error in eps*peak units.

    python tools/mdct_bench.py [--out gpurun_out/mdct_bench.json] [--reps 30]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mdctgan_b200.models.mdct import IMDCT4, MDCT4  # noqa: E402
from mdctgan_b200.models.pix2pixHD_model import Audio2MDCT, default_audio_opt  # noqa: E402
from mdctgan_b200.util.util import kbdwin  # noqa: E402

EPS = 2.0 ** -23


def timed(fn, reps, dev):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "mdct_bench.json"))
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--flavours", default="fp32,mixed")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = 6454.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    w = kbdwin(512)
    res = {"hbm_peak_gbs": peak, "cases": []}
    for (B, T) in ((8192, 8192), (64, 32512)):
        torch.manual_seed(0)
        x = 0.1 * torch.randn(B, T, device=dev)
        F = T // 256 + 1
        for prec in args.flavours.split(","):
            fwd = MDCT4(512, 256, 512, w, device=dev, precision=prec)
            inv = IMDCT4(512, 256, 512, w, device=dev, precision=prec)
            a2m = Audio2MDCT(default_audio_opt(arcsinh_gain=1000.0, src_range=(-5.0, 5.0), norm_range=(-1.0, 1.0), gpu_ids=[0]), device=dev,
                             precision=prec)
            spec, _ = fwd(x)
            y, _ = inv(spec)
            err = (y.reshape(B, -1)[:, :T].double() - x[:, :y.shape[-1]].double()).abs().amax(dim=1) / x.abs().amax(dim=1).double() / EPS
            esz = spec.element_size()
            t_f = timed(lambda: fwd(x), args.reps, dev)
            t_i = timed(lambda: inv(spec), args.reps, dev)
            out2 = torch.empty(B, 2, F, 256, device=dev)
            t_ff = timed(lambda: a2m.to_spectro(x, channels=2, out=out2), args.reps, dev)
            s1 = out2[:, 0].contiguous()
            t_fi = timed(lambda: a2m.to_audio(s1), args.reps, dev)
            n = B * T
            rows = {"raw_forward": (t_f, 4 * n + esz * B * F * 256), "raw_inverse": (t_i, esz * B * F * 256 + esz * n),
                    "fused_forward_2ch": (t_ff, 4 * n + 8 * B * F * 256), "fused_inverse": (t_fi, 4 * B * F * 256 + 4 * n)}
            case = {"B": B, "T": T, "precision": prec, "round_trip_max_err_eps_peak": float(err.max()),
                    "round_trip_mean_err_eps_peak": float(err.mean())}
            for k, (ms, nbytes) in rows.items():
                case[k] = {"ms": ms, "gsamp_per_s": n / ms / 1e6, "algorithmic_gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak}
            res["cases"].append(case)
            print(json.dumps(case), flush=True)
            del out2, s1, spec, y
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
