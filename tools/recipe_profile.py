"""Per-kernel table of the train.sh recipe train step (ngf 56, resconv / interpolate, 3 attention layers 6 x 128, 128 frames, batch B):
python tools/recipe_profile.py [batch]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
dev = torch.device("cuda:0")
extra = [("--ngf", "56"), ("--n_blocks_global", "4"), ("--n_blocks_attn_g", "3"), ("--heads_g", "6"), ("--dim_head_g", "128"),
         ("--segment_length", "32512"), ("--bins", "128"), ("--lr_sampling_rate", "16000"), ("--upsample_type", "interpolate"),
         ("--downsample_type", "resconv"), ("--lr", "0.00015")]
model = bench._build_model(extra, dev)
model.train()
lr, hr = bench.make_lr_audio(B, 32512, 11).to(dev), bench.make_hr_audio(B, 32512, 11).to(dev)
for _ in range(2):
    model.train_step(lr, hr)
with bench.LaunchProfiler(dev) as prof:
    model.train_step(lr, hr)
table = prof.table()
tot = sum(v[1] for v in table.values())
print(f"eager sum of kernels {tot:.1f} ms")
for k, v in sorted(table.items(), key=lambda kv: -kv[1][1])[:400]:
    tf = f"{v[2] / (v[1] / v[0] * 1e-3) / 1e12:6.1f} TF/s" if v[2] else ""
    print(f"{k:58s} n={v[0]:3d} {v[1]:8.2f} ms {100 * v[1] / tot:5.1f}%  {tf}")

# ---- the same recipe step by the unmodified reference on this GPU (torch eager, cuDNN): TF32 default policy, and the recipe's own --fp16 AMP
if "--ref" in sys.argv:
    import time
    from baseline import ref_runner as R
    del model
    torch.cuda.empty_cache()
    torch.backends.cudnn.benchmark = True
    args = [a for a in bench.OPT_ARGS]
    for k, v in extra:
        if k in args:
            args[args.index(k) + 1] = v
        else:
            args += [k, v]
    for name, fp16 in (("tf32_default", False), ("amp_fp16", True)):
        try:
            step, m = R.make_stepper(args, lr.cpu(), hr.cpu(), device=dev, seed=1234, fp16=fp16)
            first = step()
            for _ in range(4):
                step(False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 8
            for _ in range(n):
                step(False)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / n * 1e3
            print(json.dumps({"reference_eager_" + name: {"ms_per_step": ms, "audio_sec_per_sec": B * 32512 / 48000 / (ms * 1e-3), "first_losses": first}}))
            del step, m
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"reference_eager_" + name: {"error": repr(e)[:300]}}))
