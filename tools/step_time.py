"""CUDA-graph train-step time of the bench workload under the current MDCTGAN_CONV_ENGINE (umma = 3xTF32 default | tf32 | direct)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mdctgan_b200 import nn_ops as ops
from mdctgan_b200.models.models import create_model
from mdctgan_b200.options.train_options import TrainOptions
from mdctgan_b200.runtime import GraphedTrainStep

dev = torch.device("cuda:0")
opt = TrainOptions().parse(save=False, args=bench.OPT_ARGS + ["--gpu_ids", "0"])
torch.manual_seed(1234); torch.cuda.manual_seed(1234)
model = create_model(opt); model.train()
lr = bench.make_lr_audio(bench.BATCH, bench.SEG, 42).to(dev); hr = bench.make_hr_audio(bench.BATCH, bench.SEG, 42).to(dev)
first = model.train_step(lr, hr).cpu().tolist()
gts = GraphedTrainStep(model, bench.BATCH, bench.SEG)
gts.lr_in.copy_(lr); gts.hr_in.copy_(hr); gts.recapture()
for _ in range(5): gts.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(100): gts.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 100
print(f"engine={ops.CONV_ENGINE} ms_per_step={ms:.3f} audio_s_per_s={bench.BATCH * bench.SEG / bench.SR / (ms * 1e-3):.1f} first_losses={first}")
