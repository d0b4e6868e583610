"""N eager train steps of the bench workload (cfg4 per-GPU slice) for profiling under ncu: python tools/train_step_once.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mdctgan_b200.models.models import create_model
from mdctgan_b200.options.train_options import TrainOptions

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
opt = TrainOptions().parse(save=False, args=bench.OPT_ARGS + ["--gpu_ids", "0"])
torch.manual_seed(1234)
model = create_model(opt)
model.train()
lr = bench.make_lr_audio(bench.BATCH, bench.SEG, 42).to(dev)
hr = bench.make_hr_audio(bench.BATCH, bench.SEG, 42).to(dev)
for i in range(steps):
    torch.cuda.nvtx.range_push(f"step{i}")
    losses = model.train_step(lr, hr)
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("losses", losses.cpu().tolist())
