"""Phase breakdown of one eager train step (serialised: no side / branch streams), CUDA events at the phase boundaries."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mdctgan_b200 import nn_ops as ops, train_ops as T
from mdctgan_b200.models.models import create_model
from mdctgan_b200.options.train_options import TrainOptions

ops.SIDE_STREAM_WGRAD = ops.PARALLEL_BRANCHES = False
dev = torch.device("cuda:0")
opt = TrainOptions().parse(save=False, args=bench.OPT_ARGS + ["--gpu_ids", "0"])
torch.manual_seed(1234)
model = create_model(opt); model.train()
lr = bench.make_lr_audio(bench.BATCH, bench.SEG, 42).to(dev); hr = bench.make_hr_audio(bench.BATCH, bench.SEG, 42).to(dev)
for _ in range(3):
    model.train_step(lr, hr)
st = torch.cuda.current_stream()
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(st); return e
acc = {}
N = 5
for it in range(N):
    marks = [("start", ev())]
    g = T.GanGraph(model)
    with ops.stats_pass(dev):
        model.grad_all.zero_()
        # forward pieces
        m = model
        L = ops._L()
        lr_spectro, lr_input, _, _ = m._lr_input(lr); hr_spectro, _, _ = m.preprocess.hr_forward(hr)
        marks.append(("mdct x2", ev()))
        g.forward(lr, hr)
        marks.append(("forward G + D + losses (incl. mdct again)", ev()))
        # backward_G split: D part / G part is internal; time whole
        g.backward_G(join=False)
        marks.append(("sweep G (D dgrad on fake half + G dgrad/wgrad)", ev()))
        half = model._half_scalar()
        g.backward_D(half, half, join=False)
        marks.append(("sweep D (dgrad + wgrad)", ev()))
        model.optimizer_G.step(); model.optimizer_D.step()
        marks.append(("adam x2", ev()))
        model.packer.refresh()
        marks.append(("pack", ev()))
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
        acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) / N
for k, v in acc.items():
    print(f"{v:8.3f} ms  {k}")
# finer: G forward only / D forward only
from mdctgan_b200.nn_ops import Feat
with ops.stats_pass(dev), torch.no_grad():
    lr_spectro, lr_input, _, _ = model._lr_input(lr)
    for _ in range(2):
        e0 = ev(); out = model.netG.run(ops.to_nhwc(lr_input)); e1 = ev()
        din = torch.randn(8, 32, 256, 3, device=dev)
        e2 = ev(); f = model.netD.run_features(Feat(din)); e3 = ev()
    torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1):8.3f} ms  G forward alone (no tape)\n{e2.elapsed_time(e3):8.3f} ms  D forward alone on 2B (no tape)")
