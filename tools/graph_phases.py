"""Where the captured train step spends its time: CUDA graphs of growing prefixes of the step (real stream structure), replayed and timed.
python tools/graph_phases.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mdctgan_b200 import nn_ops as ops, train_ops as T
from mdctgan_b200.models import pix2pixHD_model as PM
from mdctgan_b200.models.models import create_model
from mdctgan_b200.options.train_options import TrainOptions

dev = torch.device("cuda:0")
opt = TrainOptions().parse(save=False, args=bench.OPT_ARGS + ["--gpu_ids", "0"])
torch.manual_seed(1234)
model = create_model(opt); model.train()
lr = bench.make_lr_audio(bench.BATCH, bench.SEG, 42).to(dev); hr = bench.make_hr_audio(bench.BATCH, bench.SEG, 42).to(dev)
for _ in range(3):
    model.train_step(lr, hr)


def prefix(stage):
    """stage 1: forward; 2: + both sweeps (no update)"""
    g = T.GanGraph(model)
    with ops.stats_pass(dev):
        model.grad_all.zero_()
        g.forward(lr, hr)
        if stage >= 2:
            half = model._half_scalar()
            main = torch.cuda.current_stream(dev)
            sD = ops.aux_stream(dev, "sweep_D")
            sD.wait_stream(main)
            with torch.cuda.stream(sD):
                g.backward_D(half, half, join=False)
            g.backward_G(join=False)
            main.wait_stream(sD)
            ops.join_side_work(dev)
    g.release()


def timed(fn, name, n=30):
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn(); fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    for _ in range(3):
        gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1) / n:8.3f} ms  {name}", flush=True)


def g_only():
    with ops.stats_pass(dev):
        m = model
        lr_spectro, lr_input, _, _ = m._lr_input(lr)
        m.preprocess.hr_forward(hr)
        tape = ops.Tape()
        with ops.recording(tape):
            m.netG.run(ops.to_nhwc(lr_input))


def sweeps(which):
    g = T.GanGraph(model)
    with ops.stats_pass(dev):
        model.grad_all.zero_()
        g.forward(lr, hr)
        half = model._half_scalar()
        if which == "D":
            g.backward_D(half, half, join=False)
        else:
            g.backward_G(join=False)
        ops.join_side_work(dev)
    g.release()


timed(g_only, "2 MDCT + generator forward only")
timed(lambda: prefix(1), "forward (2 MDCT, G, D on [fake; real], losses)")
timed(lambda: sweeps("G"), "forward + generator sweep only (D dgrad on the fake half, G dgrad + wgrad)")
timed(lambda: sweeps("D"), "forward + discriminator sweep only (dgrad + wgrad)")
timed(lambda: prefix(2), "forward + both sweeps (dgrad, wgrad, norm backward), no update")
PM.PIPELINED_UPDATE = True
timed(lambda: model.train_step(lr, hr), "full step, pipelined per-bucket update")
PM.PIPELINED_UPDATE = False
timed(lambda: model.train_step(lr, hr), "full step, update after the sweeps")
ops.SIDE_STREAM_WGRAD = False
timed(lambda: prefix(2), "forward + both sweeps, weight gradients on the sweep streams (no side stream)")
ops.SIDE_STREAM_WGRAD = True
ops.PARALLEL_BRANCHES = False
timed(lambda: prefix(1), "forward, no branch streams")
