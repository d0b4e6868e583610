"""Per-tensor gradient error of the CUDA train step vs oracle/train_oracle.py (debug aid; run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
from oracle import train_oracle as TO
from test_oracle_train import flags_to_cfg, build_nets
from make_golden_nets import TRAIN_FLAGS
from test_train_gpu import _build_model

name = sys.argv[1] if len(sys.argv) > 1 else "tr_small"
dev = torch.device("cuda:0")
gold = dict(np.load(os.path.join(ROOT, "tests/golden/train_golden.npz")))
flags, batch, T, seed = TRAIN_FLAGS[name]
cfg = flags_to_cfg(flags)
model = _build_model(list(flags) + ["--mdct_precision", os.environ.get("MDCT_PREC", "fp32")], seed, dev)
G0, D0 = build_nets(cfg, seed)
model.netG.load_state_dict(G0.state_dict()); model.netD.load_state_dict(D0.state_dict())
kw = {k: cfg[k] for k in ("netG", "n_down", "n_blocks_global", "n_blocks_local", "n_attn", "heads", "dim_head", "num_D", "n_layers_D", "fit_residual", "down", "up")}
ref = TO.train_step(G0.state_dict(), D0.state_dict(), gold[f"{name}_lr_audio"], gold[f"{name}_hr_audio"], steps=1, **kw)
lr_d, hr_d = torch.from_numpy(gold[f"{name}_lr_audio"]).to(dev), torch.from_numpy(gold[f"{name}_hr_audio"]).to(dev)
from mdctgan_b200 import train_ops as T, nn_ops as ops
model.optimizer_G.zero_grad(); model.optimizer_D.zero_grad()
model._refresh_weight_images()      # load_state_dict above changed the weights behind the packer
graph = T.GanGraph(model)
with ops.stats_pass(dev):
    graph.forward(lr_d, hr_d)
    half = torch.full((), 0.5, device=dev)
    if os.environ.get("D_FIRST"):
        graph.backward_D(half, half)
        graph.backward_G()
    else:
        graph.backward_G()
        graph.backward_D(half, half)
torch.cuda.synchronize()
print("losses", graph.losses.cpu().tolist(), ref["losses"])
print("sr rel", float((graph.sr_spectro.cpu() - ref["sr_spectro"]).norm() / ref["sr_spectro"].norm()))
def rl(a, b): return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))
for k, p in model.netG.named_parameters():
    r = ref["gradG"][k]
    print(f"G {k:45s} {rl(p.grad.cpu(), r):9.2e}  |ref| {float(r.norm()):9.2e}")
for k, p in model.netD.named_parameters():
    r = ref["gradD"][k] * 1.0
    print(f"D {k:45s} {rl(p.grad.cpu(), r):9.2e}  |ref| {float(r.norm()):9.2e}")
# ---- D backward on OUR discriminator input (isolates the sweep from forward differences upstream of D)
import torch.nn.functional as F
from oracle import networks_oracle as NO
din = graph.din.x.permute(0, 3, 1, 2).contiguous().cpu()
B = din.shape[0] // 2
sd = {k: v.detach().clone().requires_grad_(True) for k, v in D0.state_dict().items()}
pf = NO.multiscale_d(sd, din[:B], cfg["num_D"], cfg["n_layers_D"])
pr = NO.multiscale_d(sd, din[B:], cfg["num_D"], cfg["n_layers_D"])
loss = 0.5 * (sum(F.mse_loss(p[-1], torch.zeros_like(p[-1])) for p in pf) + sum(F.mse_loss(p[-1], torch.ones_like(p[-1])) for p in pr))
loss.backward()
print("D grads vs oracle D on OUR input:")
for k, p in model.netD.named_parameters():
    if k.endswith("weight"):
        print(f"D' {k:45s} {rl(p.grad.cpu(), sd[k].grad):9.2e}")
print("din fake-half rel diff vs oracle sr:", rl(din[:B, 1], ref["sr_spectro"][:, 0]))
# ---- conditioning probe: oracle D gradients under a 1e-6 relative perturbation of the same input
torch.manual_seed(0)
din2 = din * (1 + 1e-6 * torch.randn_like(din))
sd2 = {k: v.detach().clone().requires_grad_(True) for k, v in D0.state_dict().items()}
pf = NO.multiscale_d(sd2, din2[:B], cfg["num_D"], cfg["n_layers_D"])
pr = NO.multiscale_d(sd2, din2[B:], cfg["num_D"], cfg["n_layers_D"])
loss = 0.5 * (sum(F.mse_loss(p[-1], torch.zeros_like(p[-1])) for p in pf) + sum(F.mse_loss(p[-1], torch.ones_like(p[-1])) for p in pr))
loss.backward()
for k in sd:
    if k.endswith("weight"):
        print(f"cond {k:45s} {rl(sd2[k].grad, sd[k].grad):9.2e}")
# forward features of scale 0 (ours vs oracle on our input)
fo = NO.multiscale_d(sd, din, cfg["num_D"], cfg["n_layers_D"])
for i, sc in enumerate(graph.feats):
    for j, f in enumerate(sc):
        print("feat", i, j, rl(f.x.permute(0, 3, 1, 2).cpu(), fo[i][j].detach()))
# ---- generator forward on OUR network input, against the oracle generator (isolates G from the transform)
with torch.no_grad():
    lr_spectro, lr_input, _, _ = model._lr_input(lr_d)
    xin = lr_input.cpu()
    fn = NO.global_generator if cfg["netG"] == "global" else NO.local_enhancer
    args = (cfg["n_down"], cfg["n_blocks_global"]) + (() if cfg["netG"] == "global" else (cfg["n_blocks_local"],))
    y_or = fn(G0.state_dict(), xin, *args, cfg["n_attn"], cfg["heads"], cfg["dim_head"], training=True, down=cfg["down"], up=cfg["up"])
    with ops.stats_pass(dev):
        y_run = model.netG.run(ops.to_nhwc(lr_input)).x.reshape(y_or.shape).cpu()
    print("G forward (train-mode BN) on our input vs oracle:", rl(y_run, y_or))
    for k, v in G0.state_dict().items():
        w = model.netG.state_dict()[k].cpu()
        if not torch.equal(w, v) and v.dtype.is_floating_point and "running" not in k and "num_batches" not in k:
            print("  weight differs from the seeded CPU init:", k, rl(w, v))
            break
