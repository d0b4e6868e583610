import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
from mdctgan_b200 import nn_ops as ops, _lib
from mdctgan_b200.models import networks as N
from oracle import networks_oracle as NO
dev = torch.device("cuda:0")
def rl(a, b): return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))
for mode in ("randn", "const"):
    torch.manual_seed(1)
    D = N.define_D(3, 8, 2, "instance", False, 2, True)
    x = torch.randn(4, 3, 16, 256) * (1.0 if mode == "randn" else 0.05)
    if mode == "const":
        x[:, 2] = x[:, 1].abs() * 2 - 1
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in D.state_dict().items()}
    feats = NO.multiscale_d(sd, x, 2, 2)
    loss = sum(F.mse_loss(f[-1], torch.zeros_like(f[-1])) for f in feats)
    loss.backward()
    D = D.to(dev)
    tape = ops.Tape()
    with ops.stats_pass(dev), ops.recording(tape):
        leaf = ops.Feat(x.permute(0, 2, 3, 1).contiguous().to(dev), needs_grad=False)
        fs = D.run_features(leaf)
        G = ops.GradMap()
        L = ops._L()
        for sc in fs:
            pred = sc[-1]
            g = torch.empty_like(pred.x); n = g.numel()
            _lib.check(L.mdctgan_mse_const_bwd(pred.x.data_ptr(), n, 0.0, 1.0 / n, None, g.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
            G.add(pred, g)
        tape.backward(G)
    torch.cuda.synchronize()
    print(mode)
    for i, sc in enumerate(fs):
        for j, f in enumerate(sc):
            print("  feat", i, j, rl(f.x.permute(0, 3, 1, 2).cpu(), feats[i][j].detach()))
    for k, p in D.named_parameters():
        print(f"  {k:30s} {rl(p.grad.cpu(), sd[k].grad):9.2e} |ref| {float(sd[k].grad.norm()):9.2e}")
