import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mdctgan_b200 import nn_ops as ops
from mdctgan_b200.models import networks as N
from mdctgan_b200.packing import WeightPacker
from oracle import networks_oracle as NO
dev = torch.device("cuda:0")
def rl(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
torch.manual_seed(3)
kw = dict(input_nc=2, output_nc=1, ngf=8, netG="local", n_downsample_global=2, n_blocks_global=2, n_local_enhancers=1, n_blocks_local=1,
          norm="instance", input_size=(16, 256), n_attn_g=1, heads_g=2, dim_head_g=32, upsample_type="interpolate", downsample_type="resconv")
net = N.define_G(**kw)
x = (0.5 * torch.randn(2, 2, 16, 256)).clamp(-1, 1)
sd = net.state_dict()
with torch.no_grad():
    y_eval = NO.local_enhancer(sd, x, 2, 2, 1, 1, 2, 32, training=False, down="resconv", up="interpolate")
    y_train = NO.local_enhancer(sd, x, 2, 2, 1, 1, 2, 32, training=True, down="resconv", up="interpolate")
net = net.to(dev)
net.eval()
print("eval, per-layer packing:", rl(net(x.to(dev)).cpu(), y_eval))
net.train()
print("train-mode BN, per-layer packing:", rl(net(x.to(dev)).cpu(), y_train))
tape = ops.Tape()
with torch.no_grad(), ops.stats_pass(dev), ops.recording(tape):
    out = net.run(ops.to_nhwc(x.to(dev)))
print("train-mode, tape:", rl(out.x.view(2, 1, 16, 256).cpu(), y_train), "ops", len(tape.ops))
packer = WeightPacker(net)
print("train-mode, packer:", rl(net(x.to(dev)).cpu(), y_train), "generic", packer.n_desc, "tiled", packer.n_tiled)
for m in net.modules():
    if isinstance(m, (N.Conv2d, N.ConvTranspose2d)):
        sp = m._static_pack
        m2_kn = ops.pack_conv_weight(m.weight, isinstance(m, N.ConvTranspose2d))
        if sp["fwd"][0].stride(0) != 0 and not torch.equal(sp["fwd"][0], m2_kn):
            print("  fwd kn mismatch", type(m).__name__, m.in_channels, m.out_channels, m.kernel_size, rl(sp["fwd"][0], m2_kn))
