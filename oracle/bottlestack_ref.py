"""Oracle restatement of `bottleneck_transformer_pytorch==0.1.4` (BoTNet block).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the package is a pip dependency of the reference
(requirements.txt:1), imported lazily at models/networks.py:232 and :341, and is
neither vendored under /root/reference nor installed in this image.  This file
restates the published 0.1.4 semantics for the options the reference uses
(`downsample=False, rel_pos_emb=False`), cross-checked only against the
state_dict key/shape layout the reference checkpoints expect
(`net.{i}.net.{0,1,3,5,7,8}`, `...3.to_qkv.weight`, `...3.pos_emb.{height,width}`).

It is also what tests/golden/make_golden.py installs as the
`bottleneck_transformer_pytorch` module when it runs the reference.
"""
import torch
from torch import nn


class AbsPosEmb(nn.Module):
    def __init__(self, fmap_size, dim_head):
        super().__init__()
        h, w = fmap_size
        s = dim_head ** -0.5
        self.height = nn.Parameter(torch.randn(h, dim_head) * s)
        self.width = nn.Parameter(torch.randn(w, dim_head) * s)

    def forward(self, q):  # q [b, heads, L, d] -> logits [b, heads, L, L]
        emb = (self.height[:, None, :] + self.width[None, :, :]).reshape(-1, self.height.shape[-1])
        return torch.einsum("bhid,jd->bhij", q, emb)


class Attention(nn.Module):
    def __init__(self, *, dim, fmap_size, heads=4, dim_head=128, rel_pos_emb=False):
        super().__init__()
        assert not rel_pos_emb, "reference always passes rel_pos_emb=False"
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Conv2d(dim, heads * dim_head * 3, 1, bias=False)
        self.pos_emb = AbsPosEmb(fmap_size, dim_head)

    def forward(self, fmap):
        b, _, hh, ww = fmap.shape
        q, k, v = self.to_qkv(fmap).chunk(3, dim=1)

        def split(t):  # 'b (h d) x y -> b h (x y) d'
            return t.reshape(b, self.heads, -1, hh * ww).transpose(-1, -2)

        q, k, v = split(q) * self.scale, split(k), split(v)
        sim = torch.einsum("bhid,bhjd->bhij", q, k) + self.pos_emb(q)
        out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
        return out.transpose(-1, -2).reshape(b, -1, hh, ww)  # 'b h (x y) d -> b (h d) x y'


class BottleBlock(nn.Module):
    def __init__(self, *, dim, fmap_size, dim_out, proj_factor, downsample, heads=4, dim_head=128,
                 rel_pos_emb=False, activation=nn.ReLU()):
        super().__init__()
        if dim != dim_out or downsample:
            ks, st, pd = (3, 2, 1) if downsample else (1, 1, 0)
            self.shortcut = nn.Sequential(nn.Conv2d(dim, dim_out, ks, stride=st, padding=pd, bias=False),
                                          nn.BatchNorm2d(dim_out), activation)
        else:
            self.shortcut = nn.Identity()
        inner_in = dim_out // proj_factor
        inner_out = heads * dim_head
        self.net = nn.Sequential(
            nn.Conv2d(dim, inner_in, 1, bias=False),
            nn.BatchNorm2d(inner_in),
            activation,
            Attention(dim=inner_in, fmap_size=fmap_size, heads=heads, dim_head=dim_head, rel_pos_emb=rel_pos_emb),
            nn.AvgPool2d((2, 2)) if downsample else nn.Identity(),
            nn.BatchNorm2d(inner_out),
            activation,
            nn.Conv2d(inner_out, dim_out, 1, bias=False),
            nn.BatchNorm2d(dim_out),
        )
        nn.init.zeros_(self.net[-1].weight)  # library zero-gamma; networks.weights_init overrides it
        self.activation = activation

    def forward(self, x):
        sc = self.shortcut(x)
        x = self.net(x)
        x = x + sc
        return self.activation(x)


class BottleStack(nn.Module):
    def __init__(self, *, dim, fmap_size, dim_out=2048, proj_factor=4, num_layers=3, heads=4, dim_head=128,
                 downsample=True, rel_pos_emb=False, activation=nn.ReLU()):
        super().__init__()
        if isinstance(fmap_size, int):
            fmap_size = (fmap_size, fmap_size)
        self.dim = dim
        self.fmap_size = tuple(fmap_size)
        layers = []
        for i in range(num_layers):
            first = i == 0
            div = 2 if (downsample and not first) else 1
            layers.append(BottleBlock(dim=dim if first else dim_out,
                                      fmap_size=tuple(s // div for s in self.fmap_size),
                                      dim_out=dim_out, proj_factor=proj_factor, heads=heads, dim_head=dim_head,
                                      downsample=first and downsample, rel_pos_emb=rel_pos_emb,
                                      activation=activation))
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        _, c, h, w = x.shape
        assert c == self.dim, f"channels of feature map {c} must match channels given at init {self.dim}"
        assert (h, w) == self.fmap_size, \
            f"height and width ({h} {w}) of feature map must match the fmap_size given at init {self.fmap_size}"
        return self.net(x)
