"""BoTNet `BottleStack` as the reference uses it (third-party bottleneck_transformer_pytorch==0.1.4, called at
models/networks.py:232-235 and :341-344 with downsample=False, rel_pos_emb=False, activation=ReLU): the module
tree owns parameters under the package's state_dict names (`net.{i}.net.{0,1,3,5,7,8}`, `...3.to_qkv.weight`,
`...3.pos_emb.{height,width}`); execution is by the kernels of libmdctgan_b200.so.

PARITY UNPINNED: the package is not vendored in the reference tree; the arithmetic follows its published
semantics as restated in oracle/bottlestack_ref.py."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import nn_ops as ops
from ..nn_ops import Feat


class AbsPosEmb(nn.Module):
    def __init__(self, fmap_size, dim_head):
        super().__init__()
        h, w = fmap_size
        s = dim_head ** -0.5
        self.height = nn.Parameter(torch.randn(h, dim_head) * s)
        self.width = nn.Parameter(torch.randn(w, dim_head) * s)


class Attention(nn.Module):
    def __init__(self, *, dim, fmap_size, heads=4, dim_head=128, rel_pos_emb=False):
        super().__init__()
        from .networks import Conv2d

        if rel_pos_emb:
            raise NotImplementedError("rel_pos_emb=True is never used by the reference (networks.py:235,344)")
        self.heads, self.dim_head = heads, dim_head
        self.scale = dim_head ** -0.5
        self.to_qkv = Conv2d(dim, heads * dim_head * 3, 1, bias=False)
        self.pos_emb = AbsPosEmb(fmap_size, dim_head)

    def run(self, f: Feat) -> Feat:
        qkv = self.to_qkv.run(f)                       # [B, H, W, 3*heads*d], plain
        return ops.attention(qkv, self.pos_emb.height, self.pos_emb.width, self.heads, self.dim_head, self.scale)


class BottleBlock(nn.Module):
    def __init__(self, *, dim, fmap_size, dim_out, proj_factor, downsample, heads=4, dim_head=128, rel_pos_emb=False, activation=None):
        super().__init__()
        from .networks import BatchNorm2d, Conv2d, ReLU

        if downsample:
            raise NotImplementedError("BottleBlock(downsample=True) is never built by the reference (networks.py:235,344)")
        activation = activation if activation is not None else ReLU()
        if dim != dim_out:       # projection shortcut (the local-branch stack of networks.py:232-235: dim ngf -> dim_out 2 ngf)
            self.shortcut = nn.Sequential(Conv2d(dim, dim_out, 1, stride=1, padding=0, bias=False), BatchNorm2d(dim_out), activation)
        else:
            self.shortcut = nn.Identity()
        inner_in, inner_out = dim_out // proj_factor, heads * dim_head
        self.net = nn.Sequential(
            Conv2d(dim, inner_in, 1, bias=False), BatchNorm2d(inner_in), activation,
            Attention(dim=inner_in, fmap_size=fmap_size, heads=heads, dim_head=dim_head, rel_pos_emb=rel_pos_emb),
            nn.Identity(), BatchNorm2d(inner_out), activation,
            Conv2d(inner_out, dim_out, 1, bias=False), BatchNorm2d(dim_out))
        nn.init.zeros_(self.net[-1].weight)            # library zero-gamma; define_G's weights_init overrides it
        self.activation = activation

    def run(self, f: Feat) -> Feat:
        from .networks import run_layers

        n = self.net
        h = run_layers([n[0], n[1], n[2]], f)           # conv1x1 -> BN -> ReLU (pending)
        h = n[3].run(h)                                  # attention (plain) + statistics for the BN that follows
        h = run_layers([n[5], n[6], n[7], n[8]], h)      # BN -> ReLU -> conv1x1 -> BN (pending)
        sc = f if isinstance(self.shortcut, nn.Identity) else run_layers(list(self.shortcut), f)      # conv1x1 -> BN -> ReLU (pending)
        return ops.combine(h, sc, act_out=ops.ACT_RELU)  # relu(net(x) + shortcut(x))


class BottleStack(nn.Module):
    def __init__(self, *, dim, fmap_size, dim_out=2048, proj_factor=4, num_layers=3, heads=4, dim_head=128, downsample=True,
                 rel_pos_emb=False, activation=None):
        super().__init__()
        if isinstance(fmap_size, int):
            fmap_size = (fmap_size, fmap_size)
        self.dim, self.fmap_size = dim, tuple(fmap_size)
        if downsample:
            raise NotImplementedError("BottleStack(downsample=True) is never used by the reference")
        self.net = nn.Sequential(*[BottleBlock(dim=dim if i == 0 else dim_out, fmap_size=self.fmap_size, dim_out=dim_out,
                                               proj_factor=proj_factor, heads=heads, dim_head=dim_head, downsample=False,
                                               rel_pos_emb=rel_pos_emb, activation=activation) for i in range(num_layers)])

    def run(self, f: Feat) -> Feat:
        B, H, W, C = f.x.shape
        assert C == self.dim, f"channels of feature map {C} must match channels given at init {self.dim}"
        assert (H, W) == self.fmap_size, f"height and width ({H} {W}) of feature map must match the fmap_size given at init {self.fmap_size}"
        for blk in self.net:
            f = blk.run(f)
        return f
