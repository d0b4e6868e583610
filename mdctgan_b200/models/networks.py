"""Generator / discriminator stacks with the reference's constructors, module tree and state_dict keys
(models/networks.py), executed by the hand-written CUDA kernels of libmdctgan_b200.so.

The module tree exists to own parameters under the reference's names (`model.1.weight`,
`model.13.conv_block.1.bias`, `scale0_layer2.0.weight`, ...) so that the published `*_net_G.pth` files load
unchanged (models/base_model.py:43-111).  Leaf layers never compute: `forward` of a container walks its
children with `run_layers`, which fuses

    ReflectionPad2d -> Conv2d -> InstanceNorm2d -> ReLU

into: one convolution launch (reflection handled by its gather, (sum, sumsq) taken in its epilogue), one
tiny statistics->scale/shift launch, and nothing else -- the normalisation and the activation are applied
by whichever kernel reads the tensor next (nn_ops.Feat).

`module.forward(x)` is the inference path (no_grad, NCHW in / out).  Training runs the same layers through `run()` while an
`nn_ops.Tape` records them; the backward sweeps (dgrad / wgrad / norm backward) are driven by train_ops.GanGraph.
"""
from __future__ import annotations

import functools
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .. import nn_ops as ops
from ..nn_ops import Feat

_NO_DIRECT = "mdctgan_b200 leaf layers hold parameters only; call the enclosing network (no cuDNN / eager fallback)"


# ------------------------------------------------------------------------------------------- leaf layers
class _PackedWeight:
    """Caches the kernel-side layout of a parameter until the parameter changes."""

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, p: torch.Tensor, fn):
        key = (p.data_ptr(), p._version, getattr(p, "_version_bump", 0), p.device)   # _version_bump: FusedAdam.step (optim.py)
        if key != self._key:
            self._val = fn(p)
            self._key = key
        return self._val


class Conv2d(nn.Conv2d):
    """nn.Conv2d as a parameter holder (same init, same names); executed by run_layers."""

    def forward(self, x):  # pragma: no cover
        raise RuntimeError(_NO_DIRECT)

    def packed(self):
        sp = self.__dict__.get("_static_pack")
        if sp is not None:
            return sp["fwd"][0]
        if not hasattr(self, "_pk"):
            self._pk = _PackedWeight()
        return self._pk.get(self.weight, lambda w: ops.pack_conv_weight(w, False))

    def packed_umma(self):
        """tcgen05 kernel's weight image, or None for the shapes that stay on the direct kernels."""
        if not ops.umma_supported(self.in_channels, self.out_channels):
            return None
        sp = self.__dict__.get("_static_pack")
        if sp is not None:
            return sp["fwd"][1]
        if not hasattr(self, "_pku"):
            self._pku = _PackedWeight()
        return self._pku.get(self.weight, lambda w: ops.pack_conv_weight_umma(self.packed()))

    def run(self, f: Feat, pad_reflect: int = 0, act: int = ops.ACT_NONE, want_stats: bool = False) -> Feat:
        kh, kw = self.kernel_size
        assert self.stride[0] == self.stride[1] and self.padding[0] == self.padding[1] and self.dilation == (1, 1) and self.groups == 1
        pad, mode = (pad_reflect, ops.PAD_REFLECT) if pad_reflect else (self.padding[0], ops.PAD_ZERO)
        if pad_reflect and self.padding[0]:
            raise RuntimeError("ReflectionPad2d in front of a zero-padded convolution is not a reference configuration")
        return ops.conv2d(f, self.packed(), self.bias, kh=kh, kw=kw, stride=self.stride[0], pad=pad, pad_mode=mode, act=act,
                          want_stats=want_stats, w_umma=self.packed_umma(), owner=self)

    def packed_dgrad(self):
        """(weights of the input-gradient convolution [kh*kw*Cout][Cin], their tcgen05 image or None, flipped).
        stride 1: a plain convolution with the taps flipped; stride > 1: the transposed-geometry kernels."""
        flipped = self.stride[0] == 1
        sp = self.__dict__.get("_static_pack")
        if sp is not None and "dgrad" in sp:
            kn, um, fl = sp["dgrad"]
            return kn, (um if ops.CONV_ENGINE != "direct" else None), fl
        if not hasattr(self, "_pkd"):
            self._pkd, self._pkdu = _PackedWeight(), _PackedWeight()
        if flipped:
            wd = self._pkd.get(self.weight, lambda w: ops.pack_conv_weight(w.detach().permute(1, 0, 2, 3).flip(2, 3), False))
        else:
            wd = self._pkd.get(self.weight, lambda w: ops.pack_conv_weight(w, True))
        wu = None
        if ops.umma_supported(self.out_channels, self.in_channels):
            wu = self._pkdu.get(self.weight, lambda w: ops.pack_conv_weight_umma(wd))
        return wd, wu, flipped


class ConvTranspose2d(nn.ConvTranspose2d):
    def forward(self, x, output_size=None):  # pragma: no cover
        raise RuntimeError(_NO_DIRECT)

    def packed(self):
        sp = self.__dict__.get("_static_pack")
        if sp is not None:
            return sp["fwd"][0]
        if not hasattr(self, "_pk"):
            self._pk = _PackedWeight()
        return self._pk.get(self.weight, lambda w: ops.pack_conv_weight(w, True))

    def packed_umma(self):
        if not ops.umma_supported(self.in_channels, self.out_channels):
            return None
        sp = self.__dict__.get("_static_pack")
        if sp is not None:
            return sp["fwd"][1]
        if not hasattr(self, "_pku"):
            self._pku = _PackedWeight()
        return self._pku.get(self.weight, lambda w: ops.pack_conv_weight_umma(self.packed()))

    def run(self, f: Feat, pad_reflect: int = 0, act: int = ops.ACT_NONE, want_stats: bool = False) -> Feat:
        assert not pad_reflect
        kh, kw = self.kernel_size
        return ops.conv2d(f, self.packed(), self.bias, kh=kh, kw=kw, stride=self.stride[0], pad=self.padding[0], transposed=True,
                          output_padding=self.output_padding[0], act=act, want_stats=want_stats, w_umma=self.packed_umma(), owner=self)

    def packed_dgrad(self):
        """Input gradient of a ConvTranspose2d = plain strided convolution with the same weight tensor read as
        [Cout_c = in_channels][Cin_c = out_channels][kh][kw]."""
        sp = self.__dict__.get("_static_pack")
        if sp is not None and "dgrad" in sp:
            kn, um, fl = sp["dgrad"]
            return kn, (um if ops.CONV_ENGINE != "direct" else None), fl
        if not hasattr(self, "_pkd"):
            self._pkd, self._pkdu = _PackedWeight(), _PackedWeight()
        wd = self._pkd.get(self.weight, lambda w: ops.pack_conv_weight(w, False))
        wu = None
        if ops.umma_supported(self.out_channels, self.in_channels):
            wu = self._pkdu.get(self.weight, lambda w: ops.pack_conv_weight_umma(wd))
        return wd, wu, False


class InstanceNorm2d(nn.InstanceNorm2d):
    def forward(self, x):  # pragma: no cover
        raise RuntimeError(_NO_DIRECT)


class BatchNorm2d(nn.BatchNorm2d):
    def forward(self, x):  # pragma: no cover
        raise RuntimeError(_NO_DIRECT)


class ReflectionPad2d(nn.Module):
    def __init__(self, padding: int):
        super().__init__()
        self.padding = int(padding)

    def extra_repr(self):
        return str((self.padding,) * 4)


class ReLU(nn.Module):
    def __init__(self, inplace: bool = False):
        super().__init__()
        self.inplace = inplace


class LeakyReLU(nn.Module):
    def __init__(self, negative_slope: float = 0.2, inplace: bool = False):
        super().__init__()
        if abs(negative_slope - 0.2) > 1e-12:
            raise NotImplementedError("LeakyReLU slope is 0.2 everywhere in the reference (networks.py:650)")
        self.negative_slope, self.inplace = negative_slope, inplace


class Tanh(nn.Module):
    pass


class Sigmoid(nn.Module):
    """The --no_lsgan PatchGAN head (networks.py:671-672)."""


class AvgPool3s2(nn.Module):
    """nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False) (networks.py:249-250, :525-526)."""


_ACT_CODE = {ReLU: ops.ACT_RELU, LeakyReLU: ops.ACT_LEAKY, Tanh: ops.ACT_TANH, Sigmoid: ops.ACT_SIGMOID}


def _act_of(layer) -> Optional[int]:
    return _ACT_CODE.get(type(layer))


# ------------------------------------------------------------------------------------------- executor
def run_layers(layers: Sequence[nn.Module], f: Feat) -> Feat:
    """Execute a reference-ordered layer list on a Feat, fusing pad / norm / activation into the convolutions."""
    i, n = 0, len(layers)
    pad_reflect = 0
    while i < n:
        L = layers[i]
        nxt = layers[i + 1] if i + 1 < n else None
        if ops._marks:
            ops.mark(L)
        if isinstance(L, ReflectionPad2d):
            pad_reflect = L.padding
            i += 1
        elif isinstance(L, (Conv2d, ConvTranspose2d, ConvResBlock, InterpolateUpsample)):
            want_stats = isinstance(nxt, (InstanceNorm2d, BatchNorm2d))
            act, skip = ops.ACT_NONE, 0
            if not want_stats and nxt is not None and _act_of(nxt) is not None:
                act, skip = _act_of(nxt), 1          # conv -> activation with no norm in between: epilogue
            f = L.run(f, pad_reflect=pad_reflect, act=act, want_stats=want_stats)
            pad_reflect = 0
            i += 1 + skip
        elif isinstance(L, InstanceNorm2d):
            if L.affine or L.track_running_stats:
                raise NotImplementedError("InstanceNorm2d(affine=False, track_running_stats=False) only (networks.py:26)")
            f = ops.finalize_norm(f, eps=L.eps, mode=0)
            i += 1
        elif isinstance(L, BatchNorm2d):
            mode = 1 if (L.training or not L.track_running_stats) else 2
            f = ops.finalize_norm(f, eps=L.eps, mode=mode, gamma=L.weight, beta=L.bias, running_mean=L.running_mean,
                                  running_var=L.running_var, momentum=L.momentum if L.momentum is not None else 0.1)
            if mode == 1 and L.track_running_stats and L.num_batches_tracked is not None:
                L.num_batches_tracked += 1
            i += 1
        elif _act_of(L) is not None:
            f = ops.with_act(f, _act_of(L))
            i += 1
        elif isinstance(L, AvgPool3s2):
            f = ops.avgpool3s2(f)
            i += 1
        elif hasattr(L, "run"):
            if pad_reflect:
                raise RuntimeError("ReflectionPad2d must be followed by a convolution")
            f = L.run(f)
            i += 1
        elif isinstance(L, nn.Sequential):
            f = run_layers(list(L), f)
            i += 1
        else:
            raise NotImplementedError(f"run_layers: no kernel mapping for {type(L).__name__}")
    return f


def _forward_nchw(layers, x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError(f"expected a CUDA tensor, got device={x.device}; mdctgan_b200 has no CPU path")
    with torch.no_grad(), ops.stats_pass(x.device):
        return ops.to_nchw(run_layers(layers, ops.to_nhwc(x.to(torch.float32).contiguous())))


# ------------------------------------------------------------------------------------------- init helpers
def weights_init(m):
    """Same rule as the reference (networks.py:13-19): N(0, 0.02) on every class whose name contains
    'Conv2d' (so NOT ConvTranspose2d), N(1, 0.02) / 0 on BatchNorm2d."""
    classname = m.__class__.__name__
    if classname.find("Conv2d") != -1:
        m.weight.data.normal_(0.0, 0.02)
    elif classname.find("BatchNorm2d") != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def get_norm_layer(norm_type="instance"):
    if norm_type == "batch":
        return functools.partial(BatchNorm2d, affine=True)
    if norm_type == "instance":
        return functools.partial(InstanceNorm2d, affine=False)
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


# ------------------------------------------------------------------------------------------- blocks
class ResnetBlock(nn.Module):
    """x + [ReflPad1, Conv3x3, IN, ReLU, ReflPad1, Conv3x3, IN](x)   (networks.py:421-463)."""

    def __init__(self, dim, padding_type, norm_layer, activation=None, use_dropout=False):
        super().__init__()
        if padding_type != "reflect":
            raise NotImplementedError("ResnetBlock: padding_type 'reflect' is the only one the reference constructs (networks.py:174,302)")
        if use_dropout:
            raise NotImplementedError("ResnetBlock: use_dropout is never enabled by the reference")
        activation = activation if activation is not None else ReLU(True)
        self.conv_block = nn.Sequential(ReflectionPad2d(1), Conv2d(dim, dim, kernel_size=3, padding=0), norm_layer(dim), activation,
                                        ReflectionPad2d(1), Conv2d(dim, dim, kernel_size=3, padding=0), norm_layer(dim))

    def run(self, f: Feat) -> Feat:
        return ops.combine(f, run_layers(list(self.conv_block), f))

    def forward(self, x):
        return _forward_nchw([self], x)


class ConvResBlock(nn.Module):
    """downsample_type='resconv' (networks.py:403-417): conv1 (k, s, p; C->C) then conv2 5x5 p2 + conv_res 3x3 p1."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding):
        super().__init__()
        self.conv1 = Conv2d(in_channels, in_channels, kernel_size, stride, padding)
        self.conv2 = Conv2d(in_channels, out_channels, 5, padding=2)
        self.conv_res = Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)

    def run(self, f: Feat, pad_reflect: int = 0, act: int = ops.ACT_NONE, want_stats: bool = False) -> Feat:
        assert not pad_reflect
        x = self.conv1.run(f)
        y = ops.combine(self.conv2.run(x), self.conv_res.run(x), act_out=act, want_stats=want_stats)
        return y


class InterpolateUpsample(nn.Module):
    """upsample_type='interpolate' (networks.py:375-400): nearest x2, then conv1 5x5 p1 -> conv2 3x3 p2, plus conv_res 3x3 p1."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = kwargs["in_channels"], kwargs["out_channels"]
        self.conv1 = Conv2d(self.in_channels, self.out_channels, 5, padding=1)
        self.conv2 = Conv2d(self.out_channels, self.out_channels, 3, padding=2)
        self.conv_res = Conv2d(self.in_channels, self.out_channels, 3, padding=1)

    def run(self, f: Feat, pad_reflect: int = 0, act: int = ops.ACT_NONE, want_stats: bool = False) -> Feat:
        assert not pad_reflect
        assert f.x.shape[-1] == self.in_channels
        x = ops.upsample_nearest2x(f)
        res = self.conv_res.run(x)
        y = self.conv2.run(self.conv1.run(x))
        return ops.combine(y, res, act_out=act, want_stats=want_stats)


# ------------------------------------------------------------------------------------------- generators
def _sampling_layers(downsample_type, upsample_type):
    """networks.py:189-205, :311-325"""
    if downsample_type == "conv":
        down = Conv2d
    elif downsample_type == "resconv":
        down = ConvResBlock
    else:
        raise NotImplementedError("downsample layer [{:s}] is not found".format(downsample_type))
    if upsample_type == "transconv":
        up = ConvTranspose2d
    elif upsample_type == "interpolate":
        up = InterpolateUpsample
    else:
        raise NotImplementedError("upsample layer [{:s}] is not found".format(upsample_type))
    return down, up


class GlobalGenerator(nn.Module):
    def __init__(self, input_nc, output_nc, ngf=64, n_downsampling=3, n_blocks=9, norm_layer=None, padding_type="reflect",
                 upsample_type="transconv", downsample_type="conv", n_attn_g=0, input_size=(128, 256), proj_factor_g=4, heads_g=4,
                 dim_head_g=128):
        assert n_blocks >= 0
        super().__init__()
        norm_layer = norm_layer if norm_layer is not None else functools.partial(BatchNorm2d, affine=True)
        activation = ReLU(True)   # ONE shared instance, as in the reference (keeps the positional state_dict keys)
        downsample_layer, upsample_layer = _sampling_layers(downsample_type, upsample_type)
        model: List[nn.Module] = [ReflectionPad2d(3), Conv2d(input_nc, ngf, kernel_size=7, padding=0), norm_layer(ngf), activation]
        for i in range(n_downsampling):
            mult = 2 ** i
            model += [downsample_layer(ngf * mult, ngf * mult * 2, kernel_size=3, stride=2, padding=1), norm_layer(ngf * mult * 2), activation]
        mult = 2 ** n_downsampling
        bottle_neck: List[nn.Module] = [ResnetBlock(ngf * mult, padding_type=padding_type, activation=activation, norm_layer=norm_layer)
                                        for _ in range(n_blocks)]
        if n_attn_g > 0:
            from .bottleneck import BottleStack

            fmap = tuple(s // mult for s in input_size)
            bottle_neck.insert(n_blocks // 2, BottleStack(dim=ngf * mult, fmap_size=fmap, dim_out=ngf * mult, num_layers=n_attn_g,
                                                          proj_factor=proj_factor_g, downsample=False, heads=heads_g,
                                                          dim_head=dim_head_g, activation=activation, rel_pos_emb=False))
        model += bottle_neck
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            model += [upsample_layer(in_channels=ngf * mult, out_channels=int(ngf * mult / 2), kernel_size=3, stride=2, padding=1,
                                     output_padding=1), norm_layer(int(ngf * mult / 2)), activation]
        model += [ReflectionPad2d(3), Conv2d(ngf, output_nc, kernel_size=7, padding=0), Tanh()]
        self.model = nn.Sequential(*model)
        self.freeze = False

    def run(self, f: Feat) -> Feat:
        return run_layers(list(self.model), f)

    def forward(self, input):
        return _forward_nchw(list(self.model), input)

    def set_freeze(self, freeze=True, *unused):
        """Reference: networks.py:359-372.  Accepts (and ignores) the extra positional arguments that
        Pix2PixHDModel.initialize passes (pix2pixHD_model.py:241-242) -- the reference itself raises there."""
        if self.freeze == freeze:
            return
        self.freeze = freeze
        for name, layer in self.model.named_children():
            module_name = layer.__class__.__name__
            if "ResnetBlock" in module_name or "BottleStack" in module_name:
                break
            for param in layer.parameters():
                param.requires_grad = not freeze


class LocalEnhancer(nn.Module):
    def __init__(self, input_nc, output_nc, ngf=32, n_downsample_global=3, n_blocks_global=9, n_local_enhancers=1, n_blocks_local=3,
                 norm_layer=None, padding_type="reflect", downsample_type="conv", upsample_type="transconv", n_attn_g=0, n_attn_l=0,
                 input_size=(128, 256), proj_factor_g=4, heads_g=4, dim_head_g=128, proj_factor_l=4, heads_l=4, dim_head_l=128):
        super().__init__()
        norm_layer = norm_layer if norm_layer is not None else functools.partial(BatchNorm2d, affine=True)
        self.n_local_enhancers = n_local_enhancers
        downsample_layer, upsample_layer = _sampling_layers(downsample_type, upsample_type)
        ngf_global = ngf * (2 ** n_local_enhancers)
        g = GlobalGenerator(input_nc, output_nc, ngf_global, n_downsample_global, n_blocks_global, norm_layer,
                            downsample_type=downsample_type, upsample_type=upsample_type, input_size=tuple(s // 2 for s in input_size),
                            n_attn_g=n_attn_g, proj_factor_g=proj_factor_g, heads_g=heads_g, dim_head_g=dim_head_g).model
        self.model = nn.Sequential(*[g[i] for i in range(len(g) - 3)])       # drop ReflPad, Conv7x7, Tanh
        ngf_global = ngf * (2 ** (n_local_enhancers - 1))
        model_downsample = [ReflectionPad2d(3), Conv2d(input_nc, ngf_global, kernel_size=7, padding=0), norm_layer(ngf_global), ReLU(True),
                            downsample_layer(ngf_global, ngf_global * 2, kernel_size=3, stride=2, padding=1), norm_layer(ngf_global * 2), ReLU(True)]
        model_upsample: List[nn.Module] = [ResnetBlock(ngf_global * 2, padding_type=padding_type, norm_layer=norm_layer)
                                           for _ in range(n_blocks_local)]
        if n_attn_l > 0:
            # attention bottleneck in the local branch (networks.py:218-237): 8x down (the second / third step and the three up steps are
            # the SAME layer objects applied repeatedly -- the reference builds them with `[...] * 2` / `[...] * 3`, so the weights are
            # shared and their state_dict entries appear once per position), a BottleStack whose first block has a projection shortcut
            # (dim ngf -> 2 ngf), 8x up.  Construction order = the reference's (the position embeddings are drawn at construction).
            from .bottleneck import BottleStack

            middle = n_blocks_local // 2
            down = [downsample_layer(ngf_global * 2, ngf_global, kernel_size=3, stride=2, padding=1), norm_layer(ngf_global), ReLU(True)]
            down += [downsample_layer(ngf_global, ngf_global, kernel_size=3, stride=2, padding=1), norm_layer(ngf_global), ReLU(True)] * 2
            model_upsample.insert(middle, nn.Sequential(*down))
            attn_block = BottleStack(dim=ngf_global, fmap_size=tuple(s // 16 for s in input_size), dim_out=ngf_global * 2, num_layers=n_attn_l,
                                     proj_factor=proj_factor_l, downsample=False, heads=heads_l, dim_head=dim_head_l, activation=ReLU(True),
                                     rel_pos_emb=False)
            model_upsample.insert(middle + 1, attn_block)
            model_upsample += [upsample_layer(in_channels=ngf_global * 2, out_channels=ngf_global * 2, kernel_size=3, stride=2, padding=1,
                                              output_padding=1), norm_layer(ngf_global), ReLU(True)] * 3
        model_upsample += [upsample_layer(in_channels=ngf_global * 2, out_channels=ngf_global, kernel_size=3, stride=2, padding=1,
                                          output_padding=1), norm_layer(ngf_global), ReLU(True)]
        model_upsample += [ReflectionPad2d(3), Conv2d(ngf, output_nc, kernel_size=7, padding=0), Tanh()]
        self.model1_1 = nn.Sequential(*model_downsample)
        self.model1_2 = nn.Sequential(*model_upsample)
        self.downsample = AvgPool3s2()
        self.freeze = False

    def run(self, f: Feat) -> Feat:
        pyramid = [f]
        for _ in range(self.n_local_enhancers):
            pyramid.append(ops.avgpool3s2(pyramid[-1]))
        # the global trunk (on the pooled input) and the local down-sampling branch are independent until their sum: two streams,
        # forward and backward (nn_ops.run_branches)
        coarse, fine = ops.run_branches([lambda: run_layers(list(self.model), pyramid[-1]),
                                         lambda: run_layers(list(self.model1_1), pyramid[0])], f.x.device)
        return run_layers(list(self.model1_2), ops.combine(fine, coarse))   # only one enhancer level runs (networks.py:260-267)

    def forward(self, input):
        return _forward_nchw([self], input)

    def set_freeze(self, freeze_global_d=True, freeze_global_u=False, freeze_local_d=True, freeze_local_u=False):
        for name, layer in self.model.named_children():
            module_name = layer.__class__.__name__
            if "Conv2d" in module_name or "ConvResBlock" in module_name:
                for param in layer.parameters():
                    param.requires_grad = not freeze_global_d
            elif any(k in module_name for k in ("InterpolateUpsample", "ConvTranspose2d", "ResnetBlock", "BottleStack")):
                for param in layer.parameters():
                    param.requires_grad = not freeze_global_u
        for param in self.model1_1.parameters():
            param.requires_grad = not freeze_local_d
        for param in self.model1_2.parameters():
            param.requires_grad = not freeze_local_u


def define_G(input_nc, output_nc, ngf, netG, n_downsample_global=3, n_blocks_global=9, n_local_enhancers=1, n_blocks_local=3,
             norm="instance", gpu_ids=[], upsample_type="transconv", downsample_type="conv", input_size=(128, 256), n_attn_g=0,
             n_attn_l=0, proj_factor_g=4, heads_g=4, dim_head_g=128, proj_factor_l=4, heads_l=4, dim_head_l=128):
    """Reference: networks.py:33-56 (same signature; the module is not printed)."""
    norm_layer = get_norm_layer(norm_type=norm)
    if netG == "global":
        net = GlobalGenerator(input_nc, output_nc, ngf, n_downsample_global, n_blocks_global, norm_layer, downsample_type=downsample_type,
                              upsample_type=upsample_type, input_size=input_size, n_attn_g=n_attn_g, proj_factor_g=proj_factor_g,
                              heads_g=heads_g, dim_head_g=dim_head_g)
    elif netG == "local":
        net = LocalEnhancer(input_nc, output_nc, ngf, n_downsample_global, n_blocks_global, n_local_enhancers, n_blocks_local, norm_layer,
                            downsample_type=downsample_type, upsample_type=upsample_type, input_size=input_size, n_attn_g=n_attn_g,
                            proj_factor_g=proj_factor_g, heads_g=heads_g, dim_head_g=dim_head_g, n_attn_l=n_attn_l,
                            proj_factor_l=proj_factor_l, heads_l=heads_l, dim_head_l=dim_head_l)
    else:
        raise NotImplementedError("generator [%s] is not implemented (reference: 'global' | 'local'; 'encoder' is dead code)" % netG)
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.cuda(gpu_ids[0])
    net.apply(weights_init)
    return net


# ------------------------------------------------------------------------------------------- discriminator
class NLayerDiscriminator(nn.Module):
    """PatchGAN (networks.py:641-692): Conv4x4 s2 p2 + LReLU; (n_layers-1) x [Conv4x4 s2 p2, IN, LReLU];
    Conv4x4 s1 p2, IN, LReLU; Conv4x4 s1 p2 -> 1."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=None, use_sigmoid=False, getIntermFeat=False):
        super().__init__()
        norm_layer = norm_layer if norm_layer is not None else functools.partial(BatchNorm2d, affine=True)
        self.getIntermFeat, self.n_layers = getIntermFeat, n_layers
        kw, padw = 4, 2
        sequence = [[Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), LeakyReLU(0.2, True)]]
        nf = ndf
        for n in range(1, n_layers):
            nf_prev, nf = nf, min(nf * 2, 512)
            sequence += [[Conv2d(nf_prev, nf, kernel_size=kw, stride=2, padding=padw), norm_layer(nf), LeakyReLU(0.2, True)]]
        nf_prev, nf = nf, min(nf * 2, 512)
        sequence += [[Conv2d(nf_prev, nf, kernel_size=kw, stride=1, padding=padw), norm_layer(nf), LeakyReLU(0.2, True)]]
        sequence += [[Conv2d(nf, 1, kernel_size=kw, stride=1, padding=padw)]]
        if use_sigmoid:
            # like the reference (networks.py:671-672): with getIntermFeat the forward walks models 0 .. n_layers + 1 only (:686), so the
            # sigmoid stage is constructed but never applied -- BCELoss then needs --no_ganFeat_loss (getIntermFeat False) to see probabilities
            sequence += [[Sigmoid()]]
        if getIntermFeat:
            for n in range(len(sequence)):
                setattr(self, "model" + str(n), nn.Sequential(*sequence[n]))
        else:
            self.model = nn.Sequential(*[m for s in sequence for m in s])

    def stages(self):
        if self.getIntermFeat:
            return [getattr(self, "model" + str(n)) for n in range(self.n_layers + 2)]
        return [self.model]

    def run(self, f: Feat) -> List[Feat]:
        outs = []
        for st in self.stages():
            f = run_layers(list(st), f)
            outs.append(f)
        return outs if self.getIntermFeat else outs[-1:]

    def forward(self, input):
        with torch.no_grad(), ops.stats_pass(input.device):
            outs = self.run(ops.to_nhwc(input.to(torch.float32).contiguous()))
            res = [ops.to_nchw(o) for o in outs]
        return res if self.getIntermFeat else res[0]


class MultiscaleDiscriminator(nn.Module):
    """num_D PatchGANs on an AvgPool pyramid (networks.py:507-550); returns list[num_D] of list[n_layers+2]."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=None, use_sigmoid=False, num_D=3, getIntermFeat=False):
        super().__init__()
        self.num_D, self.n_layers, self.getIntermFeat = num_D, n_layers, getIntermFeat
        for i in range(num_D):
            netD = NLayerDiscriminator(input_nc, ndf, n_layers, norm_layer, use_sigmoid, getIntermFeat)
            if getIntermFeat:
                for j in range(n_layers + 2):
                    setattr(self, "scale" + str(i) + "_layer" + str(j), getattr(netD, "model" + str(j)))
            else:
                setattr(self, "layer" + str(i), netD.model)
        self.downsample = AvgPool3s2()

    def _stages(self, i):
        if self.getIntermFeat:
            return [getattr(self, "scale" + str(i) + "_layer" + str(j)) for j in range(self.n_layers + 2)]
        return [getattr(self, "layer" + str(i))]

    def run_features(self, f: Feat):
        """Forward on a Feat (NHWC), as `forward` but staying on the device layout: list[num_D] of list[n_layers + 2]
        Feats; the intermediate features are materialised (IN + LeakyReLU applied) because the feature-matching
        loss reads them (pix2pixHD_model.py:447-451); the next stage still consumes the deferred view."""
        pyramid = [f]
        for i in range(self.num_D - 1):
            pyramid.append(ops.avgpool3s2(pyramid[-1]))

        def scale(i):
            def run():
                outs, g = [], pyramid[i]
                for st in self._stages(self.num_D - 1 - i):
                    g = run_layers(list(st), g)
                    outs.append(ops.materialize(g))
                return outs
            return run

        # the scales are independent given the pyramid: one CUDA stream each (forward and both backward sweeps)
        return ops.run_branches([scale(i) for i in range(self.num_D)], f.x.device)

    def forward(self, input):
        with torch.no_grad(), ops.stats_pass(input.device):
            f = ops.to_nhwc(input.to(torch.float32).contiguous())
            result = []
            for i in range(self.num_D):
                outs, g = [], f
                for st in self._stages(self.num_D - 1 - i):
                    g = run_layers(list(st), g)
                    outs.append(ops.to_nchw(g))
                result.append(outs if self.getIntermFeat else [outs[-1]])
                if i != self.num_D - 1:
                    f = ops.avgpool3s2(f)
        return result


def define_D(input_nc, ndf, n_layers_D, norm="instance", use_sigmoid=False, num_D=1, getIntermFeat=False, gpu_ids=[]):
    """Reference: networks.py:59-68."""
    netD = MultiscaleDiscriminator(input_nc, ndf, n_layers_D, get_norm_layer(norm_type=norm), use_sigmoid, num_D, getIntermFeat)
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        netD.cuda(gpu_ids[0])
    netD.apply(weights_init)
    return netD
