"""Oracle (CPU, torch functional) for the network half of the hot path.  TEST INFRASTRUCTURE ONLY.

Restates the forward passes of the reference's generator / discriminator stacks as plain functions of a
reference-layout state_dict (citations into /root/reference/models/networks.py):

  global_generator   GlobalGenerator.forward        :301-357   (conv / transconv configuration)
  local_enhancer     LocalEnhancer.forward          :173-267   (n_local_enhancers = 1; n_attn_l >= 0: the attention sandwich of :218-237)
  resnet_block       ResnetBlock.forward            :421-463
  bottle_stack       bottleneck_transformer_pytorch==0.1.4 BottleStack (third party, PARITY UNPINNED:
                     see oracle/bottlestack_ref.py), eval or train-mode BatchNorm
  multiscale_d       MultiscaleDiscriminator.forward :507-550 + NLayerDiscriminator :641-692

Pinned against tests/golden/nets_golden.npz (outputs of the reference itself on seeded weights and inputs,
tests/golden/make_golden_nets.py) by tests/test_oracle_nets.py.
"""
import torch
import torch.nn.functional as F


def _in(x, eps=1e-5):
    return F.instance_norm(x, eps=eps)


def _conv(sd, key, x, stride=1, padding=0, reflect=0):
    if reflect:
        x = F.pad(x, (reflect,) * 4, mode="reflect")
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=padding)


def _convt(sd, key, x):
    return F.conv_transpose2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=2, padding=1, output_padding=1)


def conv_res_block(sd, key, x):
    """ConvResBlock.forward (networks.py:403-417) as the stride-2 downsampling layer (kernel 3, stride 2, padding 1)."""
    x = _conv(sd, key + ".conv1", x, stride=2, padding=1)
    return _conv(sd, key + ".conv2", x, padding=2) + _conv(sd, key + ".conv_res", x, padding=1)


def interpolate_upsample(sd, key, x):
    """InterpolateUpsample.forward (networks.py:394-400)."""
    x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    return _conv(sd, key + ".conv2", _conv(sd, key + ".conv1", x, padding=1), padding=2) + _conv(sd, key + ".conv_res", x, padding=1)


def _down(sd, key, x, kind):
    return conv_res_block(sd, key, x) if kind == "resconv" else _conv(sd, key, x, stride=2, padding=1)


def _up(sd, key, x, kind):
    return interpolate_upsample(sd, key, x) if kind == "interpolate" else _convt(sd, key, x)


def resnet_block(sd, prefix, x):
    h = F.relu(_in(_conv(sd, prefix + ".conv_block.1", x, reflect=1)))
    return x + _in(_conv(sd, prefix + ".conv_block.5", h, reflect=1))


def _bn(sd, key, x, training):
    if training:
        return F.batch_norm(x, None, None, sd[key + ".weight"], sd[key + ".bias"], True, 0.1, 1e-5)
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"], False, 0.1, 1e-5)


def attention(sd, prefix, x, heads, dim_head):
    b, _, hh, ww = x.shape
    q, k, v = F.conv2d(x, sd[prefix + ".to_qkv.weight"]).chunk(3, dim=1)

    def split(t):
        return t.reshape(b, heads, dim_head, hh * ww).transpose(-1, -2)

    q, k, v = split(q) * dim_head ** -0.5, split(k), split(v)
    emb = (sd[prefix + ".pos_emb.height"][:, None, :] + sd[prefix + ".pos_emb.width"][None, :, :]).reshape(-1, dim_head)
    sim = q @ k.transpose(-1, -2) + q @ emb.t()
    out = sim.softmax(dim=-1) @ v
    return out.transpose(-1, -2).reshape(b, heads * dim_head, hh, ww)


def bottle_stack(sd, prefix, x, num_layers, heads, dim_head, training=False):
    for i in range(num_layers):
        p = f"{prefix}.net.{i}.net"
        h = F.relu(_bn(sd, p + ".1", F.conv2d(x, sd[p + ".0.weight"]), training))
        h = attention(sd, p + ".3", h, heads, dim_head)
        h = F.relu(_bn(sd, p + ".5", h, training))
        h = _bn(sd, p + ".8", F.conv2d(h, sd[p + ".7.weight"]), training)
        sc = f"{prefix}.net.{i}.shortcut"
        if sc + ".0.weight" in sd:       # projection shortcut (dim != dim_out): conv1x1 -> BN -> activation
            x = F.relu(_bn(sd, sc + ".1", F.conv2d(x, sd[sc + ".0.weight"]), training))
        x = F.relu(h + x)
    return x


def _global_trunk(sd, prefix, x, n_down, n_blocks, n_attn=0, heads=4, dim_head=128, training=False, with_head=True, down="conv",
                  up="transconv"):
    i = 1
    x = F.relu(_in(_conv(sd, f"{prefix}.{i}", x, reflect=3)))
    i = 4
    for _ in range(n_down):
        x = F.relu(_in(_down(sd, f"{prefix}.{i}", x, down)))
        i += 3
    for blk in range(n_blocks + (1 if n_attn else 0)):
        if n_attn and blk == n_blocks // 2:
            x = bottle_stack(sd, f"{prefix}.{i}", x, n_attn, heads, dim_head, training)
        else:
            x = resnet_block(sd, f"{prefix}.{i}", x)
        i += 1
    for _ in range(n_down):
        x = F.relu(_in(_up(sd, f"{prefix}.{i}", x, up)))
        i += 3
    if with_head:
        x = torch.tanh(_conv(sd, f"{prefix}.{i + 1}", x, reflect=3))
    return x


def global_generator(sd, x, n_down=3, n_blocks=9, n_attn=0, heads=4, dim_head=128, training=False, down="conv", up="transconv"):
    return _global_trunk(sd, "model", x, n_down, n_blocks, n_attn, heads, dim_head, training, down=down, up=up)


def avgpool(x):
    return F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False)


def local_enhancer(sd, x, n_down=3, n_blocks_global=9, n_blocks_local=3, n_attn=0, heads=4, dim_head=128, training=False, down="conv",
                   up="transconv", n_attn_l=0, heads_l=4, dim_head_l=128):
    coarse = _global_trunk(sd, "model", avgpool(x), n_down, n_blocks_global, n_attn, heads, dim_head, training, with_head=False, down=down,
                           up=up)
    h = F.relu(_in(_conv(sd, "model1_1.1", x, reflect=3)))
    h = F.relu(_in(_down(sd, "model1_1.4", h, down)))
    h = h + coarse
    if n_attn_l == 0:
        for b in range(n_blocks_local):
            h = resnet_block(sd, f"model1_2.{b}", h)
        i = n_blocks_local
    else:
        # networks.py:218-237: [res x middle, down (8x: three stride-2 layers, the 2nd and 3rd share weights), BottleStack, res ...,
        # three applications of ONE up layer, up].  The state_dict lists the shared layers once per position.
        middle = n_blocks_local // 2
        for b in range(middle):
            h = resnet_block(sd, f"model1_2.{b}", h)
        for j in (0, 3, 3):        # positions 3 and 6 are ONE layer object: its first state_dict entry carries the (summed) gradient
            h = F.relu(_in(_down(sd, f"model1_2.{middle}.{j}", h, down)))
        h = bottle_stack(sd, f"model1_2.{middle + 1}", h, n_attn_l, heads_l, dim_head_l, training)
        for b in range(middle, n_blocks_local):
            h = resnet_block(sd, f"model1_2.{b + 2}", h)
        i = n_blocks_local + 2
        for _ in range(3):         # one layer object at positions i, i + 3, i + 6
            h = F.relu(_in(_up(sd, f"model1_2.{i}", h, up)))
        i += 9
    h = F.relu(_in(_up(sd, f"model1_2.{i}", h, up)))
    return torch.tanh(_conv(sd, f"model1_2.{i + 4}", h, reflect=3))


def multiscale_d(sd, x, num_D=3, n_layers=3, interm=True, sigmoid=False):
    """getIntermFeat=True layout (default): list[num_D] of list[n_layers+2] feature maps, the sigmoid stage of --no_lsgan never applied
    (networks.py:686).  interm=False (--no_ganFeat_loss): list[num_D] of [output], keys `layer{s}.{index in the flat Sequential}`, the
    sigmoid applied when the discriminator was built with it (networks.py:671-672)."""
    result = []
    for i in range(num_D):
        s = num_D - 1 - i
        feats, h = [], x
        flat = 0
        for j in range(n_layers + 2):
            key = f"scale{s}_layer{j}.0" if interm else f"layer{s}.{flat}"
            flat += 2 if (j == 0 or j == n_layers + 1) else 3          # [conv, lrelu] | [conv, norm, lrelu] | [conv]
            stride = 2 if j < n_layers else 1
            h = _conv(sd, key, h, stride=stride, padding=2)
            if 0 < j < n_layers + 1:
                h = _in(h)
            if j < n_layers + 1:
                h = F.leaky_relu(h, 0.2)
            feats.append(h)
        if not interm:
            feats = [torch.sigmoid(feats[-1]) if sigmoid else feats[-1]]
        result.append(feats)
        if i != num_D - 1:
            x = avgpool(x)
    return result
