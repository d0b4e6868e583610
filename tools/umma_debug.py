"""Debug driver for the tcgen05 convolution: one case per process (a trapped kernel poisons the context)."""
import subprocess
import sys

CASES = {
    "tiny1x1": (1, 32, 8, 16, 32, 1, 1, 0, False, False),      # M=128, K=32 (1 chunk), BN=32, no split
    "k128": (1, 128, 8, 16, 64, 1, 1, 0, False, False),        # 4 chunks, splits up to 4
    "res": (4, 256, 4, 32, 256, 3, 1, 1, True, False),
    "span": (8, 512, 2, 16, 512, 3, 1, 1, True, False),
    "convT": (2, 64, 8, 16, 32, 3, 2, 1, False, True),
    "big": (5, 32, 32, 256, 32, 3, 1, 1, True, False),
}


def run(name, engine):
    import torch
    import torch.nn.functional as F
    from mdctgan_b200 import nn_ops as ops
    ops.CONV_ENGINE = engine
    B, Cin, H, W, Cout, k, stride, pad, reflect, transposed = CASES[name]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn((Cin, Cout, k, k) if transposed else (Cout, Cin, k, k), generator=g) * 0.1
    bias = torch.randn(Cout, generator=g)
    if transposed:
        ref = F.conv_transpose2d(x, w, bias, stride=stride, padding=pad, output_padding=1)
    elif reflect:
        ref = F.conv2d(F.pad(x, (pad,) * 4, mode="reflect"), w, bias, stride=stride)
    else:
        ref = F.conv2d(x, w, bias, stride=stride, padding=pad)
    dev = torch.device("cuda:0")
    w_kn = ops.pack_conv_weight(w.to(dev), transposed)
    w_um = ops.pack_conv_weight_umma(w_kn)
    f = ops.Feat(x.permute(0, 2, 3, 1).contiguous().to(dev))
    y = ops.conv2d(f, w_kn, bias.to(dev), kh=k, kw=k, stride=stride, pad=pad, pad_mode=1 if reflect else 0, transposed=transposed,
                   output_padding=1 if transposed else 0, want_stats=True, w_umma=w_um)
    torch.cuda.synchronize()
    got = y.x.cpu().permute(0, 3, 1, 2)
    err = ((got - ref).norm() / ref.norm()).item()
    st = y.stats.cpu()
    serr = ((st[..., 0] - ref.double().sum(dim=(2, 3))).abs().max() / ref.double().sum(dim=(2, 3)).abs().max()).item()
    print(f"{name:8s} {engine:5s} rel_l2={err:.3e} stats_rel={serr:.3e} max|got|={got.abs().max():.3f} max|ref|={ref.abs().max():.3f}", flush=True)
    if err > 1e-2:
        d = (got - ref).abs()
        idx = d.flatten().argmax().item()
        print("   worst index", idx, "got", got.flatten()[idx].item(), "ref", ref.flatten()[idx].item())
        # error pattern per channel block / pixel block
        e = d.permute(0, 2, 3, 1).reshape(-1, Cout)
        print("   err by 32-row block:", [round(v, 3) for v in e.reshape(-1, 32, Cout).amax(dim=(1, 2))[:16].tolist()])
        print("   err by 8-col block:", [round(v, 3) for v in e.reshape(e.shape[0], -1, 8).amax(dim=(0, 2))[:16].tolist()])


if __name__ == "__main__":
    if len(sys.argv) == 3:
        run(sys.argv[1], sys.argv[2])
    else:
        for engine in ("tf32", "umma"):
            for name in CASES:
                r = subprocess.run([sys.executable, __file__, name, engine], capture_output=True, text=True, timeout=180)
                out = (r.stdout + r.stderr).strip().splitlines()
                print("\n".join(out[-6:]) if r.returncode else r.stdout.strip(), flush=True)
                if r.returncode:
                    print(f"   -> {name} {engine} exit {r.returncode}", flush=True)
