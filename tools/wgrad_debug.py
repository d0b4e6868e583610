"""Debug probes for the tcgen05 weight-gradient kernel (run on the GPU box): structured inputs that reveal the operand mapping."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mdctgan_b200 import _lib
from mdctgan_b200 import nn_ops as ops

np.set_printoptions(linewidth=200, precision=3, suppress=True)
dev = torch.device("cuda:0")
L = ops._L()


def run(x, dy, Cin, Cout, k, eng, layout="kn", H=None, W=None, B=1, pad=0):
    taps = k * k
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    s = (Cin * taps, taps, 1) if layout == "param" else (1, Cout, Cin * Cout)
    dw = torch.zeros(Cout * Cin * taps, device=dev)
    db = torch.zeros(Cout, device=dev)
    _lib.check(L.mdctgan_conv2d_wgrad(x.data_ptr(), B, H, W, Cin, dy.data_ptr(), Ho, Wo, Cout, k, k, 1, pad, 0, 0, None, None, 1, 0, None, 0.0, 1e-5,
                                      dw.data_ptr(), s[0], s[1], s[2], db.data_ptr(), eng, torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
    dw = dw.cpu().numpy()
    if layout == "kn":
        return dw.reshape(taps, Cin, Cout), db.cpu().numpy()
    return dw.reshape(Cout, Cin, taps).transpose(2, 1, 0), db.cpu().numpy()


print("DEBUG", os.environ.get("MDCTGAN_WGRAD_DEBUG"))
for (Cin, Cout, H, W) in ((32, 32, 8, 8),):
    P = H * W
    print(f"=== Cin {Cin} Cout {Cout} P {P}")
    one_x = torch.ones(1, H, W, Cin, device=dev)
    one_y = torch.ones(1, H, W, Cout, device=dev)
    for eng in (0, 2):
        dw, db = run(one_x, one_y, Cin, Cout, 1, eng, H=H, W=W)
        print(f"eng {eng} ones: dW unique {np.unique(dw)[:8]} db unique {np.unique(db)[:4]}")
    xc = torch.arange(Cin, device=dev, dtype=torch.float32).view(1, 1, 1, Cin).expand(1, H, W, Cin).contiguous()
    yc = torch.arange(Cout, device=dev, dtype=torch.float32).view(1, 1, 1, Cout).expand(1, H, W, Cout).contiguous()
    for eng in (0, 2):
        dw, _ = run(xc, one_y, Cin, Cout, 1, eng, H=H, W=W)
        print(f"eng {eng} x=c: dW[0,:,0]/P {dw[0, :, 0][:40] / P}")
        print(f"          dW[0,3,:]/P {dw[0, 3, :][:16] / P}")
        dw, _ = run(one_x, yc, Cin, Cout, 1, eng, H=H, W=W)
        print(f"eng {eng} dy=co: dW[0,0,:]/P {dw[0, 0, :][:40] / P}")
        print(f"           dW[0,:,5]/P {dw[0, :, 5][:16] / P}")
    # pixel probe: x[p, c] = 1 if p == 3 else 0; dy[p, co] = p  -> dW = 3
    xp = torch.zeros(1, H, W, Cin, device=dev); xp.view(P, Cin)[3] = 1
    yp = torch.arange(P, device=dev, dtype=torch.float32).view(1, H, W, 1).expand(1, H, W, Cout).contiguous()
    for eng in (0, 2):
        dw, _ = run(xp, yp, Cin, Cout, 1, eng, H=H, W=W)
        print(f"eng {eng} pixel probe: unique {np.unique(dw)[:10]}")
