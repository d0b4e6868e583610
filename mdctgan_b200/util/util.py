"""Host-side helpers with the reference's names (util/util.py)."""
import math

import torch


def kbdwin(N: int, beta: float = 12.0, device="cpu") -> torch.Tensor:
    """MATLAB-style Kaiser-Bessel-derived window, fp32 (reference: util/util.py:179-186).

    Built with the same torch library routine on the same device as the reference builds it
    (CPU by default, then moved), so the fp32 bits -- which set the MDCT round-trip error floor --
    are identical.
    """
    assert N % 2 == 0, "N must be even"
    k = torch.kaiser_window(window_length=N // 2 + 1, beta=beta * math.pi, periodic=False, device=device)
    half = torch.sqrt(torch.cumsum(k, dim=0) / k.sum())[:-1]
    return torch.cat((half, half.flip(dims=(0,))), dim=0)
