"""mdctgan_b200 -- B200-native (sm_100a) engine for the mdctGAN MDCT -> generator -> IMDCT hot path.

Python host code over PyTorch tensors calling hand-written CUDA through the C ABI of
``libmdctgan_b200.so`` (``include/mdctgan_b200.h``).  The sub-packages mirror the reference's module
paths (``models.mdct``, ``models.pix2pixHD_model``, ``util.util``) so callers switch by import path only.

There is no CPU fallback: every transform raises if the CUDA library is missing or the tensors are
not on a CUDA device.
"""
from ._lib import Plan, NormSpec, lib, lib_path, launch_count  # noqa: F401

__all__ = ["Plan", "NormSpec", "lib", "lib_path", "launch_count"]
