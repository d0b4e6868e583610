"""ctypes binding of libmdctgan_b200.so (C ABI: include/mdctgan_b200.h).  No torch types cross it."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p
from dataclasses import dataclass

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libmdctgan_b200.so"
_lib = None

F32, F64, MIXED = 0, 1, 2     # MIXED: fp64 butterflies on fp32 tensors
MODE_RAW, MODE_ARCSINH, MODE_DB, MODE_EXPLICIT = 0, 1, 2, 3   # DB / EXPLICIT: element-wise codec kernels only (csrc/spectro_codec.cuh)


class _Norm(ctypes.Structure):
    _fields_ = [("mode", c_int32), ("gain", c_float), ("src_lo", c_float), ("src_hi", c_float),
                ("norm_lo", c_float), ("norm_hi", c_float)]


@dataclass(frozen=True)
class NormSpec:
    """The in-kernel spectrogram encoding (Audio2MDCT.normalize, pix2pixHD_model.py:83-125, abs_norm branch)."""
    mode: int = MODE_ARCSINH
    gain: float = 500.0
    src_range: tuple = (-5.0, 5.0)
    norm_range: tuple = (0.0, 1.0)

    def c(self) -> _Norm:
        return _Norm(self.mode, self.gain, self.src_range[0], self.src_range[1], self.norm_range[0], self.norm_range[1])


def lib_path() -> str:
    # MDCTGAN_LIB: another build of the same library (tuning sweeps, tools/build_variants.sh); never a different implementation
    return os.environ.get("MDCTGAN_LIB") or os.path.join(_HERE, _LIB_NAME)


def lib():
    """Load the CUDA library; fail loudly (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  mdctgan_b200 has no CPU / PyTorch fallback.")
    L = ctypes.CDLL(path)
    L.mdctgan_abi_version.restype = c_int
    L.mdctgan_last_error.restype = c_char_p
    L.mdctgan_launch_count.restype = c_int64
    L.mdctgan_frame_count.restype = c_int64
    L.mdctgan_frame_count.argtypes = [c_int64, c_int64, c_int, c_int, c_int]
    L.mdctgan_plan_create.argtypes = [POINTER(c_void_p), c_int, c_int, c_int, c_void_p]
    L.mdctgan_plan_destroy.argtypes = [c_void_p]
    L.mdctgan_mdct4_forward.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p]
    L.mdctgan_audio2mdct_forward.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, POINTER(_Norm), c_void_p,
                                             c_int, c_int64, c_int64, c_int, c_void_p]
    L.mdctgan_imdct4_inverse.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p]
    L.mdctgan_mdct2audio_inverse.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, POINTER(_Norm), c_void_p, c_int64,
                                             c_int64, c_int, c_void_p]
    L.mdctgan_mdct4_forward_host.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int]
    L.mdctgan_audio2mdct_forward_host.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, POINTER(_Norm), c_void_p, c_int, c_int]
    L.mdctgan_imdct4_inverse_host.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int]
    L.mdctgan_mdct2audio_inverse_host.argtypes = [c_void_p, c_void_p, c_int64, c_int64, POINTER(_Norm), c_void_p, c_int64, c_int]
    if L.mdctgan_abi_version() != 1:
        raise RuntimeError("libmdctgan_b200.so ABI version mismatch")
    _lib = L
    return L


def launch_count() -> int:
    return int(lib().mdctgan_launch_count())


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().mdctgan_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libmdctgan_b200: error {rc}: {msg}")


def frame_count(T: int, dim0: int, hop: int, win: int, center: bool = True) -> int:
    return int(lib().mdctgan_frame_count(T, dim0, hop, win, 1 if center else 0))


class Plan:
    """Owns a mdctgan_plan (twiddle / window tables on the current CUDA device)."""

    def __init__(self, n_fft: int, hop_length: int, win_length: int, window_host_f32):
        import numpy as np

        w = np.ascontiguousarray(window_host_f32, dtype=np.float32)
        if w.shape != (win_length,):
            raise ValueError(f"window must have win_length={win_length} values, got {w.shape}")
        self._h = c_void_p()
        check(lib().mdctgan_plan_create(ctypes.byref(self._h), n_fft, hop_length, win_length, w.ctypes.data_as(c_void_p)))
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.mdctgan_plan_destroy(h)

    @property
    def handle(self):
        return self._h
