"""Kernel-side weight images for a whole model, refreshed by ONE kernel launch per optimiser step.

The convolution kernels read their weights as [kh*kw*Cin][Cout] (direct fp32 kernels) or as the pre-swizzled TF32
hi|lo shared-memory image of the tcgen05 kernel (csrc/conv_umma.cuh); the input-gradient convolutions need the same
weights on the transposed geometry.  At inference the images are cached per weight version (models/networks.py
_PackedWeight); in training the weights change every step, so `WeightPacker` lays all images of all layers out in one
buffer, installs views of it on the layers, and `refresh()` rewrites everything from the flat parameter buffer with
`mdctgan_pack_weights_multi` (one launch, ~25 bytes of traffic per parameter)."""
from __future__ import annotations

import ctypes
from ctypes import c_int32, c_int64, c_void_p

import torch

from . import _lib
from . import nn_ops as ops


class _Desc(ctypes.Structure):
    _fields_ = [("src", c_void_p), ("dst_kn", c_void_p), ("dst_umma", c_void_p), ("K", c_int32), ("N", c_int32), ("Kch", c_int32),
                ("taps", c_int32), ("flip", c_int32), ("kchunks", c_int32), ("s_kch", c_int64), ("s_n", c_int64), ("s_tap", c_int64),
                ("work_begin", c_int64)]


class WeightPacker:
    def __init__(self, *modules: torch.nn.Module, dgrad: bool = True):
        from .models.networks import Conv2d, ConvTranspose2d

        L = ops._L()
        L.mdctgan_pack_weights_multi.argtypes = [c_void_p, ctypes.c_int, c_int64, c_void_p]
        L.mdctgan_pack_weights_tiled.argtypes = [c_void_p, c_void_p, ctypes.c_int, c_int64, ctypes.c_int, c_void_p]
        layers = [m for mod in modules for m in mod.modules() if isinstance(m, (Conv2d, ConvTranspose2d))]
        layers = list(dict.fromkeys(layers))
        if not layers:
            raise ValueError("WeightPacker: no convolution layers")
        dev = layers[0].weight.device
        plans, total_floats = [], 0

        def plan(layer, role, K, N, Kch, taps, flip, s_kch, s_n, s_tap, strided_transposed=False):
            nonlocal total_floats
            kchunks = (K + 31) // 32
            use_umma = bool(L.mdctgan_conv2d_umma_supported(Kch, N))
            if strided_transposed and (Kch % 32 or min(layer.kernel_size) < layer.stride[0]):
                use_umma = False                    # the tensor-core kernel runs strided transposed geometry per parity class only
            # the [K][N] image is only read by the direct fp32 kernels: skip it where the tcgen05 kernel runs the layer
            want_kn = (not use_umma) or ops.CONV_ENGINE == "direct" or ops._ROLE_ENGINE.get("dgrad" if role == "dgrad" else "fwd", "") == "direct"
            # [kh][kw][Cin][Cout] storage (optim.FlatBucket): the master copy already IS the forward [K][N] image
            alias_kn = want_kn and not flip and s_n == 1 and s_kch == N and s_tap == Kch * N
            off_kn = None
            if want_kn and not alias_kn:
                off_kn = total_floats
                total_floats += (K * N + 3) // 4 * 4
            off_um = None
            if use_umma:
                total_floats = (total_floats + 31) // 32 * 32          # 128-byte aligned: the image is fetched by TMA bulk copies
                off_um = total_floats
                total_floats += kchunks * 2 * ((N + 31) // 32 * 32) * 32       # N padded to 32 rows: the padding stays zero (buffer is zeroed)
            if alias_kn and off_um is None:          # nothing to derive: the layer reads its weights in place
                st = layer.__dict__.setdefault("_static_pack", {})
                st[role] = (layer.weight.detach().permute(2, 3, 0, 1).reshape(K, N) if isinstance(layer, ConvTranspose2d)
                            else layer.weight.detach().permute(2, 3, 1, 0).reshape(K, N), None, bool(flip))
                assert st[role][0].data_ptr() == layer.weight.data_ptr() and st[role][0].is_contiguous()
                return
            plans.append(dict(layer=layer, role=role, K=K, N=N, Kch=Kch, taps=taps, flip=flip, s_kch=s_kch, s_n=s_n, s_tap=s_tap, kchunks=kchunks,
                              off_kn=off_kn, off_um=off_um, alias_kn=alias_kn))

        for m in layers:
            kh, kw = m.kernel_size
            taps = kh * kw
            ci, co = m.in_channels, m.out_channels
            st = m.weight.stride()                  # the parameter's actual element strides (reference layout or FlatBucket's storage order)
            if kh > 1 and st[2] != kw * st[3]:
                raise RuntimeError(f"WeightPacker: unsupported weight strides {st}")
            if isinstance(m, ConvTranspose2d):      # weight [Cin][Cout][kh][kw]
                s_ci, s_co, s_tap = st[0], st[1], st[3]
                plan(m, "fwd", taps * ci, co, ci, taps, 0, s_ci, s_co, s_tap, strided_transposed=m.stride[0] > 1)
                if dgrad:
                    plan(m, "dgrad", taps * co, ci, co, taps, 0, s_co, s_ci, s_tap)
            else:                                   # weight [Cout][Cin][kh][kw]
                s_co, s_ci, s_tap = st[0], st[1], st[3]
                plan(m, "fwd", taps * ci, co, ci, taps, 0, s_ci, s_co, s_tap)
                if dgrad:
                    plan(m, "dgrad", taps * co, ci, co, taps, 1 if m.stride[0] == 1 else 0, s_co, s_ci, s_tap, strided_transposed=m.stride[0] > 1)
        self.buf = torch.zeros(total_floats + 32, dtype=torch.float32, device=dev)
        base_off = (-(self.buf.data_ptr() // 4)) % 32                     # align the buffer itself to 128 bytes
        # descriptors that qualify for the tiled (coalesced) kernel: tensor-core image only, 32-aligned channel counts
        def tiled_ok(p):
            return p["off_um"] is not None and p["off_kn"] is None and p["Kch"] % 32 == 0 and p["N"] % 32 == 0 and p["taps"] <= 49
        plans = [p for p in plans if not tiled_ok(p)] + [p for p in plans if tiled_ok(p)]
        n_generic = sum(0 if tiled_ok(p) else 1 for p in plans)
        descs = (_Desc * len(plans))()
        work = 0
        for i, p in enumerate(plans):
            m = p["layer"]
            if p["off_kn"] is not None:
                kn = self.buf[base_off + p["off_kn"]: base_off + p["off_kn"] + p["K"] * p["N"]].view(p["K"], p["N"])
                kn_ptr = kn.data_ptr()
            elif p["alias_kn"]:                     # the master copy in [kh][kw][Cin][Cout] order is the image
                kn = (m.weight.detach().permute(2, 3, 0, 1) if isinstance(m, ConvTranspose2d) else m.weight.detach().permute(2, 3, 1, 0)).reshape(p["K"], p["N"])
                assert kn.data_ptr() == m.weight.data_ptr() and kn.is_contiguous()
                kn_ptr = None
            else:                                   # shape-only placeholder (stride 0): never dereferenced, _conv_launch checks
                kn = self.buf[:1].as_strided((p["K"], p["N"]), (0, 0))
                kn_ptr = None
            um = None
            if p["off_um"] is not None:
                n_um = p["kchunks"] * 2 * ((p["N"] + 31) // 32 * 32) * 32
                um = self.buf[base_off + p["off_um"]: base_off + p["off_um"] + n_um]
                assert um.data_ptr() % 128 == 0
            descs[i] = _Desc(m.weight.data_ptr(), kn_ptr, um.data_ptr() if um is not None else None, p["K"], p["N"], p["Kch"],
                             p["taps"], p["flip"], p["kchunks"], p["s_kch"], p["s_n"], p["s_tap"], work)
            work += p["kchunks"] * 32 * p["N"]
            if i == n_generic - 1:
                self.total_work = work                                    # the generic kernel stops here
            st = m.__dict__.setdefault("_static_pack", {})
            st[p["role"]] = (kn, um, bool(p["flip"]))
        self.n_desc = n_generic
        if n_generic == 0:
            self.total_work = 0
        tiled = plans[n_generic:]
        self.n_tiled = len(tiled)
        tb, acc_t = [], 0
        for p in tiled:
            tb.append(acc_t)
            acc_t += (p["N"] // 32) * (p["Kch"] // 32)
        self.total_tiles = acc_t
        self.max_taps = max([p["taps"] for p in tiled], default=1)
        self.tile_begin = torch.tensor(tb if tb else [0], dtype=torch.int64, device=dev)
        self._desc_size = ctypes.sizeof(_Desc)
        self.layers = layers
        self._ptrs = [m.weight.data_ptr() for m in layers]
        raw = bytes(descs)
        self.descs = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self.refresh()

    def refresh(self):
        """Rewrite every image from the current parameter values (one launch on the current stream)."""
        for m, ptr in zip(self.layers, self._ptrs):
            if m.weight.data_ptr() != ptr:
                raise RuntimeError("WeightPacker: a parameter was re-allocated after the packer was built (call it after FlatBucket)")
        with torch.cuda.device(self.buf.device):
            st = torch.cuda.current_stream(self.buf.device).cuda_stream
            if self.n_desc:
                _lib.check(ops._L().mdctgan_pack_weights_multi(self.descs.data_ptr(), self.n_desc, self.total_work, st))
            if self.n_tiled:
                _lib.check(ops._L().mdctgan_pack_weights_tiled(self.descs.data_ptr() + self.n_desc * self._desc_size, self.tile_begin.data_ptr(),
                                                               self.n_tiled, self.total_tiles, self.max_taps, st))

    def detach(self):
        for m in self.layers:
            m.__dict__.pop("_static_pack", None)
