"""The drop-in boundary on CPU: the C-ABI library loads, exports every symbol include/*.h declares, the
host-only entry points behave like the reference, and the product never touches the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import mdct_oracle as O


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build_lib()
    import mdctgan_b200

    return mdctgan_b200.lib()


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(mdctgan_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_exports_every_declared_symbol(lib):
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "mdctgan_b200", "libmdctgan_b200.so")],
                         capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (mdctgan_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert lib.mdctgan_abi_version() == 1


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "mdctgan_b200", "libmdctgan_b200.so")],
                         capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_frame_count_matches_reference_quirk(lib, mdct_golden):
    from mdctgan_b200 import _lib

    for T, dim0 in ((8192, 8192), (8193, 8193), (8193, 4), (1000, 3), (1000, 1000), (7936, 4), (32512, 64), (100, 100), (0, 1)):
        assert _lib.frame_count(T, dim0, 256, 512) == O.frame_count(T, 256, 512, dim0)[2], (T, dim0)
    assert _lib.frame_count(8193, 4, 256, 512) == mdct_golden["q4_spec"].shape[1]
    assert _lib.frame_count(8193, 8193, 256, 512) == mdct_golden["q1_spec"].shape[0]


def test_argument_errors_are_reported_not_fatal(lib):
    from mdctgan_b200 import _lib

    h = ctypes.c_void_p()
    w = O.kbdwin(512)
    # unsupported transform size -> -2 with a message, before any CUDA call
    rc = lib.mdctgan_plan_create(ctypes.byref(h), 1024, 512, 1024, w.ctypes.data_as(ctypes.c_void_p))
    assert rc == -2 and b"unsupported" in lib.mdctgan_last_error()
    rc = lib.mdctgan_mdct4_forward(None, None, 1, 8192, 8192, 33, None, 33 * 256, 0, None)
    assert rc == -1
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_product_never_imports_oracle_and_has_no_cpu_path():
    pkg = os.path.join(ROOT, "mdctgan_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "/root/reference" not in src, fn
    import torch

    from mdctgan_b200.models.mdct import MDCT4
    from mdctgan_b200.util.util import kbdwin

    m = MDCT4(512, 256, 512, kbdwin(512))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        m(torch.zeros(8192))


def test_missing_library_fails_loudly(tmp_path):
    code = ("import sys; sys.path.insert(0, %r); import mdctgan_b200._lib as L; L._HERE = %r\n"
            "try:\n    L.lib()\nexcept RuntimeError as e:\n    print('RAISED', e)\n" % (ROOT, str(tmp_path)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    assert "RAISED" in out and "no CPU" in out


def test_kbdwin_bits(mdct_golden):
    from mdctgan_b200.util.util import kbdwin

    for n in (64, 512, 1024):
        assert np.array_equal(kbdwin(n).numpy(), mdct_golden[f"kbdwin{n}"])


def test_host_side_shape_functions(lib):
    """Pure host arithmetic of the C ABI (no kernel launch): segment count of seg_pad_audio (data/audio_dataset.py:153-167) and frame
    count of the LSD spectrogram (util/util.py:170-171) against the oracle restatements."""
    import ctypes

    import torch

    from oracle import longform_oracle as LO

    lib.mdctgan_segment_count.restype = ctypes.c_int64
    lib.mdctgan_segment_count.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    for L in (1, 100, 3839, 3840, 3841, 7936, 20000, 100000, 2880000):
        for seg, ov in ((3840, 0), (3840, 128), (7936, 256), (32512, 0), (32512, 4096)):
            want = LO.seg_pad_audio(torch.zeros(1, L), seg, ov).shape[0]
            assert lib.mdctgan_segment_count(L, seg, ov) == want, (L, seg, ov)
    lib.mdctgan_lsd_frame_count.restype = ctypes.c_int64
    lib.mdctgan_lsd_frame_count.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    for T in (1024, 4000, 7936, 32512):
        for center in (0, 1):
            z = torch.stft(torch.zeros(T), 1024, 512, 1024, window=torch.ones(1024), center=bool(center), return_complex=True)
            assert lib.mdctgan_lsd_frame_count(T, 1024, 512, center) == z.shape[-1], (T, center)


def test_option_surface_matches_reference():
    """TrainOptions().parse() returns the same Namespace as the reference's (options/base_options.py:11-127, train_options.py:5-73) for
    the default and the train.sh command lines (tests/golden/options_golden.json, generated from the reference); our only extra field
    is `mdct_precision`."""
    import json
    import os
    import sys

    from conftest import GOLDEN
    from mdctgan_b200.options.train_options import TrainOptions

    sys.path.insert(0, GOLDEN)
    from make_golden import OPTION_ARGV

    gold = json.load(open(os.path.join(GOLDEN, "options_golden.json")))
    base = ["--name", "g", "--checkpoints_dir", "/tmp/x", "--gpu_ids", "-1", "--lr_sampling_rate", "12000", "--sr_sampling_rate", "48000",
            "--arcsinh_transform", "--abs_spectro", "--arcsinh_gain", "1000", "--center", "--norm_range", "-1", "1", "--abs_norm",
            "--src_range", "-5", "5"]
    for extra, ref in zip(OPTION_ARGV, gold):
        ours = vars(TrainOptions().parse(save=False, args=base + list(extra)))
        assert set(ours) - set(ref) == {"mdct_precision", "checkpoints_dir"}
        assert set(ref) <= set(ours)
        for k, v in ref.items():
            o = ours[k]
            o = list(o) if isinstance(o, (list, tuple)) else o
            if isinstance(v, float) and v != v:
                continue
            assert o == v or str(o) == v, (k, o, v)
