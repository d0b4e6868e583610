"""Pin the CPU oracle against outputs of the reference itself (tests/golden/*.npz).

These fixtures were produced by tests/golden/make_golden.py importing the
unmodified reference from /root/reference.  CPU-only.
"""
import numpy as np

from oracle import mdct_oracle as O
from conftest import rel_l2

ARC = dict(arcsinh_transform=True, arcsinh_gain=1000.0, abs_norm=True, src_range=(-5.0, 5.0), norm_range=(-1.0, 1.0))


def test_kbdwin_bits(mdct_golden):
    for n in (64, 512, 1024):
        assert np.array_equal(O.kbdwin(n), mdct_golden[f"kbdwin{n}"])
    # fp64 window agrees with the fp32 one to fp32 rounding and is Princen-Bradley to 1e-15
    w = O.kbdwin_f64(512)
    assert np.abs(w - mdct_golden["kbdwin512"]).max() < 2e-7
    assert np.abs(w[:256] ** 2 + w[256:] ** 2 - 1).max() < 1e-14


def test_frame_count_quirk(mdct_golden):
    # 1-D: ceil(T/hop)+1 ; 2-D: the pad is derived from the batch size (mdct.py:394-402)
    assert O.frame_count(8192, 256, 512, 8192)[2] == 33
    assert O.frame_count(8193, 256, 512, 8193)[2] == 34 == mdct_golden["q1_spec"].shape[0]
    assert O.frame_count(8193, 256, 512, 4)[2] == 33 == mdct_golden["q4_spec"].shape[1]
    assert O.frame_count(1000, 256, 512, 3)[2] == 5 == mdct_golden["r3_spec"].shape[1]


def test_mdct4_matches_reference(mdct_golden):
    g = mdct_golden
    w = g["kbdwin512"]
    for xk, sk in (("c1_x", "c1_spec"), ("b4_x", "b4_spec"), ("r3_x", "r3_spec"), ("q4_x", "q4_spec")):
        spec, _ = O.mdct4(g[xk], w)
        assert spec.shape == g[sk].shape
        assert np.abs(spec - g[sk]).max() <= 1e-13 * np.abs(g[sk]).max(), sk
    spec, frames = O.mdct4(g["c1_x"], w)
    assert np.array_equal(frames, g["c1_frames"])
    assert np.abs(O.mdct4(g["r3_x"][0], w)[0] - g["r1_spec"]).max() < 1e-13
    assert np.abs(O.mdct4(g["q4_x"][0], w)[0] - g["q1_spec"]).max() < 1e-13
    s2, _ = O.mdct4(g["n1024_x"], g["kbdwin1024"], 1024, 512)
    assert np.abs(s2 - g["n1024_spec"]).max() < 1e-12


def test_mdct4_closed_form(mdct_golden):
    g = mdct_golden
    c = O.mdct4_closed(g["c1_x"], g["kbdwin512"])
    assert np.abs(c - g["c1_spec"]).max() <= 1e-12 * np.abs(g["c1_spec"]).max()
    c2 = O.mdct4_closed(g["n1024_x"], g["kbdwin1024"], 1024, 512)
    assert np.abs(c2 - g["n1024_spec"]).max() < 1e-11


def test_imdct4_matches_reference(mdct_golden):
    g = mdct_golden
    w = g["kbdwin512"]
    a = O.imdct4(g["c1_spec"][None], w)
    assert a.shape == g["c1_audio"].shape
    assert np.abs(a - g["c1_audio"]).max() < 1e-15
    assert np.abs(O.imdct4(g["b4_spec"], w) - g["b4_audio"]).max() < 1e-15
    assert np.abs(O.imdct4_closed(g["b4_spec"], w) - g["b4_audio"]).max() < 1e-13
    assert np.abs(O.imdct4(g["r3_spec"], w, out_length=1000) - g["r3_audio_crop"]).max() < 1e-15
    assert np.abs(O.imdct4(g["n1024_spec"], g["kbdwin1024"], 1024, 512) - g["n1024_audio"]).max() < 1e-15


def test_round_trip_floor(mdct_golden):
    """reference round trip is 1.41 eps*peak, entirely from the fp32 window."""
    g = mdct_golden
    x = g["c1_x"].astype(np.float64)
    e = np.abs(g["c1_audio"].ravel() - x).max()
    peak = np.abs(x).max()
    assert 1.2 * 2.0 ** -23 * peak < e < 2.0 * 2.0 ** -23 * peak
    w64 = O.kbdwin_f64(512)
    s = O.mdct4_closed(g["c1_x"], w64)
    a = O.imdct4_closed(s[None], w64)
    assert np.abs(a.ravel() - x).max() < 1e-13


def test_to_spectro_to_audio(mdct_golden):
    g = mdct_golden
    w = g["kbdwin512"]
    ls, sign, hi, lo = O.to_spectro(g["b4_x"], w, **ARC)
    assert ls.dtype == np.float32 and ls.shape == g["a2m_log_spectro"].shape
    assert np.array_equal(ls, g["a2m_log_spectro"])
    assert np.array_equal(hi, g["a2m_max"]) and np.array_equal(lo, g["a2m_min"])
    _, _, _, mean, std = O.compress(g["b4_spec"][:, None], **ARC)
    assert abs(mean - g["a2m_mean"]) < 1e-6 and abs(std - g["a2m_std"]) < 1e-6
    audio = O.to_audio(ls, lo, hi, w, arcsinh_transform=True, arcsinh_gain=1000.0, norm_range=(-1.0, 1.0))
    assert rel_l2(audio, g["a2m_audio"]) < 1e-14
    # the normalise->denormalise round trip is only fp32-exact: ~1e-7 relative to the clip
    assert rel_l2(audio.reshape(4, -1), g["b4_x"]) < 2e-6


def test_minmax_and_raw_branches(mdct_golden):
    g = mdct_golden
    w = g["kbdwin512"]
    kw = dict(ARC, abs_norm=False)
    ls, _, hi, lo = O.to_spectro(g["b4_x"], w, **kw)
    assert np.array_equal(hi, g["mm_max"]) and np.array_equal(lo, g["mm_min"])
    assert np.abs(ls - g["mm_log_spectro"]).max() <= 6e-8
    a = O.to_audio(ls, lo, hi, w, arcsinh_transform=True, arcsinh_gain=1000.0, norm_range=(-1.0, 1.0))
    assert rel_l2(a, g["mm_audio"]) < 1e-6
    kw = dict(ARC, arcsinh_transform=False, raw_mdct=True)
    ls, _, hi, lo = O.to_spectro(g["b4_x"], w, **kw)
    assert np.abs(ls - g["raw_log_spectro"]).max() <= 1e-6 * np.abs(g["raw_log_spectro"]).max()
    a = O.to_audio(ls, lo, hi, w, arcsinh_transform=False, raw_mdct=True, norm_range=(-1.0, 1.0))
    assert rel_l2(a, g["raw_audio"]) < 1e-6


def test_torch_port_matches_reference(mdct_golden):
    """oracle/torch_port.py (the torch-CPU restatement bench.py times as the CPU baseline)."""
    import torch

    from oracle import torch_port as P

    g = mdct_golden
    w = P.kbdwin(512)
    assert np.array_equal(w.numpy(), g["kbdwin512"])
    fwd, inv = P.MDCT4Port(512, 256, w), P.IMDCT4Port(512, 256, w)
    s, fr = fwd(torch.from_numpy(g["c1_x"]), True)
    assert np.abs(s.numpy() - g["c1_spec"]).max() <= 1e-13 * np.abs(g["c1_spec"]).max()
    assert np.array_equal(fr.numpy(), g["c1_frames"])
    assert np.abs(inv(s[None]).numpy() - g["c1_audio"]).max() < 1e-15
    for xk, sk in (("b4_x", "b4_spec"), ("r3_x", "r3_spec"), ("q4_x", "q4_spec")):
        assert np.abs(fwd(torch.from_numpy(g[xk]))[0].numpy() - g[sk]).max() <= 1e-13 * np.abs(g[sk]).max()
    a2m = P.Audio2MDCTPort()
    ls, pha, prm = a2m.to_spectro(torch.from_numpy(g["b4_x"]))
    assert np.array_equal(ls.numpy(), g["a2m_log_spectro"])
    assert rel_l2(a2m.to_audio(ls, prm, pha).numpy(), g["a2m_audio"]) < 1e-14


ENCODING_CASES = {
    "db_abs": dict(arcsinh_transform=False, raw_mdct=False, explicit_encoding=False, abs_norm=True, src_range=(-180.0, 20.0)),
    "db_minmax": dict(arcsinh_transform=False, raw_mdct=False, explicit_encoding=False, abs_norm=False),
    "explicit_minmax": dict(arcsinh_transform=False, raw_mdct=False, explicit_encoding=True, abs_norm=False),
    "explicit_abs": dict(arcsinh_transform=False, raw_mdct=False, explicit_encoding=True, abs_norm=True, src_range=(-180.0, 20.0)),
    "arcsinh_minmax": dict(arcsinh_transform=True, raw_mdct=False, explicit_encoding=False, abs_norm=False),
    "raw_minmax": dict(arcsinh_transform=False, raw_mdct=True, explicit_encoding=False, abs_norm=False),
}
ENC_COMMON = dict(arcsinh_gain=1000.0, alpha=0.6, min_value=1e-7, norm_range=(-1.0, 1.0))


def test_secondary_encodings_match_reference(mdct_golden):
    """dB / explicit_encoding / per-plane min-max branches of Audio2MDCT (pix2pixHD_model.py:83-163) against the reference's own
    to_spectro / to_audio outputs (tests/golden/encodings_golden.npz, make_golden.py gen_encodings)."""
    import os

    from conftest import GOLDEN

    g = dict(np.load(os.path.join(GOLDEN, "encodings_golden.npz")))
    w = mdct_golden["kbdwin512"]
    keep = 30 * 256       # samples the last frame (random pseudo phase in the reference, :150-157) does not touch
    for tag, kw in ENCODING_CASES.items():
        kw = dict(ENC_COMMON, **kw)
        ls, sign, hi, lo = O.to_spectro(g["audio"], w, **kw)
        assert ls.shape == g[f"{tag}_spectro"].shape, tag
        assert np.array_equal(hi.reshape(-1), g[f"{tag}_max"].reshape(-1)) and np.array_equal(lo.reshape(-1), g[f"{tag}_min"].reshape(-1)), tag
        assert np.abs(ls - g[f"{tag}_spectro"]).max() <= 2e-7 * max(1.0, np.abs(g[f"{tag}_spectro"]).max()), tag
        assert np.array_equal(sign.astype(np.float32), g["sign"]), tag
        C = ls.shape[1]
        dkw = {k: v for k, v in kw.items() if k not in ("abs_norm", "src_range")}
        a = O.to_audio(g["log_spectro"][:, :C], lo, hi, w, pha=g["sign"], **dkw)
        ref = g[f"{tag}_decode"].reshape(2, -1)
        assert rel_l2(a.reshape(2, -1)[:, :keep], ref[:, :keep]) < 1e-12, tag
