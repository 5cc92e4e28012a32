"""Pins the oracle (oracle/bigvgan_oracle.py) against the committed outputs of the REAL
reference module (tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import bigvgan_oracle as O


def _cfg(name):
    return O.small_config() if str(name) == "small" else O.indextts15_config()


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def test_taps_match_reference_buffers(golden_dir):
    g = _load(golden_dir, "act1d_cases")
    taps = O.act1d_taps()
    assert np.array_equal(taps, taps[::-1])                   # symmetric
    assert abs(taps.sum() - 1.0) < 1e-7
    np.testing.assert_array_equal(taps.astype(np.float32), g["T12.filt"])   # bit-exact pin
    derived = O.kaiser_sinc_filter1d(0.25, 0.3, 12)           # float64 re-derivation
    np.testing.assert_allclose(derived, taps, rtol=0, atol=1e-7)
    assert abs(derived.sum() - 1.0) < 1e-12
    np.testing.assert_allclose(g["T12.filt"], g["T12.filt_down"], rtol=0, atol=0)
    # SURVEY.md §8(a5) quotes these
    np.testing.assert_allclose(taps[:6], [0.0020289647, 0.0093894657, -0.0255434588,
                                          -0.0576573834, 0.1285725832, 0.4432097971], atol=1e-9, rtol=0)


@pytest.mark.parametrize("T", [1, 2, 3, 5, 6, 11, 12, 13, 31, 32, 33, 100, 257, 1000, 4095, 4096, 4097])
def test_act1d_closed_form_vs_reference(golden_dir, T):
    g = _load(golden_dir, "act1d_cases")
    x = torch.from_numpy(g[f"T{T}.x"])
    y = O.act1d(x, torch.from_numpy(g[f"T{T}.alpha"]), torch.from_numpy(g[f"T{T}.beta"]))
    np.testing.assert_allclose(y.numpy(), g[f"T{T}.y"], rtol=0, atol=2e-6)
    # float64 restatement is even closer to the fp32 reference output (fp32 rounding only)
    y64 = O.act1d(x.double(), torch.from_numpy(g[f"T{T}.alpha"]).double(),
                  torch.from_numpy(g[f"T{T}.beta"]).double())
    np.testing.assert_allclose(y64.numpy(), g[f"T{T}.y"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("T", [1, 2, 5, 12, 33])
def test_act1d_numpy_loops_vs_reference(golden_dir, T):
    g = _load(golden_dir, "act1d_cases")
    x = g[f"T{T}.x"]
    for b in range(x.shape[0]):
        y = O.act1d_numpy(x[b].astype(np.float64), g[f"T{T}.alpha"], g[f"T{T}.beta"])
        np.testing.assert_allclose(y, g[f"T{T}.y"][b], rtol=0, atol=2e-6)


def test_layers_vs_reference(golden_dir):
    g = _load(golden_dir, "layer_cases")
    h = O.small_config()
    sd = O.make_state_dict(h, 11, "wild")
    assert O.state_dict_digest(sd) == str(g["digest"])
    f = O.fold_weight_norm(sd)
    import torch.nn.functional as F
    for n, k in ((0, 3), (1, 7), (2, 11)):
        y = O.amp_block1(torch.from_numpy(g["amp_x"]), f, f"resblocks.{n}", h, k, [1, 3, 5])
        np.testing.assert_allclose(y.numpy(), g[f"amp{n}_y"], rtol=0, atol=2e-5)
    y = F.conv_transpose1d(torch.from_numpy(g["ups0_x"]), f["ups.0.0.weight"], f["ups.0.0.bias"], stride=4, padding=2)
    np.testing.assert_allclose(y.numpy(), g["ups0_y"], rtol=0, atol=1e-5)
    y = F.conv_transpose1d(torch.from_numpy(g["ups4_x"]), f["ups.4.0.weight"], f["ups.4.0.bias"], stride=2, padding=1)
    np.testing.assert_allclose(y.numpy(), g["ups4_y"], rtol=0, atol=1e-5)
    y = F.conv1d(torch.from_numpy(g["pre_x"]), f["conv_pre.weight"], f["conv_pre.bias"], padding=3)
    np.testing.assert_allclose(y.numpy(), g["pre_y"], rtol=0, atol=1e-5)
    y = torch.tanh(F.conv1d(O._act(torch.from_numpy(g["post_x"]), f, "activation_post", h),
                            f["conv_post.weight"], f["conv_post.bias"], padding=3))
    np.testing.assert_allclose(y.numpy(), g["post_y"], rtol=0, atol=1e-5)
    y = O.ecapa_forward(torch.from_numpy(g["ecapa_mel"]), f)
    np.testing.assert_allclose(y.numpy(), g["ecapa_y"], rtol=0, atol=2e-5)


def test_speaker_encoder_real_shape_vs_reference(golden_dir):
    """ECAPA_TDNN.forward of the IndexTTS-1.5 configuration on a 281-frame reference mel (the benchmark's shape), golden made
    by the reference module itself (tests/golden/make_golden.py::ecapa_real_case)."""
    g = _load(golden_dir, "ecapa_real")
    h = O.indextts15_config()
    sd = O.make_state_dict(h, int(g["wseed"]), str(g["mode"]))
    assert O.state_dict_digest(sd) == str(g["digest"])
    _, mel = O.synthetic_inputs(h, int(g["Bm"]), 8, int(g["Tm"]), seed=int(g["iseed"]))
    with torch.no_grad():
        y = O.ecapa_forward(mel, O.fold_weight_norm(sd))
    assert y.shape == g["y"].shape
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5 * max(1.0, float(np.abs(g["y"]).max())))


@pytest.mark.parametrize("name", ["full15_tame_T12", "full15_wild_T9", "small_wild_T17_bcast", "small_tame_T1"])
def test_full_forward_vs_reference(golden_dir, name):
    g = _load(golden_dir, name)
    h = _cfg(g["config"])
    sd = O.make_state_dict(h, int(g["wseed"]), str(g["mode"]))
    assert O.state_dict_digest(sd) == str(g["digest"]), "synthetic weights differ from the golden run"
    latent, mel = O.synthetic_inputs(h, int(g["B"]), int(g["T0"]), int(g["Tm"]), seed=int(g["iseed"]), Bm=int(g["Bm"]))
    with torch.no_grad():
        wav, spk = O.bigvgan_forward(latent, mel, sd, h, return_spk=True)
    assert wav.shape == g["wav"].shape
    np.testing.assert_allclose(spk.numpy(), g["spk"], rtol=0, atol=2e-5)
    # fp32 noise floor of the reference itself is ~1e-7; gate is 1e-4
    np.testing.assert_allclose(wav.numpy(), g["wav"], rtol=0, atol=5e-6)


def test_schema_counts():
    h = O.indextts15_config()
    sch = O.state_dict_schema(h)
    assert len(sch) == 1029
    n = sum(int(np.prod(s)) for k, s, kind in sch if kind not in ("g", "bn_n", "filt", "bn_m", "bn_v"))
    assert abs(n - 133.9e6) / 133.9e6 < 0.01
    folded = O.fold_weight_norm({k: torch.zeros(s) + 1 for k, s, _ in sch})
    assert len(folded) == 913


def _c_oracle():
    import ctypes
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "oracle", "act1d_oracle.c")
    out = os.path.join(root, "oracle", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libact1d_oracle.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lm"])
    lib = ctypes.CDLL(so)
    lib.act1d_oracle.restype = ctypes.c_int
    return lib


@pytest.mark.parametrize("T", [1, 2, 5, 6, 12, 13, 33, 257, 4097])
def test_c_restatement_vs_reference_and_closed_form(golden_dir, T):
    """The plain-C oracle materialises the padded / upsampled signals like the PyTorch ops; it must agree
    with the reference golden and with the polyphase closed form (two independent restatements)."""
    import ctypes
    g = _load(golden_dir, "act1d_cases")
    lib = _c_oracle()
    x = g[f"T{T}.x"].astype(np.float64)
    a = g[f"T{T}.alpha"].astype(np.float64)
    b = g[f"T{T}.beta"].astype(np.float64)
    taps = O.act1d_taps()
    dp = ctypes.POINTER(ctypes.c_double)
    for bi in range(x.shape[0]):
        xc = np.ascontiguousarray(x[bi])
        y = np.empty_like(xc)
        rc = lib.act1d_oracle(xc.ctypes.data_as(dp), y.ctypes.data_as(dp), a.ctypes.data_as(dp), b.ctypes.data_as(dp),
                              taps.ctypes.data_as(dp), xc.shape[0], xc.shape[1])
        assert rc == 0
        np.testing.assert_allclose(y, g[f"T{T}.y"][bi], rtol=0, atol=2e-6)
        ref = O.act1d(torch.from_numpy(xc)[None], torch.from_numpy(a), torch.from_numpy(b))[0].numpy()
        np.testing.assert_allclose(y, ref, rtol=0, atol=1e-12)
