"""Generate the committed golden fixtures from the REAL reference module.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference's own tests pin nothing for the vocoder, so these fixtures are
outputs of the unmodified reference `indextts.BigVGAN.models.BigVGAN` (PyTorch path,
use_cuda_kernel=False) on seeded inputs and on the deterministic synthetic state dicts that
`oracle/bigvgan_oracle.py::make_state_dict` produces (loaded with strict load_state_dict, so
the key schema is checked too).  Weights are NOT stored (134 M params); the fixture stores the
sha256 of the generated state dict so a box that generates different bytes is detected.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# matplotlib is absent here and only imported for plotting helpers (BigVGAN/utils.py:7-13)
_m = types.ModuleType("matplotlib")
_p = types.ModuleType("matplotlib.pylab")
_m.use = lambda *a, **k: None
_m.pylab = _p
sys.modules["matplotlib"] = _m
sys.modules["matplotlib.pylab"] = _p
sys.path.insert(0, "/root/reference")

from indextts.BigVGAN.models import BigVGAN  # noqa: E402
from indextts.BigVGAN.alias_free_torch import Activation1d  # noqa: E402
from indextts.BigVGAN import activations  # noqa: E402

from oracle import bigvgan_oracle as O  # noqa: E402


def ref_model(h, sd):
    m = BigVGAN(O.AttrDict(h), use_cuda_kernel=False).eval()
    m.load_state_dict(sd, strict=True)          # infer.py:64
    m.remove_weight_norm()                      # infer.py:66
    return m


def full_forward_cases():
    out = {}
    cases = [
        # name, config, mode, wseed, B, T0, Tm, Bm
        ("full15_tame_T12", O.indextts15_config(), "tame", 0, 1, 12, 64, 1),
        ("full15_wild_T9", O.indextts15_config(), "wild", 3, 2, 9, 50, 2),
        ("small_wild_T17_bcast", O.small_config(), "wild", 5, 3, 17, 33, 1),
        ("small_tame_T1", O.small_config(), "tame", 6, 1, 1, 12, 1),
    ]
    for name, h, mode, wseed, B, T0, Tm, Bm in cases:
        sd = O.make_state_dict(h, wseed, mode)
        m = ref_model(h, sd)
        latent, mel = O.synthetic_inputs(h, B, T0, Tm, seed=wseed + 100, Bm=Bm)
        with torch.no_grad():
            spk = m.speaker_encoder(mel, None)
            wav, closs = m(latent, mel)
        assert closs is None
        out[name] = dict(
            digest=np.array(O.state_dict_digest(sd)),
            mode=np.array(mode), wseed=np.array(wseed), B=np.array(B), T0=np.array(T0),
            Tm=np.array(Tm), Bm=np.array(Bm), iseed=np.array(wseed + 100),
            config=np.array("small" if h.gpt_dim == 40 else "indextts15"),
            spk=spk.numpy(), wav=wav.numpy())
        print(name, tuple(wav.shape), "absmax", float(wav.abs().max()), "rms", float(wav.pow(2).mean().sqrt()))
    return out


def act1d_cases():
    """Activation1d module (alias_free_torch/act.py:9-29) on edge-case lengths."""
    out = {}
    g = torch.Generator().manual_seed(77)
    for T in (1, 2, 3, 5, 6, 11, 12, 13, 31, 32, 33, 100, 257, 1000, 4095, 4096, 4097):
        C = 5 if T < 1000 else 3
        act = activations.SnakeBeta(C, alpha_logscale=True)
        act.alpha.data = torch.randn(C, generator=g) * 0.5
        act.beta.data = torch.randn(C, generator=g) * 0.5
        mod = Activation1d(activation=act).eval()
        x = torch.randn(2, C, T, generator=g) * 1.5
        with torch.no_grad():
            y = mod(x)
        out[f"T{T}"] = dict(x=x.numpy(), alpha=act.alpha.data.numpy(), beta=act.beta.data.numpy(),
                            y=y.numpy(), filt=mod.upsample.filter.reshape(-1).numpy(),
                            filt_down=mod.downsample.lowpass.filter.reshape(-1).numpy())
    return out


def layer_cases():
    """Per-layer-type goldens from the reference module's own submodules (wild small config)."""
    h = O.small_config()
    sd = O.make_state_dict(h, 11, "wild")
    m = ref_model(h, sd)
    g = torch.Generator().manual_seed(12)
    out = {"digest": np.array(O.state_dict_digest(sd))}
    with torch.no_grad():
        x = torch.randn(2, 96, 37, generator=g)
        out["amp_x"] = x.numpy()
        for n in (0, 1, 2):                      # stage-0 AMPBlock1 with k = 3, 7, 11
            out[f"amp{n}_y"] = m.resblocks[n](x).numpy()
        x = torch.randn(2, 192, 9, generator=g)
        out["ups0_x"] = x.numpy()
        out["ups0_y"] = m.ups[0][0](x).numpy()
        x = torch.randn(2, 12, 21, generator=g)
        out["ups4_x"] = x.numpy()
        out["ups4_y"] = m.ups[4][0](x).numpy()
        x = torch.randn(2, 40, 13, generator=g)
        out["pre_x"] = x.numpy()
        out["pre_y"] = m.conv_pre(x).numpy()
        x = torch.randn(2, 3, 50, generator=g)
        out["post_x"] = x.numpy()
        out["post_y"] = torch.tanh(m.conv_post(m.activation_post(x))).numpy()
        mel = torch.randn(3, 29, 20, generator=g) * 2.5 - 0.3
        out["ecapa_mel"] = mel.numpy()
        out["ecapa_y"] = m.speaker_encoder(mel, None).numpy()
    return out


def ecapa_real_case():
    """The speaker encoder of the REAL configuration (IndexTTS-1.5: 100 mel bins, C = 512, 1536-channel MFA, 512-d embedding)
    on a 281-frame (3 s) reference mel, B = 3 -- the shape the benchmark and infer.py:82-93 feed it."""
    h = O.indextts15_config()
    sd = O.make_state_dict(h, 0, "wild")
    m = ref_model(h, sd)
    _, mel = O.synthetic_inputs(h, 3, 8, 281, seed=281)
    with torch.no_grad():
        y = m.speaker_encoder(mel, None)
    return dict(digest=np.array(O.state_dict_digest(sd)), wseed=np.array(0), mode=np.array("wild"), Bm=np.array(3),
                Tm=np.array(281), iseed=np.array(281), y=y.numpy())


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "ecapa_real":      # (added in round 2: leaves the other fixtures untouched)
        np.savez_compressed(os.path.join(HERE, "ecapa_real.npz"), **ecapa_real_case())
        print("ok")
        return
    full = full_forward_cases()
    for name, d in full.items():
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **d)
    act = act1d_cases()
    flat = {}
    for k, d in act.items():
        for kk, v in d.items():
            flat[f"{k}.{kk}"] = v
    np.savez_compressed(os.path.join(HERE, "act1d_cases.npz"), **flat)
    np.savez_compressed(os.path.join(HERE, "layer_cases.npz"), **layer_cases())
    np.savez_compressed(os.path.join(HERE, "ecapa_real.npz"), **ecapa_real_case())
    print("ok")


if __name__ == "__main__":
    main()
