"""Mel front end + speaker-embedding cache.

Reference: MelSpectrogramFeatures.forward (indextts/utils/feature_extractors.py:24-50) + safe_log (utils/common.py:110), called
from IndexTTS.extract_features (infer.py:82-93).  Goldens: outputs of the reference class itself (tests/golden/make_mel_golden.py).
Tolerance (floating point, stated here): log-mel max-abs <= 2e-3 where the reference's mel magnitude is above 1e-4 (fp32 DFT
of 1024 terms vs torch's fp32 FFT), and the clipped floor log(1e-7) reproduced where the reference sits on it."""
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle as M

GOLD = os.path.join(os.path.dirname(__file__), "golden", "mel_cases.npz")


def _cases():
    g = np.load(GOLD)
    return [(k[6:], g[k], g["mel_" + k[6:]]) for k in g.files if k.startswith("audio_")]


def _check(mel, ref, tol):
    assert mel.shape == ref.shape, (mel.shape, ref.shape)
    loud = ref > np.log(1e-4)
    assert np.abs(mel - ref)[loud].max() <= tol, float(np.abs(mel - ref)[loud].max())
    # quiet bins: compare magnitudes (absolute), the log amplifies rounding noise of near-zero sums
    assert np.abs(np.exp(mel) - np.exp(ref))[~loud].max(initial=0.0) <= 2e-5


def test_mel_oracle_matches_reference_golden():
    for name, audio, ref in _cases():
        _check(M.log_mel(audio).astype(np.float32), ref, 5e-4)


def test_filterbank_matches_oracle_and_shape():
    import index_tts_ipex_b200 as pkg
    fb = pkg.melscale_fbanks_htk(513, 0.0, 12000.0, 100, 24000)
    assert fb.shape == (513, 100) and fb.dtype == np.float32
    assert np.abs(fb - M.melscale_fbanks_htk(513, 0.0, 12000.0, 100, 24000)).max() < 1e-7
    assert (fb >= 0).all() and (fb.sum(0) > 0).all()          # every band has support (no empty filters at 100 mels / 1024 fft)


def test_mel_module_mirrors_reference_constructor():
    import index_tts_ipex_b200 as pkg
    m = pkg.MelSpectrogramFeatures()
    assert (m.sample_rate, m.n_fft, m.hop_length, m.n_mels) == (24000, 1024, 256, 100)
    with pytest.raises(ValueError):
        pkg.MelSpectrogramFeatures(padding="reflect")
    with pytest.raises(NotImplementedError):
        pkg.MelSpectrogramFeatures(padding="same")
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4000))                               # CPU tensor: no fallback


def test_speaker_cache_lru():
    import index_tts_ipex_b200 as pkg
    c = pkg.SpeakerEmbeddingCache(max_voices=2)
    c.put("a", torch.zeros(1)); c.put("b", torch.ones(1)); assert c.get("a") is not None
    c.put("c", torch.ones(1))                                 # evicts "b" (least recently used)
    assert c.get("b") is None and c.get("a") is not None and len(c) == 2


@pytest.mark.gpu
def test_mel_cuda_matches_reference_golden_and_oracle():
    import index_tts_ipex_b200 as pkg
    m = pkg.MelSpectrogramFeatures().cuda()
    for name, audio, ref in _cases():
        mel = m(torch.from_numpy(audio).cuda()).cpu().numpy()
        _check(mel, ref, 2e-3)
        _check(mel, M.log_mel(audio).astype(np.float32), 2e-3)
    # a batch, a 3 s prompt (281 frames, the benchmark's reference-mel length) and the [B, frames, n_mels] layout
    audio = np.stack([M.synthetic_prompt(71680, s) for s in (3, 4)])
    btc = m.forward_btc(torch.from_numpy(audio).cuda())
    assert btc.shape == (2, 281, 100)
    _check(btc.transpose(1, 2).cpu().numpy(), M.log_mel(audio).astype(np.float32), 2e-3)


@pytest.mark.gpu
def test_voice_embedding_cache_feeds_decode():
    import index_tts_ipex_b200 as pkg
    from oracle import bigvgan_oracle as O
    h = O.small_config() if hasattr(O, "small_config") else O.indextts15_config()
    m = pkg.BigVGAN(h, use_cuda_kernel=True)
    m.load_state_dict(O.make_state_dict(h, 0, "tame"), strict=True)
    m = m.cuda().eval()
    m.remove_weight_norm()
    audio = torch.from_numpy(M.synthetic_prompt(30000, 7))[None].cuda()
    e1 = m.voice_embedding(audio=audio, key="voice-1")
    e2 = m.voice_embedding(audio=audio, key="voice-1")
    assert e1 is e2 and m._spk_cache.hits == 1 and m._spk_cache.misses == 1
    mel = pkg.MelSpectrogramFeatures(n_mels=int(h.num_mels)).cuda().forward_btc(audio)
    assert torch.equal(e1, m.speaker_embed(mel))
    lat, _ = O.synthetic_inputs(h, 1, 12, 8, seed=2)
    w1 = m.decode(lat.cuda(), spk=e1)
    w2 = m.decode(lat.cuda(), mel_ref=mel)
    assert torch.equal(w1, w2)
