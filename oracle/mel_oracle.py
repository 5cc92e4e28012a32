"""TEST INFRASTRUCTURE ONLY (see oracle/bigvgan_oracle.py): CPU restatement of the reference's mel front end.

MelSpectrogramFeatures.forward (indextts/utils/feature_extractors.py:24-50): torchaudio.transforms.MelSpectrogram(
sample_rate 24000, n_fft 1024, hop 256, win_length = n_fft, hann window (periodic), center = True -> reflect pad n_fft/2,
power = 1, norm None, mel_scale 'htk', f_min 0, f_max sr/2, n_mels 100) then safe_log (utils/common.py:110:
log(clip(x, min = 1e-7))).  numpy float64; pinned against outputs of the reference class itself
(tests/golden/make_mel_golden.py -> tests/golden/mel_cases.npz)."""
import math

import numpy as np


def melscale_fbanks_htk(n_freqs, f_min, f_max, n_mels, sample_rate):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk')."""
    all_freqs = np.linspace(0.0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    return np.maximum(0.0, np.minimum(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))


def log_mel(audio, sample_rate=24000, n_fft=1024, hop=256, n_mels=100, clip=1e-7):
    """audio [B, L] -> log-mel [B, n_mels, frames] (the reference's layout), float64."""
    audio = np.asarray(audio, dtype=np.float64)
    if audio.ndim == 1:
        audio = audio[None]
    B, L = audio.shape
    pad = n_fft // 2
    x = np.pad(audio, ((0, 0), (pad, pad)), mode="reflect")
    frames = L // hop + 1
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)           # torch.hann_window(periodic=True)
    idx = np.arange(frames)[:, None] * hop + np.arange(n_fft)[None, :]
    spec = np.abs(np.fft.rfft(x[:, idx] * win, axis=-1))                          # [B, frames, n_freq], power = 1
    fb = melscale_fbanks_htk(n_fft // 2 + 1, 0.0, sample_rate / 2.0, n_mels, sample_rate)
    mel = spec @ fb                                                               # [B, frames, n_mels]
    return np.log(np.maximum(mel, clip)).transpose(0, 2, 1)


def synthetic_prompt(n_samples, seed=0, sample_rate=24000):
    """A speech-like deterministic test signal: a few decaying harmonics with vibrato + noise bursts + a silent gap."""
    rng = np.random.RandomState(seed)
    t = np.arange(n_samples) / sample_rate
    f0 = 140.0 + 25.0 * np.sin(2 * np.pi * 3.1 * t)
    ph = 2 * np.pi * np.cumsum(f0) / sample_rate
    x = sum((0.6 / k) * np.sin(k * ph + rng.uniform(0, 6.28)) for k in range(1, 12))
    x = x * (0.55 + 0.45 * np.sin(2 * np.pi * 1.7 * t)) + 0.02 * rng.randn(n_samples)
    x[n_samples // 3: n_samples // 3 + 900] = 0.0                                  # silence: exercises the 1e-7 clip
    return (0.3 * x).astype(np.float32)
