"""Reference-mel front end and per-voice speaker-embedding cache.

Mirrors `indextts.utils.feature_extractors.MelSpectrogramFeatures` (feature_extractors.py:24-50) as called from
`IndexTTS.extract_features` (infer.py:82-93): same constructor arguments, `forward(audio) -> [B, n_mels, frames]`
log-mel.  The STFT / filterbank / log run in one CUDA kernel (csrc/mel.cu) behind `bvg_mel_frontend`; this file only holds
the constant mel filterbank (torchaudio.functional.melscale_fbanks restated) and the cache.  No CPU fallback."""
import collections
import hashlib
import math

import numpy as np
import torch

from . import capi


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> np.ndarray:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') -> [n_freqs, n_mels] (float32, computed in float64)."""
    all_freqs = np.linspace(0.0, sample_rate // 2, n_freqs)
    hz_to_mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)        # noqa: E731
    m_pts = np.linspace(hz_to_mel(f_min), hz_to_mel(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up)).astype(np.float32)


class MelSpectrogramFeatures(torch.nn.Module):
    """Drop-in for the reference class of the same name (feature_extractors.py:24-50); CUDA tensors only."""

    def __init__(self, sample_rate=24000, n_fft=1024, hop_length=256, win_length=None, n_mels=100, mel_fmin=0,
                 mel_fmax=None, normalize=False, padding="center"):
        super().__init__()
        if padding not in ["center", "same"]:
            raise ValueError("Padding must be 'center' or 'same'.")
        if padding != "center" or normalize or (win_length not in (None, n_fft)):
            raise NotImplementedError("only the configuration IndexTTS uses is built: padding='center', normalize=False, "
                                      "win_length=n_fft (infer.py:90 constructs MelSpectrogramFeatures() with defaults)")
        self.sample_rate, self.n_fft, self.hop_length, self.n_mels = sample_rate, n_fft, hop_length, n_mels
        fmax = float(sample_rate // 2) if mel_fmax is None else float(mel_fmax)
        self.register_buffer("fb", torch.from_numpy(melscale_fbanks_htk(n_fft // 2 + 1, float(mel_fmin), fmax, n_mels, sample_rate)),
                             persistent=False)

    def frames(self, n_samples: int) -> int:
        return int(capi.lib().bvg_mel_frames(int(n_samples), int(self.hop_length)))

    def forward_btc(self, audio: torch.Tensor) -> torch.Tensor:
        """audio [B, L] (or [L]) fp32 CUDA -> log-mel [B, frames, n_mels]: the layout the vocoder's mel_ref argument takes
        (infer.py:204 passes cond_mel.transpose(1, 2))."""
        if audio.dim() == 1:
            audio = audio[None]
        if audio.device.type != "cuda":
            raise RuntimeError("MelSpectrogramFeatures runs on CUDA tensors only (no CPU fallback in this package)")
        audio = audio.contiguous().float()
        fb = self.fb.to(audio.device)
        B, L = audio.shape
        mel = torch.empty(B, self.frames(L), self.n_mels, device=audio.device, dtype=torch.float32)
        with torch.cuda.device(audio.device):
            capi.check(capi.lib().bvg_mel_frontend(mel.data_ptr(), audio.data_ptr(), fb.data_ptr(), B, L, self.n_fft,
                                                   self.hop_length, self.n_mels, torch.cuda.current_stream().cuda_stream),
                       "bvg_mel_frontend")
        return mel

    def forward(self, audio: torch.Tensor, **kwargs) -> torch.Tensor:
        """The reference's return layout: [B, n_mels, frames]."""
        return self.forward_btc(audio).transpose(1, 2)


class SpeakerEmbeddingCache:
    """Per-voice cache of ECAPA speaker embeddings (the reference recomputes the embedding in every BigVGAN.forward,
    models.py:205-212, although it only depends on the prompt; its web UI caches the prompt mels, webui.py:199-221).
    Keys are caller-chosen voice ids, or a digest of the prompt when none is given.  LRU, bounded."""

    def __init__(self, max_voices: int = 256):
        self.max_voices = max_voices
        self._d = collections.OrderedDict()
        self.hits = 0
        self.misses = 0

    @staticmethod
    def digest(t: torch.Tensor) -> str:
        return hashlib.sha1(t.detach().to("cpu", torch.float32).contiguous().numpy().tobytes()).hexdigest()

    def get(self, key):
        v = self._d.get(key)
        if v is not None:
            self._d.move_to_end(key)
            self.hits += 1
        return v

    def put(self, key, emb: torch.Tensor):
        self._d[key] = emb
        self._d.move_to_end(key)
        while len(self._d) > self.max_voices:
            self._d.popitem(last=False)

    def __len__(self):
        return len(self._d)

    def clear(self):
        self._d.clear()
