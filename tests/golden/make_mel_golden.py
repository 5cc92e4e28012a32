"""Generates tests/golden/mel_cases.npz from the REFERENCE's own MelSpectrogramFeatures (feature_extractors.py:24-50).
Run in the build container (needs /root/reference and torchaudio):  python tests/golden/make_mel_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from indextts.utils.feature_extractors import MelSpectrogramFeatures  # noqa: E402  (the reference class)
from oracle import mel_oracle as M  # noqa: E402


def main():
    ref = MelSpectrogramFeatures()
    out = {}
    for name, n, seed in (("a", 12000, 0), ("b", 7001, 1), ("c", 600, 2)):      # 0.5 s, an odd length, barely longer than the pad
        audio = torch.from_numpy(M.synthetic_prompt(n, seed))[None]
        with torch.no_grad():
            mel = ref(audio)                                                     # [1, 100, frames]
        out[f"audio_{name}"] = audio.numpy()
        out[f"mel_{name}"] = mel.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mel_cases.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
