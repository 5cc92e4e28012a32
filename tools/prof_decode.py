"""Profiling driver: one bf16 decode of B x 10 s utterances between cudaProfilerStart/Stop.

    ncu --profile-from-start off ... python tools/prof_decode.py [--batch 32] [--precision bf16]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import index_tts_ipex_b200 as pkg  # noqa: E402
from oracle import bigvgan_oracle as O  # noqa: E402  (synthetic weights / inputs only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=235)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--steps", type=int, default=1)
    a = ap.parse_args()
    h = O.indextts15_config()
    m = pkg.BigVGAN(h, use_cuda_kernel=True)
    m.load_state_dict(O.make_state_dict(h, 0, "tame"), strict=True)
    m = m.to("cuda:0").eval()
    m.remove_weight_norm()
    m.precision = a.precision
    lat, mel = O.synthetic_inputs(h, a.batch, a.frames, 281, seed=1)
    lat, mel = lat.cuda(), mel.cuda()
    for _ in range(a.warmup):
        m.decode(lat, mel_ref=mel)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(a.steps):
        m.decode(lat, mel_ref=mel)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    main()
