"""BASELINE config 5b: standalone Activation1d HBM GB/s sweep over channels x length x dtype.

Algorithmic bytes = 2*B*C*T*sizeof(dtype) (one read + one write); B is chosen so in+out >= 512 MB (>> the
126 MB L2); alpha, beta ~ N(0, 0.5), x ~ N(0, 1).  Timing: CUDA events around `iters` back-to-back launches
after 3 warm-ups; every launch streams the whole >= 512 MB working set, so nothing survives in L2."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import index_tts_ipex_b200 as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    peak = 6545.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    Cs = [24, 96, 768] if a.quick else [24, 48, 96, 192, 384, 768, 1536]
    Ts = [1 << 12, 1 << 16, 1 << 20] if a.quick else [1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20]
    rows = []
    for dtype in (torch.float32, torch.bfloat16):
        es = 4 if dtype == torch.float32 else 2
        for C in Cs:
            for T in Ts:
                B = max(1, -(-(512 << 20) // (2 * C * T * es)))
                if B > 65535:
                    continue
                x = torch.randn(B, C, T, device="cuda", dtype=dtype)
                al = (torch.randn(C, device="cuda") * 0.5).float()
                be = (torch.randn(C, device="cuda") * 0.5).float()
                for _ in range(3):
                    y = P.anti_alias_activation_forward(x, None, None, al, be, precise=False)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.iters):
                    y = P.anti_alias_activation_forward(x, None, None, al, be, precise=False)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.iters
                gbs = 2.0 * B * C * T * es / (ms * 1e-3) / 1e9
                rows.append(dict(dtype=str(dtype).split(".")[-1], C=C, T=T, B=B, ms=ms, gbs=gbs, frac=gbs / peak))
                print(f"{rows[-1]['dtype']:9s} C={C:5d} T={T:8d} B={B:6d}  {ms:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of {peak:.0f}")
                del x, y
    if a.out:
        json.dump(dict(peak_gbs=peak, rows=rows), open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
