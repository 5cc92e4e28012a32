"""Times the c8t Activation1d implementations (CUDA-core stencil vs tensor-core FIRs) on the generator's shapes.
Usage: python tools/act_tc_bench.py  (needs a B200).  GB/s = 2 * B*C*T * 2 bytes / time (algorithmic bytes)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import index_tts_ipex_b200 as pkg  # noqa: E402

L = pkg.capi.lib()
PEAK = 6545.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def bench(Cn, T, B, impl, iters=10):
    x = (torch.randn(B, Cn, T, device="cuda") * 1.5).to(torch.bfloat16)
    a = torch.randn(Cn, device="cuda") * 0.5
    b = torch.randn(Cn, device="cuda") * 0.5
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    # the entry point converts plain <-> c8t around the kernel; time only the Activation1d class via the profiler hooks
    for _ in range(2):
        pkg.capi.check(L.bvg_act1d_c8t_impl_fwd(y.data_ptr(), x.data_ptr(), a.data_ptr(), b.data_ptr(), B, Cn, T, impl, st))
    torch.cuda.synchronize()
    pkg.capi.profile_begin()
    for _ in range(iters):
        pkg.capi.check(L.bvg_act1d_c8t_impl_fwd(y.data_ptr(), x.data_ptr(), a.data_ptr(), b.data_ptr(), B, Cn, T, impl, st))
    prof = pkg.capi.profile_end()
    ms = prof["act1d"][0] / iters
    gbs = 2.0 * B * Cn * T * 2 / (ms * 1e-3) / 1e9
    return ms, gbs


if __name__ == "__main__":
    shapes = [(768, 940, 32), (384, 3760, 32), (192, 15040, 32), (96, 60160, 32), (48, 120320, 32), (24, 240640, 32),
              (768, 940, 1), (192, 15040, 1), (24, 240640, 1)]
    for Cn, T, B in shapes:
        r1 = bench(Cn, T, B, 1)
        r2 = bench(Cn, T, B, 2)
        print(f"C={Cn:4d} T={T:7d} B={B:3d}  stencil {r1[0]*1e3:8.1f} us {r1[1]:7.0f} GB/s ({r1[1]/PEAK:.2f})   "
              f"tensor-core {r2[0]*1e3:8.1f} us {r2[1]:7.0f} GB/s ({r2[1]/PEAK:.2f})", flush=True)
