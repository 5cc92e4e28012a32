"""Per-role cycle accounting of the tensor-core Activation1d kernel (csrc/act1d_tc.cu); needs a B200."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import index_tts_ipex_b200 as P  # noqa: E402

L = P.capi.lib()
names = ["iss_wait_x", "iss_wait_ufree", "iss_wait_afull", "iss_wait_yfree", "iss_total", "snk_wait_ufull", "snk_wait_afree", "-",
         "snk_total", "sto_wait_yfull", "sto_wait_bar", "sto_total", "-", "iss_ns", "prologue", "kernel_total"]
shapes = [(768, 940, 32), (384, 3760, 32), (192, 15040, 32), (96, 60160, 32), (48, 120320, 32), (24, 240640, 32)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (C, T, B) in shapes:
    x = (torch.randn(B, C, T, device="cuda") * 1.5).bfloat16()
    al = torch.randn(C, device="cuda") * 0.3
    be = torch.randn(C, device="cuda") * 0.3
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    for it in range(2):
        L.bvg_debug_set_umma_counters(dbg.data_ptr() if it else None)
        P.capi.check(L.bvg_act1d_c8t_impl_fwd(y.data_ptr(), x.data_ptr(), al.data_ptr(), be.data_ptr(), B, C, T, 2, st))
        torch.cuda.synchronize()
    L.bvg_debug_set_umma_counters(None)
    P.capi.profile_begin()
    for it in range(5):
        P.capi.check(L.bvg_act1d_c8t_impl_fwd(y.data_ptr(), x.data_ptr(), al.data_ptr(), be.data_ptr(), B, C, T, 2, st))
    torch.cuda.synchronize()
    ms = P.capi.profile_end()["act1d"][0] / 5
    m = dbg.view(148, 16).double().mean(0).tolist()
    gbs = 4.0 * B * C * T / ms / 1e6
    print(f"C={C} T={T} B={B}: {ms*1e3:.0f} us {gbs:.0f} GB/s  " +
          "  ".join(f"{n}={v/1e3:.0f}k" for n, v in zip(names, m) if n != "-") + f"  sm_clock={m[4]/max(m[13],1):.3f} GHz", flush=True)
