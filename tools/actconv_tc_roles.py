"""Times the fused Activation1d->conv kernels on the generator's narrow layer shapes (CUDA events via the profiler hooks) and
prints the tensor-core FIR kernel's per-role wait counters (actconv_tc.cu).

    python tools/actconv_tc_roles.py [quick]        (needs a B200)

GB/s = algorithmic bytes (input + output [+ residuals]) / kernel time; impl 1 = CUDA-core stencil kernel (round 1),
impl 2 = tensor-core FIR kernel."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import index_tts_ipex_b200 as P  # noqa: E402

L = P.capi.lib()
names = ["up_wx", "up_wu", "dn_wa", "dn_wy", "total", "sn_wu", "sn_wa", "st_wy", "st_wf", "cv_ws", "cv_wc", "cv_ww", "ep_wait", "ns",
         "ep_busy", "sn_total"]
PEAK = 6545.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

shapes = []
for C, T in ((96, 60160), (48, 120320), (24, 240640)):
    for K in (3, 7, 11):
        for dil, res in ((1, 0), (3, 0), (5, 0), (1, 1), (1, 2)):
            shapes.append((C, T, K, dil, 32, res))
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    shapes = [s for s in shapes if (s[3], s[5]) in ((1, 1), (5, 0))]

tot = {1: 0.0, 2: 0.0}
for (C, T, K, dil, B, res) in shapes:
    x = torch.randn(B, C, T, device="cuda").bfloat16()
    w = torch.randn(C, C, K, device="cuda") / (C * K) ** 0.5
    b = torch.randn(C, device="cuda"); al = torch.randn(C, device="cuda") * 0.3; be = torch.randn(C, device="cuda") * 0.3
    r1 = torch.randn(B, C, T, device="cuda").bfloat16() if res >= 1 else None
    r2 = torch.randn(B, C, T, device="cuda").bfloat16() if res >= 2 else None
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream

    def call(impl):
        P.capi.check(L.bvg_actconv_impl_fwd(y.data_ptr(), x.data_ptr(), al.data_ptr(), be.data_ptr(), w.data_ptr(), b.data_ptr(),
                                            r1.data_ptr() if r1 is not None else None, r2.data_ptr() if r2 is not None else None,
                                            1.0, B, C, C, T, K, dil, impl, st))
    res_ms = {}
    for impl in (1, 2):
        call(impl)
        torch.cuda.synchronize()
        P.capi.profile_begin()
        for _ in range(3):
            call(impl)
        torch.cuda.synchronize()
        res_ms[impl] = P.capi.profile_end()["actconv"][0] / 3
        tot[impl] += res_ms[impl]
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    L.bvg_debug_set_umma_counters(dbg.data_ptr())
    call(2)
    torch.cuda.synchronize()
    L.bvg_debug_set_umma_counters(None)
    m = dbg.view(148, 16).double().mean(0)
    alg = B * C * T * 2 * (2 + res) / 1e9
    print(f"C={C:3d} K={K:2d} d={dil} res={res}: stencil {res_ms[1]*1e3:6.0f} us  tc {res_ms[2]*1e3:6.0f} us ({alg/res_ms[2]:.2f} TB/s, "
          f"{alg/res_ms[2]*1e3/PEAK:.2f})  " + " ".join(f"{n}={v/1e3:.0f}k" for n, v in zip(names, m.tolist()) if n != "ns")
          + f"  clk={m[4].item()/max(m[13].item(),1):.2f}GHz", flush=True)
print(f"sum over shapes: stencil {tot[1]:.2f} ms, tc {tot[2]:.2f} ms")
