"""Per-role cycle accounting of the persistent tcgen05 conv kernel on one layer shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import index_tts_ipex_b200 as P

L = P.capi.lib()
names = ["prod_wait", "mma_wait_x", "mma_wait_w", "mma_wait_tmem", "mma_issue", "mma_total", "epi_wait", "epi_busy"]
for (C, T, K, dil, B, res) in [(768, 940, 3, 1, 32, 0), (768, 940, 7, 3, 32, 1), (768, 940, 11, 5, 32, 0), (384, 3760, 3, 1, 32, 0),
                               (384, 3760, 7, 3, 32, 1), (384, 3760, 11, 5, 32, 0), (192, 15040, 3, 1, 32, 0), (192, 15040, 7, 3, 32, 1),
                               (192, 15040, 11, 5, 32, 0), (96, 60160, 3, 1, 32, 1), (24, 240640, 11, 1, 32, 0)]:
    x = torch.randn(B, C, T, device="cuda").bfloat16()
    w = torch.randn(C, C, K, device="cuda") / (C * K) ** 0.5
    b = torch.randn(C, device="cuda")
    r1 = torch.randn(B, C, T, device="cuda").bfloat16() if res else None
    y = torch.empty_like(x)
    dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    for it in range(2):
        L.bvg_debug_set_umma_counters(dbg.data_ptr() if it else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        P.capi.check(L.bvg_conv1d_umma_fwd(y.data_ptr(), x.data_ptr(), w.data_ptr(), b.data_ptr(),
                                           r1.data_ptr() if res else None, None, 1.0, B, C, C, T, K, dil,
                                           torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
    L.bvg_debug_set_umma_counters(None)
    P.capi.profile_begin()
    for it in range(3):
        P.capi.check(L.bvg_conv1d_umma_fwd(y.data_ptr(), x.data_ptr(), w.data_ptr(), b.data_ptr(),
                                           r1.data_ptr() if res else None, None, 1.0, B, C, C, T, K, dil,
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ms = P.capi.profile_end()["conv1d"][0] / 3
    d = dbg.view(148, 8).double().cpu()
    m = d.mean(0)
    print(f"C={C} T={T} K={K} res={res}: {ms * 1e3:.0f} us  {2.0 * B * C * C * K * T / ms / 1e9:.0f} TFLOP/s  " + "  ".join(f"{n}={v/1e3:.0f}k" for n, v in zip(names, m.tolist())))
