"""Per-role cycle accounting of the fused Activation1d->conv kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
os.environ.setdefault("BVG_FUSE_MAX_NB", "256")      # also measure C = 192 (off by default in the decode path)
import index_tts_ipex_b200 as P
L = P.capi.lib()
names = ["st_wait_raw", "st_wait_x", "st_busy", "mma_wait_x", "mma_wait_tmem", "mma_total", "epi_wait", "epi_busy", "mma_wait_w", "mma_ns"]
for (C, T, K, dil, B, res) in [(192, 15040, 3, 1, 32, 1), (192, 15040, 7, 3, 32, 0), (192, 15040, 11, 5, 32, 0), (96, 60160, 3, 1, 32, 0), (96, 60160, 3, 1, 32, 1), (96, 60160, 7, 3, 32, 0), (96, 60160, 11, 5, 32, 0), (48, 120320, 3, 1, 32, 0), (48, 120320, 11, 5, 32, 1),
                               (24, 240640, 3, 1, 32, 1), (24, 240640, 11, 1, 32, 0)]:
    x = torch.randn(B, C, T, device="cuda").bfloat16()
    w = torch.randn(C, C, K, device="cuda") / (C * K) ** 0.5
    b = torch.randn(C, device="cuda"); al = torch.randn(C, device="cuda") * 0.3; be = torch.randn(C, device="cuda") * 0.3
    r1 = torch.randn(B, C, T, device="cuda").bfloat16() if res else None
    y = torch.empty_like(x)
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    for it in range(2):
        L.bvg_debug_set_umma_counters(dbg.data_ptr() if it else None)
        P.capi.check(L.bvg_actconv_umma_fwd(y.data_ptr(), x.data_ptr(), al.data_ptr(), be.data_ptr(), w.data_ptr(), b.data_ptr(),
                                            r1.data_ptr() if res else None, 1.0, B, C, C, T, K, dil, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
    L.bvg_debug_set_umma_counters(None)
    P.capi.profile_begin()
    for it in range(3):
        P.capi.check(L.bvg_actconv_umma_fwd(y.data_ptr(), x.data_ptr(), al.data_ptr(), be.data_ptr(), w.data_ptr(), b.data_ptr(),
                                            r1.data_ptr() if res else None, 1.0, B, C, C, T, K, dil, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    prof = P.capi.profile_end()
    ms = prof["actconv"][0] / 3
    alg = B * C * T * 2 * (2 + res) / 1e9
    m = dbg.view(148, 16).double().mean(0)
    print(f"C={C} T={T} K={K} d={dil} res={res}: {ms*1e3:.0f} us ({alg/ms:.2f} TB/s alg)  " + "  ".join(f"{n}={v/1e3:.0f}k" for n, v in zip(names, m.tolist())) + f"  sm_clock={m[5].item()/max(m[9].item(),1):.3f} GHz")
