"""One fused Activation1d->conv launch shape (for ncu): python tools/actconv_tc_one.py C T K dil B res impl"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import index_tts_ipex_b200 as P

L = P.capi.lib()
C, T, K, dil, B, res, impl = [int(v) for v in (sys.argv[1:8] if len(sys.argv) >= 8 else "96 60160 3 1 32 1 2".split())]
x = torch.randn(B, C, T, device="cuda").bfloat16()
w = torch.randn(C, C, K, device="cuda") / (C * K) ** 0.5
b = torch.randn(C, device="cuda"); al = torch.randn(C, device="cuda") * 0.3; be = torch.randn(C, device="cuda") * 0.3
r1 = torch.randn(B, C, T, device="cuda").bfloat16() if res >= 1 else None
r2 = torch.randn(B, C, T, device="cuda").bfloat16() if res >= 2 else None
y = torch.empty_like(x)
for it in range(2):
    P.capi.check(L.bvg_actconv_impl_fwd(y.data_ptr(), x.data_ptr(), al.data_ptr(), be.data_ptr(), w.data_ptr(), b.data_ptr(),
                                        r1.data_ptr() if r1 is not None else None, r2.data_ptr() if r2 is not None else None,
                                        1.0, B, C, C, T, K, dil, impl, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
