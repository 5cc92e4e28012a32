"""One conv_umma_kernel layer shape, a few launches (for `ncu -k regex:conv_umma_kernel`): C T K dil B res."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import index_tts_ipex_b200 as P

C, T, K, dil, B, res = (int(a) for a in sys.argv[1:7])
L = P.capi.lib()
x = torch.randn(B, C, T, device="cuda").bfloat16()
w = torch.randn(C, C, K, device="cuda") / (C * K) ** 0.5
b = torch.randn(C, device="cuda")
r1 = torch.randn(B, C, T, device="cuda").bfloat16() if res else None
y = torch.empty_like(x)
for it in range(3):
    P.capi.check(L.bvg_conv1d_umma_fwd(y.data_ptr(), x.data_ptr(), w.data_ptr(), b.data_ptr(), r1.data_ptr() if res else None, None,
                                       1.0, B, C, C, T, K, dil, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok")
