"""One standalone Activation1d launch shape (for ncu): python tools/act_one.py C T B dtype(fp32|bf16)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import index_tts_ipex_b200 as P
C, T, B = [int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else "768 1048576 1".split())]
dt = torch.float32 if (len(sys.argv) < 5 or sys.argv[4] == "fp32") else torch.bfloat16
x = torch.randn(B, C, T, device="cuda", dtype=dt)
al = (torch.randn(C, device="cuda") * 0.5).float(); be = (torch.randn(C, device="cuda") * 0.5).float()
for _ in range(3):
    y = P.anti_alias_activation_forward(x, None, None, al, be, precise=False)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
