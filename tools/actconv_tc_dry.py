"""Bottleneck experiments on actconv_tc_kernel (debug build only: BVG_DEBUG_BUILD=1): times a few layer shapes with parts of the
pipeline knocked out through BVG_TCF_DRY (1 no conv MMAs, 2 no up-FIR MMAs, 4 no down-FIR MMAs, 8 snake without the cosine,
16 epilogue without global loads / stores, 32 store warps without stmatrix).  Results of such runs are garbage by design."""
import os
import sys

import torch

os.environ["BVG_DEBUG_BUILD"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import index_tts_ipex_b200 as P  # noqa: E402

L = P.capi.lib()
shapes = [(96, 60160, 3, 1, 32, 0), (96, 60160, 11, 5, 32, 0), (96, 60160, 11, 1, 32, 1), (48, 120320, 7, 1, 32, 1), (24, 240640, 3, 1, 32, 1),
          (24, 240640, 11, 5, 32, 0)]
masks = [0, 1, 6, 7, 8, 16, 32, 1 | 16, 7 | 16, 7 | 8 | 16 | 32, 8 | 16 | 32]
for (C, T, K, dil, B, res) in shapes:
    x = torch.randn(B, C, T, device="cuda").bfloat16()
    w = torch.randn(C, C, K, device="cuda") / (C * K) ** 0.5
    b = torch.randn(C, device="cuda"); al = torch.randn(C, device="cuda") * 0.3; be = torch.randn(C, device="cuda") * 0.3
    r1 = torch.randn(B, C, T, device="cuda").bfloat16() if res >= 1 else None
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for m in masks:
        os.environ["BVG_TCF_DRY"] = str(m)
        def call():
            P.capi.check(L.bvg_actconv_impl_fwd(y.data_ptr(), x.data_ptr(), al.data_ptr(), be.data_ptr(), w.data_ptr(), b.data_ptr(),
                                                r1.data_ptr() if r1 is not None else None, None, 1.0, B, C, C, T, K, dil, 2, st))
        call()
        torch.cuda.synchronize()
        P.capi.profile_begin()
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        out.append((m, P.capi.profile_end()["actconv"][0] / 2))
    print(f"C={C} K={K} d={dil} res={res}: " + "  ".join(f"dry={m}:{ms*1e3:.0f}us" for m, ms in out), flush=True)
