"""Event trace of actconv_tc_kernel's pipeline on CTA 0 (debug build, BVG_TCF_TRACE=1): clock64 of each hop of blocks 128..191.
    python tools/actconv_tc_trace.py [C T K dil res]"""
import os
import sys

import torch

os.environ["BVG_DEBUG_BUILD"] = "1"
os.environ["BVG_TCF_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import index_tts_ipex_b200 as P  # noqa: E402

L = P.capi.lib()
C, T, K, dil, res = [int(v) for v in (sys.argv[1:6] if len(sys.argv) >= 6 else "96 60160 3 1 0".split())]
B = 32
x = torch.randn(B, C, T, device="cuda").bfloat16()
w = torch.randn(C, C, K, device="cuda") / (C * K) ** 0.5
b = torch.randn(C, device="cuda"); al = torch.randn(C, device="cuda") * 0.3; be = torch.randn(C, device="cuda") * 0.3
r1 = torch.randn(B, C, T, device="cuda").bfloat16() if res >= 1 else None
y = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
dbg = torch.zeros(148 * 16 + 16 * 64, dtype=torch.int64, device="cuda")
for it in range(2):
    L.bvg_debug_set_umma_counters(dbg.data_ptr() if it else None)
    P.capi.check(L.bvg_actconv_impl_fwd(y.data_ptr(), x.data_ptr(), al.data_ptr(), be.data_ptr(), w.data_ptr(), b.data_ptr(),
                                        r1.data_ptr() if r1 is not None else None, None, 1.0, B, C, C, T, K, dil, 2, st))
    torch.cuda.synchronize()
L.bvg_debug_set_umma_counters(None)
tr = dbg[148 * 16:].view(16, 64).cpu()
t0 = int(tr[0, 0])
names = ["up issue", "snake sees U", "snake has a_free+U regs", "snake done", "down issue", "store sees Y", "store done"]
print(f"C={C} K={K} d={dil} res={res}: cycles relative to the up-FIR issue of block 128 (CTA 0)")
print("blk " + " ".join(f"{n:>24s}" for n in names))
for i in range(0, 24):
    print(f"{128+i:3d} " + " ".join(f"{int(tr[e, i]) - t0:24d}" if int(tr[e, i]) else f"{'-':>24s}" for e in range(7)))
print("tile: conv start, conv issued, epilogue sees acc, ld issued, res loaded, ld done, slice 0 done, epilogue done (tiles 32..; block 128 = tile 31.5)")
for i in range(0, 8):
    print(f"{32+i:3d} " + " ".join(f"{int(tr[e, i]) - t0:12d}" if int(tr[e, i]) else f"{'-':>12s}" for e in (7, 8, 9, 11, 12, 13, 14, 10)))
