"""Tensor-core Activation1d (csrc/act1d_tc.cu: both anti-alias FIRs as banded-Toeplitz tcgen05.mma, SnakeBeta between
them) against the oracle (alias_free_torch/act.py:24-29 of the reference, restated in oracle/bigvgan_oracle.py).

Arithmetic of the kernel: bf16 input, up-FIR taps split in two bf16 terms (products exact, fp32 accumulation), fp32 snake
with a range-reduced MUFU cosine, activated 2x signal rounded to fp16, fp16 down-FIR taps, fp32 accumulation, bf16
output.  The tight check models the fp16 intermediate in the oracle (`mid_dtype`); the loose one is the plain oracle."""
import pytest
import torch

from oracle import bigvgan_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import index_tts_ipex_b200 as pkg
    return pkg


def _run(P, x, a, b, impl):
    B, Cn, T = x.shape
    y = torch.empty(B, Cn, T, device="cuda", dtype=torch.bfloat16)
    xd, ad, bd = x.cuda(), a.cuda(), b.cuda()
    P.capi.check(P.capi.lib().bvg_act1d_c8t_impl_fwd(y.data_ptr(), xd.data_ptr(), ad.data_ptr(), bd.data_ptr(), B, Cn, T, impl,
                                                     torch.cuda.current_stream().cuda_stream), "bvg_act1d_c8t_impl_fwd")
    torch.cuda.synchronize()
    return y.cpu()


# every lane geometry of the kernel: 16-chunk tiles (C = 768 / 384 / 256), 2 x 8 (192), 4 x 4 (96, 24 with a padding chunk),
# 2 x 6-in-8 (48), odd chunk counts (200 -> 26 chunks, 8, 16, 40); lengths that are not multiples of 32 / 128, one block only,
# many ranges per utterance, B = 1 .. 3
@pytest.mark.parametrize("Cn,T,B", [(24, 4097, 2), (48, 2040, 2), (96, 1000, 3), (192, 257, 2), (200, 520, 2), (384, 300, 1),
                                    (768, 940, 2), (24, 70001, 1), (96, 16389, 2), (8, 256, 1), (16, 999, 2), (40, 1283, 1),
                                    (256, 3760, 1), (192, 15040, 1), (3, 2000, 2), (6, 20480, 2)])
def test_act1d_tc_vs_oracle(P, Cn, T, B):
    gen = torch.Generator().manual_seed(Cn * 13 + T)
    x = (torch.randn(B, Cn, T, generator=gen) * 1.5).to(torch.bfloat16)
    a = torch.randn(Cn, generator=gen) * 0.5
    b = torch.randn(Cn, generator=gen) * 0.5
    y = _run(P, x, a, b, 2).double()
    ref16 = O.act1d(x.double(), a.double(), b.double(), mid_dtype=torch.float16)
    ref = O.act1d(x.double(), a.double(), b.double())
    err16 = (y - ref16).abs()
    amax = float(ref.abs().max())
    # bf16 output rounding (half an ulp <= 2^-8 |ref|) + fp16 taps of the down filter (2^-12 of sum |f||a|) + MUFU slack
    assert float((err16 - (ref16.abs() * 2.0 ** -8 + 2.0 ** -11 * max(1.0, amax))).max()) <= 0, float(err16.max())
    err = (y - ref).abs()
    assert float((err - (ref.abs() * 2.0 ** -8 + 2.0 ** -10 * max(1.0, amax))).max()) <= 0, float(err.max())
    # against the CUDA-core stencil: both are within half a bf16 ulp of the oracle (plus the fp16 intermediate)
    y1 = _run(P, x, a, b, 1).double()
    d = (y - y1).abs()
    assert float((d - (ref.abs() * 2.0 ** -7 + 2.0 ** -10 * max(1.0, amax))).max()) <= 0, float(d.max())
    snr_tc, snr_st = O.snr_db(ref.float(), y.float()), O.snr_db(ref.float(), y1.float())
    assert snr_tc >= snr_st - 1.0, (snr_tc, snr_st)          # the fp16 intermediate costs < 1 dB of the bf16 output's ~50 dB


def test_act1d_tc_large_alpha(P):
    """alpha ~ N(0, 1.5): arguments of the cosine up to several hundred radians (the range reduction's job)."""
    gen = torch.Generator().manual_seed(5)
    Cn, T = 96, 3000
    x = (torch.randn(2, Cn, T, generator=gen) * 2.0).to(torch.bfloat16)
    a = torch.randn(Cn, generator=gen) * 1.5
    b = torch.randn(Cn, generator=gen) * 1.0
    y = _run(P, x, a, b, 2).double()
    ref16 = O.act1d(x.double(), a.double(), b.double(), mid_dtype=torch.float16)
    err16 = (y - ref16).abs()
    amax = float(ref16.abs().max())
    assert float((err16 - (ref16.abs() * 2.0 ** -8 + 2.0 ** -11 * max(1.0, amax))).max()) <= 0, float(err16.max())


def test_act1d_tc_dispatch_and_finite(P):
    """impl 0 is what the decode path takes: the tensor-core kernel for large tensors (>= 10 M elements, T >= 256), the
    CUDA-core stencil for small ones (one utterance: the persistent kernel's fixed ~25 us would dominate)."""
    gen = torch.Generator().manual_seed(9)
    z = torch.zeros(24)
    x = (torch.randn(2, 24, 1000, generator=gen)).to(torch.bfloat16)                 # small: stencil
    assert torch.equal(_run(P, x, z, z, 0), _run(P, x, z, z, 1))
    x = (torch.randn(2, 24, 220000, generator=gen)).to(torch.bfloat16)               # 10.6 M elements: tensor cores
    y = _run(P, x, z, z, 0)
    assert torch.equal(y, _run(P, x, z, z, 2))
    assert torch.isfinite(y.float()).all()
