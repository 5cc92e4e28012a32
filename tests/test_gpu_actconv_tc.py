"""Fused Activation1d -> Conv1d with tensor-core FIRs (csrc/actconv_tc.cu) against the oracle.

Reference op chain: AMPBlock1.forward of BigVGAN/models.py:65-74 (xt = c1(a1(x)); xt = c2(a2(xt)); x = xt + x), restated in
oracle/bigvgan_oracle.py (act1d) + torch conv1d in float64.  Arithmetic of the kernel: the Activation1d of act1d_tc.cu
(bf16 input, hi + lo bf16 up-taps, fp32 snake, fp16 activated 2x signal, fp16 down-taps), result rounded to bf16 as the
conv's A operand, bf16 weights, fp32 accumulation, + bias + residual(s), * scale, bf16 output."""
import pytest
import torch
import torch.nn.functional as F

from oracle import bigvgan_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import index_tts_ipex_b200 as pkg
    return pkg


def _run(P, x, a, b, w, bias, r1, r2, scale, K, dil, impl):
    B, Cin, T = x.shape
    Cout = w.shape[0]
    y = torch.empty(B, Cout, T, device="cuda", dtype=torch.bfloat16)
    d = [t.cuda() if t is not None else None for t in (x, a, b, w, bias, r1, r2)]
    ptr = [t.data_ptr() if t is not None else None for t in d]
    P.capi.check(P.capi.lib().bvg_actconv_impl_fwd(y.data_ptr(), ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], ptr[5], ptr[6], scale,
                                                   B, Cin, Cout, T, K, dil, impl, torch.cuda.current_stream().cuda_stream),
                 "bvg_actconv_impl_fwd")
    torch.cuda.synchronize()
    return y.cpu()


def _case(Cin, T, K, B, seed, res=1):
    gen = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, Cin, T, generator=gen) * 1.5).to(torch.bfloat16)
    a = torch.randn(Cin, generator=gen) * 0.5
    b = torch.randn(Cin, generator=gen) * 0.5
    w = (torch.randn(Cin, Cin, K, generator=gen) / (Cin * K) ** 0.5).to(torch.bfloat16).float()
    bias = torch.randn(Cin, generator=gen)
    r1 = torch.randn(B, Cin, T, generator=gen).to(torch.bfloat16) if res >= 1 else None
    r2 = torch.randn(B, Cin, T, generator=gen).to(torch.bfloat16) if res >= 2 else None
    return x, a, b, w, bias, r1, r2


def _oracle(x, a, b, w, bias, r1, r2, scale, K, dil):
    act = O.act1d(x.double(), a.double(), b.double(), mid_dtype=torch.float16).to(torch.bfloat16).double()
    ref = F.conv1d(act, w.double(), bias.double(), dilation=dil, padding=dil * (K - 1) // 2)
    if r1 is not None:
        ref = ref + r1.double()
    if r2 is not None:
        ref = ref + r2.double()
    return ref * scale, act


# every geometry of the kernel: C = 96 (1 segment), 48 (2), 24 (4 segments, padding chunk); every (k, dilation) of the
# generator's AMP blocks -> conv halo classes LH = 8 / 16 / 32; resident weights and the weight ring (C = 96, k = 7 / 11);
# lengths that are not multiples of 32 / 128, several ranges per utterance, utterances that end inside a range group
@pytest.mark.parametrize("Cin,T,K,dil,B,res", [
    (96, 1000, 3, 1, 2, 1), (96, 2049, 3, 5, 1, 0), (96, 1500, 7, 1, 2, 2), (96, 3000, 7, 3, 1, 1), (96, 1027, 11, 1, 2, 1),
    (96, 4100, 11, 3, 1, 1), (96, 2500, 11, 5, 2, 2), (48, 2049, 7, 3, 2, 1), (48, 700, 11, 5, 3, 0), (48, 5000, 3, 3, 1, 2),
    (24, 4100, 11, 1, 2, 1), (24, 513, 3, 1, 2, 0), (24, 9000, 7, 5, 1, 2), (24, 3333, 11, 5, 2, 1),
    (96, 33000, 11, 5, 1, 1), (48, 40000, 7, 3, 1, 1), (24, 70001, 3, 1, 1, 1), (96, 512, 7, 1, 5, 1)])
def test_actconv_tc_vs_oracle(P, Cin, T, K, dil, B, res):
    x, a, b, w, bias, r1, r2 = _case(Cin, T, K, B, Cin * 7 + T + K + dil, res)
    y = _run(P, x, a, b, w, bias, r1, r2, 0.5, K, dil, 2).double()
    ref, act = _oracle(x, a, b, w, bias, r1, r2, 0.5, K, dil)
    err = (y - ref).abs()
    # output half-ulp + a bf16 ulp flip of the activated operand here and there (|w| ~ 1/sqrt(Cin K))
    bound = ref.abs() * 2.0 ** -8 + 2.0 ** -7 * float(act.abs().max()) * (K ** 0.5) / (Cin * K) ** 0.5 + 2e-3
    bad = (err - bound) > 0
    assert not bool(bad.any()), (float(err.max()), bad.nonzero()[:8].tolist())
    assert O.snr_db(ref.float(), y.float()) >= 45.0


@pytest.mark.parametrize("Cin,T,K,dil", [(96, 1500, 3, 3), (48, 2222, 11, 1), (24, 6000, 7, 1)])
def test_actconv_tc_vs_stencil_kernel(P, Cin, T, K, dil):
    """Same layer through the round-1 fused kernel (FIRs on the FP32 pipe): both are within the oracle's bounds, so they
    agree up to bf16 ulp flips of the activated operand."""
    x, a, b, w, bias, r1, r2 = _case(Cin, T, K, 2, 11 + T, 1)
    y2 = _run(P, x, a, b, w, bias, r1, None, 1.0, K, dil, 2).double()
    y1 = _run(P, x, a, b, w, bias, r1, None, 1.0, K, dil, 1).double()
    ref, act = _oracle(x, a, b, w, bias, r1, None, 1.0, K, dil)
    d = (y2 - y1).abs()
    bound = ref.abs() * 2.0 ** -7 + 2.0 ** -6 * float(act.abs().max()) * (K ** 0.5) / (Cin * K) ** 0.5 + 4e-3
    assert float((d - bound).max()) <= 0, float(d.max())
    assert O.snr_db(ref.float(), y2.float()) >= O.snr_db(ref.float(), y1.float()) - 1.0


def test_actconv_tc_zero_frame_and_dispatch(P):
    """impl 0 (the decode path's choice) takes the tensor-core kernel for qualifying layers and falls back otherwise."""
    x, a, b, w, bias, r1, r2 = _case(96, 1200, 3, 2, 3, 1)
    assert torch.equal(_run(P, x, a, b, w, bias, r1, None, 1.0, 3, 1, 0), _run(P, x, a, b, w, bias, r1, None, 1.0, 3, 1, 2))
    x, a, b, w, bias, r1, r2 = _case(96, 300, 3, 2, 4, 1)                 # too short for the streaming kernel
    assert torch.equal(_run(P, x, a, b, w, bias, r1, None, 1.0, 3, 1, 0), _run(P, x, a, b, w, bias, r1, None, 1.0, 3, 1, 1))
    with pytest.raises(RuntimeError):
        _run(P, x, a, b, w, bias, r1, None, 1.0, 3, 1, 2)
