"""tcgen05 implicit-GEMM conv layers (bf16 operands, fp32 accumulate) against a float64 conv of the
same bf16-rounded operands.  With exactly representable operands the only differences are the
fp32 accumulation order and the final bf16 rounding of the output."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import index_tts_ipex_b200 as pkg
    return pkg


def _bf(x):
    return x.to(torch.bfloat16)


def _check(y, ref):
    err = (y.double().cpu() - ref).abs()
    bound = ref.abs() * 2.0 ** -8 + 1e-3          # half-ulp bf16 rounding + accumulation-order slack
    assert float((err - bound).max()) <= 0, (float(err.max()), float(ref.abs().max()))


@pytest.mark.parametrize("Cin,Cout,T,K,dil", [
    (64, 64, 256, 1, 1), (64, 64, 256, 3, 1), (96, 96, 300, 3, 1), (96, 96, 1000, 11, 5), (24, 24, 700, 7, 3),
    (48, 48, 513, 3, 5), (192, 192, 130, 7, 1), (384, 384, 64, 3, 3), (768, 768, 40, 3, 1), (1280, 1536, 20, 7, 1),
    (128, 16, 257, 3, 1)])
def test_conv1d_umma_vs_float64(P, Cin, Cout, T, K, dil):
    gen = torch.Generator().manual_seed(Cin + 7 * T + K)
    B = 2
    x = _bf(torch.randn(B, Cin, T, generator=gen))
    w = _bf(torch.randn(Cout, Cin, K, generator=gen) / (Cin * K) ** 0.5).float()
    b = torch.randn(Cout, generator=gen)
    r1 = _bf(torch.randn(B, Cout, T, generator=gen))
    r2 = _bf(torch.randn(B, Cout, T, generator=gen))
    ref = (F.conv1d(x.double(), w.double(), b.double(), dilation=dil, padding=dil * (K - 1) // 2)
           + r1.double() + r2.double()) / 3
    L = P.capi.lib()
    xd, wd, bd, r1d, r2d = x.cuda(), w.cuda(), b.cuda(), r1.cuda(), r2.cuda()
    y = torch.empty(B, Cout, T, device="cuda", dtype=torch.bfloat16)
    P.capi.check(L.bvg_conv1d_umma_fwd(y.data_ptr(), xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), r1d.data_ptr(),
                                       r2d.data_ptr(), 1.0 / 3, B, Cin, Cout, T, K, dil,
                                       torch.cuda.current_stream().cuda_stream), "bvg_conv1d_umma_fwd")
    torch.cuda.synchronize()
    _check(y, ref)


@pytest.mark.parametrize("Cin,Cout,Tin,K,u", [(1536, 768, 20, 8, 4), (768, 384, 70, 8, 4), (384, 192, 129, 4, 4),
                                               (192, 96, 300, 4, 4), (96, 48, 515, 4, 2), (48, 24, 1000, 4, 2)])
def test_convtr1d_umma_vs_float64(P, Cin, Cout, Tin, K, u):
    gen = torch.Generator().manual_seed(Cin + Tin)
    B = 2
    x = _bf(torch.randn(B, Cin, Tin, generator=gen))
    w = _bf(torch.randn(Cin, Cout, K, generator=gen) / (Cin * K / u) ** 0.5).float()
    b = torch.randn(Cout, generator=gen)
    cond = torch.randn(B, Cout, generator=gen)
    ref = F.conv_transpose1d(x.double(), w.double(), b.double(), stride=u, padding=(K - u) // 2) + cond.double().unsqueeze(-1)
    L = P.capi.lib()
    xd, wd, bd, cd = x.cuda(), w.cuda(), b.cuda(), cond.cuda()
    y = torch.empty(B, Cout, Tin * u, device="cuda", dtype=torch.bfloat16)
    P.capi.check(L.bvg_convtr1d_umma_fwd(y.data_ptr(), xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), cd.data_ptr(), B,
                                         B, Cin, Cout, Tin, K, u, torch.cuda.current_stream().cuda_stream),
                 "bvg_convtr1d_umma_fwd")
    torch.cuda.synchronize()
    _check(y, ref)


@pytest.mark.parametrize("Cn,T", [(24, 1), (24, 2), (24, 5), (8, 12), (48, 13), (96, 31), (64, 100), (192, 257),
                                  (24, 4097), (48, 2040), (200, 520)])
def test_act1d_c8t_vs_oracle(P, Cn, T):
    """The channel-chunked Activation1d kernel of the bf16 path, edge lengths included; its input halo
    rows are poisoned by the entry point, so the result must not depend on them."""
    from oracle import bigvgan_oracle as O
    gen = torch.Generator().manual_seed(Cn * 13 + T)
    x = _bf(torch.randn(2, Cn, T, generator=gen) * 1.5)
    a = torch.randn(Cn, generator=gen) * 0.5
    b = torch.randn(Cn, generator=gen) * 0.5
    ref = O.act1d(x.double(), a.double(), b.double())
    y = torch.empty(2, Cn, T, device="cuda", dtype=torch.bfloat16)
    xd, ad, bd = x.cuda(), a.cuda(), b.cuda()
    P.capi.check(P.capi.lib().bvg_act1d_c8t_impl_fwd(y.data_ptr(), xd.data_ptr(), ad.data_ptr(), bd.data_ptr(), 2, Cn, T,
                                                     1, torch.cuda.current_stream().cuda_stream), "bvg_act1d_c8t_impl_fwd")
    torch.cuda.synchronize()
    err = (y.double().cpu() - ref).abs()
    assert float((err - (ref.abs() * 2.0 ** -8 + 1e-4)).max()) <= 0, float(err.max())
    # and it (the CUDA-core stencil, impl = 1) agrees with the plain-layout kernel to the last bit of the bf16 output
    y2 = P.anti_alias_activation_forward(xd, None, None, ad, bd)
    assert torch.equal(y, y2)


@pytest.mark.parametrize("Cin,Cout,T,K,dil", [(96, 96, 1000, 3, 1), (96, 96, 700, 11, 5), (96, 96, 260, 7, 3), (48, 48, 2049, 11, 5),
                                               (48, 48, 513, 3, 1), (24, 24, 4100, 7, 1), (24, 24, 300, 11, 3), (24, 24, 5, 3, 1),
                                               (96, 96, 13, 3, 1), (64, 128, 257, 3, 1), (12, 12, 700, 7, 3), (40, 24, 300, 3, 1),
                                               (6, 6, 1500, 11, 5), (96, 96, 30000, 3, 1), (48, 48, 40000, 7, 3),
                                               (24, 24, 70000, 11, 1), (96, 96, 33000, 11, 5), (192, 192, 700, 3, 1),
                                               (192, 192, 21000, 7, 3), (192, 192, 5000, 11, 5), (128, 128, 1000, 3, 5)])
def test_fused_actconv_equals_unfused(P, Cin, Cout, T, K, dil):
    """The fused Activation1d->conv kernel must reproduce the two-kernel path (same stencil, same products; bit for
    bit when the MMA order is the same), including at the sequence edges and across tile boundaries."""
    # (the test entry point accepts output blocks up to 256 channels: C = 192 qualifies too; the decode path stops at 128)
    gen = torch.Generator().manual_seed(Cin * 3 + T + K)
    B = 2
    x = _bf(torch.randn(B, Cin, T, generator=gen) * 1.5).cuda()
    a = (torch.randn(Cin, generator=gen) * 0.5).cuda()
    b = (torch.randn(Cin, generator=gen) * 0.5).cuda()
    w = _bf(torch.randn(Cout, Cin, K, generator=gen) / (Cin * K) ** 0.5).float().cuda()
    bias = torch.randn(Cout, generator=gen).cuda()
    r1 = _bf(torch.randn(B, Cout, T, generator=gen)).cuda()
    L = P.capi.lib()
    st = torch.cuda.current_stream().cuda_stream
    act = torch.empty_like(x)
    P.capi.check(L.bvg_act1d_c8t_impl_fwd(act.data_ptr(), x.data_ptr(), a.data_ptr(), b.data_ptr(), B, Cin, T, 1, st))
    y_ref = torch.empty(B, Cout, T, device="cuda", dtype=torch.bfloat16)
    P.capi.check(L.bvg_conv1d_umma_fwd(y_ref.data_ptr(), act.data_ptr(), w.data_ptr(), bias.data_ptr(), r1.data_ptr(), None,
                                       0.5, B, Cin, Cout, T, K, dil, st))
    y = torch.empty_like(y_ref)
    P.capi.check(L.bvg_actconv_umma_fwd(y.data_ptr(), x.data_ptr(), a.data_ptr(), b.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                        r1.data_ptr(), 0.5, B, Cin, Cout, T, K, dil, st), "bvg_actconv_umma_fwd")
    torch.cuda.synchronize()
    if Cin <= 48:
        assert torch.equal(y, y_ref), float((y.float() - y_ref.float()).abs().max())
    else:
        # two 48-channel blocks instead of a 64 + 32 split: same products, different fp32 summation order, so a
        # few outputs may land on the neighbouring bf16 value
        d = (y.float() - y_ref.float()).abs()
        mag = torch.maximum(y.float().abs(), y_ref.float().abs())
        assert float((d - (mag * 2.0 ** -7 + 2e-5)).max()) <= 0, float(d.max())
        assert float((d > 0).float().mean()) < 0.02
