"""CPU-side tests (no GPU): the C-ABI library loads and exports every declared symbol, the host
module keeps the reference's constructor / state-dict contract, and there is no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

from oracle import bigvgan_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def P():
    import index_tts_ipex_b200 as pkg
    return pkg


def test_library_exports_every_header_symbol(P):
    hdr = open(os.path.join(ROOT, "include", "bigvgan_b200.h")).read()
    declared = set(re.findall(r"BVG_API[^;(]*?\b(bvg_\w+)\s*\(", hdr))
    assert len(declared) >= 14
    assert declared == set(P.capi.SIGNATURES), declared ^ set(P.capi.SIGNATURES)
    lib = ctypes.CDLL(P.capi.library_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in P.capi.lib().bvg_version()


def test_config_struct_matches_header(P):
    # 1+1+1+8+8+1+4+12+1+1+1+1+1 int32 fields
    assert ctypes.sizeof(P.capi.BvgConfig) == 4 * 41


def test_arg_validation_without_gpu(P):
    L = P.capi.lib()
    rc = L.bvg_act1d_fwd(None, None, None, None, None, None, 1, 4, 16, 0, 1, None)
    assert rc != 0 and b"null" in L.bvg_last_error()
    bad = (ctypes.c_float * 12)(*([0.1] * 12))
    rc = L.bvg_act1d_fwd(ctypes.c_void_p(16), ctypes.c_void_p(32), ctypes.c_void_p(16), ctypes.c_void_p(16), bad, None, 1, 4, 16, 0, 1, None)
    assert rc != 0 and b"tap" in L.bvg_last_error()
    with pytest.raises(RuntimeError):
        P.capi.check(rc, "bvg_act1d_fwd")
    assert L.bvg_workspace_bytes(None, 1, 1, 1, 0) == 0


def test_state_dict_contract(P):
    """infer.py:61-66: Generator(h, use_cuda_kernel) -> load_state_dict(strict) -> eval -> remove_weight_norm."""
    h = O.small_config()
    m = P.Generator(h, use_cuda_kernel=True)
    assert h["use_cuda_kernel"] is True                                  # models.py:140 mutates h
    keys = [k for k, _, _ in O.state_dict_schema(h)]
    assert list(m.state_dict().keys()) == keys and len(keys) == 1029
    sd = O.make_state_dict(h, 11, "wild")
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert m.eval() is m
    bad = dict(sd)
    bad.pop("conv_pre.bias")
    with pytest.raises(RuntimeError):
        P.Generator(O.small_config()).load_state_dict(bad, strict=True)
    m.remove_weight_norm()
    folded = O.fold_weight_norm(sd)
    got = m.state_dict()
    assert list(got.keys()) == list(folded.keys()) and len(got) == 913
    for k in ("conv_pre.weight", "ups.2.0.weight", "resblocks.7.convs2.1.weight", "conv_post.weight"):
        assert torch.allclose(got[k], folded[k], atol=1e-7, rtol=1e-6), k
    # a post-fold state dict loads into a post-fold module (Checkpoint/resume row of SURVEY.md §5)
    m2 = P.Generator(O.small_config())
    m2.remove_weight_norm()
    m2.load_state_dict(folded, strict=True)
    assert torch.equal(m2.state_dict()["conv_pre.weight"], folded["conv_pre.weight"])


def test_unsupported_configs_raise(P):
    h = O.small_config(); h.activation = "relu"
    with pytest.raises(NotImplementedError):
        P.BigVGAN(h)
    h = O.small_config(); h.resblock = "2"
    with pytest.raises(NotImplementedError):
        P.BigVGAN(h)


def test_no_cpu_fallback(P):
    m = P.BigVGAN(O.small_config()).eval()
    latent, mel = O.synthetic_inputs(O.small_config(), 1, 4, 20)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(latent, mel)
    with pytest.raises(RuntimeError, match="CUDA"):
        P.anti_alias_activation_forward(torch.randn(1, 2, 8), None, None, torch.zeros(2), torch.zeros(2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "index-tts-ipex_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f


def test_dtype_codes_match_header(P):
    hdr = open(os.path.join(ROOT, "include", "bigvgan_b200.h")).read()
    m = re.search(r"enum \{ BVG_F32 = (\d+), BVG_BF16 = (\d+), BVG_F16 = (\d+), BVG_F32X3 = (\d+) \};", hdr)
    assert m, "dtype enum not found in the header"
    assert tuple(int(v) for v in m.groups()) == (P.capi.BVG_F32, P.capi.BVG_BF16, P.capi.BVG_F16, P.capi.BVG_F32X3)
    h = O.small_config()
    m_ = P.BigVGAN(h, use_cuda_kernel=True)
    for prec, code in (("fp32", P.capi.BVG_F32), ("bf16", P.capi.BVG_BF16), ("fp32x3", P.capi.BVG_F32X3)):
        m_.precision = prec
        assert m_._dtype_code() == code
    m_.precision = "fp16"
    with pytest.raises(RuntimeError):
        m_._dtype_code()


def test_decode_ragged_grouping_logic(P):
    """Host logic of decode_ragged without a GPU: equal lengths are batched together, one speaker-encoder pass per
    distinct reference, outputs come back in input order."""
    h = O.small_config()
    m = P.BigVGAN(h, use_cuda_kernel=True)
    calls = {"spk": 0, "decode": []}

    def fake_spk(mel):
        calls["spk"] += 1
        return mel.mean(dim=(1, 2), keepdim=True).expand(-1, 1, int(h.speaker_embedding_dim)).clone()

    def fake_decode(x, mel_ref=None, spk=None, pcm16=False, halo=(0, 0), workspace=None):
        calls["decode"].append(tuple(x.shape))
        # "waveform" = per-utterance latent mean + speaker mean, repeated over T0 * upsample samples
        v = x.mean(dim=(1, 2)) + spk.reshape(x.shape[0], -1).mean(dim=1)
        return v.view(-1, 1, 1).expand(-1, 1, x.shape[1] * m.total_upsample).clone()

    m.speaker_embed, m.decode = fake_spk, fake_decode
    gen = torch.Generator().manual_seed(0)
    lens = [4, 7, 4, 2, 7]
    lats = [torch.randn(t, int(h.gpt_dim), generator=gen) for t in lens]
    mel_a, mel_b = torch.randn(9, int(h.num_mels), generator=gen), torch.randn(11, int(h.num_mels), generator=gen)
    mels = [mel_a, mel_b, mel_a, mel_b, mel_a]
    outs = m.decode_ragged(lats, mels)
    assert calls["spk"] == 2
    assert sorted(calls["decode"]) == sorted([(2, 4, int(h.gpt_dim)), (2, 7, int(h.gpt_dim)), (1, 2, int(h.gpt_dim))])
    for x, mel, y in zip(lats, mels, outs):
        assert y.shape == (1, x.shape[0] * m.total_upsample)
        assert torch.allclose(y[0, 0], x.mean() + mel.mean())
    assert m.decode_ragged([], mels) == []
    with pytest.raises(RuntimeError):
        m.decode_ragged(lats, mels[:2])


def test_tc_kernels_issue_mma_from_uniform_registers():
    """actconv_tc_kernel's speed depends on ptxas proving the tcgen05.mma operands warp-uniform (uniform-datapath issue, ~10
    instructions per MMA); when it cannot, every MMA goes through an ELECT / R2UR.BROADCAST loop and the kernel runs 2x
    slower (profiles/README.md, round 2).  Unrelated code changes have flipped this, so the built library is checked."""
    import re
    import shutil
    import subprocess
    import pytest
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    import index_tts_ipex_b200 as pkg
    sass = subprocess.run(["cuobjdump", "-sass", pkg.build.lib_path()], stdout=subprocess.PIPE, text=True, check=True).stdout
    cur, counts = None, {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = [0, 0]
        elif cur:
            counts[cur][0] += len(re.findall(r"\bUTC\w*MMA", line))
            counts[cur][1] += len(re.findall(r"\bR2UR\.BROADCAST", line))
    for name in ("actconv_tc_kernel", "act1d_tc_kernel"):
        k = [v for n, v in counts.items() if name in n]
        assert k and k[0][0] >= 6, (name, k)
        assert k[0][1] <= 8, f"{name}: {k[0][1]} R2UR.BROADCAST -- the MMA issue loops fell off the uniform datapath"
