"""CPU oracle for the BigVGAN vocoder decode path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement, in plain PyTorch functional ops, of the
reference's `bigvgan(latent, mel_ref)` path (reference = cunkai/index-tts-ipex,
`indextts/BigVGAN/models.py:201-250`).  It exists so the CUDA product path can be checked
against something that runs anywhere (the reference itself is not on the GPU box).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import this module.  The product package (`index-tts-ipex_b200/`) never imports it,
and has no CPU fallback: it fails loudly when its CUDA library is missing.

Parity status: the reference's own tests hold NO golden vectors for the vocoder
(SURVEY.md §4), so the oracle is pinned against outputs of the reference module itself,
generated in the build container by `tests/golden/make_golden.py` (imports
`/root/reference`) and committed under `tests/golden/*.npz`.  `tests/test_oracle_golden.py`
checks this oracle against every one of those fixtures.

Each function cites the reference file:line it restates (paths relative to
`/root/reference/indextts/BigVGAN/`).
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# config
# --------------------------------------------------------------------------------------
class AttrDict(dict):
    """Mutable mapping with attribute access, the duck type `models.py:132-197` needs
    (item-set at :140, `.get` at :46, attribute reads everywhere else)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def indextts15_config() -> AttrDict:
    """IndexTTS-1.5 `config.yaml::bigvgan` (external to the reference tree)."""
    return AttrDict(
        resblock="1",
        upsample_rates=[4, 4, 4, 4, 2, 2],
        upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
        upsample_initial_channel=1536,
        resblock_kernel_sizes=[3, 7, 11],
        resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
        gpt_dim=1280,
        activation="snakebeta",
        snake_logscale=True,
        feat_upsample=False,
        cond_d_vector_in_each_upsampling_layer=True,
        speaker_embedding_dim=512,
        num_mels=100,
        sampling_rate=24000,
        hop_size=256,
    )


def small_config() -> AttrDict:
    """A narrow generator with the same topology (fast CPU tests, odd channel counts)."""
    h = indextts15_config()
    h.update(upsample_initial_channel=192, gpt_dim=40, speaker_embedding_dim=64, num_mels=20)
    return h


# --------------------------------------------------------------------------------------
# kaiser-sinc taps  (alias_free_torch/filter.py:29-58)
# --------------------------------------------------------------------------------------
def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> np.ndarray:
    """float64 restatement of filter.py:29-58 (torch.kaiser_window + torch.sinc there)."""
    even = kernel_size % 2 == 0
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    n = np.arange(kernel_size, dtype=np.float64)
    # non-periodic kaiser window: I0(beta*sqrt(1-((n-(N-1)/2)/((N-1)/2))^2))/I0(beta)
    r = (n - (kernel_size - 1) / 2.0) / ((kernel_size - 1) / 2.0)
    window = np.i0(beta * np.sqrt(np.clip(1.0 - r * r, 0.0, None))) / np.i0(beta)
    if even:
        time = np.arange(-half_size, half_size, dtype=np.float64) + 0.5
    else:
        time = np.arange(kernel_size, dtype=np.float64) - half_size
    if cutoff == 0:
        return np.zeros(kernel_size)
    filt = 2 * cutoff * window * np.sinc(2 * cutoff * time)
    return filt / filt.sum()


# The exact fp32 values the reference registers as `upsample.filter` /
# `downsample.lowpass.filter` (computed there in fp32 by torch.kaiser_window * torch.sinc;
# read back from the reference module by tests/golden/make_golden.py).  The float64
# derivation above agrees to 3e-8.
REF_TAPS_F32 = tuple(float.fromhex(v) for v in (
    "0x1.09f0c2p-9", "0x1.33ac8cp-7", "-0x1.a28108p-6", "-0x1.d8544cp-5", "0x1.075110p-3",
    "0x1.c5d8cap-2", "0x1.c5d8cap-2", "0x1.075110p-3", "-0x1.d8544cp-5", "-0x1.a28108p-6",
    "0x1.33ac8cp-7", "0x1.09f0c2p-9"))


def act1d_taps() -> np.ndarray:
    """The one filter every Activation1d in the generator uses (up and down are identical):
    kaiser_sinc_filter1d(0.25, 0.3, 12)  (resample.py:19-21, :41-44), as stored (fp32)."""
    return np.asarray(REF_TAPS_F32, dtype=np.float64)


# --------------------------------------------------------------------------------------
# Activation1d closed form  (alias_free_torch/act.py:24-29, resample.py:25-33,46-49,
# filter.py:87-96, activations.py:109-122)
# --------------------------------------------------------------------------------------
def act1d(x: torch.Tensor, alpha: torch.Tensor, beta: torch.Tensor, taps=None,
          logscale: bool = True, mid_dtype=None) -> torch.Tensor:
    """x [B,C,T] -> [B,C,T].  2x kaiser-sinc upsample (replicate pad 5/5, x ratio) ->
    SnakeBeta -> 12-tap stride-2 lowpass (replicate pad 5/6 of the ACTIVATED signal).
    `mid_dtype` (tests only) rounds the activated 2x signal to that dtype before the down filter: the model of a
    kernel that keeps this intermediate in reduced precision."""
    B, C, T = x.shape
    f = torch.as_tensor(act1d_taps() if taps is None else np.asarray(taps, dtype=np.float64),
                        dtype=x.dtype, device=x.device)
    # --- UpSample1d (resample.py:25-33): pad 5 replicate, conv_transpose stride 2, x2, crop 15
    #     polyphase: u[2j]   = 2*sum_{d=-3..2} f[5-2d] x[clamp(j+d)]
    #                u[2j+1] = 2*sum_{d=-2..3} f[6-2d] x[clamp(j+d)]
    idx = torch.arange(-3, T + 3, device=x.device).clamp_(0, T - 1)
    xp = x[..., idx]                                   # [B,C,T+6], xp[j+3] = x[clamp(j)]
    even = torch.zeros_like(x)
    odd = torch.zeros_like(x)
    for d in range(-3, 3):
        even = even + f[5 - 2 * d] * xp[..., d + 3:d + 3 + T]
    for d in range(-2, 4):
        odd = odd + f[6 - 2 * d] * xp[..., d + 3:d + 3 + T]
    u = torch.stack((2.0 * even, 2.0 * odd), dim=-1).reshape(B, C, 2 * T)
    # --- SnakeBeta (activations.py:109-122)
    a_ = alpha.to(x.dtype).view(1, C, 1)
    b_ = beta.to(x.dtype).view(1, C, 1)
    if logscale:
        a_ = torch.exp(a_)
        b_ = torch.exp(b_)
    u = u + (1.0 / (b_ + 1e-9)) * torch.sin(u * a_) ** 2
    if mid_dtype is not None:
        u = u.to(mid_dtype).to(x.dtype)
    # --- DownSample1d (filter.py:87-96): replicate pad 5 left / 6 right, stride 2
    idx2 = torch.arange(-5, 2 * T + 6, device=x.device).clamp_(0, 2 * T - 1)
    ap = u[..., idx2]                                  # [B,C,2T+11]
    y = torch.zeros_like(x)
    for k in range(12):
        y = y + f[k] * ap[..., k:k + 2 * T:2]
    return y


def act1d_numpy(x: np.ndarray, alpha_log: np.ndarray, beta_log: np.ndarray) -> np.ndarray:
    """Scalar-loop float64 restatement of the same closed form (independent of torch ops);
    x [C,T].  Used for tiny-T edge cases."""
    f = act1d_taps()
    C, T = x.shape
    y = np.zeros((C, T))
    for c in range(C):
        ea = math.exp(float(alpha_log[c]))
        ib = 1.0 / (math.exp(float(beta_log[c])) + 1e-9)
        u = np.zeros(2 * T)
        for j in range(T):
            e = 0.0
            for d in range(-3, 3):
                e += f[5 - 2 * d] * x[c, min(max(j + d, 0), T - 1)]
            o = 0.0
            for d in range(-2, 4):
                o += f[6 - 2 * d] * x[c, min(max(j + d, 0), T - 1)]
            u[2 * j] = 2 * e
            u[2 * j + 1] = 2 * o
        a = u + ib * np.sin(ea * u) ** 2
        for t in range(T):
            s = 0.0
            for k in range(12):
                s += f[k] * a[min(max(2 * t + k - 5, 0), 2 * T - 1)]
            y[c, t] = s
    return y


# --------------------------------------------------------------------------------------
# state-dict schema and deterministic synthetic weights  (models.py:132-197 ctor wiring)
# --------------------------------------------------------------------------------------
def _ecapa_schema(num_mels: int, emb: int):
    """speaker_encoder.* keys  (ECAPA_TDNN.py:461-541; channels [512]*4+[1536], k [5,3,3,3,1])."""
    out = []

    def conv(p, co, ci, k):
        out.append((p + ".conv.weight", (co, ci, k), "w"))
        out.append((p + ".conv.bias", (co,), "b"))

    def bn(p, c):
        out.append((p + ".norm.weight", (c,), "bn_w"))
        out.append((p + ".norm.bias", (c,), "bn_b"))
        out.append((p + ".norm.running_mean", (c,), "bn_m"))
        out.append((p + ".norm.running_var", (c,), "bn_v"))
        out.append((p + ".norm.num_batches_tracked", (), "bn_n"))

    def tdnn(p, ci, co, k):
        conv(p + ".conv", co, ci, k)
        bn(p + ".norm", co)

    P = "speaker_encoder."
    tdnn(P + "blocks.0", num_mels, 512, 5)
    for i in (1, 2, 3):
        b = P + f"blocks.{i}"
        tdnn(b + ".tdnn1", 512, 512, 1)
        for j in range(7):
            tdnn(b + f".res2net_block.blocks.{j}", 64, 64, 3)
        tdnn(b + ".tdnn2", 512, 512, 1)
        conv(b + ".se_block.conv1", 128, 512, 1)
        conv(b + ".se_block.conv2", 512, 128, 1)
    tdnn(P + "mfa", 1536, 1536, 1)
    tdnn(P + "asp.tdnn", 4608, 128, 1)
    conv(P + "asp.conv", 1536, 128, 1)
    out.append((P + "asp_bn.norm.weight", (3072,), "bn_w"))
    out.append((P + "asp_bn.norm.bias", (3072,), "bn_b"))
    out.append((P + "asp_bn.norm.running_mean", (3072,), "bn_m"))
    out.append((P + "asp_bn.norm.running_var", (3072,), "bn_v"))
    out.append((P + "asp_bn.norm.num_batches_tracked", (), "bn_n"))
    conv(P + "fc", emb, 3072, 1)
    return out


def state_dict_schema(h):
    """Ordered [(key, shape, kind)] of `BigVGAN(h).state_dict()` before remove_weight_norm
    (what `bigvgan_generator.pth["generator"]` holds).  Order follows
    module registration order in models.py:132-197."""
    out = []
    C0 = h.upsample_initial_channel

    def wn(p, shape):
        out.append((p + ".bias", (shape[0] if "ups." not in p else shape[1],), "b"))
        out.append((p + ".weight_g", (shape[0], 1, 1), "g"))
        out.append((p + ".weight_v", shape, "v"))

    def act(p, c):
        out.append((p + ".act.alpha", (c,), "alpha"))
        if getattr(h, "activation", "snakebeta") != "snake":      # Snake has alpha only (activations.py:25-47)
            out.append((p + ".act.beta", (c,), "beta"))
        out.append((p + ".upsample.filter", (1, 1, 12), "filt"))
        out.append((p + ".downsample.lowpass.filter", (1, 1, 12), "filt"))

    wn("conv_pre", (C0, h.gpt_dim, 7))
    for i, k in enumerate(h.upsample_kernel_sizes):
        wn(f"ups.{i}.0", (C0 >> i, C0 >> (i + 1), k))
    n = 0
    for i in range(len(h.upsample_rates)):
        ch = C0 >> (i + 1)
        for k in h.resblock_kernel_sizes:
            for grp in ("convs1", "convs2"):
                for m in range(3):
                    wn(f"resblocks.{n}.{grp}.{m}", (ch, ch, k))
            for m in range(6):
                act(f"resblocks.{n}.activations.{m}", ch)
            n += 1
    ch = C0 >> len(h.upsample_rates)
    act("activation_post", ch)
    wn("conv_post", (1, ch, 7))
    out += _ecapa_schema(h.num_mels, h.speaker_embedding_dim)
    out.append(("cond_layer.weight", (C0, h.speaker_embedding_dim, 1), "w"))
    out.append(("cond_layer.bias", (C0,), "b"))
    for i in range(len(h.upsample_rates)):
        out.append((f"conds.{i}.weight", (C0 >> (i + 1), h.speaker_embedding_dim, 1), "w"))
        out.append((f"conds.{i}.bias", (C0 >> (i + 1),), "b"))
    return out


def make_state_dict(h, seed: int = 0, mode: str = "tame") -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic weights in the reference's key schema (CPU generator, so the
    same bytes in the build container and on the GPU box).

    mode "tame": the distributions the reference ctor produces with default init
                 (kaiming-uniform v, g = ||v||, alpha = beta = 0, BN identity stats).
    mode "stress": "wild" with alpha ~ N(0, 1.5) and beta ~ N(0, 1) (large snake arguments).
    mode "wild": additionally randomises g (x U[0.7,1.3]), alpha/beta ~ N(0,0.5) and the BN
                 affine/running stats, so weight-norm folding, log-scale snake parameters
                 and BN folding are actually exercised (SURVEY.md §8(c) extra checks)."""
    g = torch.Generator().manual_seed(int(seed))
    stress = mode == "stress"          # "wild" with alpha ~ N(0, 1.5), beta ~ N(0, 1): snake arguments of hundreds of radians
    wild = mode == "wild" or stress
    taps = torch.tensor(act1d_taps(), dtype=torch.float32).view(1, 1, 12)
    sd = OrderedDict()
    pending_v = None
    sch = state_dict_schema(h)
    shapes = {k: s for k, s, _ in sch}
    for key, shape, kind in sch:
        if kind in ("v", "w"):
            fan_in = shape[1] * shape[2]
            bound = 1.0 / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "b":
            # bias bound uses the fan-in of the matching weight
            wkey = key[:-4] + ("weight_v" if (key[:-4] + "weight_v") in shapes else "weight")
            ws = shapes[wkey]
            bound = 1.0 / math.sqrt(ws[1] * ws[2])
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "g":
            sd[key] = None  # filled after v is known
            pending_v = key
        elif kind in ("alpha", "beta"):
            sd[key] = torch.randn(shape, generator=g) * ((1.5 if kind == "alpha" else 1.0) if stress else 0.5) if wild \
                else torch.zeros(shape)
        elif kind == "filt":
            sd[key] = taps.clone()
        elif kind == "bn_w":
            sd[key] = 1.0 + 0.2 * torch.randn(shape, generator=g) if wild else torch.ones(shape)
        elif kind == "bn_b":
            sd[key] = 0.1 * torch.randn(shape, generator=g) if wild else torch.zeros(shape)
        elif kind == "bn_m":
            sd[key] = 0.1 * torch.randn(shape, generator=g) if wild else torch.zeros(shape)
        elif kind == "bn_v":
            sd[key] = 0.5 + torch.rand(shape, generator=g) if wild else torch.ones(shape)
        elif kind == "bn_n":
            sd[key] = torch.zeros((), dtype=torch.int64)
        else:
            raise AssertionError(kind)
        if kind == "v":
            gkey = key[:-1] + "g"
            v = sd[key]
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).view(-1, 1, 1)
            if wild:
                nrm = nrm * (0.7 + 0.6 * torch.rand(nrm.shape, generator=g))
            sd[gkey] = nrm
            assert pending_v == gkey
    return sd


def state_dict_digest(sd) -> str:
    """sha256 over all tensors' bytes in key order (pins generator determinism across boxes)."""
    hsh = hashlib.sha256()
    for k, v in sd.items():
        hsh.update(k.encode())
        hsh.update(v.detach().cpu().contiguous().numpy().tobytes())
    return hsh.hexdigest()


def fold_weight_norm(sd):
    """w = g * v / ||v||, norm over dims (1,2) per dim-0 slice (old-style
    torch.nn.utils.weight_norm, dim=0; models.py:252-260 -> remove_weight_norm).
    Accepts already-folded dicts (`.weight` keys) unchanged."""
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            continue
        if k.endswith(".weight_v"):
            g = sd[k[:-1] + "g"]
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).view(-1, 1, 1)
            out[k[:-2]] = v * (g / nrm)
        else:
            out[k] = v
    return out


# --------------------------------------------------------------------------------------
# ECAPA-TDNN speaker encoder  (ECAPA_TDNN.py:543-581)
# --------------------------------------------------------------------------------------
def _conv_same_reflect(x, w, b, dilation=1):
    """nnet/CNN.py:411-456 + :458-488 + :519-545: reflect 'same' padding then Conv1d."""
    k = w.shape[-1]
    pad = dilation * (k - 1) // 2
    if pad:
        x = F.pad(x, (pad, pad), mode="reflect")
    return F.conv1d(x, w, b, dilation=dilation)


def _bn_eval(x, sd, p, eps=1e-5):
    """nnet/normalization.py:75-108 -> nn.BatchNorm1d in eval mode."""
    w, b = sd[p + ".weight"], sd[p + ".bias"]
    m, v = sd[p + ".running_mean"], sd[p + ".running_var"]
    scale = w / torch.sqrt(v + eps)
    shift = b - m * scale
    return x * scale.view(1, -1, 1) + shift.view(1, -1, 1)


def _tdnn(x, sd, p, dilation=1):
    """TDNNBlock.forward ECAPA_TDNN.py:126-128: conv -> ReLU -> BN (BN after ReLU)."""
    y = _conv_same_reflect(x, sd[p + ".conv.conv.weight"], sd[p + ".conv.conv.bias"], dilation)
    return _bn_eval(torch.relu(y), sd, p + ".norm.norm")


def ecapa_forward(mel: torch.Tensor, sd, prefix: str = "speaker_encoder.") -> torch.Tensor:
    """mel [B,Tm,num_mels] -> [B,1,emb]  (lengths=None path only, as models.py:202 calls it)."""
    P = prefix
    x = mel.transpose(1, 2)
    x = _tdnn(x, sd, P + "blocks.0", 1)
    xl = []
    for i, dil in ((1, 2), (2, 3), (3, 4)):                      # SERes2NetBlock :415-426
        b = P + f"blocks.{i}"
        res = x
        y = _tdnn(x, sd, b + ".tdnn1")
        chunks = torch.chunk(y, 8, dim=1)                          # Res2NetBlock :179-191
        ys = [chunks[0]]
        yi = None
        for j in range(1, 8):
            inp = chunks[j] if j == 1 else chunks[j] + yi
            yi = _tdnn(inp, sd, b + f".res2net_block.blocks.{j - 1}", dil)
            ys.append(yi)
        y = torch.cat(ys, dim=1)
        y = _tdnn(y, sd, b + ".tdnn2")
        s = y.mean(dim=2, keepdim=True)                            # SEBlock :228-242
        s = torch.relu(F.conv1d(s, sd[b + ".se_block.conv1.conv.weight"], sd[b + ".se_block.conv1.conv.bias"]))
        s = torch.sigmoid(F.conv1d(s, sd[b + ".se_block.conv2.conv.weight"], sd[b + ".se_block.conv2.conv.bias"]))
        x = s * y + res
        xl.append(x)
    x = torch.cat(xl, dim=1)
    x = _tdnn(x, sd, P + "mfa")
    # AttentiveStatisticsPooling :282-338 (mask all ones)
    L = x.shape[-1]
    mean = x.mean(dim=2)
    std = torch.sqrt(((x - mean.unsqueeze(2)) ** 2).mean(dim=2).clamp(1e-12))
    attn = torch.cat([x, mean.unsqueeze(2).expand(-1, -1, L), std.unsqueeze(2).expand(-1, -1, L)], dim=1)
    attn = _tdnn(attn, sd, P + "asp.tdnn")
    attn = F.conv1d(torch.tanh(attn), sd[P + "asp.conv.conv.weight"], sd[P + "asp.conv.conv.bias"])
    attn = torch.softmax(attn, dim=2)
    mean = (attn * x).sum(2)
    std = torch.sqrt((attn * (x - mean.unsqueeze(2)) ** 2).sum(2).clamp(1e-12))
    pooled = torch.cat((mean, std), dim=1).unsqueeze(2)
    pooled = _bn_eval(pooled, sd, P + "asp_bn.norm")
    out = F.conv1d(pooled, sd[P + "fc.conv.weight"], sd[P + "fc.conv.bias"])
    return out.transpose(1, 2)


# --------------------------------------------------------------------------------------
# generator  (models.py:201-250)
# --------------------------------------------------------------------------------------
def _act(x, sd, p, h):
    taps = sd[p + ".upsample.filter"].reshape(-1).double().cpu().numpy()
    # Snake (activations.py:49-60): x + 1/(alpha + 1e-9) sin^2(alpha x), i.e. SnakeBeta with beta := alpha
    beta = sd[p + ".act.alpha"] if getattr(h, "activation", "snakebeta") == "snake" else sd[p + ".act.beta"]
    return act1d(x, sd[p + ".act.alpha"], beta, taps, logscale=bool(h.snake_logscale))


def amp_block1(x, sd, p, h, k, dils):
    """AMPBlock1.forward models.py:65-74."""
    for m, d in enumerate(dils):
        xt = _act(x, sd, f"{p}.activations.{2 * m}", h)
        xt = F.conv1d(xt, sd[f"{p}.convs1.{m}.weight"], sd[f"{p}.convs1.{m}.bias"],
                      dilation=d, padding=(k * d - d) // 2)
        xt = _act(xt, sd, f"{p}.activations.{2 * m + 1}", h)
        xt = F.conv1d(xt, sd[f"{p}.convs2.{m}.weight"], sd[f"{p}.convs2.{m}.bias"],
                      dilation=1, padding=(k - 1) // 2)
        x = xt + x
    return x


def bigvgan_forward(latent: torch.Tensor, mel_ref: torch.Tensor, sd, h, spk=None,
                    return_spk: bool = False):
    """latent [B,T0,gpt_dim], mel_ref [B',Tm,num_mels] -> wav [B,1,T0*prod(upsample_rates)].

    `sd` may be pre- or post-fold.  dtype/device follow the inputs (cast `sd` yourself for
    float64 runs).  `spk` [B',1,emb] short-circuits the speaker encoder."""
    if any(k.endswith(".weight_v") for k in sd):
        sd = fold_weight_norm(sd)
    if spk is None:
        spk = ecapa_forward(mel_ref, sd)
    s = spk.transpose(1, 2)                                        # [B',emb,1]
    x = latent.transpose(1, 2)                                     # models.py:220
    x = F.conv1d(x, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    x = x + F.conv1d(s, sd["cond_layer.weight"], sd["cond_layer.bias"])
    nk = len(h.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
        x = F.conv_transpose1d(x, sd[f"ups.{i}.0.weight"], sd[f"ups.{i}.0.bias"],
                               stride=u, padding=(k - u) // 2)
        if h.cond_d_vector_in_each_upsampling_layer:
            x = x + F.conv1d(s, sd[f"conds.{i}.weight"], sd[f"conds.{i}.bias"])
        xs = None
        for j, (rk, rd) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
            y = amp_block1(x, sd, f"resblocks.{i * nk + j}", h, rk, rd)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = _act(x, sd, "activation_post", h)
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    x = torch.tanh(x)
    return (x, spk) if return_spk else x


def synthetic_inputs(h, B: int, T0: int, Tm: int = 281, seed: int = 1, Bm: int | None = None):
    """latent ~ N(0,1) (a LayerNorm output), mel ~ 2.5*N(0,1) - 0.3."""
    g = torch.Generator().manual_seed(int(seed))
    latent = torch.randn(B, T0, h.gpt_dim, generator=g)
    mel = torch.randn(B if Bm is None else Bm, Tm, h.num_mels, generator=g) * 2.5 - 0.3
    return latent, mel


def snr_db(ref: torch.Tensor, y: torch.Tensor) -> float:
    ref = ref.double().flatten()
    y = y.double().flatten()
    return float(10.0 * torch.log10((ref ** 2).sum() / ((ref - y) ** 2).sum().clamp_min(1e-300)))
