"""Drop-in for `indextts.BigVGAN.models.BigVGAN` (reference models.py:130-275) on B200.

Same constructor (`BigVGAN(h, use_cuda_kernel=...)`, infer.py:61), same state-dict key schema
(`load_state_dict(torch.load("bigvgan_generator.pth")["generator"])`, infer.py:63-64), same
`.to(device).eval()`, `.remove_weight_norm()` and `bigvgan(latent, mel_ref) -> (wav, None)` call
(infer.py:204, :498).  The arithmetic is not PyTorch: every layer runs in libbigvgan_b200
(hand-written sm_100a CUDA behind the C ABI of include/bigvgan_b200.h).  There is no CPU path."""
import ctypes as C
import math
import threading

import torch
import torch.nn as nn

from . import capi

# kaiser_sinc_filter1d(0.25, 0.3, 12) as the reference registers it (fp32 buffer values;
# alias_free_torch/filter.py:29-58, resample.py:19-22)
KAISER_TAPS = tuple(float.fromhex(v) for v in (
    "0x1.09f0c2p-9", "0x1.33ac8cp-7", "-0x1.a28108p-6", "-0x1.d8544cp-5", "0x1.075110p-3",
    "0x1.c5d8cap-2", "0x1.c5d8cap-2", "0x1.075110p-3", "-0x1.d8544cp-5", "-0x1.a28108p-6",
    "0x1.33ac8cp-7", "0x1.09f0c2p-9"))

_ECAPA_C, _ECAPA_MFA, _ECAPA_ATT = 512, 1536, 128


def _hget(h, k):
    return h[k] if isinstance(h, dict) or hasattr(h, "__getitem__") else getattr(h, k)


def _schema(h):
    """[(key, shape, kind)] in the reference's registration order (models.py:132-197 and
    ECAPA_TDNN.py:461-541), weight-normed convs as bias / weight_g / weight_v."""
    out = []
    C0 = int(h.upsample_initial_channel)
    E = int(h.speaker_embedding_dim)

    def wn(p, co, ci, k, transposed=False):
        shape = (ci, co, k) if transposed else (co, ci, k)
        out.append((p + ".bias", (co,), "b", ci * k))
        out.append((p + ".weight_g", (shape[0], 1, 1), "g", 0))
        out.append((p + ".weight_v", shape, "v", shape[1] * k))

    def act(p, c):
        out.append((p + ".act.alpha", (c,), "alpha", 0))
        if h.activation != "snake":                      # Snake has alpha only (activations.py:25-47)
            out.append((p + ".act.beta", (c,), "alpha", 0))
        out.append((p + ".upsample.filter", (1, 1, 12), "filt", 0))
        out.append((p + ".downsample.lowpass.filter", (1, 1, 12), "filt", 0))

    def conv(p, co, ci, k):
        out.append((p + ".weight", (co, ci, k), "v", ci * k))
        out.append((p + ".bias", (co,), "b", ci * k))

    def bn(p, c):
        out.append((p + ".weight", (c,), "one", 0))
        out.append((p + ".bias", (c,), "zero", 0))
        out.append((p + ".running_mean", (c,), "bzero", 0))
        out.append((p + ".running_var", (c,), "bone", 0))
        out.append((p + ".num_batches_tracked", (), "bcount", 0))

    def tdnn(p, ci, co, k):
        conv(p + ".conv.conv", co, ci, k)
        bn(p + ".norm.norm", co)

    wn("conv_pre", C0, int(h.gpt_dim), 7)
    for i, k in enumerate(h.upsample_kernel_sizes):
        wn(f"ups.{i}.0", C0 >> (i + 1), C0 >> i, int(k), transposed=True)
    n = 0
    for i in range(len(h.upsample_rates)):
        ch = C0 >> (i + 1)
        for k in h.resblock_kernel_sizes:
            for grp in ("convs1", "convs2"):
                for m in range(3):
                    wn(f"resblocks.{n}.{grp}.{m}", ch, ch, int(k))
            for m in range(6):
                act(f"resblocks.{n}.activations.{m}", ch)
            n += 1
    ch = C0 >> len(h.upsample_rates)
    act("activation_post", ch)
    wn("conv_post", 1, ch, 7)
    S = "speaker_encoder."
    tdnn(S + "blocks.0", int(h.num_mels), _ECAPA_C, 5)
    for i in (1, 2, 3):
        b = S + f"blocks.{i}"
        tdnn(b + ".tdnn1", _ECAPA_C, _ECAPA_C, 1)
        for j in range(7):
            tdnn(b + f".res2net_block.blocks.{j}", _ECAPA_C // 8, _ECAPA_C // 8, 3)
        tdnn(b + ".tdnn2", _ECAPA_C, _ECAPA_C, 1)
        conv(b + ".se_block.conv1.conv", _ECAPA_ATT, _ECAPA_C, 1)
        conv(b + ".se_block.conv2.conv", _ECAPA_C, _ECAPA_ATT, 1)
    tdnn(S + "mfa", _ECAPA_MFA, _ECAPA_MFA, 1)
    tdnn(S + "asp.tdnn", 3 * _ECAPA_MFA, _ECAPA_ATT, 1)
    conv(S + "asp.conv.conv", _ECAPA_MFA, _ECAPA_ATT, 1)
    bn(S + "asp_bn.norm", 2 * _ECAPA_MFA)
    conv(S + "fc.conv", E, 2 * _ECAPA_MFA, 1)
    conv("cond_layer", C0, E, 1)
    if h.cond_d_vector_in_each_upsampling_layer:
        for i in range(len(h.upsample_rates)):
            conv(f"conds.{i}", C0 >> (i + 1), E, 1)
    return out


class _Node(nn.Module):
    """Plain container; children/params are attached under the reference's names."""


def _attach(root: nn.Module, key: str, tensor: torch.Tensor, is_buffer: bool):
    parts = key.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    if is_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def _leaf(root: nn.Module, key: str):
    parts = key.split(".")
    mod = root
    for p in parts[:-1]:
        mod = mod._modules[p]
    return mod, parts[-1]


class BigVGAN(nn.Module):
    """B200 BigVGAN generator.  See module docstring; reference: models.py:130-275."""

    def __init__(self, h, use_cuda_kernel=False):
        super().__init__()
        self.h = h
        self.h["use_cuda_kernel"] = use_cuda_kernel          # models.py:140 (h must support item-set)
        if h.resblock != "1":
            raise NotImplementedError("only AMPBlock1 (resblock='1') is on the IndexTTS path")   # models.py:152
        if h.activation not in ("snakebeta", "snake"):
            raise NotImplementedError(
                "activation incorrectly specified. check the config file and look for 'activation'.")  # models.py:63,182
        # Snake (activations.py:49-60) is SnakeBeta with beta := alpha; its modules carry no `beta` parameter
        self.snake_only = h.activation == "snake"
        if h.feat_upsample:
            raise NotImplementedError("feat_upsample=True is not used by IndexTTS (models.py:213-218)")
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        self.cond_in_each_up_layer = bool(h.cond_d_vector_in_each_upsampling_layer)
        self.total_upsample = int(math.prod(int(u) for u in h.upsample_rates))
        for rd in h.resblock_dilation_sizes:
            if len(rd) != 3:
                raise NotImplementedError("AMPBlock1 has exactly 3 dilations per block")
        taps = torch.tensor(KAISER_TAPS, dtype=torch.float32).view(1, 1, 12)
        pending_g = {}
        for key, shape, kind, fan_in in _schema(h):
            if kind == "v":
                bound = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape) * 2 - 1) * bound
                _attach(self, key, t, False)
                if key.endswith("weight_v"):
                    g = t.reshape(shape[0], -1).norm(dim=1).view(-1, 1, 1)
                    mod, _ = _leaf(self, key)
                    mod.weight_g.data.copy_(g)
            elif kind == "g":
                _attach(self, key, torch.ones(shape), False)
            elif kind == "b":
                bound = 1.0 / math.sqrt(fan_in)
                _attach(self, key, (torch.rand(shape) * 2 - 1) * bound, False)
            elif kind == "alpha":
                _attach(self, key, torch.zeros(shape) if h.snake_logscale else torch.ones(shape), False)
            elif kind == "filt":
                _attach(self, key, taps.clone(), True)
            elif kind == "one":
                _attach(self, key, torch.ones(shape), False)
            elif kind == "zero":
                _attach(self, key, torch.zeros(shape), False)
            elif kind == "bzero":
                _attach(self, key, torch.zeros(shape), True)
            elif kind == "bone":
                _attach(self, key, torch.ones(shape), True)
            elif kind == "bcount":
                _attach(self, key, torch.zeros((), dtype=torch.long), True)
        del pending_g
        self._plans = {}                 # device index -> bvg_plan*; one plan per device, built on first use
        self._plan_gen = 0               # bumped whenever the weights may have changed: captured CUDA graphs go stale
        self._plan_lock = threading.Lock()
        # None: fp32, or bf16 under torch.autocast; "fp32" / "bf16" to force; "fp32x3": fp32 tensors with the Conv1d
        # layers on the tensor cores (3-term bf16 split, ~2e-6 max-abs against the fp32 path's 5e-7)
        self.precision = None

    # ------------------------------------------------------------------ module plumbing
    def _invalidate(self):
        """Weights (may) have changed: drop every device's plan.  CUDA graphs captured by make_graphed_decode hold raw
        pointers into a plan, so they are marked stale (their run() raises) instead of replaying freed memory."""
        plans = getattr(self, "_plans", None)
        if plans:
            for plan in plans.values():
                capi.lib().bvg_plan_destroy(plan)
            plans.clear()
        self._plan_gen = getattr(self, "_plan_gen", 0) + 1

    @staticmethod
    def _device_index(device) -> int:
        """torch.device('cuda') and torch.device('cuda:0') name the same device: plans are keyed by the index."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("BigVGAN (B200): inputs must live on a CUDA device; there is no CPU path")
        return device.index if device.index is not None else torch.cuda.current_device()

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._invalidate()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def remove_weight_norm(self):
        """models.py:252-260: fold g*v/||v|| into .weight (old-style weight_norm, dim=0)."""
        print('Removing weight norm...')
        self._invalidate()
        for key, _ in list(self.named_parameters()):
            if not key.endswith(".weight_v"):
                continue
            mod, _n = _leaf(self, key)
            v, g = mod.weight_v.data, mod.weight_g.data
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).view(-1, 1, 1)
            w = v * (g / nrm)
            del mod._parameters["weight_g"]
            del mod._parameters["weight_v"]
            mod.register_parameter("weight", nn.Parameter(w, requires_grad=False))
        return self

    def _folded_tensors(self):
        """(key, fp32 tensor) of the post-fold state dict, folding on the fly if still weight-normed."""
        sd = self.state_dict()
        for k, v in sd.items():
            if k.endswith(".weight_g") or k.endswith("num_batches_tracked"):
                continue
            if k.endswith(".weight_v"):
                g = sd[k[:-1] + "g"]
                nrm = v.reshape(v.shape[0], -1).norm(dim=1).view(-1, 1, 1)
                yield k[:-2], (v * (g / nrm)).float()
            else:
                yield k, v.float()
                if self.snake_only and k.endswith(".act.alpha"):
                    yield k[:-5] + "beta", v.float()

    def _ensure_plan(self, device):
        idx = self._device_index(device)          # validates the device BEFORE anything is torn down
        plan = self._plans.get(idx)
        if plan is not None:
            return plan
        with self._plan_lock:
            plan = self._plans.get(idx)
            if plan is not None:
                return plan
            h = self.h
            cfg = capi.BvgConfig()
            cfg.gpt_dim = int(h.gpt_dim)
            cfg.upsample_initial_channel = int(h.upsample_initial_channel)
            cfg.num_upsamples = self.num_upsamples
            for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
                cfg.upsample_rates[i] = int(u)
                cfg.upsample_kernel_sizes[i] = int(k)
            cfg.num_kernels = self.num_kernels
            for j, (k, d) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
                cfg.resblock_kernel_sizes[j] = int(k)
                for m in range(3):
                    cfg.resblock_dilation_sizes[j][m] = int(d[m])
            cfg.speaker_embedding_dim = int(h.speaker_embedding_dim)
            cfg.num_mels = int(h.num_mels)
            cfg.cond_in_each_up_layer = int(self.cond_in_each_up_layer)
            cfg.snake_logscale = int(bool(h.snake_logscale))
            cfg.device = idx
            L = capi.lib()
            plan = C.c_void_p()
            capi.check(L.bvg_plan_create(C.byref(plan), C.byref(cfg)), "bvg_plan_create")
            try:
                for key, t in self._folded_tensors():
                    t = t.detach().to("cpu", torch.float32).contiguous()
                    capi.check(L.bvg_plan_set_tensor(plan, key.encode(), t.data_ptr(), t.numel()), "bvg_plan_set_tensor")
                capi.check(L.bvg_plan_finalize(plan, 1), "bvg_plan_finalize")
            except Exception:
                L.bvg_plan_destroy(plan)
                raise
            self._plans[idx] = plan
            return plan

    # ------------------------------------------------------------------ compute
    def _dtype_code(self):
        p = self.precision
        if p is None:
            p = "bf16" if torch.is_autocast_enabled() else "fp32"
        if p not in ("fp32", "bf16", "fp32x3"):
            raise RuntimeError(f"precision must be 'fp32', 'fp32x3' or 'bf16', got {p!r}")
        return {"fp32": capi.BVG_F32, "bf16": capi.BVG_BF16, "fp32x3": capi.BVG_F32X3}[p]

    def workspace_bytes(self, B, T0, Tm, dtype_code=None):
        if not self._plans:
            raise RuntimeError("plan not built yet; call the model once or _ensure_plan(device)")
        plan = next(iter(self._plans.values()))      # sizes depend on the config only, not on the device
        return int(capi.lib().bvg_workspace_bytes(plan, B, T0, Tm, self._dtype_code() if dtype_code is None else dtype_code))

    @torch.no_grad()
    def speaker_embed(self, mel_ref: torch.Tensor) -> torch.Tensor:
        """ECAPA_TDNN.forward (ECAPA_TDNN.py:543-581): mel [B',Tm,num_mels] -> [B',1,emb]."""
        mel = mel_ref.detach().to(torch.float32).contiguous()
        plan = self._ensure_plan(mel.device)
        Bm, Tm, _ = mel.shape
        spk = torch.empty(Bm, 1, int(self.h.speaker_embedding_dim), device=mel.device, dtype=torch.float32)
        nbytes = self.workspace_bytes(Bm, 1, Tm, capi.BVG_F32)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=mel.device)
        with torch.cuda.device(mel.device):
            st = torch.cuda.current_stream().cuda_stream
            capi.check(capi.lib().bvg_speaker_embed(plan, mel.data_ptr(), Bm, Tm, spk.data_ptr(), ws.data_ptr(),
                                                    nbytes, st), "bvg_speaker_embed")
        return spk

    @torch.no_grad()
    def voice_embedding(self, mel_ref: torch.Tensor = None, audio: torch.Tensor = None, key=None, cache=None) -> torch.Tensor:
        """Speaker embedding [B',1,emb] of a voice prompt, cached per voice.  Pass the prompt either as
        `mel_ref` [B',Tm,num_mels] or as 24 kHz mono `audio` [B',L] (then the mel front end of mel.py runs first, as
        infer.py:82-93 does); `key` names the voice (default: a digest of the prompt); `cache` defaults to the model's own
        SpeakerEmbeddingCache.  The result goes to decode(..., spk=...) and removes the speaker encoder from steady state."""
        from .mel import MelSpectrogramFeatures, SpeakerEmbeddingCache
        if (mel_ref is None) == (audio is None):
            raise ValueError("pass exactly one of mel_ref / audio")
        if cache is None:
            if not hasattr(self, "_spk_cache"):
                self._spk_cache = SpeakerEmbeddingCache()
            cache = self._spk_cache
        src = mel_ref if mel_ref is not None else audio
        if key is None:
            key = SpeakerEmbeddingCache.digest(src)
        key = (key, str(src.device))
        emb = cache.get(key)
        if emb is None:
            cache.misses += 1
            if mel_ref is None:
                if not hasattr(self, "_mel_frontend"):
                    self._mel_frontend = MelSpectrogramFeatures(n_mels=int(self.h.num_mels))
                mel_ref = self._mel_frontend.forward_btc(audio)
            emb = self.speaker_embed(mel_ref)
            cache.put(key, emb)
        return emb

    @torch.no_grad()
    def decode(self, x, mel_ref=None, spk=None, pcm16=False, halo=(0, 0), workspace=None):
        """The device-resident call.  x [B,T0,gpt_dim]; exactly one of mel_ref [B',Tm,num_mels] /
        spk [B',1,emb].  Returns wav fp32 [B,1,L] (or int16 [B,L] when pcm16=True, the fused
        epilogue of infer.py:206-212,234).  `halo=(lo,hi)` marks latent frames whose samples are
        dropped (chunked long-form decode)."""
        if x.dim() != 3 or x.shape[-1] != int(self.h.gpt_dim):
            raise RuntimeError(f"expected latent [B, T0, {int(self.h.gpt_dim)}], got {tuple(x.shape)}")
        if (mel_ref is None) == (spk is None):
            raise RuntimeError("pass exactly one of mel_ref / spk")
        # the GPT hands the latent over in its autocast dtype (gpt/model.py:462-477 under infer.py:194): fp16 / bf16 / fp32
        # [B, T, C] tensors are ingested as they are, anything else is widened to fp32 first
        lat_code = {torch.float32: capi.BVG_F32, torch.bfloat16: capi.BVG_BF16, torch.float16: capi.BVG_F16}.get(x.dtype)
        lat = x.detach().contiguous() if lat_code is not None else x.detach().to(torch.float32).contiguous()
        if lat_code is None:
            lat_code = capi.BVG_F32
        dev = lat.device
        plan = self._ensure_plan(dev)
        B, T0, _ = lat.shape
        if mel_ref is not None:
            cond = mel_ref.detach().to(dev, torch.float32).contiguous()
            if cond.dim() != 3 or cond.shape[-1] != int(self.h.num_mels):
                raise RuntimeError(f"expected mel_ref [B', Tm, {int(self.h.num_mels)}], got {tuple(cond.shape)}")
            Bm, Tm = cond.shape[0], cond.shape[1]
            mel_ptr, spk_ptr = cond.data_ptr(), None
        else:
            cond = spk.detach().to(dev, torch.float32).reshape(spk.shape[0], -1).contiguous()
            Bm, Tm = cond.shape[0], 1
            mel_ptr, spk_ptr = None, cond.data_ptr()
        code = self._dtype_code()
        lo, hi = int(halo[0]), int(halo[1])
        Lout = (T0 - lo - hi) * self.total_upsample
        nbytes = int(capi.lib().bvg_workspace_bytes(plan, B, T0, Tm, code))
        ws = workspace if workspace is not None else torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if ws.numel() < nbytes:
            raise RuntimeError(f"workspace too small: {ws.numel()} < {nbytes}")
        if pcm16:
            out = torch.empty(B, Lout, device=dev, dtype=torch.int16)
            wav_ptr, pcm_ptr = None, out.data_ptr()
        else:
            out = torch.empty(B, 1, Lout, device=dev, dtype=torch.float32)
            wav_ptr, pcm_ptr = out.data_ptr(), None
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            capi.check(capi.lib().bvg_decode_lat(plan, lat.data_ptr(), lat_code, mel_ptr, spk_ptr, B, T0, Bm, Tm, code, wav_ptr,
                                                 pcm_ptr, lo, hi, ws.data_ptr(), ws.numel(), st), "bvg_decode_lat")
        return out

    def forward(self, x, mel_ref, lens=None):
        """models.py:201-250.  Returns (wav [B,1,T0*prod(upsample_rates)], None)."""
        if lens is not None:
            raise NotImplementedError("lens is never passed on the IndexTTS path (infer.py:204,498)")
        wav = self.decode(x, mel_ref=mel_ref)
        if torch.is_autocast_enabled():
            wav = wav.to(torch.get_autocast_dtype("cuda"))     # the reference's last op yields the autocast dtype
        return wav, None

    @torch.no_grad()
    def decode_host(self, latent_cpu, mel_cpu, device, pcm16=False, out=None):
        """End-to-end call with HOST tensors (pinned for async copies): H2D, decode, D2H inside
        the library (bvg_decode_host).  Returns a pinned CPU tensor; pass a pinned `out` of the right shape
        ([B, 1, T0*1024] fp32, or [B, T0*1024] int16 with pcm16) to reuse it across calls -- allocating pinned
        memory costs more than the copies themselves."""
        device = torch.device(device)
        plan = self._ensure_plan(device)
        lat = latent_cpu.to(torch.float32).contiguous()
        mel = mel_cpu.to(torch.float32).contiguous()
        B, T0, _ = lat.shape
        Bm, Tm, _ = mel.shape
        code = self._dtype_code()
        L = T0 * self.total_upsample
        nbytes = int(capi.lib().bvg_workspace_bytes(plan, B, T0, Tm, code))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        shape, dt = ((B, L), torch.int16) if pcm16 else ((B, 1, L), torch.float32)
        if out is None:
            out = torch.empty(shape, dtype=dt, pin_memory=True)
        elif tuple(out.shape) != shape or out.dtype != dt or out.device.type != "cpu" or not out.is_contiguous():
            raise RuntimeError(f"decode_host: out must be a contiguous CPU {dt} tensor of shape {shape}")
        wav_ptr, pcm_ptr = (None, out.data_ptr()) if pcm16 else (out.data_ptr(), None)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream().cuda_stream
            capi.check(capi.lib().bvg_decode_host(plan, lat.data_ptr(), mel.data_ptr(), B, T0, Bm, Tm, code, wav_ptr,
                                                  pcm_ptr, ws.data_ptr(), nbytes, st), "bvg_decode_host")
        return out

    @torch.no_grad()
    def make_graphed_decode(self, B, T0, Tm, Bm=None, device="cuda", pcm16=False, warmup=2):
        """Capture one whole decode (~300 kernel launches) into a CUDA graph for a fixed shape bucket and return
        `run(latent, mel_ref) -> wav`.  The library call allocates nothing and never synchronises, so it is
        capturable as is; inputs are copied into static buffers and the returned tensor is the static output
        (clone it if it must outlive the next call).  This is the low-latency form for B = 1 serving, where launch
        overhead is a visible part of the 8 ms a 10 s utterance takes."""
        device = torch.device(device)
        Bm = B if Bm is None else Bm
        plan = self._ensure_plan(device)
        code = self._dtype_code()
        lat = torch.zeros(B, T0, int(self.h.gpt_dim), device=device)
        mel = torch.zeros(Bm, Tm, int(self.h.num_mels), device=device)
        L = T0 * self.total_upsample
        out = torch.empty(B, L, device=device, dtype=torch.int16) if pcm16 else torch.empty(B, 1, L, device=device)
        nbytes = int(capi.lib().bvg_workspace_bytes(plan, B, T0, Tm, code))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        wav_ptr, pcm_ptr = (None, out.data_ptr()) if pcm16 else (out.data_ptr(), None)

        def enqueue():
            st = torch.cuda.current_stream(device).cuda_stream
            capi.check(capi.lib().bvg_decode(plan, lat.data_ptr(), mel.data_ptr(), None, B, T0, Bm, Tm, code, wav_ptr,
                                             pcm_ptr, 0, 0, ws.data_ptr(), nbytes, st), "bvg_decode")

        with torch.cuda.device(device):
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    enqueue()
            torch.cuda.current_stream(device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                enqueue()

        gen = self._plan_gen

        def run(latent, mel_ref):
            if self._plan_gen != gen:
                raise RuntimeError("graphed decode is stale: the model's weights / device changed after capture "
                                   "(load_state_dict, .to(), remove_weight_norm); call make_graphed_decode again")
            lat.copy_(latent, non_blocking=True)
            mel.copy_(mel_ref, non_blocking=True)
            graph.replay()
            return out

        run.graph = graph
        return run

    @torch.no_grad()
    def decode_long(self, x, mel_ref, chunk_frames=256, halo_frames=36, pcm16=False):
        """Long-form decode in overlapped chunks (BASELINE config 5).  The generator has a finite
        receptive field (<= 35 latent frames per side), so chunks of `chunk_frames`
        frames decoded with `halo_frames` extra frames per side and the halo samples dropped equal
        the unchunked decode.  The speaker embedding is global: computed once, reused."""
        B, T0, _ = x.shape
        spk = self.speaker_embed(mel_ref.to(x.device))
        outs = []
        for s in range(0, T0, chunk_frames):
            e = min(s + chunk_frames, T0)
            lo = min(halo_frames, s)
            hi = min(halo_frames, T0 - e)
            outs.append(self.decode(x[:, s - lo:e + hi], spk=spk, pcm16=pcm16, halo=(lo, hi)))
        return torch.cat(outs, dim=-1)

    @torch.no_grad()
    @torch.no_grad()
    def decode_varlen(self, latents, mel_ref=None, spk=None, pcm16=False, lens=None):
        """True batched vocoding of utterances of different lengths in ONE call (bf16 path; `bvg_decode_varlen`).
        `latents`: a sequence of `[T0_i, gpt_dim]` tensors, or a padded `[B, T0_max, gpt_dim]` tensor with `lens` (ints).
        Exactly one of `mel_ref` `[B' in {1, B}, Tm, num_mels]` / `spk` `[B', 1, emb]`.  Returns `[B, 1, T0_max * up]` fp32
        (or `[B, T0_max * up]` int16): utterance b's first `lens[b] * up` samples are what decoding it alone gives (bit for
        bit when it is long enough to take the same kernels alone: >= 64 latent frames), the rest of its row is zero."""
        if self.precision != "bf16":
            raise RuntimeError("decode_varlen: ragged batches run on the bf16 tensor-core path (set precision='bf16')")
        if (mel_ref is None) == (spk is None):
            raise RuntimeError("pass exactly one of mel_ref / spk")
        if torch.is_tensor(latents):
            if lens is None:
                raise RuntimeError("decode_varlen: a padded latent tensor needs `lens`")
            x = latents.detach()
            lens = [int(v) for v in lens]
        else:
            seq = [t.detach() if t.dim() == 2 else t.detach().squeeze(0) for t in latents]
            lens = [int(t.shape[0]) for t in seq]
            x = torch.nn.utils.rnn.pad_sequence(seq, batch_first=True)
        if x.dim() != 3 or x.shape[-1] != int(self.h.gpt_dim) or len(lens) != x.shape[0]:
            raise RuntimeError(f"decode_varlen: expected [B, T0_max, {int(self.h.gpt_dim)}] latents and B lengths")
        B, T0 = int(x.shape[0]), int(x.shape[1])
        if min(lens) < 1 or max(lens) > T0:
            raise RuntimeError(f"decode_varlen: lengths must be in [1, {T0}]")
        lat_code = {torch.float32: capi.BVG_F32, torch.bfloat16: capi.BVG_BF16, torch.float16: capi.BVG_F16}.get(x.dtype)
        if lat_code is None:
            x, lat_code = x.to(torch.float32), capi.BVG_F32
        x = x.contiguous()
        dev = x.device
        plan = self._ensure_plan(dev)
        lens_dev = torch.tensor(lens, dtype=torch.int32, device=dev)
        if mel_ref is not None:
            cond = mel_ref.detach().to(dev, torch.float32).contiguous()
            Bm, Tm = cond.shape[0], cond.shape[1]
            mel_ptr, spk_ptr = cond.data_ptr(), None
        else:
            cond = spk.detach().to(dev, torch.float32).reshape(spk.shape[0], -1).contiguous()
            Bm, Tm = cond.shape[0], 1
            mel_ptr, spk_ptr = None, cond.data_ptr()
        Lout = T0 * self.total_upsample
        nbytes = int(capi.lib().bvg_workspace_bytes(plan, B, T0, Tm, capi.BVG_BF16))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if pcm16:
            out = torch.empty(B, Lout, device=dev, dtype=torch.int16)
            wav_ptr, pcm_ptr = None, out.data_ptr()
        else:
            out = torch.empty(B, 1, Lout, device=dev, dtype=torch.float32)
            wav_ptr, pcm_ptr = out.data_ptr(), None
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            capi.check(capi.lib().bvg_decode_varlen(plan, x.data_ptr(), lat_code, lens_dev.data_ptr(), mel_ptr, spk_ptr, B, T0, Bm,
                                                    Tm, wav_ptr, pcm_ptr, ws.data_ptr(), ws.numel(), st), "bvg_decode_varlen")
        return out

    def decode_ragged(self, latents, mel_refs, pcm16=False):
        """Batched vocoding of utterances of DIFFERENT lengths (what `infer_fast` works around by concatenating two
        sentences in time at B = 1, `indextts/infer.py:480-503`).  `latents` is a sequence of `[T0_i, gpt_dim]`
        tensors, `mel_refs` one `[Tm, num_mels]` reference for all of them or a sequence with one per utterance.
        The generator's edge handling (zero-padded convs, replicate-padded anti-alias filters) depends on where each
        utterance ends, so plain padding to a common length would change the last ~35 frames.  On the bf16 path the kernels
        take the per-utterance lengths and the whole ragged batch is ONE call (`decode_varlen`); on the fp32 paths
        utterances of equal length are batched together.  Either way every waveform is the one a single-utterance call
        returns.  Speaker embeddings are computed once per distinct reference.
        Returns a list of `[1, T0_i * 1024]` waveforms (or `[T0_i * 1024]` int16 with pcm16) in input order."""
        lat = [t if t.dim() == 2 else t.squeeze(0) for t in latents]
        n = len(lat)
        if n == 0:
            return []
        dev = lat[0].device
        shared = torch.is_tensor(mel_refs)
        mels = [mel_refs] * n if shared else list(mel_refs)
        if len(mels) != n:
            raise RuntimeError(f"decode_ragged: {n} latents but {len(mels)} reference mels")
        # one speaker-encoder pass per distinct reference tensor (the web UI reuses one voice for many sentences)
        spk_of = {}
        for m in mels:
            key = (m.data_ptr(), tuple(m.shape))
            if key not in spk_of:
                mm = m if m.dim() == 3 else m.unsqueeze(0)
                spk_of[key] = self.speaker_embed(mm.to(dev))
        out = [None] * n
        if self.precision == "bf16" and len({int(t.shape[0]) for t in lat}) > 1:
            # bf16 path: ONE batched call, the kernels take the per-utterance lengths (bvg_decode_varlen)
            spk = torch.cat([spk_of[(m.data_ptr(), tuple(m.shape))] for m in mels], 0)
            y = self.decode_varlen(lat, spk=spk, pcm16=pcm16)
            up = self.total_upsample
            for i, t in enumerate(lat):
                L = int(t.shape[0]) * up
                out[i] = y[i, :L] if pcm16 else y[i, :, :L]
            return out
        groups = {}
        for i, t in enumerate(lat):
            groups.setdefault(int(t.shape[0]), []).append(i)
        for T0, idx in groups.items():
            x = torch.stack([lat[i] for i in idx], 0)
            spk = torch.cat([spk_of[(mels[i].data_ptr(), tuple(mels[i].shape))] for i in idx], 0)
            y = self.decode(x, spk=spk, pcm16=pcm16)
            for j, i in enumerate(idx):
                out[i] = y[j]
        return out

