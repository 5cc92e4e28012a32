"""Op boundary: drop-in for the reference's fused anti-aliased activation.

Mirrors `alias_free_activation/cuda/activation1d.py` of the reference: the module-level function
`forward(inputs, up_ftr, down_ftr, alpha, beta)` has the signature of the pybind op
`anti_alias_activation_cuda.forward` (anti_alias_activation.cpp:19-23), and `Activation1d` has the
constructor / forward of activation1d.py:34-76.  The arithmetic follows the PyTorch module
(alias_free_torch/act.py:24-29), edges included."""
import ctypes as C

import torch
import torch.nn as nn

from . import capi


# Filter buffers already validated against the compiled-in kaiser-sinc taps, keyed by (storage pointer, version): the
# check needs the values on the host (one blocking D2H copy), so it runs once per buffer, not once per call -- a
# generator pass makes ~100 Activation1d calls, and a per-call copy would serialise the stream and break graph capture.
_validated_taps = set()


def _taps_ptr(t):
    """-> ctypes float[12] to validate, or None when `t` is absent or was validated before."""
    if t is None:
        return None
    key = (t.data_ptr(), t._version, t.device)
    if key in _validated_taps:
        return None
    host = t.detach().reshape(-1).to("cpu", torch.float32).contiguous()
    if host.numel() != 12:
        raise RuntimeError("anti_alias_activation: the fused kernel hard-codes filter size 12 / ratio 2")
    return (C.c_float * 12)(*host.tolist())


def forward(inputs: torch.Tensor, up_ftr, down_ftr, alpha: torch.Tensor, beta: torch.Tensor,
            precise=None) -> torch.Tensor:
    """inputs [B,C,T] contiguous CUDA float/half/bf16; alpha, beta fp32 [C] LOG-scale.
    Returns a new tensor of the same shape/dtype (requires_grad False), like fwd_cuda
    (anti_alias_activation_cuda.cu:214-256).  Runs on torch's current stream."""
    if not inputs.is_cuda:
        raise RuntimeError("anti_alias_activation: inputs must be a CUDA tensor (no CPU fallback)")
    if inputs.dim() != 3:
        raise RuntimeError("anti_alias_activation: expected [B, C, T]")
    x = inputs.contiguous()
    B, Cn, T = x.shape
    a = alpha.detach().to(x.device, torch.float32).contiguous()
    b = beta.detach().to(x.device, torch.float32).contiguous()
    if a.numel() != Cn or b.numel() != Cn:
        raise RuntimeError("anti_alias_activation: alpha/beta must have C elements")
    out = torch.empty_like(x)
    if precise is None:
        precise = x.dtype == torch.float32
    up, down = _taps_ptr(up_ftr), _taps_ptr(down_ftr)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        capi.check(capi.lib().bvg_act1d_fwd(out.data_ptr(), x.data_ptr(), a.data_ptr(), b.data_ptr(), up, down,
                                            B, Cn, T, capi.dtype_code(x.dtype), int(bool(precise)), st),
                   "bvg_act1d_fwd")
    for t in (up_ftr, down_ftr):          # the library accepted them: remember, so later calls skip the host copy
        if t is not None:
            _validated_taps.add((t.data_ptr(), t._version, t.device))
    return out


class FusedAntiAliasActivation(torch.autograd.Function):
    """activation1d.py:13-31 of the reference: forward only."""

    @staticmethod
    def forward(ctx, inputs, up_ftr, down_ftr, alpha, beta):
        return forward(inputs, up_ftr, down_ftr, alpha, beta)

    @staticmethod
    def backward(ctx, output_grads):
        raise NotImplementedError


class Activation1d(nn.Module):
    """Same constructor and forward contract as the reference's cuda/activation1d.py:34-76.
    `activation` is a Snake / SnakeBeta-like module exposing `.alpha`, (`.beta`), `.alpha_logscale`."""

    def __init__(self, activation, up_ratio: int = 2, down_ratio: int = 2, up_kernel_size: int = 12,
                 down_kernel_size: int = 12, fused: bool = True):
        super().__init__()
        if (up_ratio, down_ratio, up_kernel_size, down_kernel_size) != (2, 2, 12, 12):
            raise NotImplementedError("the fused kernel assumes ratio 2 and filter size 12")
        if not fused:
            raise NotImplementedError("only the fused CUDA path exists in this package")
        self.up_ratio, self.down_ratio = up_ratio, down_ratio
        self.act = activation
        from .models import KAISER_TAPS
        taps = torch.tensor(KAISER_TAPS, dtype=torch.float32).view(1, 1, 12)
        # same buffer names as the reference so state dicts line up
        self.upsample = nn.Module()
        self.upsample.register_buffer("filter", taps.clone())
        self.downsample = nn.Module()
        self.downsample.lowpass = nn.Module()
        self.downsample.lowpass.register_buffer("filter", taps.clone())

    def forward(self, x):
        alpha = self.act.alpha.data
        beta = self.act.alpha.data if self.act.__class__.__name__ == "Snake" else self.act.beta.data
        if not self.act.alpha_logscale:      # exp is baked into the kernel: cancel it (activation1d.py:67-71)
            alpha, beta = torch.log(alpha), torch.log(beta)
        return FusedAntiAliasActivation.apply(x, self.upsample.filter, self.downsample.lowpass.filter, alpha, beta)
