"""ctypes binding of include/bigvgan_b200.h (the C-ABI drop-in boundary).

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that
fails, import of this module raises.  Every non-zero status becomes a RuntimeError carrying
bvg_last_error(), mirroring the reference where AT_ERROR surfaces as RuntimeError
(alias_free_activation/cuda/type_shim.h:41-42; load.py:51-52 raises RuntimeError without CUDA)."""
import ctypes as C
import os

from . import build as _build

BVG_F32, BVG_BF16, BVG_F16, BVG_F32X3 = 0, 1, 2, 3


class BvgConfig(C.Structure):
    _fields_ = [
        ("gpt_dim", C.c_int32),
        ("upsample_initial_channel", C.c_int32),
        ("num_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * 8),
        ("upsample_kernel_sizes", C.c_int32 * 8),
        ("num_kernels", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * 4),
        ("resblock_dilation_sizes", (C.c_int32 * 3) * 4),
        ("speaker_embedding_dim", C.c_int32),
        ("num_mels", C.c_int32),
        ("cond_in_each_up_layer", C.c_int32),
        ("snake_logscale", C.c_int32),
        ("device", C.c_int32),
    ]


_vp, _i64, _int, _f32p = C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float)

# name -> (restype, argtypes); must list every symbol include/bigvgan_b200.h declares
SIGNATURES = {
    "bvg_last_error": (C.c_char_p, []),
    "bvg_version": (C.c_char_p, []),
    "bvg_launch_count": (_i64, []),
    "bvg_launch_count_reset": (None, []),
    "bvg_profile_begin": (None, []),
    "bvg_profile_end": (_int, [_f32p, C.POINTER(C.c_int64)]),
    "bvg_act1d_fwd": (_int, [_vp, _vp, _vp, _vp, _f32p, _f32p, _i64, _i64, _i64, _int, _int, _vp]),
    "bvg_conv1d_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_float, _i64, _i64, _i64, _i64, _int, _int, _int, _int, _vp]),
    "bvg_convtr1d_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _vp]),
    "bvg_act1d_c8t_fwd": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "bvg_act1d_c8t_impl_fwd": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "bvg_debug_set_umma_counters": (None, [_vp]),
    "bvg_debug_set_tc_min_melems": (None, [_int]),
    "bvg_debug_set_multi_stream": (None, [_int]),
    "bvg_actconv_umma_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, _i64, _i64, _i64, _i64, _int, _int, _vp]),
    "bvg_actconv_impl_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, _i64, _i64, _i64, _i64, _int, _int, _int, _vp]),
    "bvg_mel_frames": (_i64, [_i64, _int]),
    "bvg_mel_frontend": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _int, _int, _vp]),
    "bvg_conv1d_umma_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_float, _i64, _i64, _i64, _i64, _int, _int, _vp]),
    "bvg_convtr1d_umma_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _int, _int, _vp]),
    "bvg_plan_create": (_int, [C.POINTER(_vp), C.POINTER(BvgConfig)]),
    "bvg_plan_destroy": (None, [_vp]),
    "bvg_plan_set_tensor": (_int, [_vp, C.c_char_p, _vp, _i64]),
    "bvg_plan_finalize": (_int, [_vp, _int]),
    "bvg_workspace_bytes": (C.c_size_t, [_vp, _i64, _i64, _i64, _int]),
    "bvg_speaker_embed": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, C.c_size_t, _vp]),
    "bvg_decode": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _vp, _vp, _i64, _i64, _vp, C.c_size_t, _vp]),
    "bvg_decode_lat": (_int, [_vp, _vp, _int, _vp, _vp, _i64, _i64, _i64, _i64, _int, _vp, _vp, _i64, _i64, _vp, C.c_size_t, _vp]),
    "bvg_decode_varlen": (_int, [_vp, _vp, _int, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, C.c_size_t, _vp]),
    "bvg_decode_host": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _vp, _vp, _vp, C.c_size_t, _vp]),
}

_lib = None


def library_path() -> str:
    return _build.lib_path()


def lib():
    """Load (building first if needed) libbigvgan_b200.so.  Raises if it cannot be had."""
    global _lib
    if _lib is None:
        path = _build.build_library() if os.environ.get("BVG_NO_BUILD") != "1" else _build.lib_path()
        if not os.path.exists(path):
            raise RuntimeError(f"libbigvgan_b200 not found at {path}: the CUDA extension is required "
                               "(there is no CPU fallback)")
        l = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().bvg_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libbigvgan_b200 {what} failed (status {status}): {msg}")


def dtype_code(torch_dtype) -> int:
    import torch
    try:
        return {torch.float32: BVG_F32, torch.bfloat16: BVG_BF16, torch.float16: BVG_F16}[torch_dtype]
    except KeyError:
        raise RuntimeError(f"libbigvgan_b200: unsupported dtype {torch_dtype}")


def launch_count() -> int:
    return int(lib().bvg_launch_count())


def launch_count_reset() -> None:
    lib().bvg_launch_count_reset()


KERNEL_CLASSES = ("act1d", "conv1d", "convtr1d", "other", "actconv")


def profile_begin() -> None:
    lib().bvg_profile_begin()


def profile_end():
    """-> {class: (ms, launches)} for kernels launched since profile_begin()."""
    ms = (C.c_float * len(KERNEL_CLASSES))()
    n = (C.c_int64 * len(KERNEL_CLASSES))()
    check(lib().bvg_profile_end(ms, n), "bvg_profile_end")
    return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_CLASSES)}
