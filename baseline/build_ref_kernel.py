"""Builds the reference's own fused Activation1d CUDA op -- the "existing GPU kernel to beat" of BASELINE.md section 4 -- for
sm_100, straight from the sources where they lie under /root/reference (never copied into this repo):

    /root/reference/indextts/BigVGAN/alias_free_activation/cuda/anti_alias_activation.cpp
    /root/reference/indextts/BigVGAN/alias_free_activation/cuda/anti_alias_activation_cuda.cu

with the reference's own flags (load.py:86-108: -O3 --use_fast_math, half operators enabled) except for the gencode, which the
reference pins to sm_80 and which is sm_100 here.  The output goes to baseline/_ref/ (git-ignored, travels to the GPU box);
bench.py times it beside this repo's kernels when it is present (bench-only: nothing in the product imports it).

    python baseline/build_ref_kernel.py          (in the build container; no GPU needed)"""
import os
import pathlib
import sys

HERE = pathlib.Path(__file__).resolve().parent
SRC = pathlib.Path("/root/reference/indextts/BigVGAN/alias_free_activation/cuda")
OUT = HERE / "_ref" / "anti_alias_activation_cuda"
NAME = "anti_alias_activation_cuda"


def so_path() -> pathlib.Path:
    return OUT / f"{NAME}.so"


def build(verbose: bool = False) -> pathlib.Path:
    os.environ["TORCH_CUDA_ARCH_LIST"] = ""           # (as load.py:16: the arch comes from the explicit gencode below)
    from torch.utils import cpp_extension
    if not SRC.exists():
        raise RuntimeError(f"{SRC} not found: the reference sources only exist in the build container")
    OUT.mkdir(parents=True, exist_ok=True)
    cpp_extension.load(
        name=NAME,
        sources=[str(SRC / "anti_alias_activation.cpp"), str(SRC / "anti_alias_activation_cuda.cu")],
        build_directory=str(OUT),
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "--use_fast_math", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                           "--expt-relaxed-constexpr", "--expt-extended-lambda", "-gencode", "arch=compute_100,code=sm_100"],
        verbose=verbose,
    )
    return so_path()


def load():
    """The built module (torch extension), or None when it was never built."""
    p = so_path()
    if not p.exists():
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(NAME, str(p))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
