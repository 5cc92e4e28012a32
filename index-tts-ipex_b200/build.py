"""Builds libbigvgan_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

The reference JIT-builds its one kernel through torch cpp_extension with an sm_80 gencode
(alias_free_activation/cuda/load.py:49-133); here the library is a plain C-ABI shared object so it
travels with the source tree and is loaded through ctypes (capi.py)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libbigvgan_b200.so"
SOURCES = ["act1d.cu", "act1d_c8t.cu", "act1d_tc.cu", "actconv_tc.cu", "conv_simt.cu", "conv_umma.cu", "conv_umma_fused.cu", "ecapa.cu", "mel.cu", "misc.cu", "plan.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", 
    "-cudart", "static",
]


if os.environ.get("BVG_DEBUG_BUILD") == "1":      # experiments only: enables the result-corrupting dry-run switches
    NVCC_FLAGS = NVCC_FLAGS + ["-DBVG_DEBUG"]
if os.environ.get("BVG_EXTRA_NVCC"):              # experiments only: extra -D switches (tile geometry sweeps)
    NVCC_FLAGS = NVCC_FLAGS + os.environ["BVG_EXTRA_NVCC"].split()


def lib_path() -> str:
    return os.path.join(LIBDIR, LIBNAME)


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; libbigvgan_b200 needs the CUDA toolkit to build")


def _src_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [
        os.path.join(os.path.dirname(HERE), "include", "bigvgan_b200.h"), os.path.abspath(__file__)]
    h.update(" ".join(NVCC_FLAGS).encode())
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def _stale() -> bool:
    # content hash, not mtimes: the tree is copied to the GPU box and mtimes need not survive
    out = lib_path()
    stamp = out + ".srchash"
    if not os.path.exists(out) or not os.path.exists(stamp):
        return True
    return open(stamp).read().strip() != _src_hash()


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return lib_path()
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
        objs.append(obj)
    tmp = lib_path() + ".tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", tmp, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, lib_path())
    with open(lib_path() + ".srchash", "w") as f:
        f.write(_src_hash())
    return lib_path()


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
