"""Builds lib/libazb.so with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIB = os.path.join(_HERE, "lib", "libazb.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",  # every contraction in the search kernels is written out (bit parity with the reference's f32)
    "-Xcompiler", "-fPIC", "-shared",
    # the CUDA runtime is linked dynamically (libcudart.so.12 is in the image's ld cache; the rpath covers a bare box):
    # a statically linked runtime embeds the name of every runtime entry point in the shipped .so
    "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
]


def lib_path() -> str:
    return os.environ.get("AZB_LIB", _LIB)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; azdopt_b200 has no CPU fallback")


def _stale() -> bool:
    if not os.path.exists(_LIB):
        return True
    t = os.path.getmtime(_LIB)
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)]
    srcs.append(os.path.join(_HERE, "..", "include", "azb.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build_variant(name: str, defines=()) -> str:
    """lib/libazb_<name>.so: the same library with extra -D flags (select it with AZB_LIB=<path>)."""
    out = os.path.join(_HERE, "lib", f"libazb_{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call([_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, os.path.join(_CSRC, "azb.cu")])
    return out


def build_profile_flavour() -> str:
    """lib/libazb_prof.so: -DAZB_PROFILE (per-phase clock64 accumulators; tools/phase_probe.py, tools/async_probe.py)."""
    return build_variant("prof", ["AZB_PROFILE"])


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree; returns its path."""
    if not force and not _stale():
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", _LIB, os.path.join(_CSRC, "azb.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return _LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
