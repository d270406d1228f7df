"""In-tree build of libpimc_b200.so (nvcc, sm_100a only).  The .so is git-ignored but travels to the GPU box."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# PIMCB_LIB_PATH selects another build of the same sources (A/B variants compiled with different -D switches by
# tools/build_variants.sh); the default is the in-tree product library.
LIB = os.environ.get("PIMCB_LIB_PATH") or os.path.join(_HERE, "libpimc_b200.so")
SOURCES = [os.path.join(CSRC, "pimcb.cu")]
HEADERS = [os.path.join(CSRC, "kernels.cuh"), os.path.join(CSRC, "kernels_ext.cuh"), os.path.join(CSRC, "kernels_pair.cuh"), os.path.join(CSRC, "table_codec.h"), os.path.join(_HERE, "..", "include", "pimc_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--fmad=true", "-cudart", "shared"]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if os.environ.get("PIMCB_LIB_PATH"):
        return LIB                      # a prebuilt variant: never rebuilt implicitly
    if force or is_stale():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        extra = os.environ.get("PIMCB_NVCC_EXTRA", "").split()
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_lib(force=True, verbose="-v" in sys.argv))
