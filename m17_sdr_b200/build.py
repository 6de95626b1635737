"""Build libm17b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.environ.get("M17B_LIB") or os.path.join(PKG, "libm17b200.so")      # M17B_LIB: kernel-tuning variants built side by side
SRC = os.path.join(PKG, "csrc", "m17b200.cu")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # parity: the reference build never contracts a*b+c (generic x86-64, makefile:6)
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc"))] + [os.path.join(ROOT, "include", "m17b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    """Compile the CUDA library if sources are newer than the .so.  Raises if nvcc is unavailable."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libm17b200.so cannot be built (there is no CPU fallback)")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("M17B_NVCC_EXTRA", "").split() + ["-I", os.path.join(ROOT, "include"), "-o", LIB, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
