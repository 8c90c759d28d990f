"""Build atropos_b200/libatropos_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m atropos_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libatropos_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unknown-pragmas", "-shared", "--expt-relaxed-constexpr",
         "-Xptxas", "-v" if os.environ.get("ATR_PTXAS_V") else "-O3"]


def sources():
    # .cu: the kernels + C ABI (nvcc); .cpp: host-only parts of the ABI (the host compiler through nvcc)
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cpp"))]


def deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(ROOT, "include", "atropos_b200.h"))
    return d


def up_to_date():
    return os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(p) for p in deps())


def build_variant(out, defines, verbose=False):
    """A second build with -D overrides (kernel-variant A/B runs: ATROPOS_B200_LIB=<out> selects it)."""
    cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + ["-I", os.path.join(ROOT, "include"), "-o", out] + sources()
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [NVCC] + FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", LIB] + sources()
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
