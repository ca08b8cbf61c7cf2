"""Builds csrc/libriptrm_b200.so in-tree with nvcc for sm_100a (B200).

    python riemannian-interior-point-trust-region-method_b200/build.py [--force]

The library is the product's only compute path; nothing falls back to the CPU.
-fmad=false: fused multiply-adds occur only where the kernels write fma() explicitly
(dot products / matrix-vector products); see csrc/common.cuh.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libriptrm_b200.so")
SOURCES = ["riptrm_api.cu"]
# No --split-compile: round 2 saw two builds of unchanged sources with `--split-compile 0` come out with different code
# (different stack frames in sphere_tmem2_kernel, the floating-point contraction of an -fmad=true build in some functions)
# and wrong results on the GPU; the single-job compile is deterministic (3.5 min on 8 cores instead of 1.3).
NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo", "-fmad=false",
    "-gencode", "arch=compute_100a,code=sm_100a",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "riptrm_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("RIPTRM_NVCC_EXTRA", "").split()   # e.g. -DRIPTRM_COLUMNS_TIMING (diagnostic builds)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [
        os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
