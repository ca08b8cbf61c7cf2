"""TEST / BENCH INFRASTRUCTURE -- builds oracle/_build/libriptrm_det.so from oracle/c/riptrm_det.c with gcc.
-ffp-contract=off: fused multiply-adds only where the source writes fma() (the arithmetic specification)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_build")
LIB = os.path.join(OUT_DIR, "libriptrm_det.so")
SRC = os.path.join(HERE, "riptrm_det.c")


LIB_FAITHFUL = os.path.join(OUT_DIR, "libriptrm_det_faithful.so")


def build(force=False):
    """libriptrm_det.so (merged-reduction tCG, the arithmetic the CUDA kernel implements) and
    libriptrm_det_faithful.so (-DFAITHFUL_TCG: the tCG loop in the reference's operation order)."""
    if not force and all(os.path.exists(l) and os.path.getmtime(l) >= os.path.getmtime(SRC) for l in (LIB, LIB_FAITHFUL)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    base = ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-mfma"]
    subprocess.run(base + ["-o", LIB, SRC, "-lm"], check=True)
    subprocess.run(base + ["-DFAITHFUL_TCG", "-o", LIB_FAITHFUL, SRC, "-lm"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
