"""TEST / BENCH INFRASTRUCTURE -- builds oracle/_build/libriptrm_det.so from oracle/c/riptrm_det.c with gcc.
-ffp-contract=off: fused multiply-adds only where the source writes fma() (the arithmetic specification)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_build")
LIB = os.path.join(OUT_DIR, "libriptrm_det.so")
SRC = os.path.join(HERE, "riptrm_det.c")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-mfma",
           "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
