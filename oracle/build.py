"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Compiles oracle/rnnt_cpu.c into oracle/_build/ (git-ignored)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rnnt_cpu.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "librnnt_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        tmp = OUT + ".tmp%d" % os.getpid()
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", tmp, SRC, "-lm"])
        os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
