"""Compiles oracle/awq_oracle.c -> oracle/libawq_oracle.so (gcc, OpenMP).  Test infrastructure.

The reference is pure Python, so there is nothing to compile into oracle/_ref/: the "real reference"
strengthening of the oracle is done by importing it (tests/golden/make_golden.py) instead."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "awq_oracle.c")
LIB = os.path.join(HERE, "libawq_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fexcess-precision=standard",
           "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
