"""Build the C restatement of the oracle (oracle/lbm_oracle_c.c) into oracle/_build/liblbm_oracle_c.so.
TEST INFRASTRUCTURE ONLY.  -ffp-contract=off and no fast-math: it must reproduce the NumPy oracle bit for bit."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lbm_oracle_c.c")
OUT = os.path.join(HERE, "_build", "liblbm_oracle_c.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["/usr/bin/gcc", "-O3", "-funroll-loops", "-ffp-contract=off", "-fno-fast-math", "-march=x86-64-v3", "-fopenmp", "-shared",
                           "-fPIC", "-o", OUT, SRC, "-lm"])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
