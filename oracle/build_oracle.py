"""ORACLE / TEST INFRASTRUCTURE ONLY.  Compiles oracle/crb_oracle.c into oracle/_ref/liboracle.so with gcc.
(oracle/_ref/ is git-ignored but travels to the GPU box.)  There is no compiled reference to build: the
reference is pure Python over TensorFlow 1.x, which is not installable here."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "liboracle.so")


def build(force=False):
    src = os.path.join(HERE, "crb_oracle.c")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", OUT, src, "-lm"])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
