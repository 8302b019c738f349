"""Compile the C half of the oracle (TEST INFRASTRUCTURE).

``python -m oracle.build`` or ``oracle.build.build()``; called from
``__graft_entry__.build()``.  Output: ``oracle/_build/liboracle_ncc.so``
(git-ignored, travels to the GPU box with the gpurun snapshot).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_ncc.so")


def build(force=False):
    src = os.path.join(HERE, "ncc_exact.c")
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(src)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    # -march=x86-64-v3 (AVX2) rather than native: the .so is built in the CPU
    # container and executed on the GPU box, whose host CPU may differ.
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC",
           "-o", LIB, src]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
