"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE, see oracle/oracle.c header).

`python -m oracle.build` or `oracle.build.build()` -> oracle/liboracle.so.

There is no oracle/_ref/: the reference needs the Chapel compiler, GHC-built
liblattice_symmetries_haskell and HDF5, none of which exist in this image (DESIGN.md "Oracle").
-march=x86-64-v3 (not native): the .so is built in the CPU container and travels to the GPU box.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle.c")
LIB = os.path.join(HERE, "liboracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    # into a private file, then an atomic rename: several ranks of one box may get here together (bench.py checks parity
    # on every rank), and none of them may ever load a half-written library
    tmp = f"{LIB}.tmp{os.getpid()}"
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-Wall",
           "-Wextra", "-o", tmp, SRC, "-lm"]
    try:
        subprocess.run(cmd, check=True)
        os.replace(tmp, LIB)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
