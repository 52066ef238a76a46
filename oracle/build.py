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
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-Wall",
           "-Wextra", "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
