#!/usr/bin/env python3
"""Regenerate tests/golden/matvec_golden.npz.

What is pinned and by what:
  /x   REFERENCE RECIPE.  The reference's generator of data/matvec/*.h5 (input_for_matvec.py:8,31,49-76) seeds numpy's
       global RandomState with 42 and draws x = rand(N, 1) - 0.5 file after file in the order of its main(); the stream
       position of every file therefore depends only on the basis dimensions N of the files before it.  This script
       replays that stream with the dimensions of our enumeration (they agree with SURVEY.md section 8: 6, 20, 70, 126,
       4096, 12870, 184756, 2704156, 28968, 924, 472, 12870, 107, 5200300), so /x of a model here is bit-for-bit the /x
       of the reference's HDF5 file -- if the reference's files ever become available, /x must match exactly.
  /y   OUR ORACLE (oracle/oracle.c, the C restatement of the reference algorithm), NOT reference output: the reference
       cannot be built in this image (Chapel + GHC-built lattice_symmetries missing, files fetched from surfdrive at
       test time).  Parity against reference artefacts therefore stays UNPINNED; these vectors pin regressions of the
       oracle and give the GPU tests fixed inputs in the reference's own format and order.
Large models (chain_20, chain_24, square_5x5) only advance the stream; for chain_20 a digest of x and y is kept.

    python tests/golden/make_golden.py        # needs gcc (builds the oracle); no GPU
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from distributed_matvec_b200 import load_config_from_yaml  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

# order of main() in the reference's input_for_matvec.py:49-76
ORDER = ["heisenberg_chain_4", "heisenberg_chain_6", "heisenberg_chain_8", "heisenberg_chain_10",
         "heisenberg_chain_12", "heisenberg_chain_16", "heisenberg_chain_20", "heisenberg_chain_24",
         "heisenberg_chain_24_symm", "heisenberg_kagome_12", "heisenberg_kagome_12_symm", "heisenberg_kagome_16",
         "heisenberg_square_4x4", "heisenberg_square_5x5"]
KNOWN_DIMENSIONS = {"heisenberg_chain_20": 184756, "heisenberg_chain_24": 2704156, "heisenberg_square_5x5": 5200300}
FULL = [n for n in ORDER if n not in KNOWN_DIMENSIONS]          # x, y and representatives stored in full
DIGEST = ["heisenberg_chain_20"]                                  # digests only


def replay():
    """Yields (name, x) for every model in the reference's order, replaying its RandomState stream."""
    rs = np.random.RandomState(42)            # np.random.seed(42) + np.random.rand: the legacy global stream
    for name in ORDER:
        if name in KNOWN_DIMENSIONS:
            n = KNOWN_DIMENSIONS[name]
            reps = None
        else:
            basis, _ = load_config_from_yaml(os.path.join(ROOT, "data", name + ".yaml"))
            reps = po.enumerate_states(basis)[0]
            n = reps.shape[0]
        x = rs.rand(n, 1)[:, 0] - 0.5
        yield name, reps, x


def main():
    out = {}
    for name, reps, x in replay():
        if name in FULL:
            _, matrix = load_config_from_yaml(os.path.join(ROOT, "data", name + ".yaml"))
            y = po.matvec_global(matrix, reps, x, 1)
            out[name + "/representatives"], out[name + "/x"], out[name + "/y"] = reps, x, y
        elif name in DIGEST:
            basis, matrix = load_config_from_yaml(os.path.join(ROOT, "data", name + ".yaml"))
            reps = po.enumerate_states(basis)[0]
            assert reps.shape[0] == x.shape[0]
            y = po.matvec_global(matrix, reps, x, 1)
            out[name + "/digest"] = np.array([x.sum(), y.sum(), np.abs(y).max(), y[0], y[-1]])
            out[name + "/x_sha256"] = np.frombuffer(hashlib.sha256(x.tobytes()).digest(), dtype=np.uint8)
        print(name, x.shape[0], flush=True)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "matvec_golden.npz"), **out)


if __name__ == "__main__":
    main()
