#!/usr/bin/env python3
"""Table of the outside pins of the CPU oracle (no GPU): lowest eigenvalue of the oracle's product for each model against
the Bethe-ansatz energy of the ring (tests/bethe.py) or the exact-diagonalisation literature.  The same numbers are asserted
by tests/test_oracle_pins.py; this prints them.   python tools/pins_table.py > profiles/r02_oracle_pins.md"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
from scipy.sparse.linalg import LinearOperator, eigsh  # noqa: E402

import bethe  # noqa: E402
from oracle import model as omodel  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

ROWS = [  # model, branch of computeOffDiag, independent value in the file's units, source
    ("heisenberg_chain_4", "a", 4 * bethe.heisenberg_ring_e0(4), "Bethe ansatz"),
    ("heisenberg_chain_6", "a", 4 * bethe.heisenberg_ring_e0(6), "Bethe ansatz"),
    ("heisenberg_chain_8", "a", 4 * bethe.heisenberg_ring_e0(8), "Bethe ansatz"),
    ("heisenberg_chain_10", "b (inversion -1)", 4 * bethe.heisenberg_ring_e0(10), "Bethe ansatz"),
    ("heisenberg_chain_12", "a (identity index)", 4 * bethe.heisenberg_ring_e0(12), "Bethe ansatz (literature: -5.387390917 J)"),
    ("heisenberg_chain_16", "a", 4 * bethe.heisenberg_ring_e0(16), "Bethe ansatz (literature: -7.142296361 J)"),
    ("heisenberg_chain_20", "a", 4 * bethe.heisenberg_ring_e0(20), "Bethe ansatz (literature: -8.90438653 J)"),
    ("heisenberg_chain_24_symm", "c (|G| = 96)", 4 * bethe.heisenberg_ring_e0(24), "Bethe ansatz (literature: -10.6700145 J)"),
    ("heisenberg_square_4x4", "c (|G| = 256)", 4 * 16 * -0.7017802, "literature: E0 / N = -0.7017802 J"),
    ("heisenberg_kagome_12_symm", "c (|G| = 2)", 12 * -0.45374, "literature: E0 / N = -0.45374 J"),
]


def main():
    print("# Outside pins of the CPU oracle: lowest eigenvalue of its product against Bethe ansatz / literature\n")
    print("Energies in the units of the model files (sigma-form files: 4 J; heisenberg_kagome_12_symm: J). Asserted by "
          "`tests/test_oracle_pins.py` (`test_ring_ground_state_equals_bethe_ansatz`, "
          "`test_ground_state_energies_from_the_literature`); at full size: `r02_oracle_chain32_bethe.log`, "
          "`r02_oracle_chain36_bethe.log`; through the CUDA kernels: `tests/test_zz_literature_gpu.py`.\n")
    print("| model | states | branch (BO:82-213) | oracle E0 | independent value | source | difference |\n|---|---|---|---|---|---|---|")
    for name, branch, want, source in ROWS:
        basis, matrix = omodel.load_model(os.path.join(ROOT, "data", name + ".yaml"))
        reps, _ = po.enumerate_states(basis)
        N = reps.shape[0]
        op = LinearOperator((N, N), dtype=np.float64,
                            matvec=lambda v: po.matvec_global(matrix, reps, np.ascontiguousarray(v.ravel()), 1))
        if N <= 600:
            val = np.linalg.eigvalsh(np.array([op.matvec(e) for e in np.eye(N)]).T)[0]
        else:
            val = eigsh(op, k=1, which="SA", tol=1e-13)[0][0]
        print(f"| {name} | {N} | {branch} | {val:.10f} | {want:.10f} | {source} | {val - want:+.1e} |", flush=True)


if __name__ == "__main__":
    main()
